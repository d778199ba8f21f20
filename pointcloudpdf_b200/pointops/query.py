"""knn_query -- mirror of libs/pointops/functions/query.py:7-24,111."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from . import _common as C


class KNNQuery(Function):
    @staticmethod
    def forward(ctx, nsample, xyz, offset, new_xyz=None, new_offset=None):
        """
        input: xyz (n, 3) f32, offset (b), new_xyz (m, 3) f32, new_offset (b)
        output: idx (m, nsample) i32, -1 is the placeholder; dist (m, nsample) f32 = sqrt(d2),
                1e5 for placeholders.  Neighbours ordered by (d2, idx); no gradient.
        """
        if new_xyz is None or new_offset is None:  # functions/query.py:14-16
            new_xyz, new_offset = xyz, offset
        nsample = int(nsample)
        if nsample < 1 or nsample > 256:
            raise ValueError(f"nsample must be in [1, 256], got {nsample}")
        C.require(xyz, "xyz", torch.float32, 2, 3)
        C.require(new_xyz, "new_xyz", torch.float32, 2, 3)
        offset, new_offset = C.offset_i32(offset, "offset"), C.offset_i32(new_offset, "new_offset")
        C.same_device(("xyz", xyz), ("new_xyz", new_xyz), ("offset", offset), ("new_offset", new_offset))
        if offset.numel() != new_offset.numel():
            raise ValueError("offset and new_offset must describe the same number of scenes")
        with _lib.device_guard(xyz.device):
            # fresh tensors per call, like the reference: an identical query (PTv1 repeats its self-kNN in
            # every block of a stage, point_transformer_seg.py:51) is answered by copying the cached result,
            # so no two callers ever share storage
            idx, dist, _ = C.cached_knn(nsample, xyz, offset, new_xyz, new_offset)
            idx, dist = idx.clone(), dist.clone()
        ctx.mark_non_differentiable(idx, dist)
        return idx, dist


knn_query = KNNQuery.apply


def _ball_args(nsample, max_radius, min_radius, xyz, offset, new_xyz, new_offset):
    if new_xyz is None or new_offset is None:  # functions/query.py:39-41, 90-92
        new_xyz, new_offset = xyz, offset
    nsample = int(nsample)
    if nsample < 1:
        raise ValueError(f"nsample must be positive, got {nsample}")
    if not float(min_radius) < float(max_radius):   # the reference asserts this (functions/query.py:43,94)
        raise ValueError("min_radius must be smaller than max_radius")
    C.require(xyz, "xyz", torch.float32, 2, 3)
    C.require(new_xyz, "new_xyz", torch.float32, 2, 3)
    offset, new_offset = C.offset_i32(offset, "offset"), C.offset_i32(new_offset, "new_offset")
    C.same_device(("xyz", xyz), ("new_xyz", new_xyz), ("offset", offset), ("new_offset", new_offset))
    if offset.numel() != new_offset.numel():
        raise ValueError("offset and new_offset must describe the same number of scenes")
    return nsample, xyz, offset, new_xyz, new_offset


class BallQuery(Function):
    @staticmethod
    def forward(ctx, nsample, max_radius, min_radius, xyz, offset, new_xyz=None, new_offset=None):
        """Mirror of functions/query.py:77-109.
        input: xyz (n, 3), new_xyz (m, 3), offset (b), new_offset (b)
        output: idx (m, nsample) i32 (-1 = none), sqrt of the kernel's dist2 (m, nsample): accepted points
        (d2 <= 1e-5 or min_r^2 <= d2 < max_r^2) in the order the reference kernel leaves them in (its
        heap_sort runs on an un-heapified list: partially ordered by distance, reproduced exactly); with more
        than nsample of them every (cnt / nsample)-th is kept and "dist2" holds the point index, as there.
        More than 2048 accepted points per query overrun the reference's stack arrays; here that raises."""
        nsample, xyz, offset, new_xyz, new_offset = _ball_args(nsample, max_radius, min_radius, xyz, offset, new_xyz, new_offset)
        m = new_xyz.shape[0]
        dev = xyz.device
        idx = torch.empty((m, nsample), dtype=torch.int32, device=dev)
        dist2 = torch.empty((m, nsample), dtype=torch.float32, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        with _lib.device_guard(dev):
            grid = C.get_grid(xyz, offset)
            _lib.run("pob_ball_query", m, nsample, float(min_radius), float(max_radius), grid.n, grid.b, _lib.ptr(xyz),
                     _lib.ptr(new_xyz), _lib.ptr(new_offset), grid.cell_pts, _lib.ptr(grid.workspace), _lib.ptr(idx),
                     _lib.ptr(dist2), _lib.ptr(flag), _lib.current_stream(dev),
                     alg_bytes=12 * grid.n + 12 * m + 8 * nsample * m)
        if not torch.cuda.is_current_stream_capturing() and int(flag.item()):
            raise RuntimeError("pointops.ball_query: a query accepted more than 2048 points (the reference kernel's "
                               "stack arrays hold 2048: undefined behaviour there); shrink max_radius")
        ctx.mark_non_differentiable(idx)
        return idx, torch.sqrt(dist2)


class RandomBallQuery(Function):
    @staticmethod
    def forward(ctx, nsample, max_radius, min_radius, xyz, offset, new_xyz=None, new_offset=None, order=None):
        """Mirror of functions/query.py:27-74: the first nsample accepted points in the order of a random
        permutation of each scene.  `order` (n, int32 global rows; additive argument) fixes the permutation,
        else it is drawn with torch.randperm per scene like the reference does.
        output: idx (m, nsample) i32 (-1 = none), dist (m, nsample) = sqrt(d2), 1e5 for placeholders."""
        nsample, xyz, offset, new_xyz, new_offset = _ball_args(nsample, max_radius, min_radius, xyz, offset, new_xyz, new_offset)
        if nsample > 256:
            raise ValueError("random_ball_query: nsample must be <= 256")
        n, m, dev = xyz.shape[0], new_xyz.shape[0], xyz.device
        if order is None:
            parts, s = [], 0
            for e in C.host_offset(offset):
                parts.append(torch.randperm(e - s, dtype=torch.int32, device=dev) + s)
                s = e
            order = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int32, device=dev)
        C.require(order, "order", torch.int32, 1)
        if order.numel() != n:
            raise ValueError("order must hold one entry per row of xyz")
        idx = torch.empty((m, nsample), dtype=torch.int32, device=dev)
        dist2 = torch.empty((m, nsample), dtype=torch.float32, device=dev)
        inv = torch.empty((n,), dtype=torch.int32, device=dev)
        with _lib.device_guard(dev):
            grid = C.get_grid(xyz, offset)
            _lib.run("pob_random_ball_query", m, nsample, float(min_radius), float(max_radius), grid.n, grid.b,
                     _lib.ptr(order), _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(new_offset), grid.cell_pts,
                     _lib.ptr(grid.workspace), _lib.ptr(inv), _lib.ptr(idx), _lib.ptr(dist2), _lib.current_stream(dev),
                     alg_bytes=20 * n + 12 * m + 8 * nsample * m)
        ctx.mark_non_differentiable(idx)
        return idx, torch.sqrt(dist2)


ball_query = BallQuery.apply
random_ball_query = RandomBallQuery.apply
