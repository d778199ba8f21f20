"""farthest_point_sampling -- mirror of libs/pointops/functions/sampling.py:7-27."""
from __future__ import annotations

import torch
from torch.autograd import Function

import os

from .. import _lib
from . import _common as C

USE_GRID = os.environ.get("POINTOPS_B200_FPS_GRID", "1") != "0"
CLUSTER_HINT = int(os.environ.get("POINTOPS_B200_FPS_CLUSTER", "0"))   # 0 = the library chooses; 1/2/4/8/16 force
# schedule of the kernel (include/pointops_b200.h POB_FPS_*): same samples, bit for bit, whatever is chosen
VARIANTS = {"auto": 0, "merge": 1, "chain": 2, "single": 3, "merge_cells": 4}
VARIANT = VARIANTS[os.environ.get("POINTOPS_B200_FPS", "auto")]
RESIDENT_MAX = 131072   # points per scene the cluster-resident kernels hold in registers
STATS = None            # diagnostics (bench.py): device int64[4] every launch accumulates {rounds, samples, distances, -} into


def fps_launch(xyz, offset, new_offset, offset_host, new_offset_host, variant=None, stats=None, cluster_hint=None):
    """The launch itself, for callers that already validated their tensors and know the host copies
    of both offset vectors (the PTv1 mirror's geometry pass): no checks, no autograd node.
    variant / stats / cluster_hint are per-call (the library keeps no tuning state)."""
    sizes = C.scene_sizes(offset_host)
    m = new_offset_host[-1]
    n_max = max(sizes) if sizes else 0
    idx = torch.empty((m,), dtype=torch.int32, device=xyz.device)
    if m == 0:
        return idx
    tmp = None
    if n_max > RESIDENT_MAX:  # beyond the register-resident capacity the kernel streams through tmp
        tmp = torch.empty((xyz.shape[0],), dtype=torch.float32, device=xyz.device)
    with _lib.device_guard(xyz.device):
        flops = sum(10 * max(mb - 1, 0) * nb for nb, mb in zip(sizes, C.scene_sizes(new_offset_host)))
        # the search grid of this cloud (built once, reused by the kNN queries that follow in
        # TransitionDown) gives the kernel cell-ordered points for exact pruning
        grid = C.get_grid(xyz, offset) if (USE_GRID and n_max > 2048) else None
        _lib.run("pob_farthest_point_sampling", offset.numel(), n_max, _lib.ptr(xyz), _lib.ptr(offset), _lib.ptr(new_offset),
                 _lib.ptr(tmp), _lib.ptr(idx), CLUSTER_HINT if cluster_hint is None else int(cluster_hint),
                 _lib.ptr(grid.workspace if grid else None), xyz.shape[0], grid.cell_pts if grid else 0.0,
                 VARIANT if variant is None else VARIANTS[variant] if isinstance(variant, str) else int(variant),
                 _lib.ptr(stats if stats is not None else STATS), _lib.current_stream(xyz.device),
                 alg_bytes=12 * xyz.shape[0] + 4 * m, alg_flops=flops)
    return idx


class FarthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, offset, new_offset):
        """
        input: xyz (n, 3) f32, offset (b), new_offset (b) cumulative sample counts
        output: idx (new_offset[-1]) i32, global row indices, scene-major
        """
        C.require(xyz, "xyz", torch.float32, 2, 3)
        offset, new_offset = C.offset_i32(offset, "offset"), C.offset_i32(new_offset, "new_offset")
        C.same_device(("xyz", xyz), ("offset", offset), ("new_offset", new_offset))
        if new_offset.numel() != offset.numel():
            raise ValueError("offset and new_offset must describe the same number of scenes")
        # the host numbers the launch needs; registered by callers that know them, else one
        # .tolist() per offset tensor (the reference syncs once per scene, sampling.py:15-18)
        idx = fps_launch(xyz, offset, new_offset, C.host_offset(offset), C.host_offset(new_offset))
        ctx.mark_non_differentiable(idx)
        return idx


farthest_point_sampling = FarthestPointSampling.apply
