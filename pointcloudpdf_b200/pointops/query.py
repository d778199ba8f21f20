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
            idx, dist, _ = C.cached_knn(nsample, xyz, offset, new_xyz, new_offset)
        ctx.mark_non_differentiable(idx, dist)
        return idx, dist


knn_query = KNNQuery.apply


def ball_query(*args, **kwargs):
    """pointops.ball_query (functions/query.py:77-109) is not on the PTv1 / openseg path
     and is not implemented in this round."""
    raise NotImplementedError("pointcloudpdf_b200.pointops.ball_query: outside the PTv1 hot path (SURVEY.md 8f)")


def random_ball_query(*args, **kwargs):
    """pointops.random_ball_query (functions/query.py:27-74): see ball_query."""
    raise NotImplementedError("pointcloudpdf_b200.pointops.random_ball_query: outside the PTv1 hot path (SURVEY.md 8f)")
