"""knn_query_and_group / query_and_group / offset helpers -- mirror of
libs/pointops/functions/utils.py."""
from __future__ import annotations

import torch

from . import _common as C
from .query import knn_query, ball_query
from .grouping import grouping, grouping2


def knn_query_and_group(feat, xyz, offset=None, new_xyz=None, new_offset=None, idx=None, nsample=None,
                        with_xyz=False):
    """functions/utils.py:5-18 -- the entry point PTv1 calls (point_transformer_seg.py:51-63,
    106-114).  Returns (grouped (m, nsample, [3+]c), idx)."""
    if idx is None:
        assert nsample is not None
        idx, _ = knn_query(nsample, xyz, offset, new_xyz, new_offset)
    return grouping(idx, feat, xyz, new_xyz, with_xyz), idx


def ball_query_and_group(feat, xyz, offset=None, new_xyz=None, new_offset=None, idx=None, max_radio=None,
                         min_radio=0, nsample=None, with_xyz=False):
    """functions/utils.py:21-39: ball query (unless idx is given) + gather with -1 masking."""
    if idx is None:
        idx, _ = ball_query(nsample, max_radio, min_radio, xyz, offset, new_xyz, new_offset)
    return grouping(idx, feat, xyz, new_xyz, with_xyz), idx


def query_and_group(nsample, xyz, new_xyz, feat, idx, offset, new_offset, dilation=0, with_feat=True,
                    with_xyz=True):
    """functions/utils.py:42-99: kNN (optionally dilated) + gather, without -1 masking (a
    placeholder -1 wraps to the last row, as torch indexing does there).
    output: (m, nsample, 3 + c) and idx (m, nsample)."""
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        total = 1 + (nsample - 1) * (dilation + 1)
        full, _ = knn_query(total, xyz, offset, new_xyz, new_offset)
        off = C.host_offset(C.offset_i32(offset))
        noff = C.host_offset(C.offset_i32(new_offset))
        parts, s_n, s_m = [], 0, 0
        for e_n, e_m in zip(off, noff):
            n_b = e_n - s_n
            soft = (n_b - 1) / (nsample - 1) - 1 if n_b < total else dilation
            cols = [int((soft + 1) * i) for i in range(nsample)]
            parts.append(full[s_m:e_m][:, cols])
            s_n, s_m = e_n, e_m
        idx = torch.cat(parts, dim=0).contiguous()
    if not with_feat:
        return idx
    n = xyz.shape[0]
    wrapped = torch.where(idx < 0, idx + n, idx).contiguous()
    if feat.dtype == torch.float32:
        grouped = grouping(wrapped, feat.contiguous(), xyz, new_xyz, with_xyz)
    else:  # dtype-generic in the reference (plain indexing): keep the dtype
        gf = feat[wrapped.long()]
        grouped = torch.cat((xyz[wrapped.long()] - new_xyz.unsqueeze(1), gf), -1) if with_xyz else gf
    return grouped, idx


def offset2batch(offset):
    """functions/utils.py:102-116: scene id per row, int64, on offset's device."""
    sizes = torch.diff(offset, prepend=torch.zeros(1, dtype=offset.dtype, device=offset.device))
    return torch.repeat_interleave(torch.arange(offset.numel(), device=offset.device), sizes.long()).long()


def batch2offset(batch):
    """functions/utils.py:119-120."""
    return torch.cumsum(batch.bincount(), dim=0).int()
