"""grouping / grouping2 -- mirror of libs/pointops/functions/grouping.py."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from . import _common as C

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class Grouping(Function):
    """pointops.grouping2 (functions/grouping.py:7-33): plain gather, f32, idx >= 0."""

    @staticmethod
    def forward(ctx, input, idx):
        """input: (n, c) f32, idx: (m, nsample) i32 -> (m, nsample, c)"""
        C.require(input, "input", torch.float32, 2)
        C.require(idx, "idx", torch.int32, 2)
        C.same_device(("input", input), ("idx", idx))
        m, nsample = idx.shape
        n, c = input.shape
        output = torch.empty((m, nsample, c), dtype=torch.float32, device=input.device)
        with _lib.device_guard(input.device):
            _lib.run("pob_grouping_forward", m, nsample, c, _lib.ptr(input), _lib.ptr(idx), _lib.ptr(output),
                     _lib.current_stream(input.device), alg_bytes=4 * (n * c + m * nsample + m * nsample * c))
        ctx.n = n
        ctx.save_for_backward(idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        grad_output = grad_output.contiguous().float()
        m, nsample, c = grad_output.shape
        grad_input = torch.zeros((ctx.n, c), dtype=torch.float32, device=grad_output.device)
        with _lib.device_guard(grad_output.device):
            _lib.run("pob_grouping_backward", m, nsample, c, _lib.ptr(grad_output), _lib.ptr(idx),
                     _lib.ptr(grad_input), _lib.current_stream(grad_output.device),
                     alg_bytes=4 * (ctx.n * c + m * nsample + m * nsample * c))
        return grad_input, None


grouping2 = Grouping.apply


class _GroupXYZ(Function):
    """The torch-level pointops.grouping (functions/grouping.py:36-60) as one kernel each way."""

    @staticmethod
    def forward(ctx, feat, idx, xyz, new_xyz, with_xyz):
        m, nsample = idx.shape
        n, c = feat.shape
        width = c + (3 if with_xyz else 0)
        out = torch.empty((m, nsample, width), dtype=torch.float32, device=feat.device)
        with _lib.device_guard(feat.device):
            x3 = 3 if with_xyz else 0
            _lib.run("pob_group_xyz_forward", m, nsample, c, 1 if with_xyz else 0, _lib.ptr(feat),
                     _DTYPE_CODE[feat.dtype], _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(idx), _lib.ptr(out),
                     _lib.current_stream(feat.device),
                     alg_bytes=feat.element_size() * n * c + 4 * (x3 * n + x3 * m + m * nsample + m * nsample * width))
        ctx.shape = (n, c, bool(with_xyz), feat.dtype)
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        n, c, with_xyz, dtype = ctx.shape
        grad_out = grad_out.contiguous().float()
        m, nsample = idx.shape
        grad_feat = torch.zeros((n, c), dtype=torch.float32, device=grad_out.device)
        with _lib.device_guard(grad_out.device):
            _lib.run("pob_group_xyz_backward", m, nsample, c, 1 if with_xyz else 0, _lib.ptr(grad_out),
                     _lib.ptr(idx), _lib.ptr(grad_feat), _lib.current_stream(grad_out.device),
                     alg_bytes=4 * (n * c + m * nsample + m * nsample * (c + (3 if with_xyz else 0))))
        return grad_feat.to(dtype), None, None, None, None


def grouping(idx, feat, xyz, new_xyz=None, with_xyz=False):
    """pointops.grouping(idx, feat, xyz, new_xyz=None, with_xyz=False)
    (functions/grouping.py:36-60): gather with idx -1 -> zero row; with_xyz prepends
    (xyz[idx] - new_xyz[m]) masked the same way.  feat may be f32/f16/bf16; the result is f32
    (the reference's cat with an f32 zero row promotes).  Differentiable w.r.t. feat."""
    if new_xyz is None:
        new_xyz = xyz
    C.require(idx, "idx", torch.int32, 2)
    C.require(feat, "feat", tuple(_DTYPE_CODE), 2)
    C.require(xyz, "xyz", torch.float32, 2, 3)
    C.same_device(("idx", idx), ("feat", feat), ("xyz", xyz), ("new_xyz", new_xyz))
    if with_xyz:
        C.require(new_xyz, "new_xyz", torch.float32, 2, 3)
        if new_xyz.shape[0] != idx.shape[0]:
            raise ValueError("new_xyz and idx must have the same number of rows")
    if feat.shape[0] != xyz.shape[0]:
        raise ValueError("feat and xyz must have the same number of rows")
    return _GroupXYZ.apply(feat, idx, xyz, new_xyz, bool(with_xyz))


def grouping_split(idx, feat, xyz, new_xyz=None):
    """Same values as ``grouping(idx, feat, xyz, new_xyz, with_xyz=True)`` but as two tensors,
    ``(rel_xyz (m, ns, 3), grouped_feat (m, ns, c))``: both stay contiguous and 16-byte aligned,
    so callers that slice the concatenated form apart again (PointTransformerLayer,
    point_transformer_seg.py:64) skip the interleave and the strided re-reads.  Additive API."""
    if new_xyz is None:
        new_xyz = xyz
    grouped = grouping(idx, feat, xyz, new_xyz, with_xyz=False)
    C.require(new_xyz, "new_xyz", torch.float32, 2, 3)
    m, nsample = idx.shape
    rel = torch.empty((m, nsample, 3), dtype=torch.float32, device=xyz.device)
    with _lib.device_guard(xyz.device):
        _lib.run("pob_group_relxyz_forward", m, nsample, _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(idx), _lib.ptr(rel),
                 _lib.current_stream(xyz.device), alg_bytes=4 * (3 * xyz.shape[0] + 3 * m + m * nsample + 3 * m * nsample))
    return rel, grouped
