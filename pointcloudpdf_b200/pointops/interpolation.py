"""interpolation / interpolation2 -- mirror of libs/pointops/functions/interpolation.py."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from . import _common as C


class _InterpolateRows(Function):
    """out[n,:] = sum_i input[idx[n,i],:] * weight[n,i]; grad only to input."""

    @staticmethod
    def forward(ctx, input, idx, weight):
        n, k = idx.shape
        m, c = input.shape
        output = torch.empty((n, c), dtype=torch.float32, device=input.device)
        with _lib.device_guard(input.device):
            _lib.run("pob_interpolation_forward", n, c, k, _lib.ptr(input), _lib.ptr(idx), _lib.ptr(weight),
                     _lib.ptr(output), _lib.current_stream(input.device), alg_bytes=4 * (m * c + 2 * n * k + n * c))
        ctx.m = m
        ctx.save_for_backward(idx, weight)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        idx, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous().float()
        n, c = grad_output.shape
        k = idx.shape[1]
        grad_input = torch.zeros((ctx.m, c), dtype=torch.float32, device=grad_output.device)
        with _lib.device_guard(grad_output.device):
            _lib.run("pob_interpolation_backward", n, c, k, _lib.ptr(grad_output), _lib.ptr(idx), _lib.ptr(weight),
                     _lib.ptr(grad_input), _lib.current_stream(grad_output.device),
                     alg_bytes=4 * (ctx.m * c + 2 * n * k + n * c))
        return grad_input, None, None


def _neighbours_and_weights(xyz, new_xyz, offset, new_offset, k):
    C.require(xyz, "xyz", torch.float32, 2, 3)
    C.require(new_xyz, "new_xyz", torch.float32, 2, 3)
    offset, new_offset = C.offset_i32(offset, "offset"), C.offset_i32(new_offset, "new_offset")
    with _lib.device_guard(xyz.device):
        idx, _, weight = C.cached_knn(int(k), xyz, offset, new_xyz, new_offset, want_weight=True)
    return idx, weight


def _wrap_placeholders(idx: torch.Tensor, m: int) -> torch.Tensor:
    # quirk C6 (functions/interpolation.py:21): a placeholder -1 indexes feat[-1] in torch
    return torch.where(idx < 0, idx + m, idx)


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """pointops.interpolation(xyz, new_xyz, feat, offset, new_offset, k=3)
    (functions/interpolation.py:8-22).  xyz/feat: the coarse (source) cloud (m rows); new_xyz:
    the fine cloud (n rows).  out (n, c) = inverse-distance weighted sum over the k nearest coarse
    points.  kNN, the weights (1/(dist+1e-8), normalised) and the gather run in two kernels.
    Differentiable w.r.t. feat."""
    C.require(feat, "feat", (torch.float32, torch.float16, torch.bfloat16), 2)
    if feat.shape[0] != xyz.shape[0]:
        raise ValueError("feat and xyz must have the same number of rows")
    idx, weight = _neighbours_and_weights(xyz, new_xyz, offset, new_offset, k)
    idx = _wrap_placeholders(idx, feat.shape[0])  # only bites when a scene holds < k coarse points
    src = feat if feat.dtype == torch.float32 else feat.float()
    return _InterpolateRows.apply(src.contiguous(), idx, weight)


class Interpolation(Function):
    """pointops.interpolation2 (functions/interpolation.py:25-59)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, input, offset, new_offset, k=3):
        C.require(input, "input", torch.float32, 2)
        idx, weight = _neighbours_and_weights(xyz, new_xyz, offset, new_offset, k)
        idx = _wrap_placeholders(idx, input.shape[0])
        n, c, m = new_xyz.shape[0], input.shape[1], input.shape[0]
        output = torch.empty((n, c), dtype=torch.float32, device=input.device)
        with _lib.device_guard(input.device):
            _lib.run("pob_interpolation_forward", n, c, int(k), _lib.ptr(input), _lib.ptr(idx), _lib.ptr(weight),
                     _lib.ptr(output), _lib.current_stream(input.device), alg_bytes=4 * (m * c + 2 * n * int(k) + n * c))
        ctx.m, ctx.k = m, int(k)
        ctx.save_for_backward(idx, weight)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        idx, weight = ctx.saved_tensors
        grad_output = grad_output.contiguous().float()
        n, c = grad_output.shape
        grad_input = torch.zeros((ctx.m, c), dtype=torch.float32, device=grad_output.device)
        with _lib.device_guard(grad_output.device):
            _lib.run("pob_interpolation_backward", n, c, ctx.k, _lib.ptr(grad_output), _lib.ptr(idx), _lib.ptr(weight),
                     _lib.ptr(grad_input), _lib.current_stream(grad_output.device),
                     alg_bytes=4 * (ctx.m * c + 2 * n * ctx.k + n * c))
        return None, None, grad_input, None, None, None


interpolation2 = Interpolation.apply
