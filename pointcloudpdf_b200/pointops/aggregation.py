"""aggregation -- mirror of libs/pointops/functions/aggregation.py:7-57."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from . import _common as C


class Aggregation(Function):
    @staticmethod
    def forward(ctx, input, position, weight, idx):
        """input (n, c), position (n, nsample, c), weight (n, nsample, c'), idx (n, nsample) i32
        -> (n, c): out[n,c] = sum_s (input[idx[n,s],c] + position[n,s,c]) * weight[n,s,c % c']"""
        C.require(input, "input", torch.float32, 2)
        C.require(position, "position", torch.float32, 3)
        C.require(weight, "weight", torch.float32, 3)
        C.require(idx, "idx", torch.int32, 2)
        C.same_device(("input", input), ("position", position), ("weight", weight), ("idx", idx))
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        if weight.shape[:2] != (n, nsample) or idx.shape != (n, nsample) or input.shape[1] != c:
            raise ValueError("aggregation: inconsistent shapes")
        if c % w_c:
            raise ValueError("aggregation: weight channels must divide feature channels")
        output = torch.empty((n, c), dtype=torch.float32, device=input.device)
        with _lib.device_guard(input.device):
            _lib.run("pob_aggregation_forward", n, nsample, c, w_c, _lib.ptr(input), _lib.ptr(position),
                     _lib.ptr(weight), _lib.ptr(idx), _lib.ptr(output), _lib.current_stream(input.device),
                     alg_bytes=4 * (input.shape[0] * c + n * nsample * c + n * nsample * w_c + n * nsample + n * c))
        ctx.save_for_backward(input, position, weight, idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, position, weight, idx = ctx.saved_tensors
        grad_output = grad_output.contiguous().float()
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        dev = grad_output.device
        grad_input = torch.zeros_like(input)
        grad_position = torch.empty_like(position)
        grad_weight = torch.empty_like(weight)
        with _lib.device_guard(dev):
            fwd_in = input.shape[0] * c + n * nsample * c + n * nsample * w_c + n * nsample
            _lib.run("pob_aggregation_backward", n, nsample, c, w_c, _lib.ptr(input), _lib.ptr(position),
                     _lib.ptr(weight), _lib.ptr(idx), _lib.ptr(grad_output), _lib.ptr(grad_input),
                     _lib.ptr(grad_position), _lib.ptr(grad_weight), _lib.current_stream(dev),
                     alg_bytes=4 * (fwd_in + n * c + input.shape[0] * c + n * nsample * c + n * nsample * w_c))
        return grad_input, grad_position, grad_weight, None


aggregation = Aggregation.apply
