"""subtraction -- mirror of libs/pointops/functions/subtraction.py:7-38."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from . import _common as C


class Subtraction(Function):
    @staticmethod
    def forward(ctx, input1, input2, idx):
        """input1, input2: (n, c) f32, idx: (n, nsample) i32 -> (n, nsample, c):
        out[n,s,:] = input1[n,:] - input2[idx[n,s],:]"""
        C.require(input1, "input1", torch.float32, 2)
        C.require(input2, "input2", torch.float32, 2)
        C.require(idx, "idx", torch.int32, 2)
        C.same_device(("input1", input1), ("input2", input2), ("idx", idx))
        n, c = input1.shape
        nsample = idx.shape[-1]
        output = torch.empty((n, nsample, c), dtype=torch.float32, device=input1.device)
        with _lib.device_guard(input1.device):
            _lib.run("pob_subtraction_forward", n, nsample, c, _lib.ptr(input1), _lib.ptr(input2), _lib.ptr(idx),
                     _lib.ptr(output), _lib.current_stream(input1.device),
                     alg_bytes=4 * (2 * n * c + n * nsample + n * nsample * c))
        ctx.n2 = input2.shape[0]
        ctx.save_for_backward(idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (idx,) = ctx.saved_tensors
        grad_output = grad_output.contiguous().float()
        n, nsample, c = grad_output.shape
        dev = grad_output.device
        grad_input1 = torch.empty((n, c), dtype=torch.float32, device=dev)
        grad_input2 = torch.zeros((ctx.n2, c), dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            _lib.run("pob_subtraction_backward", n, nsample, c, _lib.ptr(idx), _lib.ptr(grad_output),
                     _lib.ptr(grad_input1), _lib.ptr(grad_input2), _lib.current_stream(dev),
                     alg_bytes=4 * (n * nsample * c + n * nsample + 2 * n * c))
        return grad_input1, grad_input2, None


subtraction = Subtraction.apply
