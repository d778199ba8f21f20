"""attention_relation_step / attention_fusion_step -- mirror of libs/pointops/functions/attention.py.

Point Transformer v2's grouped vector attention operators; not called by PTv1 or the recognizers
 but part of the package surface, so they are provided with the same
signatures, autograd behaviour (including the reference's: no gradient is returned for ``weight`` of the
relation step, functions/attention.py:62) and float32-only kernels.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib
from . import _common as C


def _idx(t, name, m=None):
    C.require(t, name, (torch.int32, torch.int64), 1)
    if m is not None and t.shape[0] != m:
        raise ValueError("index_target and index_refer must have the same length")
    return t if t.dtype == torch.int32 else t.int()   # the reference casts with .int() at every call


class AttentionRelationStep(Function):
    @staticmethod
    def forward(ctx, query, key, weight, index_target, index_refer):
        """
        input - query: (n, g, c), key: (n, g, c), weight: (c)  1_c for scatter attention,
                index_target: (m), index_refer: (m)
        output - relation: (m, g)
        """
        C.require(query, "query", torch.float32, 3)
        C.require(key, "key", torch.float32, 3)
        C.require(weight, "weight", torch.float32, 1)
        it = _idx(index_target, "index_target")
        ir = _idx(index_refer, "index_refer", it.shape[0])
        _, g, c = query.shape
        if key.shape[1:] != (g, c) or weight.shape[0] != c:
            raise ValueError("query / key must be (n, g, c) and weight (c)")
        m = it.shape[0]
        out = torch.empty((m, g), dtype=torch.float32, device=query.device)
        with _lib.device_guard(query.device):
            _lib.run("pob_attention_relation_step_forward", m, g, c, _lib.ptr(query), _lib.ptr(key), _lib.ptr(weight),
                     _lib.ptr(it), _lib.ptr(ir), _lib.ptr(out), _lib.current_stream(query.device),
                     alg_bytes=4 * (2 * m * g * c + c + 2 * m + m * g))
        ctx.save_for_backward(query, key, weight, it, ir)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        query, key, weight, it, ir = ctx.saved_tensors
        n, g, c = query.shape
        m = it.shape[0]
        grad_output = grad_output.float().contiguous()
        grad_query, grad_key = torch.zeros_like(query), torch.zeros_like(key)
        grad_weight = torch.zeros_like(weight)
        with _lib.device_guard(query.device):
            _lib.run("pob_attention_relation_step_backward", m, g, c, _lib.ptr(query), _lib.ptr(grad_query), _lib.ptr(key),
                     _lib.ptr(grad_key), _lib.ptr(weight), _lib.ptr(grad_weight), _lib.ptr(it), _lib.ptr(ir),
                     _lib.ptr(grad_output), _lib.current_stream(query.device),
                     alg_bytes=4 * (4 * m * g * c + 2 * c + 2 * m + m * g))
        return grad_query, grad_key, None, None, None   # functions/attention.py:62 drops grad_weight too


class AttentionFusionStep(Function):
    @staticmethod
    def forward(ctx, weight, value, index_target, index_refer):
        """
        input - weight: (m, g), value: (n, g, c), index_target: (m), index_refer: (m)
        output - output: (n, g, c)
        """
        C.require(weight, "weight", torch.float32, 2)
        C.require(value, "value", torch.float32, 3)
        it = _idx(index_target, "index_target")
        ir = _idx(index_refer, "index_refer", it.shape[0])
        n, g, c = value.shape
        m = ir.shape[0]
        if weight.shape != (m, g):
            raise ValueError("weight must be (m, g)")
        out = torch.zeros((n, g, c), dtype=torch.float32, device=value.device)
        with _lib.device_guard(value.device):
            _lib.run("pob_attention_fusion_step_forward", m, g, c, _lib.ptr(weight), _lib.ptr(value), _lib.ptr(it),
                     _lib.ptr(ir), _lib.ptr(out), _lib.current_stream(value.device),
                     alg_bytes=4 * (m * g + 2 * m * g * c + 2 * m))
        ctx.save_for_backward(weight, value, it, ir)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        """grad_output (n, g, c) -> grad_weight (m, g), grad_value (n, g, c)."""
        weight, value, it, ir = ctx.saved_tensors
        n, g, c = value.shape
        m = it.shape[0]
        grad_output = grad_output.float().contiguous()
        grad_weight = torch.empty_like(weight)
        grad_value = torch.zeros_like(value)
        with _lib.device_guard(value.device):
            _lib.run("pob_attention_fusion_step_backward", m, g, c, _lib.ptr(weight), _lib.ptr(grad_weight), _lib.ptr(value),
                     _lib.ptr(grad_value), _lib.ptr(it), _lib.ptr(ir), _lib.ptr(grad_output),
                     _lib.current_stream(value.device), alg_bytes=4 * (2 * m * g + 3 * m * g * c + 2 * m))
        return grad_weight, grad_value, None, None


attention_relation_step = AttentionRelationStep.apply
attention_fusion_step = AttentionFusionStep.apply
