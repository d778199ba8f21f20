"""Inference-only fused entry points (additive to the reference's pointops API; no autograd).

They serve the eval plan of the PTv1 mirror (``pointcloudpdf_b200/ptv1.py``): with every
BatchNorm in inference mode the eager glue between the cuBLAS linears of
``pointcept/models/point_transformer/point_transformer_seg.py`` collapses into a few kernels.
Each wrapper validates, allocates the output and calls one C-ABI entry point of
``include/pointops_b200.h``; unsupported shapes raise (callers keep the unfused operator path).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib
from . import _common as C

PT_LAYER_CHANNELS = (32, 64, 128, 256, 512)
PT_LAYER_NSAMPLE = (8, 16)


def pt_layer_supported(c: int, share_planes: int, nsample: int) -> bool:
    return c in PT_LAYER_CHANNELS and share_planes == 8 and nsample in PT_LAYER_NSAMPLE


def pack_pt_layer_params(A, cvec, wp, bp, aw, bw, w1, b1, w2, b2, oa, ob) -> torch.Tensor:
    """Pack folded PointTransformerLayer parameters into the block pob_pt_layer_forward reads:
    A (3,3), cvec (3): Linear(3,3)+BN; wp (C,3), bp (C): Linear(3,C); aw, bw (C): first BN of linear_w;
    w1 (C/8, C), b1 (C/8): Linear+BN folded; w2 (C/8, C/8), b2: last Linear; oa, ob (C): output affine."""
    c, wc = wp.shape[0], w1.shape[0]
    dev = wp.device
    f = lambda t: t.detach().to(device=dev, dtype=torch.float32).reshape(-1)
    head = torch.zeros(16, dtype=torch.float32, device=dev)
    head[:9] = f(A)
    head[9:12] = f(cvec)
    parts = [head, f(wp[:, 0]), f(wp[:, 1]), f(wp[:, 2]), f(bp), f(aw), f(bw), f(oa), f(ob),
             f(w1), f(b1), f(w2.t().contiguous()), f(b2)]
    out = torch.cat(parts).contiguous()
    expect = int(_lib.load().pob_pt_layer_param_floats(c, wc))
    if out.numel() != expect:
        raise ValueError(f"pt_layer parameter block has {out.numel()} floats, expected {expect}")
    return out


def pt_layer_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, xyz: torch.Tensor, idx: torch.Tensor,
                     params: torch.Tensor, out_affine: bool = True, split: int = 0) -> torch.Tensor:
    """Eval-mode PointTransformerLayer (+ optional BN/ReLU tail) on q, k, v (n, C) -- which may be
    column slices of one (n, 3C) tensor -- coordinates xyz (n, 3) and the self-kNN idx (n, ns) i32.
    split (per call, tests / tuning): 0 = the CTA-tiled kernel, 1 / 4 / 8 / 16 / -1 = the warp-per-point forms."""
    n, c = q.shape
    ns = idx.shape[1]
    for name, t in (("q", q), ("k", k), ("v", v)):
        if not t.is_cuda or t.dtype != torch.float32 or t.shape != (n, c) or t.stride(1) != 1:
            raise ValueError(f"{name} must be a CUDA f32 (n, C) tensor with unit channel stride")
    C.require(xyz, "xyz", torch.float32, 2, 3)
    C.require(idx, "idx", torch.int32, 2)
    C.require(params, "params", torch.float32, 1)
    if idx.shape[0] != n or xyz.shape[0] != n:
        raise ValueError("pt_layer_forward: q, xyz and idx must have the same number of rows")
    if not pt_layer_supported(c, 8, ns):
        raise ValueError(f"pt_layer_forward: unsupported C={c} / nsample={ns}")
    out = torch.empty((n, c), dtype=torch.float32, device=q.device)
    wc = c // 8
    with _lib.device_guard(q.device):
        _lib.run("pob_pt_layer_forward", n, ns, c, wc, _lib.ptr(q), q.stride(0), _lib.ptr(k), k.stride(0),
                 _lib.ptr(v), v.stride(0), _lib.ptr(xyz), _lib.ptr(idx), _lib.ptr(params), int(bool(out_affine)),
                 _lib.ptr(out), c, int(split), _lib.current_stream(q.device),
                 # compulsory traffic: q, k, v tables, coordinates, indices, output, once each
                 alg_bytes=4 * (4 * n * c + 3 * n + n * ns) + 4 * params.numel(),
                 alg_flops=2 * n * ns * (9 + 3 * c + c * wc + wc * wc + 2 * c),
                 # SURVEY.md 8(d) bytes of the three reference operators this launch stands for: knn_query_and_group
                 # (k rows with xyz), grouping (v rows), aggregation forward -- the (n, ns, .) tensors they move
                 unfused_bytes=4 * ((n * c + 6 * n + n * ns + n * ns * (3 + c)) + (n * c + n * ns + n * ns * c)
                                    + (2 * n * c + n * ns * c + n * ns * wc + n * ns)))
    return out


def linear(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, relu: bool = False, config: int = 0) -> torch.Tensor:
    """act(x @ wt + bias + residual): FP32 linear with the epilogue applied on the accumulators
    (``pob_linear_forward``).  x (M, K) f32 with unit column stride (a column block of a wider tensor is
    fine), wt (K, N) dense -- the Linear weight transposed once --, bias (N), residual (M, N).
    config (per call, tests / tuning): 0 = tile picked from the shape, 1..16 = force one instantiation."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("linear: x must be a CUDA f32 (M, K) tensor with unit column stride")
    C.require(wt, "wt", torch.float32, 2)
    m, k = x.shape
    n = wt.shape[1]
    if wt.shape[0] != k:
        raise ValueError(f"linear: x is (M, {k}) but wt is {tuple(wt.shape)}")
    if bias is not None:
        C.require(bias, "bias", torch.float32, 1)
        if bias.shape[0] != n:
            raise ValueError("linear: bias must be (N,)")
    ldr = 0
    if residual is not None:
        if (not residual.is_cuda or residual.dtype != torch.float32 or residual.shape != (m, n)
                or residual.stride(1) != 1):
            raise ValueError("linear: residual must be a CUDA f32 (M, N) tensor with unit column stride")
        ldr = residual.stride(0)
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    with _lib.device_guard(x.device):
        _lib.run("pob_linear_forward", m, k, n, _lib.ptr(x), x.stride(0) if m > 1 else max(x.stride(0), k),
                 _lib.ptr(wt), _lib.ptr(bias), _lib.ptr(residual), ldr if m > 1 else max(ldr, n), int(bool(relu)),
                 _lib.ptr(out), n, int(config), _lib.current_stream(x.device),
                 alg_bytes=4 * (m * k + k * n + m * n * (1 + (residual is not None)) + (n if bias is not None else 0)),
                 alg_flops=2 * m * k * n)
    return out


def affine_act(x: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
               residual: Optional[torch.Tensor] = None, relu: bool = True, inplace: bool = False) -> torch.Tensor:
    """relu?(x * scale + shift + residual) over (rows, c)."""
    C.require(x, "x", torch.float32, 2)
    rows, c = x.shape
    out = x if inplace else torch.empty_like(x)
    with _lib.device_guard(x.device):
        _lib.run("pob_affine_act", rows, c, _lib.ptr(x), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(residual),
                 int(bool(relu)), _lib.ptr(out), _lib.current_stream(x.device),
                 alg_bytes=4 * rows * c * (2 + (residual is not None)))
    return out


def interpolation_add(feat: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor,
                      base: Optional[torch.Tensor] = None, inplace: bool = False) -> torch.Tensor:
    """base + sum_i feat[idx[:, i]] * weight[:, i]  (TransitionUp's interpolation + skip in one kernel)."""
    C.require(feat, "feat", torch.float32, 2)
    C.require(idx, "idx", torch.int32, 2)
    C.require(weight, "weight", torch.float32, 2)
    n, k = idx.shape
    c = feat.shape[1]
    if base is not None:
        C.require(base, "base", torch.float32, 2)
        if base.shape != (n, c):
            raise ValueError("interpolation_add: base must be (n, c)")
    out = base if (inplace and base is not None) else torch.empty((n, c), dtype=torch.float32, device=feat.device)
    with _lib.device_guard(feat.device):
        _lib.run("pob_interpolation_add_forward", n, c, k, _lib.ptr(feat), _lib.ptr(idx), _lib.ptr(weight),
                 _lib.ptr(base), _lib.ptr(out), _lib.current_stream(feat.device),
                 alg_bytes=4 * (feat.shape[0] * c + 2 * n * k + n * c * (1 + (base is not None))))
    return out


def transition_down_pool(z: torch.Tensor, xyz: torch.Tensor, new_xyz: torch.Tensor, idx: torch.Tensor,
                         wxyz: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """out[m, c] = max_s relu(scale[c] * (z[idx[m,s], c] + wxyz[c] . (xyz[idx[m,s]] - new_xyz[m])) + shift[c]).

    TransitionDown (point_transformer_seg.py:106-119) with its Linear(3 + C, C') split by
    linearity: z = x @ W[:, 3:].T is computed on the N ungathered points (4x fewer GEMM rows than
    on the (m, ns) grouped tensor), the 3 coordinate columns W[:, :3] = wxyz are applied here on
    the relative positions, then BatchNorm (eval: scale / shift), ReLU and MaxPool1d(ns).
    Placeholder neighbours (idx < 0) contribute a zero grouped row, as pointops.grouping masks them."""
    C.require(z, "z", torch.float32, 2)
    C.require(xyz, "xyz", torch.float32, 2, 3)
    C.require(new_xyz, "new_xyz", torch.float32, 2, 3)
    C.require(idx, "idx", torch.int32, 2)
    C.require(wxyz, "wxyz", torch.float32, 2, 3)
    m, ns = idx.shape
    c = z.shape[1]
    if new_xyz.shape[0] != m or xyz.shape[0] != z.shape[0] or wxyz.shape[0] != c:
        raise ValueError("transition_down_pool: inconsistent shapes")
    out = torch.empty((m, c), dtype=torch.float32, device=z.device)
    with _lib.device_guard(z.device):
        _lib.run("pob_transition_down_pool", m, ns, c, _lib.ptr(z), _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(idx),
                 _lib.ptr(wxyz), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(out), _lib.current_stream(z.device),
                 alg_bytes=4 * (z.shape[0] * c + 3 * z.shape[0] + 3 * m + m * ns + m * c))
    return out
