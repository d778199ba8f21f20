"""Point Transformer v1 (Seg26/38/50) on the B200 point operators.

Host-side mirror of the reference backbone ``pointcept/models/point_transformer/
point_transformer_seg.py`` (PointTransformerLayer :22-81, TransitionDown :84-122, TransitionUp
:125-171, Bottleneck :174-195, PointTransformerSeg :198-306) and of the PDF U-decoder
``pointcept/recognizers/recognizer_model/pt_v1.py:8-44``.  The reference model files run
unmodified on ``import pointops`` from this repo; this mirror exists because they cannot travel
to the GPU box, and because the caller is where the remaining waste sits:

  * identical module tree, parameter names, construction order and maths -> ``state_dict``s are
    interchangeable and, under the same seed, identical (tests/test_cpu_ptv1.py);
  * one kNN per stage instead of one per block (the cloud does not change inside a stage;
    the reference recomputes it 18x per forward, :51) -- same indices, bit for bit;
  * vector attention ``einsum(x_v[idx] + p_r, w)`` runs as the fused ``pointops.aggregation``
    kernel (mathematically identical), so the gathered values are never stored;
  * ``LayerNorm1d`` (= BatchNorm1d over a transposed copy, point_transformer/utils.py:7-14)
    normalises the (n*ns, c) view directly: same statistics, no transposes;
  * no ``.item()`` / per-scene python loops on device tensors (:99-103, :152-164): stage sizes
    come from the host copy of ``offset``.

``fused=False`` reproduces the reference's op sequence literally (gather both k and v, einsum),
which the parity tests compare against.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, pointops
from .pointops import _common as C
from .pointops.sampling import fps_launch
from .pointops import fused as FZ


# ------------------------------------------------------------- frozen (inference) form ----
# In eval mode under torch.no_grad() every BatchNorm is a per-channel affine map, so the eager glue
# between the cuBLAS linears folds away:
#   Linear -> BN -> ReLU                      = one pob_linear_forward (FP32 FFMA GEMM, bias + ReLU on the accumulators)
#   q, k, v linears                           = one GEMM on concatenated weights
#   linear3 -> bn3 -> + identity -> ReLU      = one pob_linear_forward with the skip as epilogue operand
#   gathers, linear_p, relation, linear_w, softmax, aggregation, bn2, ReLU = pob_pt_layer_forward
#   TransitionDown's Linear(3+C, C') on the grouped tensor = GEMM on the UNGATHERED points (linearity)
#     + pob_transition_down_pool (gather, coordinate columns, BN, ReLU, max over neighbours)
#   TransitionUp's interpolation + skip      = pob_interpolation_add_forward
# ~900 kernel launches per 80k-point room become ~170; the modules, their parameters and their
# state_dict are untouched -- the folded tensors are a cache, dropped on train() / .to() /
# load_state_dict() and rebuilt lazily (call invalidate_frozen() after editing weights in place).

def _bn_affine(bn: nn.BatchNorm1d):
    a = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    b = bn.bias.detach().double() - bn.running_mean.detach().double() * a
    return a, b


def _fold(linear: nn.Linear, bn: Optional[nn.BatchNorm1d]):
    """(W', b') in f32 with bn(linear(x)) == x @ W'.T + b' in eval mode (folded in f64)."""
    W = linear.weight.detach().double()
    b = linear.bias.detach().double() if linear.bias is not None else torch.zeros(W.shape[0], dtype=torch.float64, device=W.device)
    if bn is not None:
        a, sh = _bn_affine(bn)
        W, b = W * a[:, None], b * a + sh
    return W.float().contiguous(), b.float().contiguous()


def _f32(t):
    return t.detach().float().contiguous()


_GENERATION = [0]   # bumped whenever folded weights are dropped or a backend switch changes what a forward launches


def _bump_generation() -> None:
    """Captured room graphs hold raw pointers to the folded tensors and a fixed kernel sequence: anything
    that frees / rebuilds those tensors or changes the sequence makes every captured graph stale."""
    _GENERATION[0] += 1


class _Freezable:
    """Mixin: cache of folded inference tensors, dropped whenever the parameters may have changed."""
    _frozen = None
    use_frozen = True

    def _drop_frozen(self):
        if self._frozen is not None:
            self._frozen = None
        _bump_generation()

    def train(self, mode: bool = True):
        self._drop_frozen()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self._drop_frozen()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._drop_frozen()
        return super()._load_from_state_dict(*args, **kwargs)

    def _can_freeze(self, x: torch.Tensor) -> bool:
        return (self.use_frozen and not self.training and not torch.is_grad_enabled() and x.is_cuda
                and x.dtype == torch.float32)

    def frozen(self):
        f = self._frozen
        if f is None:
            with torch.no_grad():
                f = self._frozen = self._freeze()
        return f


def invalidate_frozen(module: nn.Module) -> None:
    """Call after editing weights in place: drops the folded tensors (and with them every captured room graph)."""
    for m in module.modules():
        if isinstance(m, _Freezable):
            m._frozen = None
    _bump_generation()


# Backend of the frozen linears.  "cublas": torch.addmm / _addmm_activation (cuBLAS SIMT GEMM, a cuBLASLt pass for
# bias / ReLU, pob_affine_act for the skip); "pob": every linear runs pob_linear_forward (FP32 FFMA tiles, bias /
# skip / ReLU applied on the accumulators); "auto" (default): pob_linear_forward where it measured faster on B200
# (profiles/r01d_linear_time.txt): linears WITH an epilogue on >= 600 rows (one launch instead of two), the
# 80 000-row layers, and shapes that are not 16-byte friendly; cuBLAS for the plain q/k/v GEMMs of the deeper stages.
_LINEAR_BACKEND = os.environ.get("POINTOPS_B200_LINEAR", "auto")


def set_linear_backend(name: str) -> None:
    global _LINEAR_BACKEND
    if name not in ("pob", "cublas", "auto"):
        raise ValueError("linear backend must be 'pob', 'cublas' or 'auto'")
    if name != _LINEAR_BACKEND:
        _bump_generation()
    _LINEAR_BACKEND = name


def _use_pob_linear(m: int, k: int, n: int, epilogue: bool) -> bool:
    if _LINEAR_BACKEND != "auto":
        return _LINEAR_BACKEND == "pob"
    if k % 4 or n % 4:
        return True
    if m >= 40000:
        return n <= 96
    if m >= 10000 and n <= 192:
        return True          # 3xTF32 tensor-core tiles: 15.2 vs 18.7 us (cuBLAS) at 20 000 x 64 x 192
    return epilogue and m >= 600


GEO_PRIORITY = int(os.environ.get("POINTOPS_B200_GEO_PRIORITY", "0"))

_OWN = object()


class _Lin:
    """Frozen linear  y = act(x @ W'.T + b' [+ residual])  with (W', b') from _fold()."""
    __slots__ = ("w", "wt", "b")

    def __init__(self, W: torch.Tensor, b: Optional[torch.Tensor]):
        self.w = W.contiguous()             # (N, K); .t() is the view cuBLAS consumes
        self.wt = self.w.t().contiguous()   # (K, N) dense, what pob_linear_forward streams
        self.b = b

    def __call__(self, x, relu: bool = False, residual=None, bias=_OWN):
        b = self.b if bias is _OWN else bias
        if x.stride(-1) != 1:               # both routes take rows with unit column stride
            x = x.contiguous()
        if residual is not None and residual.stride(-1) != 1:
            residual = residual.contiguous()
        if _use_pob_linear(x.shape[0], self.wt.shape[0], self.wt.shape[1], b is not None or relu or residual is not None):
            return FZ.linear(x, self.wt, b, residual, relu)
        wt = self.w.t()
        if residual is not None:
            z = torch.addmm(residual, x, wt)
            return FZ.affine_act(z, None, b, None, relu=relu, inplace=True) if (b is not None or relu) else z
        if b is None:
            z = torch.mm(x, wt)
            return torch.relu_(z) if relu else z
        return torch._addmm_activation(b, x, wt) if relu else torch.addmm(b, x, wt)


class Level:
    """Coordinate-only state of one resolution level, computed ahead of the features on a side
    stream (PointTransformerSeg.geometry): self-kNN index, and -- towards the next coarser level --
    the FPS selection, the cross kNN index and the 3-NN interpolation index/weights."""

    __slots__ = ("p", "o", "o_host", "knn", "knn_ev", "down", "down_ev", "coarser")

    def __init__(self, p, o, o_host):
        self.p, self.o, self.o_host = p, o, list(o_host)
        self.knn, self.knn_ev = {}, None
        self.down, self.down_ev, self.coarser = None, None, None


def _wait(ev):
    if ev is not None:
        _lib.current_stream_obj().wait_event(ev)


class Cloud:
    """One resolution level: coordinates, features, cumulative offsets (device + host copy) and
    the self-kNN index shared by every block of the level."""

    __slots__ = ("p", "x", "o", "o_host", "_knn", "level")

    def __init__(self, p, x, o, o_host, level=None):
        self.p, self.x, self.o, self.o_host = p, x, o, list(o_host)
        self._knn = {}
        self.level = level

    def with_feat(self, x) -> "Cloud":
        c = Cloud(self.p, x, self.o, self.o_host, self.level)
        c._knn = self._knn
        return c

    def knn(self, nsample: int) -> torch.Tensor:
        idx = self._knn.get(nsample)
        if idx is None:
            if self.level is not None and nsample in self.level.knn:
                _wait(self.level.knn_ev)
                idx = self.level.knn[nsample]
            else:
                idx, _ = pointops.knn_query(nsample, self.p, self.o)
            self._knn[nsample] = idx
        return idx


class LayerNorm1d(nn.BatchNorm1d):
    """BatchNorm over the channel axis of (n, ns, c); same parameters/buffers as the reference's
    transposing subclass (point_transformer/utils.py:7-14), applied on the flattened view."""

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if input.dim() != 3:
            return super().forward(input)
        n, ns, c = input.shape
        return super().forward(input.reshape(n * ns, c)).view(n, ns, c)


class PointTransformerLayer(nn.Module):
    def __init__(self, in_planes, out_planes, share_planes=8, nsample=16):
        super().__init__()
        self.mid_planes = mid_planes = out_planes // 1
        self.out_planes = out_planes
        self.share_planes = share_planes
        self.nsample = nsample
        self.linear_q = nn.Linear(in_planes, mid_planes)
        self.linear_k = nn.Linear(in_planes, mid_planes)
        self.linear_v = nn.Linear(in_planes, out_planes)
        self.linear_p = nn.Sequential(nn.Linear(3, 3), LayerNorm1d(3), nn.ReLU(inplace=True), nn.Linear(3, out_planes))
        self.linear_w = nn.Sequential(
            LayerNorm1d(mid_planes), nn.ReLU(inplace=True), nn.Linear(mid_planes, out_planes // share_planes),
            LayerNorm1d(out_planes // share_planes), nn.ReLU(inplace=True),
            nn.Linear(out_planes // share_planes, out_planes // share_planes))
        self.softmax = nn.Softmax(dim=1)
        self.fused = True

    def forward(self, cloud: Cloud) -> torch.Tensor:
        p, x, o = cloud.p, cloud.x, cloud.o
        x_q, x_k, x_v = self.linear_q(x), self.linear_k(x), self.linear_v(x)
        if self.fused:
            idx = cloud.knn(self.nsample)
        else:  # the reference's literal sequence: kNN inside every layer
            idx, _ = pointops.knn_query(self.nsample, p, o)
        if self.fused:  # same values as the interleaved tensor, kept as two aligned ones
            p_r, x_kg = pointops.grouping_split(idx, x_k, p, p)
        else:
            g = pointops.grouping(idx, x_k, p, p, with_xyz=True)  # (n, ns, 3 + c)
            p_r, x_kg = g[:, :, 0:3], g[:, :, 3:]
        p_r = self.linear_p(p_r)                                # (n, ns, out)
        n, ns, _ = p_r.shape
        # the reference reduces p_r over "(i j) -> j" with j = mid_planes; i == 1 here: the identity
        p_mid = p_r if self.out_planes == self.mid_planes else p_r.view(n, ns, -1, self.mid_planes).sum(2)
        r_qk = x_kg - x_q.unsqueeze(1) + p_mid
        w = self.softmax(self.linear_w(r_qk))                   # (n, ns, out // share), softmax over neighbours
        if self.fused:
            return pointops.aggregation(x_v.float().contiguous(), p_r.float().contiguous(), w.float().contiguous(), idx)
        x_vg = pointops.grouping(idx, x_v, p, p, with_xyz=False)
        s = self.share_planes
        out = torch.einsum("ntsi,nti->nsi", (x_vg + p_r).view(n, ns, s, self.out_planes // s), w)
        return out.reshape(n, self.out_planes)


def strided_offsets(o_host: Sequence[int], stride: int) -> List[int]:
    """Cumulative per-scene sample counts n_b // stride (point_transformer_seg.py:99-103), on the host."""
    out, acc = [], 0
    for n_b in C.scene_sizes(o_host):
        acc += n_b // stride
        out.append(acc)
    return out


class TransitionDown(_Freezable, nn.Module):
    def __init__(self, in_planes, out_planes, stride=1, nsample=16):
        super().__init__()
        self.stride, self.nsample = stride, nsample
        if stride != 1:
            self.linear = nn.Linear(3 + in_planes, out_planes, bias=False)
            self.pool = nn.MaxPool1d(nsample)
        else:
            self.linear = nn.Linear(in_planes, out_planes, bias=False)
        self.bn = nn.BatchNorm1d(out_planes)
        self.relu = nn.ReLU(inplace=True)

    def _freeze(self):
        if self.stride == 1:
            W, b = _fold(self.linear, self.bn)
            return dict(lin=_Lin(W, b))
        W = self.linear.weight.detach()
        a, sh = _bn_affine(self.bn)
        return dict(wxyz=_f32(W[:, :3]), wfeat=_Lin(_f32(W[:, 3:]), None), scale=a.float().contiguous(), shift=sh.float().contiguous())

    def forward(self, cloud: Cloud) -> Cloud:
        frozen = self._can_freeze(cloud.x)
        if self.stride == 1:
            if frozen:
                return cloud.with_feat(self.frozen()["lin"](cloud.x, relu=True))
            return cloud.with_feat(self.relu(self.bn(self.linear(cloud.x))))
        p, x, o = cloud.p, cloud.x, cloud.o
        lvl = cloud.level
        if lvl is not None and lvl.down is not None:   # computed ahead on the geometry stream
            _wait(lvl.down_ev)
            n_p, n_o, n_o_host, cross_idx = lvl.down["n_p"], lvl.down["n_o"], lvl.down["n_o_host"], lvl.down["cross"]
            nxt = lvl.coarser
        else:
            n_o_host = strided_offsets(cloud.o_host, self.stride)
            n_o = C.const_offset(n_o_host, p.device)
            C.register_host_offset(n_o, n_o_host)
            C.register_host_offset(o, cloud.o_host)
            idx = pointops.farthest_point_sampling(p, o, n_o)          # (m)
            n_p = p[idx.long(), :]                                     # (m, 3)
            cross_idx, _ = pointops.knn_query(self.nsample, p, o, n_p, n_o)
            nxt = None
        if frozen and x.shape[1] % 4 == 0 and self.linear.out_features % 4 == 0:
            f = self.frozen()
            z = f["wfeat"](x)                                          # (n, c'): the GEMM on ungathered points
            y = FZ.transition_down_pool(z, p, n_p, cross_idx, f["wxyz"], f["scale"], f["shift"])
            return Cloud(n_p, y, n_o, n_o_host, nxt)
        g = pointops.grouping(cross_idx, x, p, n_p, with_xyz=True)     # (m, ns, 3 + c)
        m, ns, w = g.shape
        y = self.relu(self.bn(self.linear(g).view(m * ns, -1)))    # BN over (m, ns) per channel, no transpose
        y = y.view(m, ns, -1).max(dim=1)[0]                        # MaxPool1d(nsample)
        return Cloud(n_p, y, n_o, n_o_host, nxt)


class TransitionUp(_Freezable, nn.Module):
    def __init__(self, in_planes, out_planes=None):
        super().__init__()
        if out_planes is None:
            self.linear1 = nn.Sequential(nn.Linear(2 * in_planes, in_planes), nn.BatchNorm1d(in_planes),
                                         nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, in_planes), nn.ReLU(inplace=True))
        else:
            self.linear1 = nn.Sequential(nn.Linear(out_planes, out_planes), nn.BatchNorm1d(out_planes),
                                         nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, out_planes), nn.BatchNorm1d(out_planes),
                                         nn.ReLU(inplace=True))

    def _freeze(self):
        W1, b1 = _fold(self.linear1[0], self.linear1[1])
        if isinstance(self.linear2[1], nn.BatchNorm1d):
            W2, b2 = _fold(self.linear2[0], self.linear2[1])
            return dict(l1=_Lin(W1, b1), l2=_Lin(W2, b2))
        c = W1.shape[0]   # head: linear1 takes cat(x, tiled scene mean); linear2 is Linear + ReLU
        W2, b2 = _fold(self.linear2[0], None)
        return dict(l1a=_Lin(W1[:, :c], None), l1b=_Lin(W1[:, c:], b1), l2=_Lin(W2, b2))

    def forward(self, fine: Cloud, coarse: Optional[Cloud] = None) -> torch.Tensor:
        frozen = self._can_freeze(fine.x)
        if coarse is None:
            # head: concatenate every point with its scene's mean feature (:152-164), segmented
            x = fine.x
            sizes = C.scene_sizes(fine.o_host)
            b = len(sizes)
            if b == 1 and frozen:
                # cat(x, tiled) @ W.T = x @ Wa.T + (mean-feature row) @ Wb.T: the second term is a bias
                f = self.frozen()
                t = f["l2"](x.mean(0, keepdim=True), relu=True)
                bias = f["l1b"](t).view(-1)
                return f["l1a"](x, relu=True, bias=bias)
            if b == 1:
                mean = x.sum(0, keepdim=True) / sizes[0]
                tiled = self.linear2(mean).expand(x.shape[0], -1)
            else:
                counts = C.const_offset(sizes, x.device).long()
                batch = torch.repeat_interleave(torch.arange(b, device=x.device), counts, output_size=x.shape[0])
                sums = torch.zeros((b, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
                tiled = self.linear2(sums / counts.to(x.dtype).unsqueeze(1))[batch]
            return self.linear1(torch.cat((x, tiled), 1))
        lvl = fine.level
        ahead = lvl is not None and lvl.down is not None and coarse.level is lvl.coarser and coarse.level is not None
        if frozen and fine.x.shape[1] % 4 == 0:
            f = self.frozen()
            feat = f["l2"](coarse.x, relu=True)
            base = f["l1"](fine.x, relu=True)
            if ahead:
                _wait(lvl.down_ev)
                up_idx, up_w = lvl.down["up_idx"], lvl.down["up_w"]
            else:
                from .pointops.interpolation import _neighbours_and_weights, _wrap_placeholders
                up_idx, up_w = _neighbours_and_weights(coarse.p, fine.p, coarse.o, fine.o, 3)
                up_idx = _wrap_placeholders(up_idx, feat.shape[0])
            return FZ.interpolation_add(feat, up_idx, up_w, base=base, inplace=True)
        feat = self.linear2(coarse.x)
        if ahead:
            _wait(lvl.down_ev)
            from .pointops.interpolation import _InterpolateRows
            up = _InterpolateRows.apply(feat.float().contiguous(), lvl.down["up_idx"], lvl.down["up_w"])
        else:
            up = pointops.interpolation(coarse.p, fine.p, feat, coarse.o, fine.o)
        return self.linear1(fine.x) + up


class Bottleneck(_Freezable, nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, share_planes=8, nsample=16):
        super().__init__()
        self.linear1 = nn.Linear(in_planes, planes, bias=False)
        self.bn1 = nn.BatchNorm1d(planes)
        self.transformer = PointTransformerLayer(planes, planes, share_planes, nsample)
        self.bn2 = nn.BatchNorm1d(planes)
        self.linear3 = nn.Linear(planes, planes * self.expansion, bias=False)
        self.bn3 = nn.BatchNorm1d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)

    def _freeze(self):
        t = self.transformer
        W1, b1 = _fold(self.linear1, self.bn1)
        Wqkv = torch.cat([t.linear_q.weight, t.linear_k.weight, t.linear_v.weight]).detach().float().contiguous()
        bqkv = torch.cat([t.linear_q.bias, t.linear_k.bias, t.linear_v.bias]).detach().float().contiguous()
        A, cvec = _fold(t.linear_p[0], t.linear_p[1])
        aw, bw = _bn_affine(t.linear_w[0])
        w1, b1w = _fold(t.linear_w[2], t.linear_w[3])
        oa, ob = _bn_affine(self.bn2)
        pack = lambda bw_, ob_: FZ.pack_pt_layer_params(A, cvec, t.linear_p[3].weight, t.linear_p[3].bias, aw, bw_, w1, b1w,
                                                        t.linear_w[5].weight, t.linear_w[5].bias, oa, ob_)
        # The q / k / v biases can leave the GEMM (whose bias epilogue is a second pass over (n, 3c) in cuBLASLt):
        #   k_j + bk - (q_i + bq) + p_r   enters relu(aw * . + bw)  ->  bw' = bw + aw (bk - bq)
        #   sum_s w_s (v_j + bv + p_r) = sum_s w_s (v_j + p_r) + bv (softmax weights sum to 1)  ->  ob' = ob + oa bv
        # valid when no neighbour slot is a placeholder (those group to ZERO rows, not to the bias), i.e. when
        # every scene has at least nsample points; the caller checks that on the host and else keeps `params`.
        bq, bk, bv = (x.detach().double() for x in (t.linear_q.bias, t.linear_k.bias, t.linear_v.bias))
        W3, b3 = _fold(self.linear3, self.bn3)
        return dict(l1=_Lin(W1, b1), qkv=_Lin(Wqkv, bqkv), params=pack(bw, ob),
                    params_nobias=pack(bw + aw * (bk - bq), ob + oa * bv), l3=_Lin(W3, b3))

    def _frozen_ok(self, x) -> bool:
        t = self.transformer
        return (self._can_freeze(x) and t.fused and t.mid_planes == t.out_planes
                and FZ.pt_layer_supported(t.out_planes, t.share_planes, t.nsample))

    def forward(self, cloud: Cloud) -> Cloud:
        identity = cloud.x
        if self._frozen_ok(identity):
            f = self.frozen()
            c = self.transformer.out_planes
            h = f["l1"](identity, relu=True)
            if min(C.scene_sizes(cloud.o_host)) >= self.transformer.nsample:   # no placeholder neighbours: biases folded
                qkv, params = f["qkv"](h, bias=None), f["params_nobias"]                      # (n, 3c)
            else:
                qkv, params = f["qkv"](h), f["params"]
            y = FZ.pt_layer_forward(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], cloud.p, cloud.knn(self.transformer.nsample),
                                    params, out_affine=True)                                  # ... bn2 + relu
            return cloud.with_feat(f["l3"](y, relu=True, residual=identity))                  # linear3 + bn3 + skip + ReLU
        x = self.relu(self.bn1(self.linear1(cloud.x)))
        x = self.relu(self.bn2(self.transformer(cloud.with_feat(x))))
        x = self.bn3(self.linear3(x))
        x = self.relu(x + identity)
        return cloud.with_feat(x)


class PointTransformerSeg(_Freezable, nn.Module):
    """Encoder-decoder of point_transformer_seg.py:198-306; ``forward(data_dict)`` takes the
    reference's dict (coord, feat, offset) and returns per-point logits."""

    def __init__(self, block, blocks, in_channels=6, num_classes=13):
        super().__init__()
        self.in_channels = in_channels
        self.in_planes, planes = in_channels, [32, 64, 128, 256, 512]
        share_planes = 8
        stride, nsample = [1, 4, 4, 4, 4], [8, 16, 16, 16, 16]
        for i in range(5):
            setattr(self, f"enc{i + 1}", self._make_enc(block, planes[i], blocks[i], share_planes, stride[i], nsample[i]))
        for i in (4, 3, 2, 1, 0):
            setattr(self, f"dec{i + 1}", self._make_dec(block, planes[i], 1, share_planes, nsample[i], is_head=(i == 4)))
        self.cls = nn.Sequential(nn.Linear(planes[0], planes[0]), nn.BatchNorm1d(planes[0]), nn.ReLU(inplace=True),
                                 nn.Linear(planes[0], num_classes))
        self.taps = None  # filled per forward: what the reference's ModelHook would capture
        self.overlap_geometry = True  # coordinate-only work on a side stream (fused path only)
        self._geo_stream = None
        self._all_fused = True

    def _make_enc(self, block, planes, blocks, share_planes=8, stride=1, nsample=16):
        layers = [TransitionDown(self.in_planes, planes * block.expansion, stride, nsample)]
        self.in_planes = planes * block.expansion
        for _ in range(blocks):
            layers.append(block(self.in_planes, self.in_planes, share_planes, nsample=nsample))
        return nn.Sequential(*layers)

    def _make_dec(self, block, planes, blocks, share_planes=8, nsample=16, is_head=False):
        layers = [TransitionUp(self.in_planes, None if is_head else planes * block.expansion)]
        self.in_planes = planes * block.expansion
        for _ in range(blocks):
            layers.append(block(self.in_planes, self.in_planes, share_planes, nsample=nsample))
        return nn.Sequential(*layers)

    def set_fused(self, fused: bool) -> "PointTransformerSeg":
        for m in self.modules():
            if isinstance(m, PointTransformerLayer):
                m.fused = bool(fused)
            if isinstance(m, _Freezable):
                m.use_frozen = bool(fused)
        self._all_fused = bool(fused)
        _bump_generation()
        return self

    @torch.no_grad()
    def geometry(self, p0, o0, offset_host, stream=None) -> List[Level]:
        """Everything that depends on coordinates only -- 4 FPS, 5 self-kNN, 4 cross-kNN, 4 three-NN
        with weights -- issued on `stream` so that it overlaps the feature path (FPS is a serial,
        16-SM kernel; the linears run on the other SMs meanwhile).  Each item carries the event
        the feature path waits on.  Results are bit-identical to computing them inline."""
        C.require(p0, "coord", torch.float32, 2, 3)   # validated once here; the launches below skip the per-op checks
        o0 = C.offset_i32(o0, "offset")
        main = _lib.current_stream_obj(p0.device)
        stream = stream if stream is not None else main
        strides = [m[0].stride for m in (self.enc1, self.enc2, self.enc3, self.enc4, self.enc5)]
        nsamples = [m[0].nsample for m in (self.enc1, self.enc2, self.enc3, self.enc4, self.enc5)]
        stream.wait_stream(main)
        levels: List[Level] = []
        keep = []
        with torch.cuda.stream(stream):
            C.register_host_offset(o0, offset_host)
            lvl = Level(p0, o0, offset_host)
            for s in range(5):
                levels.append(lvl)
                lvl.knn[nsamples[s]] = C.cached_knn(nsamples[s], lvl.p, lvl.o, lvl.p, lvl.o)[0]
                lvl.knn_ev = torch.cuda.Event()
                lvl.knn_ev.record(stream)
                keep.append(lvl.knn[nsamples[s]])
                if s == 4:
                    break
                n_o_host = strided_offsets(lvl.o_host, strides[s + 1])
                n_o = C.const_offset(n_o_host, lvl.p.device)
                C.register_host_offset(n_o, n_o_host)
                sel = fps_launch(lvl.p, lvl.o, n_o, lvl.o_host, n_o_host)
                n_p = lvl.p.index_select(0, sel)
                cross = C.cached_knn(nsamples[s + 1], lvl.p, lvl.o, n_p, n_o)[0]
                up_idx, _, up_w = C.cached_knn(3, n_p, n_o, lvl.p, lvl.o, want_weight=True)
                if min(C.scene_sizes(n_o_host)) < 3:  # quirk C6, as pointops.interpolation: -1 indexes feat[-1]
                    up_idx = torch.where(up_idx < 0, up_idx + n_p.shape[0], up_idx)
                lvl.down = dict(n_p=n_p, n_o=n_o, n_o_host=n_o_host, cross=cross, up_idx=up_idx, up_w=up_w)
                lvl.down_ev = torch.cuda.Event()
                lvl.down_ev.record(stream)
                keep += [n_p, n_o, cross, up_idx, up_w]
                lvl.coarser = Level(n_p, n_o, n_o_host)
                lvl = lvl.coarser
        if stream is not main and not torch.cuda.is_current_stream_capturing():
            for t in keep:
                t.record_stream(main)  # consumed by kernels on the main stream
        return levels

    def forward(self, data_dict, offset_host: Optional[Sequence[int]] = None, geometry: Optional[List[Level]] = None):
        p0, x0 = data_dict["coord"], data_dict["feat"]
        o0 = data_dict["offset"].int()
        if offset_host is None:
            offset_host = C.host_offset(o0)
        else:
            C.register_host_offset(o0, offset_host)
        level0 = None
        if geometry is not None:      # computed ahead by the caller (OpenSegPTv1.infer_stream)
            level0 = geometry[0]
        elif self.overlap_geometry and p0.is_cuda and self._all_fused:
            if self._geo_stream is None:
                # POINTOPS_B200_GEO_PRIORITY=-1: the coordinate branch (FPS clusters, kNN) on a high-priority stream, so
                # that a pending 16-CTA cluster is placed ahead of the other rooms' queued CTAs (captured into the room
                # graphs as the kernel nodes' priority)
                self._geo_stream = torch.cuda.Stream(device=p0.device, priority=GEO_PRIORITY)
            level0 = self.geometry(p0.contiguous(), o0, offset_host, self._geo_stream)[0]
        c1 = self.enc1(Cloud(p0, x0, o0, offset_host, level0))
        c2 = self.enc2(c1)
        c3 = self.enc3(c2)
        c4 = self.enc4(c3)
        c5 = self.enc5(c4)
        d5 = self.dec5[1:](c5.with_feat(self.dec5[0](c5)))
        d4 = self.dec4[1:](c4.with_feat(self.dec4[0](c4, d5)))
        d3 = self.dec3[1:](c3.with_feat(self.dec3[0](c3, d4)))
        d2 = self.dec2[1:](c2.with_feat(self.dec2[0](c2, d3)))
        d1 = self.dec1[1:](c1.with_feat(self.dec1[0](c1, d2)))
        self.taps = dict(enc=[c1, c2, c3, c4, c5], dec=[d1, d2, d3, d4, d5])
        if level0 is not None and geometry is None:
            torch.cuda.current_stream().wait_stream(self._geo_stream)
        if self._can_freeze(d1.x):
            f = self.frozen()
            return f["out"](f["hid"](d1.x, relu=True))
        return self.cls(d1.x)

    def _freeze(self):
        W, b = _fold(self.cls[0], self.cls[1])
        Wo, bo = _fold(self.cls[3], None)
        return dict(hid=_Lin(W, b), out=_Lin(Wo, bo))


class PointTransformerSeg26(PointTransformerSeg):
    def __init__(self, **kwargs):
        super().__init__(Bottleneck, [1, 1, 1, 1, 1], **kwargs)


class PointTransformerSeg38(PointTransformerSeg):
    def __init__(self, **kwargs):
        super().__init__(Bottleneck, [1, 2, 2, 2, 2], **kwargs)


class PointTransformerSeg50(PointTransformerSeg):
    def __init__(self, **kwargs):
        super().__init__(Bottleneck, [1, 2, 3, 5, 2], **kwargs)


class PTRecognizer(_Freezable, nn.Module):
    """PDF U-decoder (recognizers/recognizer_model/pt_v1.py:8-44): five TransitionUp over the
    backbone's hooked encoder/decoder activations, then a confidence head -> conf (n, 1)."""

    def __init__(self):
        super().__init__()
        planes = [32, 64, 128, 256, 512]
        self.dec5 = TransitionUp(planes[4], planes[4])
        self.dec4 = TransitionUp(planes[4], planes[3])
        self.dec3 = TransitionUp(planes[3], planes[2])
        self.dec2 = TransitionUp(planes[2], planes[1])
        self.dec1 = TransitionUp(planes[1], planes[0])
        self.confidence = nn.Sequential(nn.Linear(planes[0], planes[0]), nn.BatchNorm1d(planes[0]),
                                        nn.ReLU(inplace=True), nn.Linear(planes[0], 1))

    def forward(self, taps) -> torch.Tensor:
        enc, dec = taps["enc"], taps["dec"]
        c1, c2, c3, c4, c5 = enc
        d1, d2, d3, d4, d5 = dec
        r5 = self.dec5(d5, c5)                       # ([p5, x5_dec, o5], [p5, x5_enc, o5])
        r4 = self.dec4(d4, c5.with_feat(r5))
        r3 = self.dec3(d3, c4.with_feat(r4))
        r2 = self.dec2(d2, c3.with_feat(r3))
        r1 = self.dec1(d1, c2.with_feat(r2))
        if self._can_freeze(r1):
            f = self.frozen()
            return f["out"](f["hid"](r1, relu=True))
        return self.confidence(r1)

    def _freeze(self):
        W, b = _fold(self.confidence[0], self.confidence[1])
        Wo, bo = _fold(self.confidence[3], None)
        return dict(hid=_Lin(W, b), out=_Lin(Wo, bo))


class _RoomGraph:
    """One stream slot of OpenSegPTv1.infer_stream for one size signature: static input buffers, the
    captured forward (geometry branch on a forked stream + feature branch, joined), static outputs.
    Replaying it costs the host a few calls instead of ~220 kernel launches plus their python."""

    def __init__(self, model: "OpenSegPTv1", offset_host, in_channels: int, device):
        self.offset_host = [int(v) for v in offset_host]
        n = self.offset_host[-1]
        self.method = model.method
        self.stream = torch.cuda.Stream(device=device)
        self.coord = torch.empty((n, 3), dtype=torch.float32, device=device)
        self.feat = torch.empty((n, in_channels), dtype=torch.float32, device=device)
        self.offset = C.const_offset(self.offset_host, device)
        self.coord.zero_()
        self.feat.zero_()
        d = dict(coord=self.coord, feat=self.feat, offset=self.offset)
        torch.cuda.synchronize(device)
        with C.record_constants() as consts:
            with torch.cuda.stream(self.stream):
                # eager dry run: folds the BatchNorms, creates the constant offset tensors, the geometry
                # stream, cuBLAS handles / workspaces -- nothing lazy may be left for the capture
                model.forward(d, self.offset_host)
            torch.cuda.synchronize(device)
            pointops.clear_caches()      # the capture must launch every kernel itself
            self.graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            # thread_local: other threads of the process (NCCL watchdog, data loaders) may keep calling CUDA
            with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="thread_local"):
                out = model.forward(d, self.offset_host)
        self.launches = _lib.launch_count() - before   # kernels of ours inside one replay
        pointops.clear_caches()      # entries keyed on the graph's private buffers are of no use to anyone
        self.score, self.pred = out["score"], out["pred"]
        model.backbone.taps = None   # the taps of the capture pass point into the graph's private pool
        # everything outside the graph's private pool that the captured kernels read: the folded weights of every
        # module and the constant offset tensors.  Held here so that nothing the replay dereferences can be freed
        # while this graph is alive; `generation` tells infer_stream when they are no longer the model's current ones.
        self.keep = [m._frozen for m in model.modules() if isinstance(m, _Freezable)] + list(consts.items)
        self.generation = _GENERATION[0]

    def run(self, coord: torch.Tensor, feat: torch.Tensor):
        """Queue one room on this slot's stream; returns (event, score, pred) with score / pred in
        pinned host memory, valid once the event has completed."""
        n = self.coord.shape[0]
        score = torch.empty((n,), dtype=torch.float32, pin_memory=True)
        pred = torch.empty((n,), dtype=torch.int32, pin_memory=True)
        with torch.cuda.stream(self.stream):
            self.coord.copy_(coord, non_blocking=True)
            self.feat.copy_(feat, non_blocking=True)
            self.graph.replay()
            _lib.add_replayed_launches(self.launches)
            score.copy_(self.score, non_blocking=True)
            pred.copy_(self.pred, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return done, score, pred


class OpenSegPTv1(nn.Module):
    """What an ``openseg-pt-v1-0-{msp,ml,pointpdf}`` evaluation step computes per batch
    (hooks/evaluator.py:39-72): backbone logits, then the recognizer's open-set score.

    method: "msp" / "max_logits" (MaxProbability) or "pdf" (U-decoder + softmax score)."""

    def __init__(self, in_channels=6, num_classes=13, method="msp", blocks=(1, 2, 3, 5, 2)):
        super().__init__()
        self.backbone = PointTransformerSeg(Bottleneck, list(blocks), in_channels=in_channels, num_classes=num_classes)
        self.method = method
        self.recognizer = PTRecognizer() if method == "pdf" else None
        self._streams = []
        self._graphs = {}   # (scene sizes, feature width) -> [_RoomGraph per stream slot]

    def forward(self, data_dict, offset_host=None, geometry=None):
        from .scoring import fused_scores
        logits = self.backbone(data_dict, offset_host, geometry).float().contiguous()
        if self.method == "pdf":
            conf = self.recognizer(self.backbone.taps).float().contiguous()
            r = fused_scores(logits, conf, want=("pdf_score", "pred"))
            score = r["pdf_score"]
        else:
            key = "msp_score" if self.method == "msp" else "ml_score"
            r = fused_scores(logits, want=(key, "pred"))
            score = r[key]
        return dict(seg_logits=logits, score=score, pred=r["pred"])  # pred = argmax of the logits (first maximum), i32

    @torch.no_grad()
    def infer(self, coord: torch.Tensor, feat: torch.Tensor, offset: torch.Tensor, device=None):
        """Host tensors in (pinned memory recommended), host score out: the end-to-end call."""
        dev = device if device is not None else next(self.parameters()).device
        offset_host = offset.tolist()
        d = dict(coord=coord.to(dev, non_blocking=True), feat=feat.to(dev, non_blocking=True),
                 offset=offset.to(dev, non_blocking=True))
        out = self.forward(d, offset_host)
        return out["score"].cpu(), out["pred"].cpu()

    @torch.no_grad()
    def infer_stream(self, rooms, depth: int = 2, device=None, graphs="auto"):
        """Serve a sequence of rooms ``(coord, feat, offset)`` (host tensors, pinned recommended),
        yielding ``(score, pred)`` host tensors per room, in order.

        Same arithmetic as calling ``infer`` room by room; what changes is the schedule: up to
        ``depth`` rooms are in flight, so the serial FPS chains (16 SMs each) of the next rooms run
        under the feature path of the current one.

        graphs: "auto" (default) replays a CUDA graph of the whole room -- coordinate branch and
        feature branch with their fork/join, ~220 launches -- when every room of the call has the
        same scene sizes (a graph is captured per size signature and stream slot, once); rooms of
        differing sizes are launched eagerly.  False forces eager launches, True requires graphs."""
        dev = device if device is not None else next(self.parameters()).device
        rooms = list(rooms)
        if not rooms:
            return
        use_graphs = False
        if graphs and not self.training:
            sigs = {(tuple(C.host_offset(r[2]) if r[2].is_cuda else r[2].tolist()), int(r[1].shape[1])) for r in rooms}
            use_graphs = len(sigs) == 1 and (graphs is True or len(rooms) >= 2 or next(iter(sigs)) in self._graphs)
            if graphs is True and len(sigs) != 1:
                raise ValueError("infer_stream(graphs=True) needs rooms of identical scene sizes")
        if use_graphs:
            yield from self._infer_stream_graphs(rooms, next(iter(sigs)), depth, dev)
        else:
            yield from self._infer_stream_eager(rooms, depth, dev)

    MAX_GRAPH_SIGNATURES = 4   # captured size signatures kept alive (each holds `depth` private memory pools)

    def _infer_stream_graphs(self, rooms, sig, depth, dev):
        # graphs captured before the weights / the method / a backend switch changed are stale: drop them
        for k in [k for k, slots in self._graphs.items()
                  if any(g.generation != _GENERATION[0] or g.method != self.method for g in slots)]:
            del self._graphs[k]
        if sig in self._graphs:
            self._graphs[sig] = self._graphs.pop(sig)          # most recently used last
        while len(self._graphs) >= self.MAX_GRAPH_SIGNATURES and sig not in self._graphs:
            self._graphs.pop(next(iter(self._graphs)))         # drop the least recently used signature's graphs
        slots = self._graphs.setdefault(sig, [])
        while len(slots) < min(depth, len(rooms)):
            slots.append(_RoomGraph(self, sig[0], sig[1], dev))
        n_slots = len(slots)
        inflight = {}

        def launch(i):
            coord, feat, _ = rooms[i]
            inflight[i] = slots[i % n_slots].run(coord, feat)

        for i in range(min(n_slots, len(rooms))):
            launch(i)
        for i in range(len(rooms)):
            done, score, pred = inflight.pop(i)
            done.synchronize()
            if i + n_slots < len(rooms):
                launch(i + n_slots)   # the slot's previous results are already in their own host buffers
            yield score, pred

    def _infer_stream_eager(self, rooms, depth, dev):
        """One room after the other on the main stream; the host->device copy and the coordinate-only
        work (FPS, kNN) of the next ``depth`` rooms run ahead on their own streams, and the host only
        blocks on a room's result after the next room has been queued."""
        if len(self._streams) < depth:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        main = torch.cuda.current_stream(dev)
        prepared = {}

        def prepare(i):
            coord, feat, offset = rooms[i]
            st = self._streams[i % depth]
            # run-ahead is bounded: geometry() makes `st` wait for what the main stream has queued
            # so far, i.e. room i's coordinate work starts once room i - depth has left the GPU
            pointops.clear_caches()  # nothing is reusable across rooms; drop what the last ones pinned
            offset_host = offset.tolist()
            with torch.cuda.stream(st):
                d = dict(coord=coord.to(dev, non_blocking=True), feat=feat.to(dev, non_blocking=True),
                         offset=offset.to(dev, non_blocking=True).int())
                ev = torch.cuda.Event()
                ev.record(st)
            for t in d.values():
                t.record_stream(main)
            levels = self.backbone.geometry(d["coord"], d["offset"], offset_host, st)
            prepared[i] = (d, offset_host, levels, ev)

        for i in range(min(depth, len(rooms))):
            prepare(i)
        pending = None
        for i in range(len(rooms)):
            d, offset_host, levels, ev = prepared.pop(i)
            main.wait_event(ev)
            out = self.forward(d, offset_host, levels)
            n_i = out["score"].shape[0]
            score = torch.empty((n_i,), dtype=torch.float32, pin_memory=True)
            pred = torch.empty((n_i,), dtype=torch.int32, pin_memory=True)
            score.copy_(out["score"], non_blocking=True)
            pred.copy_(out["pred"], non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if i + depth < len(rooms):
                prepare(i + depth)
            if pending is not None:
                pending[0].synchronize()
                yield pending[1], pending[2]
            pending = (done, score, pred)
        if pending is not None:
            pending[0].synchronize()
            yield pending[1], pending[2]
