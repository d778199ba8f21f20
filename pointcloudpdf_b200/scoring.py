"""Open-set uncertainty scoring, fused into one pass over the logits.

Host-side mirror of the reference recognizers' scoring arithmetic:
  * ``MaxProbability`` (pointcept/recognizers/max_probability/max_probability_v1m1_base.py:7-32)
  * ``PointPdfV1`` eval score and the scoring prefix of ``pseudo_labeling``
    (pointcept/recognizers/ours/pointpdf_v1m1_base.py:106-113, 199-227)
all served by ``pob_score_fused`` (csrc/score.cu).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .pointops import _common as C


def fused_scores(logits: torch.Tensor, conf: Optional[torch.Tensor] = None, offset: Optional[torch.Tensor] = None,
                 beta: float = 1.5, want=("msp_score", "ml_score")) -> Dict[str, torch.Tensor]:
    """One kernel pass over ``logits`` (n, K) f32 [and ``conf`` (n,) or (n, 1)].

    ``want`` picks per-point outputs among msp_score (= -max log_softmax), ml_score (= -max
    logit), pdf_score (= softmax(cat[logits, conf])[:, K]), msp_prob (= max softmax), max_logit,
    pred (argmax, int32), ml_norm (min-max normalised max logit per scene).  With ``offset`` the
    result also holds ``scene`` (b, 8): msp mean/std/stop, ml_norm mean/std/stop, min, max.
    """
    C.require(logits, "logits", torch.float32, 2)
    n, K = logits.shape
    dev = logits.device
    want = set(want)
    unknown = want - {"msp_score", "ml_score", "pdf_score", "msp_prob", "max_logit", "pred", "ml_norm"}
    if unknown:
        raise ValueError(f"unknown outputs requested: {sorted(unknown)}")
    if conf is not None:
        C.require(conf, "conf", torch.float32)
        if conf.numel() != n:
            raise ValueError("conf must hold one value per row of logits")
    if "pdf_score" in want and conf is None:
        raise ValueError("pdf_score needs conf")
    stats = offset is not None
    if "ml_norm" in want:
        if not stats:
            raise ValueError("ml_norm needs offset")
        want.add("max_logit")
    out: Dict[str, torch.Tensor] = {}
    for name in ("msp_score", "ml_score", "pdf_score", "msp_prob", "max_logit", "ml_norm"):
        if name in want:
            out[name] = torch.empty((n,), dtype=torch.float32, device=dev)
    if "pred" in want:
        out["pred"] = torch.empty((n,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    b, ws, ws_bytes, off32 = 0, None, 0, None
    if stats:
        off32 = C.offset_i32(offset)
        b = off32.numel()
        ws_bytes = lib.pob_score_workspace_bytes(b)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out["scene"] = torch.empty((b, 8), dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        n_out = sum(k != "scene" for k in out)
        _lib.run("pob_score_fused", n, K, b, _lib.ptr(logits), _lib.ptr(conf), _lib.ptr(off32), float(beta),
                 _lib.ptr(out.get("msp_score")), _lib.ptr(out.get("ml_score")), _lib.ptr(out.get("pdf_score")),
                 _lib.ptr(out.get("msp_prob")), _lib.ptr(out.get("max_logit")), _lib.ptr(out.get("pred")),
                 _lib.ptr(out.get("ml_norm")), _lib.ptr(out.get("scene")), _lib.ptr(ws), ws_bytes,
                 _lib.current_stream(dev), alg_bytes=4 * n * (K + (conf is not None) + n_out))
    return out


class MaxProbability(object):
    """Mirror of the reference recognizer ``MaxProbability`` (max_probability_v1m1_base.py:7-32):
    same constructor, ``__call__(input_dict) -> dict(score=...)``, reads the hooked backbone
    logits from ``self.model_hooks["backbone"]["forward_output"]``."""

    def __init__(self, method=None):
        if method == "msp":
            self.prob_func = self.msp
        elif method == "max_logits":
            self.prob_func = self.ml
        else:
            raise ValueError(f"Unknown MaxProbability method {method}")
        self.method = method
        self.model_hooks = None

    def __call__(self, input_dict):
        seg_logits = self.model_hooks["backbone"]["forward_output"]
        return dict(score=self.score(seg_logits))

    def score(self, seg_logits: torch.Tensor) -> torch.Tensor:
        key = "msp_score" if self.method == "msp" else "ml_score"
        return fused_scores(seg_logits.float().contiguous(), want=(key,))[key]

    def msp(self, seg_logits):  # = max log_softmax, like the reference helper (score is its negation)
        return -fused_scores(seg_logits.float().contiguous(), want=("msp_score",))["msp_score"]

    def ml(self, seg_logits):
        return -fused_scores(seg_logits.float().contiguous(), want=("ml_score",))["ml_score"]

    def set_epoch(self, epoch):
        self.epoch = epoch


def pdf_score(seg_logits: torch.Tensor, conf: torch.Tensor) -> torch.Tensor:
    """PointPdfV1 eval score (pointpdf_v1m1_base.py:110-113): unknown-class probability."""
    return fused_scores(seg_logits.float().contiguous(), conf.float().contiguous(), want=("pdf_score",))["pdf_score"]


def pseudo_label_prefix(seg_logits: torch.Tensor, offset: torch.Tensor, beta: float, condition_from: str = "msp",
                        seed_from: str = "ml", seed_range: float = 0.15):
    """Scoring prefix of PointPdfV1.pseudo_labeling (pointpdf_v1m1_base.py:199-227) for a whole
    batch: per-point msp / normalised ml, per-scene stop threshold, and the per-scene seed pool
    (lowest ``seed_range`` fraction by the seed score; the sort stays on torch).  The random seed
    draw and the graph growth after it are outside the kernel contract."""
    r = fused_scores(seg_logits.float().contiguous(), offset=offset, beta=beta, want=("msp_prob", "ml_norm"))
    scene = r["scene"]
    col = 2 if condition_from == "msp" else 5
    seed = r["msp_prob"] if seed_from == "msp" else r["ml_norm"]
    off = C.host_offset(C.offset_i32(offset))
    pools, s = [], 0
    for e in off:
        order = torch.sort(seed[s:e], dim=-1)[1]
        pools.append(order[: int(seed_range * (e - s))])
        s = e
    return dict(msp=r["msp_prob"], ml=r["ml_norm"], stop=scene[:, col], scene=scene, pools=pools)
