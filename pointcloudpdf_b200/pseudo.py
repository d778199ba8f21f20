"""Device-side front half of PDF's pseudo-labelling: the neighbour graph and the region growth.

Reference: ``PointPdfV1.get_pseudo_mask`` / ``pseudo_labeling``
(pointcept/recognizers/ours/pointpdf_v1m1_base.py:118-185, 199-304).  After the fused scoring pass (a11,
``scoring.pseudo_label_prefix``) the reference

  1. builds a fixed-radius neighbour graph with ``tp.ball_query(radius, max_neighbor, coord, coord,
     mode="partial_dense", batch_x, batch_y)[0]`` -- torch-points-kernels, a dependency the reference neither vendors
     nor pins (and which this image does not have).  ``ball_query_partial_dense`` below returns the same thing (the
     first ``max_neighbor`` points of the query's scene, in index order, with d2 < radius^2, -1 padded) from the kNN
     search grid this library already builds: a radius query whose top-k key is the point index
     (``pob_random_ball_query`` with the identity permutation) -- one warp per query over the covering cells instead
     of one thread per query over the whole scene;
  2. grows a region from random low-confidence seeds: per iteration ``unique`` of the members' neighbours, ``isin``
     against the members, a similarity (distance to the region's centroid, score against a windowed mean), ``topk`` of
     40 %, ``unique`` of the union -- every one of them shape-dependent, i.e. a host sync, on 4 joblib threads.
     ``grow_unknown_region`` keeps the region as a membership MASK, so an iteration is a fixed sequence of O(n)
     device ops (scatter of the members' neighbour rows, masked reductions, one sort for the top 40 %) with no
     shape-dependent step; a finished scene turns its remaining iterations into no-ops through a device flag that
     the host reads only every few iterations.

MST / GMM / connected components after the growth stay on the CPU as in the reference (out of scope).
There is no CPU path: CUDA tensors only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib
from .pointops import _common as C


def ball_query_partial_dense(radius: float, max_neighbor: int, x: torch.Tensor, y: torch.Tensor,
                             batch_x: Optional[torch.Tensor] = None, batch_y: Optional[torch.Tensor] = None,
                             offset_x: Optional[torch.Tensor] = None, offset_y: Optional[torch.Tensor] = None):
    """``tp.ball_query(radius, max_neighbor, x, y, mode="partial_dense", batch_x=..., batch_y=...)``:
    idx (M, max_neighbor) int64 into x, -1 padded; dist2 (M, max_neighbor) f32 squared distances, -1 padded.
    Scenes are given as sorted batch vectors (like tp) or directly as cumulative offsets."""
    C.require(x, "x", torch.float32, 2, 3)
    C.require(y, "y", torch.float32, 2, 3)
    max_neighbor = int(max_neighbor)
    if max_neighbor < 1 or max_neighbor > 256:
        raise ValueError("max_neighbor must be in [1, 256]")
    dev = x.device

    def to_offset(batch, offset, n):
        if offset is not None:
            return C.offset_i32(offset)
        if batch is None:
            return C.const_offset([n], dev)
        b = int(batch[-1].item()) + 1 if n else 1   # tp needs sorted batch vectors; one sync, like its own host code
        return torch.cumsum(torch.bincount(batch, minlength=b), 0).to(torch.int32)

    off_x, off_y = to_offset(batch_x, offset_x, x.shape[0]), to_offset(batch_y, offset_y, y.shape[0])
    if off_x.numel() != off_y.numel():
        raise ValueError("x and y must describe the same number of scenes")
    n, m = x.shape[0], y.shape[0]
    idx = torch.empty((m, max_neighbor), dtype=torch.int32, device=dev)
    dist2 = torch.empty((m, max_neighbor), dtype=torch.float32, device=dev)
    inv = torch.empty((n,), dtype=torch.int32, device=dev)
    order = torch.arange(n, dtype=torch.int32, device=dev)   # identity permutation: "first accepted in index order"
    with _lib.device_guard(dev):
        grid = C.get_grid(x, off_x)
        _lib.run("pob_random_ball_query", m, max_neighbor, 0.0, float(radius), grid.n, grid.b, _lib.ptr(order), _lib.ptr(x),
                 _lib.ptr(y), _lib.ptr(off_y), grid.cell_pts, _lib.ptr(grid.workspace), _lib.ptr(inv), _lib.ptr(idx),
                 _lib.ptr(dist2), _lib.current_stream(dev), alg_bytes=20 * n + 12 * m + 8 * max_neighbor * m)
    found = idx >= 0
    return idx.long(), torch.where(found, dist2, torch.full_like(dist2, -1.0))


def scene_neighbors(neighbors: torch.Tensor, offset_host: Sequence[int]) -> List[torch.Tensor]:
    """Per scene, LOCAL neighbour indices -- what get_pseudo_mask hands to pseudo_labeling (:153-159)."""
    out, s = [], 0
    for e in offset_host:
        nn = neighbors[s:e]
        out.append(torch.where(nn >= 0, nn - s, nn))
        s = e
    return out


@torch.no_grad()
def grow_unknown_region(coord: torch.Tensor, score: torch.Tensor, neighbors: torch.Tensor, seeds: torch.Tensor,
                        stop_condition, slide_window: bool = False, check_every: int = 4, max_iter: int = 100000) -> torch.Tensor:
    """The growth loop of PointPdfV1.pseudo_labeling (:230-304) for one scene, on the device.

    coord (n, 3), score (n) = the condition score (msp or normalised max logit), neighbors (n, K) local indices
    (-1 = none), seeds (num_seed) = ``argsort(seed_score)[randint(...)]`` (may repeat), stop_condition = mean - beta *
    std of the condition score.  Returns the region's point indices, ascending (the reference's ``torch.unique``
    order); if the seeds already satisfy the stop test they are returned as they are, like the reference does."""
    for t, name in ((coord, "coord"), (score, "score"), (neighbors, "neighbors"), (seeds, "seeds")):
        if not t.is_cuda:
            raise ValueError(f"{name} must be a CUDA tensor (pointcloudpdf_b200 has no CPU path)")
    n, dev = coord.shape[0], coord.device
    stop = torch.as_tensor(stop_condition, dtype=score.dtype, device=dev)
    nbr = torch.where(neighbors >= 0, neighbors, torch.full_like(neighbors, n)).long()   # -1 -> a dump slot
    # iteration 0 sees the seed LIST (duplicates count twice in its means and its length), later ones a set
    w = torch.zeros(n, dtype=score.dtype, device=dev).index_add_(0, seeds.long(), torch.ones(seeds.numel(), dtype=score.dtype, device=dev))
    in_graph = w > 0
    done = torch.zeros((), dtype=torch.bool, device=dev)
    stopped_at_seeds = torch.zeros((), dtype=torch.bool, device=dev)
    ar = torch.arange(n, device=dev)
    inf = torch.tensor(float("inf"), dtype=score.dtype, device=dev)
    for it in range(max_iter):
        cnt = w.sum()
        mean_score = (w * score).sum() / cnt
        stop_now = (mean_score > stop) & (cnt > 0.01 * n) & (cnt > 50)
        if it == 0:
            stopped_at_seeds = stop_now.clone()
        active = ~done & ~stop_now
        center = (w[:, None] * coord).sum(0) / cnt
        # candidates: neighbours of members that are not members
        mark = torch.zeros(n + 1, dtype=torch.bool, device=dev)
        mark[torch.where(in_graph[:, None], nbr, torch.full_like(nbr, n)).reshape(-1)] = True
        cand = mark[:n] & ~in_graph
        d = torch.norm(coord - center, dim=-1)
        dmin = torch.where(cand, d, inf).min()
        dmax = torch.where(cand, d, -inf).max()
        dist_sim = 1 - (d - dmin) / (dmax - dmin + 1e-3)
        if slide_window:
            # kthvalue over the member scores, members counted with their multiplicity (iteration 0)
            gs, order = torch.sort(torch.where(in_graph, score, inf))
            cw = torch.cumsum(w[order], 0)                                   # cw[i] = members among the i + 1 smallest
            k1 = torch.floor(cnt.double() * 0.1).to(cw.dtype)
            k2 = torch.floor(cnt.double() * 0.6).to(cw.dtype)
            s_lo = gs[torch.searchsorted(cw, torch.clamp(k1, min=1.0)).clamp(max=n - 1)]
            s_hi = gs[torch.searchsorted(cw, torch.clamp(k2, min=1.0)).clamp(max=n - 1)]
            win = in_graph & (score >= s_lo) & (score <= s_hi)
            wmean = (w * score * win).sum() / (w * win).sum()
        else:
            wmean = mean_score
        conf_sim = torch.exp(-torch.abs(score - wmean))
        sim = 0.4 * dist_sim + 0.6 * conf_sim
        k = torch.floor(cand.sum().double() * 0.4).long()
        order = torch.sort(torch.where(cand, sim, -inf), descending=True, stable=True)[1]
        rank = torch.empty_like(order)
        rank[order] = ar
        new_in = in_graph | (cand & (rank < k) & active)
        new_cnt = new_in.sum().to(cnt.dtype)
        grew = new_cnt != cnt                                               # the reference compares lengths (:300)
        done = done | stop_now | (active & ~grew)
        in_graph = torch.where(active & grew, new_in, in_graph)
        w = torch.where(active & grew, new_in.to(w.dtype), w)
        if (it + 1) % check_every == 0 and bool(done):                      # the only host read
            break
    if bool(stopped_at_seeds) or bool((w > 1).any()):                       # never left the seed list
        return seeds
    return torch.nonzero(in_graph).reshape(-1)
