"""Multi-GPU forms of the path: one process per GPU, torch.distributed plumbing.

* ``shard_scenes``       -- which scenes of a batch a rank owns (the reference's DistributedSampler
                            + ``batch_size // world_size`` split, pointcept/engines/defaults.py:136-139):
                            scenes are independent units, so there is no data-path collective.
* ``sharded_knn_query``  -- one very large scene: every rank holds the whole reference set, answers a
                            contiguous slice of the queries with the local kernel, and the index
                            slices are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests).
* ``allreduce_gradients``-- the one training collective (what DDP does for the 31 MB of PTv1
                            gradients); kept explicit for the bench harness.
FPS of a single scene is one dependent chain and does not shard: replicas only.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_scenes(num_scenes: int, rank: int, world: int) -> List[int]:
    """Scene ids owned by `rank`: contiguous blocks of ceil(B/W), like batch_size // world_size."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-num_scenes // world)
    return list(range(min(rank * per, num_scenes), min((rank + 1) * per, num_scenes)))


def slice_batch(coord: torch.Tensor, feat: torch.Tensor, offset_host: Sequence[int], scenes: Sequence[int]):
    """Rows and re-based offsets of the given scenes of a concatenated batch."""
    starts = [0] + list(offset_host[:-1])
    rows, new_off, acc = [], [], 0
    for s in scenes:
        rows.append((starts[s], offset_host[s]))
        acc += offset_host[s] - starts[s]
        new_off.append(acc)
    if not rows:
        return coord[:0], feat[:0], []
    c = torch.cat([coord[a:b] for a, b in rows])
    f = torch.cat([feat[a:b] for a, b in rows])
    return c, f, new_off


def query_slices(offset_host: Sequence[int], rank: int, world: int) -> Tuple[List[Tuple[int, int]], List[int]]:
    """Per scene, the contiguous slice of query rows rank `rank` answers, and the cumulative
    new_offset of that slice set.  Scene s with rows [a, b) is cut into `world` near-equal parts."""
    starts = [0] + list(offset_host[:-1])
    spans, new_off, acc = [], [], 0
    for a, b in zip(starts, offset_host):
        n = b - a
        lo = a + (n * rank) // world
        hi = a + (n * (rank + 1)) // world
        spans.append((lo, hi))
        acc += hi - lo
        new_off.append(acc)
    return spans, new_off


def sharded_knn_query(nsample: int, xyz: torch.Tensor, offset: torch.Tensor, offset_host: Sequence[int],
                      knn_fn: Optional[Callable] = None, group=None):
    """kNN of every point of (xyz, offset) among its own scene, with the QUERIES sharded over the
    ranks of `group` and the reference set replicated.  Returns the full (N, nsample) idx and dist on
    every rank, bit-identical to the single-GPU result (queries are independent).

    knn_fn(nsample, xyz, offset, new_xyz, new_offset) -> (idx, dist); defaults to the B200 kernel.
    """
    if knn_fn is None:
        from .pointops import knn_query as knn_fn  # noqa: N813
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    spans, new_off = query_slices(offset_host, rank, world)
    q = torch.cat([xyz[a:b] for a, b in spans]).contiguous()
    new_offset = torch.tensor(new_off, dtype=torch.int32, device=xyz.device)
    idx_loc, dist_loc = knn_fn(nsample, xyz, offset, q, new_offset)
    if world == 1:
        return idx_loc, dist_loc
    # shards differ in length by at most B rows: pad to the maximum, all-gather, cut back
    all_spans = [query_slices(offset_host, r, world)[0] for r in range(world)]
    lens = [sum(b - a for a, b in sp) for sp in all_spans]
    m = max(lens)
    pad_i = torch.full((m, nsample), -1, dtype=idx_loc.dtype, device=xyz.device)
    pad_d = torch.zeros((m, nsample), dtype=dist_loc.dtype, device=xyz.device)
    pad_i[: idx_loc.shape[0]] = idx_loc
    pad_d[: dist_loc.shape[0]] = dist_loc
    gi = [torch.empty_like(pad_i) for _ in range(world)]
    gd = [torch.empty_like(pad_d) for _ in range(world)]
    dist.all_gather(gi, pad_i, group=group)
    dist.all_gather(gd, pad_d, group=group)
    n = xyz.shape[0]
    idx = torch.empty((n, nsample), dtype=idx_loc.dtype, device=xyz.device)
    dst = torch.empty((n, nsample), dtype=dist_loc.dtype, device=xyz.device)
    for r in range(world):
        pos = 0
        for a, b in all_spans[r]:
            idx[a:b] = gi[r][pos:pos + (b - a)]
            dst[a:b] = gd[r][pos:pos + (b - a)]
            pos += b - a
    return idx, dst


# ------------------------------------------------------------ fused query + all-gather --
# The all-gather above moves every result row twice (kernel -> local shard -> NCCL -> assembled result) and costs
# more than the search itself once the queries are split 8 ways (2 M x 32 results = 512 MB of idx + dist).  The
# fused form gives the query kernel the peer-mapped result buffers of ALL ranks (torch symmetric memory:
# cudaMalloc'ed, exchanged once through CUDA IPC / fabric handles) and lets it store each result row straight into
# every rank's (n, k) result over NVLink / NVSwitch -- the transfer overlaps the search warp by warp, and no
# assembly copy is left.  One device-side barrier over the symmetric signal pads closes the step.

_SYMM = {}


def _symm_buffers(n: int, k: int, device, group):
    """(idx, dist, handle_idx, handle_dist) symmetric (n, k) result buffers, rendezvoused once per shape."""
    import torch.distributed._symmetric_memory as symm_mem
    gname = group.group_name if group is not None else dist.group.WORLD.group_name
    key = (n, k, device.index, gname)
    hit = _SYMM.get(key)
    if hit is None:
        idx = symm_mem.empty((n, k), dtype=torch.int32, device=device)
        dst = symm_mem.empty((n, k), dtype=torch.float32, device=device)
        hit = _SYMM[key] = (idx, dst, symm_mem.rendezvous(idx, gname), symm_mem.rendezvous(dst, gname))
    return hit


def sharded_knn_query_fused(nsample: int, xyz: torch.Tensor, offset: torch.Tensor, offset_host: Sequence[int], group=None):
    """Same contract as sharded_knn_query (full (n, nsample) idx and dist on every rank, bit-identical to the
    single-GPU result) with the all-gather fused into the query kernel: P2P stores into every rank's result
    buffer (pob_knn_grid_query_scatter).  Single-scene clouds (the cfg5 case); the returned tensors are the
    symmetric buffers themselves and are overwritten by the next call of the same shape."""
    import ctypes
    from . import _lib
    from .pointops import _common as C
    if len(offset_host) != 1:
        raise ValueError("the fused form shards one large scene; batches shard by scene (shard_scenes)")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n, dev = int(xyz.shape[0]), xyz.device
    if world == 1:
        return C.get_grid(xyz, offset).query(nsample, xyz, offset, True, False)[:2]
    idx, dst, h_idx, h_dst = _symm_buffers(n, int(nsample), dev, group)
    (lo, hi), = query_slices(offset_host, rank, world)[0]
    grid = C.get_grid(xyz, offset)
    q = xyz[lo:hi]                                  # a contiguous slice: no copy
    new_offset = C.const_offset([hi - lo], dev)
    PtrArr = ctypes.c_void_p * world
    pi = PtrArr(*[int(p) for p in h_idx.buffer_ptrs])
    pd = PtrArr(*[int(p) for p in h_dst.buffer_ptrs])
    h_idx.barrier(channel=0)                        # nobody is still reading the previous result
    _lib.run("pob_knn_grid_query_scatter", hi - lo, int(nsample), grid.n, grid.b, _lib.ptr(xyz), _lib.ptr(q),
             _lib.ptr(new_offset), grid.cell_pts, _lib.ptr(grid.workspace), lo, world,
             ctypes.cast(pi, ctypes.c_void_p), ctypes.cast(pd, ctypes.c_void_p), 1, _lib.current_stream(dev),
             alg_bytes=12 * n + 12 * (hi - lo) + 8 * nsample * (hi - lo) * world)
    h_idx.barrier(channel=1)                        # every rank's rows have landed everywhere
    return idx, dst


def allreduce_gradients(params, group=None, bucket_bytes: int = 32 << 20) -> None:
    """Average gradients over the ranks in a few flat buckets (what DDP's hooks do)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, group=group)
        flat /= world
        pos = 0
        for g in bucket:
            g.copy_(flat[pos:pos + g.numel()].view_as(g))
            pos += g.numel()
        bucket, size = [], 0
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
