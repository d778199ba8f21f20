"""Device-side data path either side of the model: what the reference does in numpy inside dataloader
workers (GridSample, SphereCrop), in ``collate_fn`` and with ``torch_scatter.scatter_mean`` in the tester.

  * ``grid_sample``   -- ``GridSample.__call__`` (pointcept/datasets/transform.py:813-907): voxel coordinates, FNV / ravel
                         key (``pob_grid_hash``), sort, run lengths; train mode picks one point per voxel, test mode
                         returns the ``count.max()`` index parts.  numpy's default ``argsort`` is unstable, so WHICH point of
                         a voxel comes first is unspecified there; here the sort is stable (lowest original index first).
                         Everything that does not depend on that -- the voxel partition, ``count``, ``inverse``,
                         ``grid_coord`` of the voxels -- is identical to the reference's.
  * ``sphere_crop``   -- ``SphereCrop`` modes "random" / "center" (:995-1023): the ``point_max`` nearest points of a centre,
                         ascending distance (ties: lower index).
  * ``collate``       -- ``collate_fn`` for the dict batches the configs use (pointcept/datasets/utils.py:15-41): tensors are
                         concatenated, ``offset`` keys become cumulative sums.
  * ``scatter_mean``  -- fragment score averaging of the tester (pointcept/engines/test.py:243-248).
CUDA tensors only (no CPU path)."""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional, Sequence

import torch

from . import _lib


def _cuda(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (pointcloudpdf_b200 has no CPU path)")
    return t


def voxel_keys(coord: torch.Tensor, grid_size=0.05, hash_type: str = "fnv"):
    """(grid_coord (n, 3) i32, key (n) i64 ordered like the reference's uint64 key, min_cell (3) i64)."""
    _cuda(coord, "coord")
    if coord.dtype != torch.float32 or coord.dim() != 2 or coord.shape[1] != 3 or not coord.is_contiguous():
        raise ValueError("coord must be a contiguous (n, 3) f32 tensor")
    gs = [float(grid_size)] * 3 if not isinstance(grid_size, (list, tuple)) else [float(g) for g in grid_size]
    n, dev = coord.shape[0], coord.device
    div = torch.tensor(gs, dtype=torch.float64, device=dev)
    min_cell = torch.floor(coord.double() / div).amin(0).long() if n else torch.zeros(3, dtype=torch.int64, device=dev)
    grid_coord = torch.empty((n, 3), dtype=torch.int32, device=dev)
    key = torch.empty((n,), dtype=torch.int64, device=dev)
    with _lib.device_guard(dev):
        _lib.run("pob_grid_hash", n, _lib.ptr(coord), gs[0], gs[1], gs[2], _lib.ptr(min_cell), _lib.ptr(grid_coord), _lib.ptr(key),
                 _lib.current_stream(dev), alg_bytes=12 * n + 12 * n + 8 * n)
    if hash_type != "fnv":   # ravel_hash_vec (:911-925): Fortran-style ravel of the non-negative cell coordinates
        g = grid_coord.long()
        mx = g.amax(0) + 1
        key = (g[:, 0] * mx[1] + g[:, 1]) * mx[2] + g[:, 2]
    return grid_coord, key, min_cell


def grid_sample(coord: torch.Tensor, grid_size=0.05, mode: str = "train", hash_type: str = "fnv",
                rand: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> Dict[str, object]:
    """GridSample on the device.  Returns
         idx_sort (n), count (V), inverse (n: voxel id of every point, voxels numbered by ascending key),
         grid_coord (n, 3) i32, min_coord (3) f64 (= min_cell * grid_size),
         train: idx_unique (V) = one point per voxel: idx_sort[start + rand % count], rand = randint(0, count.max(), V)
                (pass `rand` to fix the draw, as the parity test does);
         test:  parts = list of count.max() index tensors, part i = idx_sort[start + i % count]."""
    if mode not in ("train", "test"):
        raise ValueError("mode must be 'train' or 'test'")
    grid_coord, key, min_cell = voxel_keys(coord, grid_size, hash_type)
    dev = coord.device
    key_sort, idx_sort = torch.sort(key, stable=True)
    _, inverse_sorted, count = torch.unique_consecutive(key_sort, return_inverse=True, return_counts=True)
    inverse = torch.empty_like(inverse_sorted)
    inverse[idx_sort] = inverse_sorted                      # data_dict["inverse"][idx_sort] = inverse  (:844-845)
    start = torch.cumsum(count, 0) - count                  # np.cumsum(np.insert(count, 0, 0)[0:-1])
    gs = [float(grid_size)] * 3 if not isinstance(grid_size, (list, tuple)) else [float(g) for g in grid_size]
    out = dict(idx_sort=idx_sort, count=count, inverse=inverse, grid_coord=grid_coord,
               min_coord=min_cell.double() * torch.tensor(gs, dtype=torch.float64, device=dev))
    if mode == "train":
        if rand is None:
            hi = int(count.max().item()) if count.numel() else 1
            rand = torch.randint(0, max(hi, 1), (count.numel(),), device=dev, generator=generator)
        out["idx_unique"] = idx_sort[start + rand.to(dev) % count]
    else:
        cmax = int(count.max().item()) if count.numel() else 0
        out["parts"] = [idx_sort[start + i % count] for i in range(cmax)]
    return out


def sphere_crop(coord: torch.Tensor, point_max: int = 80000, mode: str = "random", center_index: Optional[int] = None,
                generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """SphereCrop "random" / "center": indices of the point_max points nearest to the centre, ascending squared
    distance (f32, sum of squares like np.sum(np.square(.), 1)); all indices in order when the cloud is small enough."""
    _cuda(coord, "coord")
    n = coord.shape[0]
    if n <= point_max:
        return torch.arange(n, device=coord.device)
    if center_index is None:
        if mode == "center":
            center_index = n // 2
        elif mode == "random":
            center_index = int(torch.randint(0, n, (1,), generator=generator).item())
        else:
            raise ValueError("mode must be 'random' or 'center'")
    sq = torch.square(coord - coord[center_index])
    d2 = (sq[:, 0] + sq[:, 1]) + sq[:, 2]      # the summation order of np.sum(np.square(.), 1) on rows of three: bit-identical f32
    return torch.sort(d2, stable=True)[1][:point_max]


def collate(batch: Sequence[Mapping[str, object]]) -> Dict[str, object]:
    """collate_fn for a list of per-scene dicts already on the device: tensors concatenated along dim 0, keys containing
    "offset" turned into cumulative sums (scenes contribute their own offset entries, e.g. tensor([n_i])), strings listed."""
    if not batch or not isinstance(batch[0], Mapping):
        raise TypeError("collate expects a non-empty sequence of dicts")
    out: Dict[str, object] = {}
    for key in batch[0]:
        vals = [d[key] for d in batch]
        if isinstance(vals[0], torch.Tensor):
            out[key] = torch.cat([_cuda(v, key) for v in vals])
        elif isinstance(vals[0], str):
            out[key] = list(vals)
        else:
            out[key] = torch.utils.data.dataloader.default_collate(vals)
    for key in out:
        if "offset" in key:
            out[key] = torch.cumsum(out[key], dim=0)
    return out


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter.scatter_mean(src, index, dim=0, dim_size=dim_size) for src (rows,) or (rows, c) f32 and index (rows)
    int64: mean of the rows sharing an index, 0 where none do."""
    _cuda(src, "src"); _cuda(index, "index")
    one_d = src.dim() == 1
    s = src.reshape(src.shape[0], -1).float().contiguous()
    rows, c = s.shape
    out = torch.zeros((dim_size, c), dtype=torch.float32, device=src.device)
    cnt = torch.zeros((dim_size,), dtype=torch.float32, device=src.device)
    with _lib.device_guard(src.device):
        _lib.run("pob_scatter_mean", rows, c, _lib.ptr(s), _lib.ptr(index.long().contiguous()), int(dim_size), _lib.ptr(out), _lib.ptr(cnt),
                 _lib.current_stream(src.device), alg_bytes=4 * rows * c + 8 * rows + 8 * dim_size * c)
    return out.reshape(dim_size) if one_d else out
