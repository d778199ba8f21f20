"""The reference's data-path code, run as it is -- TEST INFRASTRUCTURE ONLY.

`reference_transforms()` imports the UNMODIFIED pointcept/datasets/transform.py (GridSample, SphereCrop) and
pointcept/datasets/utils.py (collate_fn) where the reference tree is mounted (the registry module it needs is the
reference's own pointcept/utils/registry.py).  torch_scatter is not installed: `scatter_mean` restates its documented
semantics (sum of the rows per index / max(count, 1)).  `grid_sample_stable` is GridSample with `kind="stable"` on its
argsort -- the one freedom numpy's default (unstable) sort leaves open -- for index-exact comparison."""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"


def available() -> bool:
    return os.path.exists(os.path.join(REF, "pointcept", "datasets", "transform.py"))


def reference_transforms():
    saved = {k: v for k, v in sys.modules.items() if k == "pointcept" or k.startswith("pointcept.")}
    for k in saved:
        del sys.modules[k]
    try:
        for name in ("pointcept", "pointcept.utils"):
            m = types.ModuleType(name); m.__path__ = []; sys.modules[name] = m
        mods = {}
        for name, rel in (("pointcept.utils.misc", "pointcept/utils/misc.py"), ("pointcept.utils.registry", "pointcept/utils/registry.py"),
                          ("ref_transform", "pointcept/datasets/transform.py"),
                          ("ref_dsutils", "pointcept/datasets/utils.py")):
            sp = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            mod = importlib.util.module_from_spec(sp)
            sys.modules[name] = mod
            sp.loader.exec_module(mod)
            mods[name] = mod
        return mods["ref_transform"], mods["ref_dsutils"]
    finally:
        for k in [k for k in sys.modules if k == "pointcept" or k.startswith("pointcept.") or k in ("ref_transform", "ref_dsutils")]:
            del sys.modules[k]
        sys.modules.update(saved)


def fnv_hash_vec(arr):
    """transform.py:911-925, statement by statement."""
    arr = arr.copy().astype(np.uint64, copy=False)
    hashed = np.uint64(14695981039346656037) * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        hashed *= np.uint64(1099511628211)
        hashed = np.bitwise_xor(hashed, arr[:, j])
    return hashed


def grid_sample_stable(coord, grid_size, rand=None):
    """GridSample.__call__ (:813-857) up to idx_unique, with a STABLE argsort; returns what the device version returns."""
    scaled = coord / np.array(grid_size)
    grid_coord = np.floor(scaled).astype(int)
    min_coord = grid_coord.min(0)
    grid_coord -= min_coord
    key = fnv_hash_vec(grid_coord)
    idx_sort = np.argsort(key, kind="stable")
    key_sort = key[idx_sort]
    _, inverse, count = np.unique(key_sort, return_inverse=True, return_counts=True)
    start = np.cumsum(np.insert(count, 0, 0)[0:-1])
    inv = np.zeros_like(inverse)
    inv[idx_sort] = inverse
    out = dict(idx_sort=idx_sort, count=count, inverse=inv, grid_coord=grid_coord, min_coord=min_coord * np.array(grid_size), start=start)
    if rand is not None:
        out["idx_unique"] = idx_sort[start + rand % count]
    return out


def scatter_mean(src, index, dim_size):
    src2 = src.reshape(src.shape[0], -1).astype(np.float64)
    out = np.zeros((dim_size, src2.shape[1]))
    cnt = np.zeros(dim_size)
    np.add.at(out, index, src2)
    np.add.at(cnt, index, 1)
    out /= np.maximum(cnt, 1)[:, None]
    return out.reshape(dim_size) if src.ndim == 1 else out
