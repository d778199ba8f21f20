"""Seeded synthetic inputs shaped like the reference's S3DIS / ScanNet batches.

No dataset is available offline, so tests and bench.py use these generators.  They imitate what
the reference's data pipeline hands to the model:
  * points on the surfaces of a room (floor, ceiling, 4 walls, axis-aligned furniture boxes);
  * one point kept per occupied voxel with its CONTINUOUS coordinate, like
    ``GridSample(mode="train")`` (pointcept/datasets/transform.py:825-857), grid 0.04 m for S3DIS
    and 0.02 m for ScanNet (configs/s3dis/openseg-pt-v1-0-msp.py:83-89);
  * ``PositiveShift`` (min -> 0, transform.py:138-144) or ``CenterShift``, then ``ShufflePoint``;
  * feat = cat(coord, colour[, normal]) (Collect feat_keys), offset = cumulative counts
    (pointcept/datasets/utils.py:34-39).
Continuous coordinates keep exact d2 ties at measure zero.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch


def _surface_points(rng: np.random.Generator, n_cand: int, dims: Sequence[float], n_boxes: int):
    """Uniform samples on the room shell and on furniture boxes; returns (xyz, normal)."""
    X, Y, Z = dims
    faces = []  # (origin, u, v, normal)
    def box_faces(lo, hi, inward):
        lo, hi = np.asarray(lo, float), np.asarray(hi, float)
        d = hi - lo
        sgn = -1.0 if inward else 1.0
        for ax in range(3):
            u, v = [a for a in range(3) if a != ax]
            for side in (0, 1):
                o = lo.copy()
                if side:
                    o[ax] = hi[ax]
                eu, ev = np.zeros(3), np.zeros(3)
                eu[u], ev[v] = d[u], d[v]
                nrm = np.zeros(3)
                nrm[ax] = sgn * (1.0 if side else -1.0)
                faces.append((o, eu, ev, nrm))
    box_faces([0, 0, 0], [X, Y, Z], inward=True)
    for _ in range(n_boxes):
        size = rng.uniform([0.06 * X, 0.08 * Y, 0.15 * Z], [0.25 * X, 0.25 * Y, 0.55 * Z])
        lo = np.array([rng.uniform(0.02 * X, 0.98 * X - size[0]), rng.uniform(0.02 * Y, 0.98 * Y - size[1]), 0.0])
        box_faces(lo, lo + size, inward=False)
    areas = np.array([np.linalg.norm(np.cross(f[1], f[2])) for f in faces])
    pick = rng.choice(len(faces), size=n_cand, p=areas / areas.sum())
    a, b = rng.random(n_cand), rng.random(n_cand)
    O = np.stack([f[0] for f in faces])[pick]
    U = np.stack([f[1] for f in faces])[pick]
    V = np.stack([f[2] for f in faces])[pick]
    N = np.stack([f[3] for f in faces])[pick]
    xyz = O + a[:, None] * U + b[:, None] * V
    xyz += rng.normal(0.0, 0.004, xyz.shape)  # sensor noise: keeps points off exact planes
    return xyz, N


def room(n_points: int, seed: int, voxel: float = 0.04, shift: str = "positive", n_boxes: int = 10,
         base_dims: Sequence[float] = (8.0, 6.0, 3.0)) -> Dict[str, np.ndarray]:
    """One scene with exactly ``n_points`` points (float32 coord, colour in [0,1], normal)."""
    rng = np.random.default_rng(seed)
    # scale the footprint so the voxelised shell holds ~1.5x the requested points
    X, Y, Z = base_dims
    shell = 2 * (X * Y) + 2 * Z * (X + Y)
    need = 1.5 * n_points * voxel * voxel
    s = max(np.sqrt(need / shell), 0.2)
    dims = (X * s, Y * s, max(Z * min(s, 1.0), 2.0 * voxel * 8))
    coord = np.zeros((0, 3))
    normal = np.zeros((0, 3))
    for attempt in range(6):
        xyz, nrm = _surface_points(rng, int(n_points * (4 + 2 * attempt)), dims, n_boxes)
        key = np.floor(xyz / voxel).astype(np.int64) + (1 << 19)
        packed = (key[:, 0] << 42) | (key[:, 1] << 21) | key[:, 2]
        _, first = np.unique(packed, return_index=True)  # one point per occupied voxel
        coord, normal = xyz[first], nrm[first]
        if coord.shape[0] >= n_points:
            break
    if coord.shape[0] < n_points:
        raise RuntimeError("synthetic room too small for the requested point count")
    perm = rng.permutation(coord.shape[0])[:n_points]  # ShufflePoint + crop to point_max
    coord, normal = coord[perm], normal[perm]
    if shift == "positive":
        coord = coord - coord.min(0)
    else:
        c = (coord.min(0) + coord.max(0)) / 2
        coord = coord - np.array([c[0], c[1], coord[:, 2].min()])
    color = rng.random((n_points, 3))
    return dict(coord=coord.astype(np.float32), color=color.astype(np.float32), normal=normal.astype(np.float32))


def s3dis_batch(sizes: Sequence[int], seed: int = 2024, num_classes: int = 13) -> Dict[str, torch.Tensor]:
    """S3DIS-shaped batch: feat = cat(coord, colour) (6 ch), labels U{0..12} (host tensors)."""
    scenes = [room(n, seed + 17 * i, voxel=0.04, shift="positive") for i, n in enumerate(sizes)]
    coord = torch.from_numpy(np.concatenate([s["coord"] for s in scenes]))
    color = torch.from_numpy(np.concatenate([s["color"] for s in scenes]))
    g = torch.Generator().manual_seed(seed)
    segment = torch.randint(0, num_classes, (coord.shape[0],), generator=g)
    offset = torch.tensor(np.cumsum(sizes), dtype=torch.int32)
    return dict(coord=coord, feat=torch.cat([coord, color], 1), offset=offset, segment=segment)


def scannet_batch(sizes: Sequence[int], seed: int = 2027, num_classes: int = 20) -> Dict[str, torch.Tensor]:
    """ScanNet-shaped batch: 0.02 m voxels, CenterShift, feat = coord + colour in [-1,1] +
    normal (9 ch), 20 classes."""
    scenes = [room(n, seed + 17 * i, voxel=0.02, shift="center", base_dims=(5.0, 4.0, 2.6)) for i, n in enumerate(sizes)]
    coord = torch.from_numpy(np.concatenate([s["coord"] for s in scenes]))
    color = torch.from_numpy(np.concatenate([s["color"] for s in scenes])) * 2 - 1
    normal = torch.from_numpy(np.concatenate([s["normal"] for s in scenes]))
    g = torch.Generator().manual_seed(seed)
    segment = torch.randint(0, num_classes, (coord.shape[0],), generator=g)
    offset = torch.tensor(np.cumsum(sizes), dtype=torch.int32)
    return dict(coord=coord, feat=torch.cat([coord, color, normal], 1), offset=offset, segment=segment)


def openset_logits(n: int, num_classes: int, seed: int = 2028, unknown_frac: float = 0.15):
    """Logits (n, K) = randn * 3 with planted low-confidence rows, conf (n, 1) and labels whose
    'unknown' rows correlate with low confidence (so AUROC / AUPR are informative)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, num_classes, generator=g) * 3
    unknown = torch.rand(n, generator=g) < unknown_frac
    logits[unknown] *= 0.35
    conf = torch.randn(n, 1, generator=g) + unknown.float().unsqueeze(1) * 1.5
    label = torch.randint(0, num_classes, (n,), generator=g)
    return logits.contiguous(), conf.contiguous(), unknown, label
