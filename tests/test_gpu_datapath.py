"""f-4 on the GPU: device GridSample / SphereCrop / collate / scatter_mean against the reference's own classes (imported
unmodified where the tree is mounted: this container) and against their restatements (oracle/datapath_oracle.py, which
travel).  Integer work is exact.  numpy's default argsort is unstable, so which point of a voxel is first is the one thing
the reference leaves open: voxel partition, counts, inverse and voxel coordinates are compared with the REAL GridSample,
picked indices with its stable-sort restatement."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cloud(n, seed):
    from pointcloudpdf_b200 import synthetic as S
    b = S.scannet_batch([n], seed=seed)
    g = torch.Generator().manual_seed(seed)
    c = b["coord"] + (torch.rand(n, 3, generator=g) - 0.5) * 0.05 - 1.3      # negative coordinates too, several points per voxel
    return c.contiguous()


@pytest.mark.parametrize("n,grid", [(50000, 0.05), (20000, 0.02), (3000, 0.3), (1, 0.05)])
def test_grid_sample_equals_stable_restatement(cuda, n, grid):
    from oracle import datapath_oracle as DO
    from pointcloudpdf_b200.datapath import grid_sample
    coord = cloud(n, 7 + n)
    rng = np.random.default_rng(n)
    ref = DO.grid_sample_stable(coord.numpy(), grid)
    rand = rng.integers(0, max(int(ref["count"].max()), 1), ref["count"].size)
    ref = DO.grid_sample_stable(coord.numpy(), grid, rand)
    out = grid_sample(coord.to(cuda), grid, mode="train", rand=torch.from_numpy(rand))
    assert torch.equal(out["idx_sort"].cpu(), torch.from_numpy(ref["idx_sort"]))
    assert torch.equal(out["count"].cpu(), torch.from_numpy(ref["count"]))
    assert torch.equal(out["inverse"].cpu(), torch.from_numpy(ref["inverse"]))
    assert torch.equal(out["grid_coord"].cpu().long(), torch.from_numpy(ref["grid_coord"]).long())
    assert torch.equal(out["idx_unique"].cpu(), torch.from_numpy(ref["idx_unique"]))
    assert np.allclose(out["min_coord"].cpu().numpy(), ref["min_coord"])
    parts = grid_sample(coord.to(cuda), grid, mode="test")["parts"]
    assert len(parts) == int(ref["count"].max())
    for i in (0, len(parts) - 1):
        assert torch.equal(parts[i].cpu(), torch.from_numpy(ref["idx_sort"][ref["start"] + i % ref["count"]]))
    # every point is covered by the union of the test parts, each part holds one point per voxel
    assert torch.unique(torch.cat(parts)).numel() == n and all(p.numel() == ref["count"].size for p in parts)


def test_grid_sample_partition_equals_the_reference_class(cuda):
    from oracle import datapath_oracle as DO
    from pointcloudpdf_b200.datapath import grid_sample
    if not DO.available():
        pytest.skip("reference tree not mounted (the GPU box): covered by the restatement test")
    T, _ = DO.reference_transforms()
    coord = cloud(30000, 3)
    seg = np.arange(30000)
    d = T.GridSample(grid_size=0.05, mode="train", keys=("coord", "segment"), return_inverse=True, return_grid_coord=True)(
        dict(coord=coord.numpy().copy(), segment=seg.copy()))
    out = grid_sample(coord.to(cuda), 0.05, mode="train")
    assert torch.equal(out["inverse"].cpu(), torch.from_numpy(d["inverse"]))           # same voxel id for every point
    picked = torch.from_numpy(d["segment"])                                              # the reference's one point per voxel
    assert torch.equal(out["inverse"].cpu()[picked], torch.arange(picked.numel()))       # ... one from each voxel, in voxel order
    mine = out["idx_unique"].cpu()
    assert torch.equal(out["inverse"].cpu()[mine], torch.arange(mine.numel()))
    assert torch.equal(out["grid_coord"].cpu().long()[picked], torch.from_numpy(d["grid_coord"]).long())


@pytest.mark.parametrize("n,pmax", [(60000, 20000), (5000, 8000)])
def test_sphere_crop_center(cuda, n, pmax):
    from pointcloudpdf_b200.datapath import sphere_crop
    coord = cloud(n, 11)
    idx = sphere_crop(coord.to(cuda), pmax, mode="center").cpu()
    c = coord.numpy()
    if n <= pmax:
        assert torch.equal(idx, torch.arange(n))
        return
    d2 = np.sum(np.square(c - c[n // 2]), 1)                                             # transform.py:1003-1005
    ref = np.argsort(d2, kind="stable")[:pmax]
    assert torch.equal(idx, torch.from_numpy(ref))


def test_collate_equals_the_reference_collate_fn(cuda):
    from oracle import datapath_oracle as DO
    from pointcloudpdf_b200.datapath import collate
    g = torch.Generator().manual_seed(0)
    scenes = [dict(coord=torch.rand(n, 3, generator=g), feat=torch.rand(n, 6, generator=g), segment=torch.randint(0, 13, (n,), generator=g),
                   offset=torch.tensor([n]), name=f"room{n}") for n in (500, 1, 1200)]
    out = collate([{k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in s.items()} for s in scenes])
    assert out["offset"].tolist() == [500, 501, 1701] and out["name"] == ["room500", "room1", "room1200"]
    assert torch.equal(out["coord"].cpu(), torch.cat([s["coord"] for s in scenes]))
    if DO.available():
        _, U = DO.reference_transforms()
        ref = U.collate_fn([dict(s) for s in scenes])
        for k in ("coord", "feat", "segment", "offset"):
            assert torch.equal(out[k].cpu(), ref[k])
        assert out["name"] == ref["name"]


@pytest.mark.parametrize("rows,c,dim", [(200000, 1, 50000), (30000, 13, 4000), (1000, 32, 5000)])
def test_scatter_mean(cuda, rows, c, dim):
    from oracle import datapath_oracle as DO
    from pointcloudpdf_b200.datapath import scatter_mean
    g = torch.Generator().manual_seed(rows)
    src = torch.rand(rows, generator=g) if c == 1 else torch.rand(rows, c, generator=g)
    index = torch.randint(0, dim, (rows,), generator=g)
    out = scatter_mean(src.to(cuda), index.to(cuda), dim).cpu().numpy()
    ref = DO.scatter_mean(src.numpy(), index.numpy(), dim)
    assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-5
