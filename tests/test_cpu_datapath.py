"""f-4, host side (this container has the reference tree): the restatements in oracle/datapath_oracle.py against the
reference's own GridSample -- same voxel partition, counts and inverse; its FNV key function bit for bit."""
import numpy as np
import pytest


def test_restatement_equals_reference_gridsample():
    from oracle import datapath_oracle as DO
    if not DO.available():
        pytest.skip("reference tree not mounted")
    T, _ = DO.reference_transforms()
    rng = np.random.default_rng(5)
    coord = (rng.random((20000, 3)) * np.array([6.0, 4.0, 3.0]) - 2.0).astype(np.float32)
    g = np.floor(coord / np.array(0.05)).astype(int)
    g -= g.min(0)
    assert np.array_equal(DO.fnv_hash_vec(g), T.GridSample.fnv_hash_vec(g))
    d = T.GridSample(grid_size=0.05, mode="train", keys=("coord", "segment"), return_inverse=True, return_grid_coord=True)(
        dict(coord=coord.copy(), segment=np.arange(20000)))
    mine = DO.grid_sample_stable(coord, 0.05)
    assert np.array_equal(mine["inverse"], d["inverse"])
    assert np.array_equal(mine["inverse"][d["segment"]], np.arange(d["segment"].size))
    parts = T.GridSample(grid_size=0.05, mode="test", keys=("coord",))(dict(coord=coord.copy()))
    assert len(parts) == int(mine["count"].max())


def test_scatter_mean_restatement():
    from oracle import datapath_oracle as DO
    out = DO.scatter_mean(np.array([1.0, 3.0, 5.0, 7.0]), np.array([2, 2, 0, 2]), 4)
    assert np.allclose(out, [5.0, 0.0, 11.0 / 3.0, 0.0])
