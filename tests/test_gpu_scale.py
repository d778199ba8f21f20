"""Parity at the sizes of BASELINE configs[2] / configs[4] (the small-size tests live in test_gpu_parity.py).

* cfg5: grid kNN at 250 000 and 1 000 000 points, k = 16 and 32, against the exhaustive kernel of the same
  library (same key (d2, idx), same d2 chain; itself pinned to the oracle at small sizes) -- bit for bit.  The
  oracle's O(n^2) scan would take minutes here; a strided subset of the queries IS checked against it.
* cfg3: backward of grouping / subtraction / aggregation / interpolation on one 100 000-point scene with real
  kNN indices (high fan-in scatter-adds, the RED.v4 paths) against the oracle's C restatement, 1e-5 relative.
* large-scene FPS (> 131 072 points: the grid-wide kernel) against the oracle over thousands of dependent samples.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REL = 1e-5


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def room(sizes, seed):
    from pointcloudpdf_b200 import synthetic as S
    b = S.s3dis_batch(sizes, seed=seed)
    return b["coord"], b["offset"]


@pytest.mark.parametrize("n", [250_000, 1_000_000])
@pytest.mark.parametrize("k", [16, 32])
def test_knn_grid_equals_bruteforce_kernel_at_cfg5_sizes(cuda, oracle, n, k):
    import pointops
    from pointcloudpdf_b200 import _lib
    xyz, offset = room([n], 2029)
    xyz_d, off_d = xyz.to(cuda), offset.to(cuda)
    idx, dist = pointops.knn_query(k, xyz_d, off_d)
    assert (idx[:, 0] == torch.arange(n, device=cuda)).all() and (dist[:, 1:] >= dist[:, :-1]).all()
    # exhaustive kernel on a slice of the queries (every 1 / stride-th block of rows keeps it to seconds at 1 M)
    m = 65536
    sel = torch.arange(0, n, n // m, device=cuda)[:m]
    q = xyz_d[sel].contiguous()
    qoff = torch.tensor([m], dtype=torch.int32, device=cuda)
    bidx = torch.empty((m, k), dtype=torch.int32, device=cuda)
    bdist = torch.empty((m, k), dtype=torch.float32, device=cuda)
    ws = torch.empty(64, dtype=torch.uint8, device=cuda)
    rc = _lib.load().pob_knn_query_bruteforce(m, k, 1, _lib.ptr(xyz_d), _lib.ptr(q), _lib.ptr(off_d), _lib.ptr(qoff),
                                              _lib.ptr(bidx), _lib.ptr(bdist), 1, _lib.ptr(ws), 64, _lib.current_stream(cuda))
    assert rc == 0
    assert torch.equal(idx[sel], bidx) and torch.equal(dist[sel], bdist)
    # and the oracle itself on a few hundred of those queries
    few = sel[:: m // 256].cpu()
    o_idx, o_dist = oracle.knn_query(k, xyz, offset, xyz[few].contiguous(), torch.tensor([few.numel()], dtype=torch.int32))
    assert torch.equal(idx[few.to(cuda)].cpu(), o_idx) and torch.equal(dist[few.to(cuda)].cpu(), o_dist)


def test_backward_kernels_at_cfg3_scene_size(cuda, oracle):
    import pointops
    n, ns, c, wc = 100_000, 16, 64, 8
    xyz, offset = room([n], 2027)
    xyz_d, off_d = xyz.to(cuda), offset.to(cuda)
    idx_d, _ = pointops.knn_query(ns, xyz_d, off_d)
    idx = idx_d.cpu()
    g = torch.Generator().manual_seed(5)
    feat, feat2 = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    pos = torch.randn(n, ns, c, generator=g)
    w = torch.softmax(torch.randn(n, ns, wc, generator=g), 1)
    gout3 = torch.randn(n, ns, c, generator=g)
    gout2 = torch.randn(n, c, generator=g)
    # grouping2 / grouping
    x = feat.to(cuda).requires_grad_(True)
    pointops.grouping2(x, idx_d).backward(gout3.to(cuda))
    assert rel_err(x.grad, oracle.grouping2_bwd(gout3, idx, n)) <= REL
    x = feat.to(cuda).requires_grad_(True)
    gx = torch.randn(n, ns, 3 + c, generator=g)
    pointops.grouping(idx_d, x, xyz_d, xyz_d, with_xyz=True).backward(gx.to(cuda))
    assert rel_err(x.grad, oracle.grouping2_bwd(gx[:, :, 3:].contiguous(), idx, n)) <= REL
    # subtraction
    a, b = feat.to(cuda).requires_grad_(True), feat2.to(cuda).requires_grad_(True)
    pointops.subtraction(a, b, idx_d).backward(gout3.to(cuda))
    g1, g2 = oracle.subtraction_bwd(idx, gout3)
    assert rel_err(a.grad, g1) <= REL and rel_err(b.grad, g2) <= REL
    # aggregation
    t = [v.to(cuda).requires_grad_(True) for v in (feat, pos, w)]
    out = pointops.aggregation(t[0], t[1], t[2], idx_d)
    assert rel_err(out, oracle.aggregation(feat, pos, w, idx)) <= REL
    out.backward(gout2.to(cuda))
    gi, gp, gw = oracle.aggregation_bwd(feat, pos, w, idx, gout2)
    assert rel_err(t[0].grad, gi) <= REL and rel_err(t[1].grad, gp) <= REL and rel_err(t[2].grad, gw) <= REL
    # interpolation (25 000 coarse -> 100 000 fine)
    new_offset = torch.tensor([n // 4], dtype=torch.int32)
    sel = pointops.farthest_point_sampling(xyz_d, off_d, new_offset.to(cuda)).long()
    coarse = xyz_d[sel].contiguous()
    cf = torch.randn(n // 4, c, generator=g)
    f_ref = cf.clone().requires_grad_(True)
    ref = oracle.interpolation(coarse.cpu(), xyz, f_ref, new_offset, offset)
    ref.backward(gout2)
    f = cf.to(cuda).requires_grad_(True)
    up = pointops.interpolation(coarse, xyz_d, f, new_offset.to(cuda), off_d)
    assert rel_err(up, ref) <= REL
    up.backward(gout2.to(cuda))
    assert rel_err(f.grad, f_ref.grad) <= REL


@pytest.mark.parametrize("n,m,kind", [(250_000, 6000, "room"), (400_000, 3000, "volume"), (1_300_000, 1200, "room")])
def test_fps_large_scene_thousands_of_samples(cuda, oracle, n, m, kind):
    """Scenes beyond the cluster-resident capacity (131 072 points) against the oracle, thousands of dependent
    samples deep; 1.3 M points exceeds what 148 SMs hold in registers and takes the shared-memory-points form."""
    import pointops
    if kind == "room":
        xyz, offset = room([n], 2031)
    else:
        g = torch.Generator().manual_seed(n)
        xyz = torch.rand(n, 3, generator=g) * torch.tensor([9.0, 7.0, 3.0])
        offset = torch.tensor([n], dtype=torch.int32)
    new_offset = torch.tensor([m], dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("sizes,ms", [([180_000, 3000, 40_000], [2500, 750, 2000]),      # one cluster each (48 points per thread)
                                      ([230_000, 3000, 150_000], [2000, 750, 1500])])     # grid-wide + two cluster scenes in one call
def test_fps_large_scene_in_a_batch_with_small_ones(cuda, oracle, sizes, ms):
    import pointops
    xyz, offset = room(sizes, 2033)
    new_offset = torch.tensor(np.cumsum(ms), dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert torch.equal(out.cpu(), ref)
