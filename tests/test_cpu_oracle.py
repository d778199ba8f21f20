"""CPU tests (run with -m "not gpu"): the oracle against itself (C restatement vs pure-PyTorch
restatement vs literal reference mechanics), against the committed golden fixtures, and -- where
/root/reference exists -- against the reference's own Python wrappers running on top of it."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cloud(sizes, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(sum(sizes), 3, generator=g) * torch.tensor([5.0, 4.0, 3.0])
    return xyz, torch.tensor(np.cumsum(sizes), dtype=torch.int32)


@pytest.mark.parametrize("sizes,k", [([1500, 700], 16), ([900], 3), ([40, 5, 300], 8), ([200], 128)])
def test_knn_three_restatements_agree(oracle, sizes, k):
    xyz, offset = cloud(sizes, k)
    a = oracle.knn_query(k, xyz, offset)                 # C, contract key (d2, idx)
    b = oracle.knn_query(k, xyz, offset, tie="ref")      # C, literal heap of the reference
    c = oracle.knn_query_torch(k, xyz, offset)           # pure PyTorch
    for x, y in ((a, b), (a, c)):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])


def test_knn_invariants(oracle):
    xyz, offset = cloud([800, 6, 500], 3)
    idx, dist = oracle.knn_query(16, xyz, offset)
    off = [0] + offset.tolist()
    for s in range(3):
        blk = idx[off[s]:off[s + 1]]
        real = blk[blk >= 0]
        assert (real >= off[s]).all() and (real < off[s + 1]).all()          # scene isolation
    assert (idx[800:806, 6:] == -1).all() and (dist[800:806, 6:] == 1e5).all()  # placeholders
    assert (idx[:800, 0] == torch.arange(800)).all() and (dist[:800, 0] == 0).all()
    assert (dist[:, 1:] >= dist[:, :-1]).all()


def test_knn_tie_rule_is_lower_index(oracle):
    xyz = torch.tensor([[0.0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1]])
    offset = torch.tensor([6], dtype=torch.int32)
    idx, dist = oracle.knn_query(4, xyz, offset)
    assert idx[0].tolist() == [0, 1, 2, 3]      # five points at distance 1: the lowest indices win
    tidx, _ = oracle.knn_query_torch(4, xyz, offset)
    assert torch.equal(idx, tidx)


@pytest.mark.parametrize("sizes,stride", [([600, 300], 4), ([1024], 4), ([50, 7], 2)])
def test_fps_three_restatements_agree(oracle, sizes, stride):
    xyz, offset = cloud(sizes, stride)
    new_offset = torch.tensor(np.cumsum([max(s // stride, 1) for s in sizes]), dtype=torch.int32)
    a = oracle.farthest_point_sampling(xyz, offset, new_offset)
    b = oracle.farthest_point_sampling(xyz, offset, new_offset, tie="ref")
    c = oracle.farthest_point_sampling_torch(xyz, offset, new_offset)
    assert torch.equal(a, b) and torch.equal(a, c)
    off = [0] + offset.tolist()
    noff = [0] + new_offset.tolist()
    for s in range(len(sizes)):
        blk = a[noff[s]:noff[s + 1]]
        assert blk[0] == off[s]                                              # first sample = scene start
        assert (blk >= off[s]).all() and (blk < off[s + 1]).all()
        assert blk.unique().numel() == blk.numel()                            # distinct points -> distinct samples


def test_fps_tie_rule_lowest_index(oracle):
    xyz = torch.tensor([[0.0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, 1, 0]])
    offset = torch.tensor([5], dtype=torch.int32)
    out = oracle.farthest_point_sampling(xyz, offset, torch.tensor([3], dtype=torch.int32))
    assert out.tolist()[:2] == [0, 1]           # points 1,2,3,4 tie at distance 1: lowest index


def test_gather_ops_c_vs_torch_forms(oracle):
    g = torch.Generator().manual_seed(0)
    n, ns, c, w_c = 300, 8, 16, 4
    inp, inp2 = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    pos, w = torch.randn(n, ns, c, generator=g), torch.randn(n, ns, w_c, generator=g)
    idx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32)
    gout = torch.randn(n, c, generator=g)
    t = [x.clone().requires_grad_(True) for x in (inp, pos, w)]
    out = oracle.aggregation(t[0], t[1], t[2], idx)
    assert torch.allclose(out, oracle.aggregation_exact(inp, pos, w, idx), rtol=1e-5, atol=1e-5)
    out.backward(gout)
    gi, gp, gw = oracle.aggregation_bwd(inp, pos, w, idx, gout)
    for a, b in zip((t[0].grad, t[1].grad, t[2].grad), (gi, gp, gw)):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
    g3 = torch.randn(n, ns, c, generator=g)
    a, b = inp.clone().requires_grad_(True), inp2.clone().requires_grad_(True)
    oracle.subtraction(a, b, idx).backward(g3)
    g1, g2 = oracle.subtraction_bwd(idx, g3)
    assert torch.allclose(a.grad, g1, rtol=1e-4, atol=1e-5) and torch.allclose(b.grad, g2, rtol=1e-4, atol=1e-5)
    x = inp.clone().requires_grad_(True)
    oracle.grouping2(x, idx).backward(g3)
    assert torch.allclose(x.grad, oracle.grouping2_bwd(g3, idx, n), rtol=1e-4, atol=1e-5)


def test_aggregation_gradcheck_f64(oracle):
    g = torch.Generator().manual_seed(1)
    n, ns, c, w_c = 6, 3, 8, 2
    inp = torch.randn(n, c, generator=g, dtype=torch.float64, requires_grad=True)
    pos = torch.randn(n, ns, c, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(n, ns, w_c, generator=g, dtype=torch.float64, requires_grad=True)
    idx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32)
    assert torch.autograd.gradcheck(lambda a, b, c_: oracle.aggregation(a, b, c_, idx), (inp, pos, w))


def test_scores_closed_forms(oracle):
    g = torch.Generator().manual_seed(2)
    logits, conf = torch.randn(500, 13, generator=g) * 3, torch.randn(500, 1, generator=g)
    assert torch.allclose(oracle.msp_score(logits), torch.logsumexp(logits, -1) - logits.max(-1)[0], atol=1e-6)
    p = oracle.pdf_score(logits, conf)
    assert ((p > 0) & (p < 1)).all()
    lab = torch.randint(0, 13, (500,), generator=g)
    aupr, auroc = oracle.aupr_and_auroc(oracle.msp_score(logits), lab, [5, 9])
    assert 0 <= aupr <= 1 and 0 <= auroc <= 1
    assert oracle.aupr_and_auroc(p, torch.zeros(500, dtype=torch.long), [5, 9]) == (None, None)


# ----------------------------------------------------------------- golden fixtures --------

def _gold(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated")
    return torch.load(path)


def test_golden_ops_oracle_matches_reference_wrappers(oracle):
    """tests/golden/ops_small.pt holds outputs of the REFERENCE's functions/*.py (run over the
    literal C restatement by tests/golden/make_golden.py); the oracle's own forms must agree."""
    G = _gold("ops_small.pt")
    xyz, offset, feat = G["xyz"], G["offset"], G["feat"]
    idx, dist = oracle.knn_query(16, xyz, offset)
    assert torch.equal(idx, G["knn_idx"]) and torch.equal(dist, G["knn_dist"])
    fps = oracle.farthest_point_sampling(xyz, offset, G["new_offset"])
    assert torch.equal(fps, G["fps_idx"])
    assert torch.equal(oracle.grouping(idx, feat, xyz, xyz, True), G["grouped_xyz"])
    out, idx2 = oracle.knn_query_and_group(feat, xyz, offset, xyz[fps.long()].contiguous(), G["new_offset"],
                                           nsample=16, with_xyz=True)
    assert torch.equal(out, G["cross_grouped"]) and torch.equal(idx2, G["cross_idx"])
    interp = oracle.interpolation(xyz[fps.long()].contiguous(), xyz, feat[fps.long()].contiguous(), G["new_offset"], offset)
    assert torch.allclose(interp, G["interp"], rtol=1e-6, atol=1e-6)
    agg = oracle.aggregation(feat, G["pos"], G["w"], idx.clamp(min=0))
    assert float((agg - G["aggregation"]).abs().max() / G["aggregation"].abs().max()) <= 1e-5
    assert torch.equal(oracle.aggregation_exact(feat, G["pos"], G["w"], idx.clamp(min=0)), G["aggregation"])
    assert torch.equal(oracle.subtraction(feat, G["feat2"], idx.clamp(min=0)), G["subtraction"])
    assert torch.equal(oracle.grouping2(feat, idx.clamp(min=0)), G["grouping2"])
    q, qi = oracle.query_and_group(8, xyz, xyz, feat, None, offset, offset, dilation=1)
    assert torch.equal(q, G["qag"]) and torch.equal(qi, G["qag_idx"])


def test_golden_scores(oracle):
    G = _gold("scores_small.pt")
    assert torch.allclose(oracle.msp_score(G["logits"]), G["msp"], atol=1e-7)
    assert torch.equal(oracle.ml_score(G["logits"]), G["ml"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/libs/pointops/functions"), reason="reference tree absent")
def test_reference_wrappers_live(oracle):
    """Where the reference tree exists: run its Python wrappers now and compare (not a fixture)."""
    from oracle import ref_glue
    xyz, offset = cloud([700, 5, 400], 9)
    feat = torch.randn(xyz.shape[0], 32, generator=torch.Generator().manual_seed(1))
    with ref_glue.reference_modules() as R:
        ridx, rdist = R.pointops.knn_query(16, xyz, offset)
        rg = R.pointops.grouping(ridx, feat, xyz, xyz, with_xyz=True)
        new_offset = torch.tensor([175, 176, 276], dtype=torch.int32)
        rf = R.pointops.farthest_point_sampling(xyz, offset, new_offset)
    idx, dist = oracle.knn_query(16, xyz, offset)
    assert torch.equal(idx, ridx) and torch.equal(dist, rdist)
    assert torch.equal(oracle.grouping(idx, feat, xyz, xyz, True), rg)
    assert torch.equal(oracle.farthest_point_sampling(xyz, offset, new_offset), rf)


def test_radius_query_restatements_closed_forms(oracle):
    """ball_query / random_ball_query restatements on cases with known answers."""
    xyz = torch.tensor([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [10, 0, 0]], dtype=torch.float32)
    offset = torch.tensor([5], dtype=torch.int32)
    S = lambda *a: oracle.ball_query_dist2(*a, order="sorted")
    idx, d2 = S(4, 2.5, 0.0, xyz, offset)
    assert idx[0].tolist() == [0, 1, 2, -1] and d2[0].tolist() == [0.0, 1.0, 4.0, 1e10]
    assert idx[4].tolist() == [4, -1, -1, -1]
    # inner radius: the query itself (d2 <= 1e-5) is always accepted
    idx, _ = S(4, 2.5, 1.5, xyz, offset)
    assert idx[0].tolist() == [0, 2, -1, -1]
    # more candidates than nsample: every (cnt / nsample)-th of the list, "dist2" = the index (reference quirk)
    idx, d2 = S(2, 3.5, 0.0, xyz, offset)
    assert idx[0].tolist() == [0, 2] and d2[0].tolist() == [0.0, 2.0]
    # the reference's mechanics (default): heap_sort on the un-heapified scan-ordered list [0, 1, 4] -> [1, 4, 0]
    idx, dist = oracle.ball_query(4, 2.5, 0.0, xyz, offset)
    assert idx[0].tolist() == [1, 2, 0, -1] and dist[0].tolist() == [1.0, 2.0, 0.0, 1e5]
    # random ball query with the identity order = first nsample accepted rows in index order
    order = torch.arange(5, dtype=torch.int32)
    ridx, rd2 = oracle.random_ball_query_dist2(2, 3.5, 0.0, order, xyz, offset)
    assert ridx[3].tolist() == [0, 1] and rd2[3].tolist() == [9.0, 4.0]
    ridx, _ = oracle.random_ball_query_dist2(2, 3.5, 0.0, torch.tensor([4, 3, 2, 1, 0], dtype=torch.int32), xyz, offset)
    assert ridx[3].tolist() == [3, 2]
