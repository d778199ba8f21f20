"""GPU parity: every operator of the PTv1 path, through the pointops API (ctypes -> C ABI ->
sm_100a kernels), against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): kNN / FPS indices and grouping bit-exact; aggregation,
interpolation and every backward within 1e-5 relative (f32); scores within 1e-6.
"""
import numpy as np
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-5  # stated tolerance for f32 accumulation-order differences


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_cloud(sizes, seed, kind="room"):
    from pointcloudpdf_b200 import synthetic as S
    if kind == "room":
        b = S.s3dis_batch(sizes, seed=seed)
        return b["coord"], b["offset"]
    g = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    xyz = torch.rand(n, 3, generator=g) * torch.tensor([4.0, 3.0, 2.0])
    return xyz, torch.tensor(np.cumsum(sizes), dtype=torch.int32)


# ------------------------------------------------------------------------------- kNN --

@pytest.mark.parametrize("sizes,k,kind", [
    ([24000], 16, "room"),            # BASELINE config 1
    ([5000, 300, 7, 12000], 16, "room"),   # ragged batch incl. a scene with < k points
    ([9000, 4000], 8, "volume"),
    ([3000], 3, "volume"),
    ([2500], 32, "room"),
    ([1500], 48, "volume"),           # two registers per lane
    ([400], 128, "volume"),           # reference maximum (stack arrays [128])
    ([1], 4, "volume"),
])
def test_knn_self_bit_exact(cuda, oracle, sizes, k, kind):
    import pointops
    xyz, offset = make_cloud(sizes, 100 + k, kind)
    ref_idx, ref_dist = oracle.knn_query(k, xyz, offset)
    idx, dist = pointops.knn_query(k, xyz.to(cuda), offset.to(cuda))
    assert idx.dtype == torch.int32 and dist.dtype == torch.float32
    assert torch.equal(idx.cpu(), ref_idx)
    assert torch.equal(dist.cpu(), ref_dist)  # sqrt(d2) with identical d2 bits


def test_knn_cross_query_and_placeholders(cuda, oracle):
    import pointops
    xyz, offset = make_cloud([6000, 2, 3000], 7, "room")
    g = torch.Generator().manual_seed(3)
    new_xyz = torch.cat([xyz[:6000][torch.randperm(6000, generator=g)[:1500]] + 0.01,
                         torch.rand(5, 3, generator=g), xyz[6002:][:700] * 1.3 - 0.2])
    new_offset = torch.tensor([1500, 1505, 2205], dtype=torch.int32)
    ref_idx, ref_dist = oracle.knn_query(16, xyz, offset, new_xyz, new_offset)
    idx, dist = pointops.knn_query(16, xyz.to(cuda), offset.to(cuda), new_xyz.to(cuda), new_offset.to(cuda))
    assert torch.equal(idx.cpu(), ref_idx) and torch.equal(dist.cpu(), ref_dist)
    assert (idx[1500:1505, 2:] == -1).all() and (dist[1500:1505, 2:] == 1e5).all()


def test_knn_exact_ties_lower_index_wins(cuda, oracle):
    """Lattice points: many exactly equal distances.  Contract: key (d2, idx)."""
    import pointops
    r = torch.arange(12, dtype=torch.float32)
    xyz = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3) * 0.5
    xyz = torch.cat([xyz, xyz[:100]])  # plus exact duplicates
    offset = torch.tensor([xyz.shape[0]], dtype=torch.int32)
    ref_idx, ref_dist = oracle.knn_query(16, xyz, offset)
    idx, dist = pointops.knn_query(16, xyz.to(cuda), offset.to(cuda))
    assert torch.equal(idx.cpu(), ref_idx) and torch.equal(dist.cpu(), ref_dist)


def test_knn_grid_equals_bruteforce_kernel_at_full_size(cuda):
    """Size-independent property at BASELINE's 80k: the binned search returns exactly what the
    exhaustive kernel (same key, same d2) returns."""
    import pointops
    from pointcloudpdf_b200 import _lib
    xyz, offset = make_cloud([80000], 2025, "room")
    xyz, offset = xyz.to(cuda), offset.to(cuda)
    for k in (8, 16):
        idx, dist = pointops.knn_query(k, xyz, offset)
        bidx = torch.empty_like(idx)
        bdist = torch.empty_like(dist)
        ws = torch.empty(64, dtype=torch.uint8, device=cuda)
        rc = _lib.load().pob_knn_query_bruteforce(80000, k, 1, _lib.ptr(xyz), _lib.ptr(xyz), _lib.ptr(offset),
                                                  _lib.ptr(offset), _lib.ptr(bidx), _lib.ptr(bdist), 1, _lib.ptr(ws),
                                                  64, _lib.current_stream(cuda))
        assert rc == 0
        assert torch.equal(idx, bidx) and torch.equal(dist, bdist)
        assert (idx[:, 0] == torch.arange(80000, device=cuda)).all()  # self is its own nearest
        assert (dist[:, 1:] >= dist[:, :-1]).all()                     # ascending


def test_knn_degenerate_geometry(cuda, oracle):
    import pointops
    g = torch.Generator().manual_seed(11)
    plane = torch.rand(4000, 3, generator=g); plane[:, 2] = 1.25          # zero extent in z
    line = torch.rand(1000, 3, generator=g); line[:, 1:] = 0.5            # zero extent in y, z
    same = torch.ones(600, 3) * 3.0                                        # all points identical
    far = torch.rand(3000, 3, generator=g); far[0] = torch.tensor([500.0, -300.0, 40.0])  # outlier
    xyz = torch.cat([plane, line, same, far])
    offset = torch.tensor([4000, 5000, 5600, 8600], dtype=torch.int32)
    ref_idx, ref_dist = oracle.knn_query(16, xyz, offset)
    idx, dist = pointops.knn_query(16, xyz.to(cuda), offset.to(cuda))
    assert torch.equal(idx.cpu(), ref_idx) and torch.equal(dist.cpu(), ref_dist)


# ------------------------------------------------------------------------------- FPS --

@pytest.mark.parametrize("sizes,stride", [
    ([24000], 4),
    ([5000, 300, 9, 12000], 4),
    ([700], 2),
    ([40000, 30000], 4),   # 16-CTA clusters, one per scene
    ([1250, 1250], 4),
])
def test_fps_bit_exact(cuda, oracle, sizes, stride):
    import pointops
    xyz, offset = make_cloud(sizes, 50 + stride, "room")
    new_offset = torch.tensor(np.cumsum([max(s // stride, 1) for s in sizes]), dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert out.dtype == torch.int32
    assert torch.equal(out.cpu(), ref)


def test_fps_full_size_80k_to_20k(cuda, oracle):
    """BASELINE's stage-1 size: 19 999 dependent iterations, 16-CTA cluster, cell-ordered pruning.
    Exact f32 ties between running minima do occur at this size; the contract (lowest index) decides."""
    import pointops
    xyz, offset = make_cloud([80000], 2026, "room")
    new_offset = torch.tensor([20000], dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("variant", ["merge", "merge_cells", "chain", "single"])
@pytest.mark.parametrize("sizes,stride", [([80000], 4), ([52000, 45000], 4), ([100000], 16), ([131072], 64),
                                          ([20000], 4), ([5000, 3000, 2049], 4)])
def test_fps_variants_same_result(cuda, oracle, variant, sizes, stride):
    """The three schedules of the resident kernel -- merged lists (default), one candidate per CTA (round 1),
    one sample per exchange -- walk the points in cell order with exact pruning (grid given) or strided (no
    grid): same arithmetic, same indices.  131 072 points = the largest scene the resident kernels take."""
    from pointcloudpdf_b200 import _lib
    from pointcloudpdf_b200.pointops import _common as C
    from pointcloudpdf_b200.pointops.sampling import fps_launch, VARIANTS
    xyz, offset = make_cloud(sizes, 77 + stride, "room")
    new_offset = torch.tensor(np.cumsum([max(s // stride, 1) for s in sizes]), dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    lib = _lib.load()
    xyz_d, off_d, noff_d = xyz.to(cuda), offset.to(cuda), new_offset.to(cuda)
    stats = torch.zeros(4, dtype=torch.int64, device=cuda)
    for with_grid in (True, False):          # cell-ordered + pruning, and the strided layout
        C.clear_caches()
        if with_grid:
            out = fps_launch(xyz_d, off_d, noff_d, offset.tolist(), new_offset.tolist(), variant=variant, stats=stats)
        else:
            out = torch.empty(int(new_offset[-1]), dtype=torch.int32, device=cuda)
            rc = lib.pob_farthest_point_sampling(len(sizes), max(sizes), _lib.ptr(xyz_d), _lib.ptr(off_d), _lib.ptr(noff_d),
                                                 None, _lib.ptr(out), 0, None, 0, 0.0, VARIANTS[variant], None,
                                                 _lib.current_stream(cuda))
            assert rc == 0
        assert torch.equal(out.cpu(), ref)
    if variant != "single":
        rounds, samples = stats.tolist()[:2]
        assert samples == int(new_offset[-1]) - len(sizes) and 0 < rounds <= samples


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_fps_every_cluster_size_same_result(cuda, oracle, cluster):
    from pointcloudpdf_b200 import _lib
    xyz, offset = make_cloud([6000, 2000], 99, "room")
    new_offset = torch.tensor([1500, 2000], dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    xyz_d, off_d, noff_d = xyz.to(cuda), offset.to(cuda), new_offset.to(cuda)
    out = torch.empty(2000, dtype=torch.int32, device=cuda)
    from pointcloudpdf_b200.pointops import _common as C
    grid = C.NeighbourGrid(xyz_d, off_d)
    for ws, n, cp in ((None, 0, 0.0), (_lib.ptr(grid.workspace), 8000, grid.cell_pts)):  # strided / cell order + pruning
        out.zero_()
        rc = _lib.load().pob_farthest_point_sampling(2, 6000, _lib.ptr(xyz_d), _lib.ptr(off_d), _lib.ptr(noff_d), None,
                                                     _lib.ptr(out), cluster, ws, n, cp, 0, None, _lib.current_stream(cuda))
        assert rc == 0
        assert torch.equal(out.cpu(), ref)


def test_fps_ties_lowest_index(cuda, oracle):
    import pointops
    r = torch.arange(10, dtype=torch.float32)
    xyz = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)
    xyz = torch.cat([xyz, xyz[:50]])
    offset = torch.tensor([xyz.shape[0]], dtype=torch.int32)
    new_offset = torch.tensor([400], dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("sizes,distinct", [([5000, 1200], 1), ([3000], 100), ([9000, 2600], 700)])
def test_fps_duplicate_points_all_distances_zero(cuda, oracle, sizes, distinct):
    """More samples requested than there are distinct positions: once every position is taken all
    remaining min-distances are 0 and the tie rule (lowest index) decides every further sample --
    including the fully degenerate cloud (every point identical) a zero-filled buffer gives."""
    import pointops
    g = torch.Generator().manual_seed(sum(sizes) + distinct)
    parts = []
    for n in sizes:
        base = torch.rand(distinct, 3, generator=g) * 4
        parts.append(base[torch.randint(0, distinct, (n,), generator=g)])
    xyz = torch.cat(parts)
    offset = torch.tensor(sizes, dtype=torch.int32).cumsum(0).int()
    new_offset = torch.tensor([n // 4 for n in sizes], dtype=torch.int32).cumsum(0).int()
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert torch.equal(out.cpu(), ref)


def test_fps_streamed_large_scene(cuda, oracle):
    """> 131072 points in one scene: the streamed kernel; few samples so the oracle stays fast."""
    import pointops
    xyz, offset = make_cloud([150000], 5, "volume")
    new_offset = torch.tensor([64], dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("kind", ["room", "volume"])
def test_fps_streamed_many_samples_and_mixed_batch(cuda, oracle, kind):
    """The tile-pruned streamed kernel (scene > 131072 points) over thousands of dependent samples, in a
    batch whose other scene is small (every scene of a launch runs the same kernel)."""
    import pointops
    xyz, offset = make_cloud([140000, 3000], 9, kind)
    new_offset = torch.tensor([1500, 1500 + 750], dtype=torch.int32)
    ref = oracle.farthest_point_sampling(xyz, offset, new_offset)
    out = pointops.farthest_point_sampling(xyz.to(cuda), offset.to(cuda), new_offset.to(cuda))
    assert torch.equal(out.cpu(), ref)


# --------------------------------------------------------------- grouping (a3, a4, a8) --

@pytest.mark.parametrize("c,ns,with_xyz,dtype", [
    (32, 16, True, torch.float32), (32, 8, False, torch.float32), (64, 16, True, torch.float16),
    (6, 16, True, torch.float32), (128, 16, True, torch.bfloat16), (512, 16, True, torch.float32),
])
def test_grouping_with_placeholders_bit_exact(cuda, oracle, c, ns, with_xyz, dtype):
    import pointops
    xyz, offset = make_cloud([3000, 5, 1200], 21, "room")
    g = torch.Generator().manual_seed(c)
    feat = torch.randn(xyz.shape[0], c, generator=g).to(dtype)
    idx, _ = oracle.knn_query(ns, xyz, offset)  # scene of 5 points -> -1 placeholders
    assert (idx < 0).any()
    ref = oracle.grouping(idx, feat, xyz, xyz, with_xyz)
    out = pointops.grouping(idx.to(cuda), feat.to(cuda), xyz.to(cuda), xyz.to(cuda), with_xyz)
    assert out.dtype == ref.dtype == torch.float32
    assert torch.equal(out.cpu(), ref)


def test_knn_query_and_group_config1(cuda, oracle):
    """BASELINE config 1: kNN(16) + group + aggregate on a 24k S3DIS-shaped cloud, C = 32."""
    import pointops
    xyz, offset = make_cloud([24000], 2025, "room")
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(24000, 32, generator=g)
    ref, ref_idx = oracle.knn_query_and_group(feat, xyz, offset, xyz, offset, nsample=16, with_xyz=True)
    out, idx = pointops.knn_query_and_group(feat.to(cuda), xyz.to(cuda), offset.to(cuda), xyz.to(cuda),
                                            offset.to(cuda), nsample=16, with_xyz=True)
    assert torch.equal(idx.cpu(), ref_idx) and torch.equal(out.cpu(), ref)
    pos = torch.randn(24000, 16, 32, generator=g)
    w = torch.randn(24000, 16, 4, generator=g)
    ref_agg = oracle.aggregation_exact(feat, pos, w, ref_idx)
    agg = pointops.aggregation(feat.to(cuda), pos.to(cuda), w.to(cuda), idx)
    assert rel_err(agg, ref_agg) <= REL


def test_grouping_backward(cuda, oracle):
    import pointops
    xyz, offset = make_cloud([2000, 6], 33, "room")
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(xyz.shape[0], 32, generator=g)
    idx, _ = oracle.knn_query(16, xyz, offset)
    gout = torch.randn(xyz.shape[0], 16, 35, generator=g)
    f_ref = feat.clone().requires_grad_(True)
    oracle.grouping(idx, f_ref, xyz, xyz, True).backward(gout)
    f = feat.to(cuda).requires_grad_(True)
    pointops.grouping(idx.to(cuda), f, xyz.to(cuda), xyz.to(cuda), True).backward(gout.to(cuda))
    assert rel_err(f.grad, f_ref.grad) <= REL


@pytest.mark.parametrize("c", [32, 64, 20, 256])
def test_grouping2_forward_backward(cuda, oracle, c):
    import pointops
    g = torch.Generator().manual_seed(c)
    n, m, ns = 3000, 2500, 16
    inp = torch.randn(n, c, generator=g)
    idx = torch.randint(0, n, (m, ns), generator=g, dtype=torch.int32)
    gout = torch.randn(m, ns, c, generator=g)
    x = inp.to(cuda).requires_grad_(True)
    out = pointops.grouping2(x, idx.to(cuda))
    assert torch.equal(out.detach().cpu(), oracle.grouping2(inp, idx))
    out.backward(gout.to(cuda))
    assert rel_err(x.grad, oracle.grouping2_bwd(gout, idx, n)) <= REL


def test_query_and_group_dilated(cuda, oracle):
    import pointops
    xyz, offset = make_cloud([2000, 900], 44, "room")
    g = torch.Generator().manual_seed(4)
    feat = torch.randn(xyz.shape[0], 16, generator=g)
    ref, ref_idx = oracle.query_and_group(8, xyz, xyz, feat, None, offset, offset, dilation=1)
    out, idx = pointops.query_and_group(8, xyz.to(cuda), xyz.to(cuda), feat.to(cuda), None, offset.to(cuda),
                                        offset.to(cuda), dilation=1)
    assert torch.equal(idx.cpu(), ref_idx) and torch.equal(out.cpu(), ref)
    legacy = pointops.queryandgroup(8, xyz.to(cuda), xyz.to(cuda), feat.to(cuda), None, offset.to(cuda),
                                    offset.to(cuda))
    ref0, _ = oracle.query_and_group(8, xyz, xyz, feat, None, offset, offset)
    assert torch.equal(legacy.cpu(), ref0)


# ----------------------------------------------------- subtraction / aggregation (a5, a6) --

@pytest.mark.parametrize("c", [32, 64, 12, 512])
def test_subtraction_forward_backward(cuda, oracle, c):
    import pointops
    g = torch.Generator().manual_seed(c + 1)
    n, ns = 2000, 16
    a, b = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
    idx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32)
    gout = torch.randn(n, ns, c, generator=g)
    a_d, b_d = a.to(cuda).requires_grad_(True), b.to(cuda).requires_grad_(True)
    out = pointops.subtraction(a_d, b_d, idx.to(cuda))
    assert torch.equal(out.detach().cpu(), oracle.subtraction(a, b, idx))  # one rounded op: bit-exact
    out.backward(gout.to(cuda))
    g1, g2 = oracle.subtraction_bwd(idx, gout)
    assert rel_err(a_d.grad, g1) <= REL and rel_err(b_d.grad, g2) <= REL


@pytest.mark.parametrize("c,share,ns", [(32, 8, 8), (32, 8, 16), (64, 8, 16), (128, 8, 16), (256, 8, 16),
                                        (512, 8, 16), (24, 4, 5), (32, 32, 16)])
def test_aggregation_forward_backward(cuda, oracle, c, share, ns):
    import pointops
    g = torch.Generator().manual_seed(c + share)
    n = 1500
    w_c = c // share
    inp, pos = torch.randn(n, c, generator=g), torch.randn(n, ns, c, generator=g)
    w = torch.softmax(torch.randn(n, ns, w_c, generator=g), dim=1)
    idx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32)
    gout = torch.randn(n, c, generator=g)
    t = [x.to(cuda).requires_grad_(True) for x in (inp, pos, w)]
    out = pointops.aggregation(t[0], t[1], t[2], idx.to(cuda))
    assert rel_err(out, oracle.aggregation_exact(inp, pos, w, idx)) <= REL
    assert rel_err(out, oracle.aggregation(inp, pos, w, idx)) <= REL
    out.backward(gout.to(cuda))
    gi, gp, gw = oracle.aggregation_bwd(inp, pos, w, idx, gout)
    assert rel_err(t[0].grad, gi) <= REL and rel_err(t[1].grad, gp) <= REL and rel_err(t[2].grad, gw) <= REL


def test_aggregation_equals_ptv1_einsum(cuda):
    """The op is the einsum PTv1 spells in torch (point_transformer_seg.py:75-80)."""
    import einops
    import pointops
    g = torch.Generator().manual_seed(8)
    n, ns, c, share = 1000, 16, 64, 8
    x_v = torch.randn(n, c, generator=g).to(cuda)
    p_r = torch.randn(n, ns, c, generator=g).to(cuda)
    w = torch.softmax(torch.randn(n, ns, c // share, generator=g), 1).to(cuda)
    idx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32).to(cuda)
    ein = torch.einsum("n t s i, n t i -> n s i",
                       einops.rearrange(x_v[idx.long()] + p_r, "n ns (s i) -> n ns s i", s=share), w)
    out = pointops.aggregation(x_v, p_r, w, idx)
    assert rel_err(out, einops.rearrange(ein, "n s i -> n (s i)")) <= REL


# ------------------------------------------------------------------ interpolation (a7) --

@pytest.mark.parametrize("c", [32, 64, 256, 13])
def test_interpolation_forward_backward(cuda, oracle, c):
    import pointops
    xyz, offset = make_cloud([4000, 1600], 61, "room")
    new_offset = torch.tensor([1000, 1400], dtype=torch.int32)
    sel = oracle.farthest_point_sampling(xyz, offset, new_offset).long()
    coarse = xyz[sel].contiguous()
    g = torch.Generator().manual_seed(c)
    feat = torch.randn(coarse.shape[0], c, generator=g)
    gout = torch.randn(xyz.shape[0], c, generator=g)
    f_ref = feat.clone().requires_grad_(True)
    ref = oracle.interpolation(coarse, xyz, f_ref, new_offset, offset)
    ref.backward(gout)
    f = feat.to(cuda).requires_grad_(True)
    out = pointops.interpolation(coarse.to(cuda), xyz.to(cuda), f, new_offset.to(cuda), offset.to(cuda))
    assert rel_err(out, ref) <= REL
    out.backward(gout.to(cuda))
    assert rel_err(f.grad, f_ref.grad) <= REL
    f2 = feat.to(cuda).requires_grad_(True)
    out2 = pointops.interpolation2(coarse.to(cuda), xyz.to(cuda), f2, new_offset.to(cuda), offset.to(cuda))
    assert rel_err(out2, ref) <= REL
    out2.backward(gout.to(cuda))
    assert rel_err(f2.grad, f_ref.grad) <= REL


def test_interpolation_fewer_than_k_coarse_points(cuda, oracle):
    """Quirk C6: placeholder -1 wraps to feat[-1] with weight ~1e-5."""
    import pointops
    xyz, offset = make_cloud([500, 400], 62, "volume")
    coarse = torch.cat([xyz[:2], xyz[500:520]]).contiguous()
    coff = torch.tensor([2, 22], dtype=torch.int32)
    feat = torch.randn(22, 32, generator=torch.Generator().manual_seed(1))
    ref = oracle.interpolation(coarse, xyz, feat, coff, offset)
    out = pointops.interpolation(coarse.to(cuda), xyz.to(cuda), feat.to(cuda), coff.to(cuda), offset.to(cuda))
    assert rel_err(out, ref) <= REL


# --------------------------------------------------------------- scoring (a9, a10, a11) --

@pytest.mark.parametrize("n,K", [(80000, 13), (150000, 20), (1000, 3), (257, 64)])
def test_scores_within_1e6(cuda, oracle, n, K):
    from pointcloudpdf_b200 import synthetic as S
    from pointcloudpdf_b200.scoring import fused_scores, MaxProbability, pdf_score
    logits, conf, unknown, label = S.openset_logits(n, K)
    r = fused_scores(logits.to(cuda), conf.to(cuda), want=("msp_score", "ml_score", "pdf_score", "pred", "msp_prob"))
    assert (r["msp_score"].cpu() - oracle.msp_score(logits)).abs().max() <= 1e-6
    assert torch.equal(r["ml_score"].cpu(), oracle.ml_score(logits))
    assert (r["pdf_score"].cpu() - oracle.pdf_score(logits, conf)).abs().max() <= 1e-6
    assert (r["msp_prob"].cpu() - torch.softmax(logits, -1).max(-1)[0]).abs().max() <= 1e-6
    assert torch.equal(r["pred"].cpu().long(), logits.argmax(-1))
    rec = MaxProbability(method="msp")
    rec.model_hooks = {"backbone": {"forward_output": logits.to(cuda)}}
    assert (rec({})["score"].cpu() - oracle.msp_score(logits)).abs().max() <= 1e-6
    assert (pdf_score(logits.to(cuda), conf.to(cuda)).cpu() - oracle.pdf_score(logits, conf)).abs().max() <= 1e-6


def test_auroc_aupr_identical_to_4_decimals(cuda, oracle):
    from pointcloudpdf_b200 import synthetic as S
    from pointcloudpdf_b200.scoring import fused_scores
    logits, conf, unknown, label = S.openset_logits(120000, 20)
    label = label.clone()
    label[unknown] = 4  # an 'unknown' class id of the ScanNet openseg config
    label[~unknown & (label == 4)] = 5
    r = fused_scores(logits.to(cuda), conf.to(cuda), want=("msp_score", "ml_score", "pdf_score"))
    for key, ref in (("msp_score", oracle.msp_score(logits)), ("ml_score", oracle.ml_score(logits)),
                     ("pdf_score", oracle.pdf_score(logits, conf))):
        a = oracle.aupr_and_auroc(r[key].cpu(), label, [4, 7, 14, 16])
        b = oracle.aupr_and_auroc(ref, label, [4, 7, 14, 16])
        assert round(a[0], 4) == round(b[0], 4) and round(a[1], 4) == round(b[1], 4)


def test_pseudo_label_prefix(cuda, oracle):
    from pointcloudpdf_b200 import synthetic as S
    from pointcloudpdf_b200.scoring import pseudo_label_prefix
    logits, _, _, _ = S.openset_logits(30000, 20, seed=5)
    offset = torch.tensor([9000, 9100, 30000], dtype=torch.int32)
    ref = oracle.pseudo_label_prefix(logits, offset, beta=1.5, condition_from="msp", seed_from="ml")
    out = pseudo_label_prefix(logits.to(cuda), offset.to(cuda), beta=1.5, condition_from="msp", seed_from="ml")
    s = 0
    for i, e in enumerate(offset.tolist()):
        assert (out["msp"][s:e].cpu() - ref[i]["msp"]).abs().max() <= 1e-6
        assert (out["ml"][s:e].cpu() - ref[i]["ml"]).abs().max() <= 1e-6
        assert abs(float(out["stop"][i]) - float(ref[i]["stop"])) <= 1e-6
        assert abs(float(out["scene"][i, 6]) - float(ref[i]["ml_min"])) == 0
        assert abs(float(out["scene"][i, 7]) - float(ref[i]["ml_max"])) == 0
        # seed pool: same points up to permutations among exactly equal scores
        a = ref[i]["ml"][out["pools"][i].cpu()]
        b = ref[i]["ml"][ref[i]["pool"]]
        assert torch.equal(torch.sort(a)[0], torch.sort(b)[0])
        s = e
    ml_cond = pseudo_label_prefix(logits.to(cuda), offset.to(cuda), beta=1.5, condition_from="ml")
    ref_ml = oracle.pseudo_label_prefix(logits, offset, beta=1.5, condition_from="ml")
    for i in range(3):
        assert abs(float(ml_cond["stop"][i]) - float(ref_ml[i]["stop"])) <= 1e-6


# ------------------------------------------------------------------- host-side contract --

def test_argument_validation(cuda):
    import pointops
    xyz = torch.rand(100, 3)
    off = torch.tensor([100], dtype=torch.int32)
    with pytest.raises(ValueError):
        pointops.knn_query(4, xyz, off)  # CPU tensors: no CPU path
    with pytest.raises(TypeError):
        pointops.knn_query(4, xyz.double().to(cuda), off.to(cuda))
    with pytest.raises(ValueError):
        pointops.knn_query(4, xyz.to(cuda).t().contiguous().t(), off.to(cuda))
    with pytest.raises(ValueError):
        pointops.knn_query(0, xyz.to(cuda), off.to(cuda))
    idx, _ = pointops.knn_query(4, xyz.to(cuda), off.long().to(cuda))  # int64 offsets accepted
    assert idx.shape == (100, 4)


def test_runs_on_non_default_stream(cuda, oracle):
    import pointops
    xyz, offset = make_cloud([5000], 71, "room")
    ref_idx, _ = oracle.knn_query(16, xyz, offset)
    s = torch.cuda.Stream()
    xyz_d, off_d = xyz.to(cuda), offset.to(cuda)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        idx, _ = pointops.knn_query(16, xyz_d, off_d)
    s.synchronize()
    assert torch.equal(idx.cpu(), ref_idx)
