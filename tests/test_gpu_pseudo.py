"""f-3 on the GPU: the neighbour graph (tp.ball_query partial_dense semantics from the search grid) against the oracle's
restatement of torch-points-kernels' kernel, and the device-side region growth against regions produced by the
REFERENCE's own function (tests/golden/pseudo_small.pt).

Index work is bit-exact.  The growth compares floating-point similarities computed on the GPU with the reference's CPU
run; a region is a chaotic function of those (one different pick changes every later iteration), so the fixtures are
scenes where no decision falls within rounding distance -- they must come out IDENTICAL."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "pseudo_small.pt")


@pytest.mark.parametrize("sizes,radius,k", [([6000, 2500], 0.1, 16), ([20000], 0.08, 32), ([300, 5, 900], 0.3, 8)])
def test_ball_query_partial_dense_equals_oracle(cuda, sizes, radius, k):
    from oracle import pseudo_oracle as PO
    from pointcloudpdf_b200 import synthetic as S
    from pointcloudpdf_b200.pseudo import ball_query_partial_dense, scene_neighbors
    b = S.scannet_batch(sizes, seed=41)
    coord, offset = b["coord"], b["offset"]
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    ref_idx, ref_d2 = PO.ball_query_partial_dense(radius, k, coord, coord, batch, batch)
    idx, d2 = ball_query_partial_dense(radius, k, coord.to(cuda), coord.to(cuda), batch.to(cuda), batch.to(cuda))
    assert idx.dtype == torch.int64
    assert torch.equal(idx.cpu(), ref_idx)
    assert torch.allclose(d2.cpu(), ref_d2, rtol=1e-5, atol=1e-9)
    idx2, _ = ball_query_partial_dense(radius, k, coord.to(cuda), coord.to(cuda), offset_x=offset.to(cuda), offset_y=offset.to(cuda))
    assert torch.equal(idx2, idx)
    loc = scene_neighbors(idx, offset.tolist())
    ref_loc = PO.scene_neighbors(ref_idx, offset.tolist())
    assert all(torch.equal(a.cpu(), r) for a, r in zip(loc, ref_loc))


def test_region_growth_equals_the_reference_regions(cuda):
    from oracle import pseudo_oracle as PO
    from pointcloudpdf_b200.pseudo import ball_query_partial_dense, grow_unknown_region
    from pointcloudpdf_b200.scoring import pseudo_label_prefix
    for c in torch.load(GOLD):
        coord, logits = c["coord"].to(cuda), c["logits"].to(cuda)
        n = coord.shape[0]
        off = torch.tensor([n], dtype=torch.int32, device=cuda)
        nbrs, _ = ball_query_partial_dense(c["radius"], c["max_neighbor"], coord, coord, offset_x=off, offset_y=off)
        r = pseudo_label_prefix(logits, off, c["beta"], c["condition_from"], c["seed_from"], c["seed_range"])   # fused scoring pass (a11)
        score = r["msp"] if c["condition_from"] == "msp" else r["ml"]
        # the reference's seed draw, reproduced with its generator state (CPU randint, like the golden run)
        msp, ml, _s, _stop = PO.scores_and_stop(c["logits"], c["condition_from"], c["beta"])
        torch.manual_seed(c["torch_seed"])
        seeds = PO.draw_seeds(ml if c["seed_from"] == "ml" else msp, c["seed_range"], c["num_seed"])
        region = grow_unknown_region(coord, score, nbrs, seeds.to(cuda), r["stop"][0], c["slide_window"])
        assert torch.equal(region.cpu(), c["region"]), (region.numel(), c["region"].numel())


def test_region_growth_no_cpu_path():
    from pointcloudpdf_b200.pseudo import grow_unknown_region
    with pytest.raises(ValueError):
        grow_unknown_region(torch.rand(10, 3), torch.rand(10), torch.zeros(10, 4, dtype=torch.long), torch.tensor([0]), 0.5)
