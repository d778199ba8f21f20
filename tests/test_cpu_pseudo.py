"""f-3 (pseudo-label neighbour graph + region growth), host side: the oracle's literal restatement of the growth loop
reproduces the regions the REFERENCE's own function produced (tests/golden/pseudo_small.pt, made by
tests/golden/make_golden.py::pseudo_small through oracle/pseudo_oracle.py::reference_growth), and -- where the reference
tree is mounted -- the reference function itself is run again and compared."""
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "pseudo_small.pt")


@pytest.fixture(scope="module")
def cases():
    return torch.load(GOLD)


def neighbours(c):
    from oracle import pseudo_oracle as PO
    n = c["coord"].shape[0]
    batch = torch.zeros(n, dtype=torch.long)
    return PO.ball_query_partial_dense(c["radius"], c["max_neighbor"], c["coord"], c["coord"], batch, batch)[0]


def test_oracle_growth_reproduces_the_reference_regions(cases):
    from oracle import pseudo_oracle as PO
    for c in cases:
        nbrs = neighbours(c)
        msp, ml, score, stop = PO.scores_and_stop(c["logits"], c["condition_from"], c["beta"])
        torch.manual_seed(c["torch_seed"])
        seeds = PO.draw_seeds(ml if c["seed_from"] == "ml" else msp, c["seed_range"], c["num_seed"])
        region = PO.grow_region(c["coord"], score, nbrs, seeds, stop, c["slide_window"])
        assert torch.equal(region, c["region"])
        assert region.numel() > 100          # the fixtures really grow (planted low-confidence blob)


def test_reference_function_itself_when_mounted(cases):
    from oracle import pseudo_oracle as PO
    if not os.path.exists("/root/reference/pointcept/recognizers/ours/pointpdf_v1m1_base.py"):
        pytest.skip("reference tree not mounted")
    c = cases[0]
    region = PO.reference_growth(c["coord"], c["logits"], neighbours(c), c["condition_from"], c["beta"], c["seed_from"],
                                 c["seed_range"], c["num_seed"], c["slide_window"], seed=c["torch_seed"])
    assert torch.equal(region, c["region"])


def test_ball_query_partial_dense_oracle_semantics():
    from oracle import pseudo_oracle as PO
    x = torch.tensor([[0.0, 0, 0], [0.05, 0, 0], [0.2, 0, 0], [0.0, 0.05, 0], [5.0, 5, 5], [5.05, 5, 5]])
    batch = torch.tensor([0, 0, 0, 0, 1, 1])
    idx, d2 = PO.ball_query_partial_dense(0.1, 2, x, x, batch, batch)
    assert idx.tolist() == [[0, 1], [0, 1], [2, -1], [0, 1], [4, 5], [4, 5]]     # first two IN INDEX ORDER, own scene only
    assert d2[2].tolist() == [0.0, -1.0]


def test_tp_compat_surface_and_no_cpu_path():
    """The stand-in for `import torch_points_kernels as tp` serves exactly the call the recognizer makes and nothing else;
    CPU tensors are rejected (no CPU path)."""
    from pointcloudpdf_b200 import tp_compat
    x = torch.rand(32, 3)
    b = torch.zeros(32, dtype=torch.long)
    with pytest.raises(NotImplementedError):
        tp_compat.ball_query(0.1, 8, x, x, mode="dense")
    with pytest.raises(NotImplementedError):
        tp_compat.ball_query(0.1, 8, x, x, mode="partial_dense", batch_x=b, batch_y=b, sort=True)
    with pytest.raises(ValueError):
        tp_compat.ball_query(0.1, 8, x, x, mode="partial_dense", batch_x=b, batch_y=b)
