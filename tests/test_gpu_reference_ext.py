"""Pin the oracle (and the product) against the REFERENCE's own CUDA kernels.

oracle/_ref/libpointops_ref.so is the reference's libs/pointops/src/**/_kernel.cu compiled
unmodified for sm_100a by oracle/Makefile (target `ref`, built where /root/reference exists; the
.so travels to the GPU box).  Its extern "C" launchers are called with raw device pointers, on
the legacy default stream they hard-code.  Inputs are continuous random coordinates, so exact d2
ties (where the reference's heap / block-reduction mechanics, not the contract, decide) do not
occur; the literal restatements in oracle_c.c are compared on tied inputs as well.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpointops_ref.so")


@pytest.fixture(scope="module")
def ref(cuda):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libpointops_ref.so not built (make -C oracle ref needs /root/reference)")
    return ctypes.CDLL(REF_SO)


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def ref_knn(ref, k, xyz, offset, new_xyz, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros((m, k), dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros((m, k), dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    ref.knn_query_cuda_launcher(ctypes.c_int(m), ctypes.c_int(k), P(xyz), P(new_xyz), P(offset), P(new_offset), P(idx), P(d2))
    torch.cuda.synchronize()
    return idx, d2


def ref_fps(ref, xyz, offset, new_offset, n_max, m):
    idx = torch.zeros((m,), dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    ref.farthest_point_sampling_cuda_launcher(ctypes.c_int(offset.numel()), ctypes.c_int(n_max), P(xyz), P(offset),
                                              P(new_offset), P(tmp), P(idx))
    torch.cuda.synchronize()
    return idx


def cloud(sizes, seed):
    from pointcloudpdf_b200 import synthetic as S
    b = S.s3dis_batch(sizes, seed=seed)
    return b["coord"], b["offset"]


@pytest.mark.parametrize("sizes,k", [([24000], 16), ([6000, 9, 3000], 16), ([5000], 8), ([2000], 3), ([700], 100)])
def test_knn_reference_kernel_equals_oracle_and_product(cuda, ref, oracle, sizes, k):
    import pointops
    xyz, offset = cloud(sizes, 300 + k)
    xyz_d, off_d = xyz.to(cuda), offset.to(cuda)
    r_idx, r_d2 = ref_knn(ref, k, xyz_d, off_d, xyz_d, off_d)
    o_idx, o_d2 = oracle.knn_query_dist2(k, xyz, offset, tie="ref")       # literal heap restatement
    c_idx, c_d2 = oracle.knn_query_dist2(k, xyz, offset, tie="contract")
    assert torch.equal(r_idx.cpu(), o_idx) and torch.equal(r_d2.cpu(), o_d2)  # restatement == reference, bit for bit
    assert torch.equal(o_idx, c_idx) and torch.equal(o_d2, c_d2)              # no ties here: contract agrees
    idx, dist = pointops.knn_query(k, xyz_d, off_d)
    assert torch.equal(idx, r_idx) and torch.equal(dist, torch.sqrt(r_d2))    # product == reference


def test_knn_reference_kernel_on_tied_input_matches_literal_restatement(cuda, ref, oracle):
    r = torch.arange(9, dtype=torch.float32)
    xyz = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3).contiguous()
    offset = torch.tensor([xyz.shape[0]], dtype=torch.int32)
    r_idx, r_d2 = ref_knn(ref, 16, xyz.to(cuda), offset.to(cuda), xyz.to(cuda), offset.to(cuda))
    o_idx, o_d2 = oracle.knn_query_dist2(16, xyz, offset, tie="ref")
    assert torch.equal(r_idx.cpu(), o_idx) and torch.equal(r_d2.cpu(), o_d2)
    # the multiset of distances is tie-rule independent
    c_idx, c_d2 = oracle.knn_query_dist2(16, xyz, offset, tie="contract")
    assert torch.equal(c_d2, o_d2)


@pytest.mark.parametrize("sizes", [[24000], [5000, 300, 12000], [1250, 600]])
def test_fps_reference_kernel_equals_oracle_and_product(cuda, ref, oracle, sizes):
    import pointops
    xyz, offset = cloud(sizes, 400 + len(sizes))
    new_offset = torch.tensor(np.cumsum([s // 4 for s in sizes]), dtype=torch.int32)
    xyz_d, off_d, noff_d = xyz.to(cuda), offset.to(cuda), new_offset.to(cuda)
    r = ref_fps(ref, xyz_d, off_d, noff_d, max(sizes), int(new_offset[-1]))
    assert torch.equal(r.cpu(), oracle.farthest_point_sampling(xyz, offset, new_offset, tie="ref"))
    assert torch.equal(r.cpu(), oracle.farthest_point_sampling(xyz, offset, new_offset, tie="contract"))
    assert torch.equal(pointops.farthest_point_sampling(xyz_d, off_d, noff_d), r)


def test_gather_kernels_reference_equals_product(cuda, ref, oracle):
    import pointops
    g = torch.Generator().manual_seed(9)
    n, ns, c, w_c = 4000, 16, 32, 4
    inp = torch.randn(n, c, generator=g).to(cuda)
    inp2 = torch.randn(n, c, generator=g).to(cuda)
    pos = torch.randn(n, ns, c, generator=g).to(cuda)
    w = torch.randn(n, ns, w_c, generator=g).to(cuda)
    idx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32).to(cuda)
    I = ctypes.c_int
    out = torch.empty(n, ns, c, device=cuda)
    ref.grouping_forward_cuda_launcher(I(n), I(ns), I(c), P(inp), P(idx), P(out)); torch.cuda.synchronize()
    assert torch.equal(pointops.grouping2(inp, idx), out)
    out = torch.zeros(n, ns, c, device=cuda)
    ref.subtraction_forward_cuda_launcher(I(n), I(ns), I(c), P(inp), P(inp2), P(idx), P(out)); torch.cuda.synchronize()
    assert torch.equal(pointops.subtraction(inp, inp2, idx), out)
    out = torch.zeros(n, c, device=cuda)
    ref.aggregation_forward_cuda_launcher(I(n), I(ns), I(c), I(w_c), P(inp), P(pos), P(w), P(idx), P(out)); torch.cuda.synchronize()
    mine = pointops.aggregation(inp, pos, w, idx)
    assert float((mine - out).abs().max() / out.abs().max()) <= 1e-5
    assert torch.equal(oracle.aggregation_exact(inp.cpu(), pos.cpu(), w.cpu(), idx.cpu()), out.cpu())  # same FMA order
    # backward: reference atomics vs product
    gout = torch.randn(n, c, generator=g).to(cuda)
    gi, gp, gw = torch.zeros_like(inp), torch.zeros_like(pos), torch.zeros_like(w)
    ref.aggregation_backward_cuda_launcher(I(n), I(ns), I(c), I(w_c), P(inp), P(pos), P(w), P(idx), P(gout), P(gi), P(gp), P(gw))
    torch.cuda.synchronize()
    t = [x.clone().requires_grad_(True) for x in (inp, pos, w)]
    pointops.aggregation(t[0], t[1], t[2], idx).backward(gout)
    for a, b in zip((t[0].grad, t[1].grad, t[2].grad), (gi, gp, gw)):
        assert float((a - b).abs().max() / b.abs().max()) <= 1e-5
    k = 3
    wi = torch.rand(n, k, generator=g).to(cuda)
    ii = torch.randint(0, n, (n, k), generator=g, dtype=torch.int32).to(cuda)
    out = torch.zeros(n, c, device=cuda)
    ref.interpolation_forward_cuda_launcher(I(n), I(c), I(k), P(inp), P(ii), P(wi), P(out)); torch.cuda.synchronize()
    from pointcloudpdf_b200.pointops.interpolation import _InterpolateRows
    assert float((_InterpolateRows.apply(inp, ii, wi) - out).abs().max() / out.abs().max()) <= 1e-5
