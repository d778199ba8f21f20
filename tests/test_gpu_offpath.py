"""GPU parity of the pointops names that are NOT on the PTv1 path: ball_query,
random_ball_query, ball_query_and_group, attention_relation_step, attention_fusion_step -- against the
oracle's restatements of the reference kernels and, where oracle/_ref is built, against the reference's
own CUDA kernels.  Index results bit-exact; float results of the atomically accumulated attention steps
within 1e-5 relative (summation order)."""
import ctypes
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpointops_ref.so")


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def cloud(sizes, seed):
    from pointcloudpdf_b200 import synthetic as S
    b = S.s3dis_batch(sizes, seed=seed)
    return b["coord"], b["offset"]


@pytest.mark.parametrize("sizes,nsample,max_r,min_r", [
    ([3000, 800], 16, 0.25, 0.0),      # most queries find more than nsample points: the strided pick (and the
    ([3000, 800], 8, 0.12, 0.05),      # reference's index-in-dist2 quirk); an inner radius
    ([2500], 32, 0.08, 0.0),           # fewer than nsample: sorted list + placeholders
    ([900, 6, 1500], 12, 0.4, 0.0),    # a 6-point scene (no grid), large balls
])
def test_ball_query_bit_exact(cuda, oracle, sizes, nsample, max_r, min_r):
    import pointops
    xyz, offset = cloud(sizes, 5)
    m_off = torch.tensor([s // 3 for s in sizes], dtype=torch.int32).cumsum(0).int()
    starts = [0] + offset.tolist()[:-1]
    new_xyz = torch.cat([xyz[a:a + s // 3] for a, s in zip(starts, sizes)]).contiguous()
    for q, qo in ((None, None), (new_xyz, m_off)):
        ref_idx, ref_dist = oracle.ball_query(nsample, max_r, min_r, xyz, offset, q, qo)
        idx, dist = pointops.ball_query(nsample, max_r, min_r, xyz.to(cuda), offset.to(cuda),
                                        None if q is None else q.to(cuda), None if qo is None else qo.to(cuda))
        assert idx.dtype == torch.int32 and dist.dtype == torch.float32
        assert torch.equal(idx.cpu(), ref_idx)
        assert torch.equal(dist.cpu(), ref_dist)
    grouped, idx2 = pointops.ball_query_and_group(torch.randn(xyz.shape[0], 8).to(cuda), xyz.to(cuda), offset.to(cuda),
                                                  max_radio=max_r, min_radio=min_r, nsample=nsample, with_xyz=True)
    assert grouped.shape == (xyz.shape[0], nsample, 11) and idx2.shape == (xyz.shape[0], nsample)


def test_ball_query_overflow_is_reported(cuda):
    import pointops
    xyz = (torch.rand(6000, 3) * 0.1).to(cuda)         # every point within 0.2 of every other: 6000 > 2048 candidates
    offset = torch.tensor([6000], dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError, match="2048"):
        pointops.ball_query(16, 0.5, 0.0, xyz, offset)


@pytest.mark.parametrize("sizes,nsample,max_r,min_r", [([3000, 800], 16, 0.25, 0.0), ([2500], 8, 0.1, 0.03),
                                                         ([700, 5], 40, 0.5, 0.0)])
def test_random_ball_query_bit_exact(cuda, oracle, sizes, nsample, max_r, min_r):
    import pointops
    xyz, offset = cloud(sizes, 6)
    g = torch.Generator().manual_seed(3)
    starts = [0] + offset.tolist()[:-1]
    order = torch.cat([torch.randperm(s, generator=g) + a for a, s in zip(starts, sizes)]).int()
    ref_idx, ref_d2 = oracle.random_ball_query_dist2(nsample, max_r, min_r, order, xyz, offset)
    idx, dist = pointops.random_ball_query(nsample, max_r, min_r, xyz.to(cuda), offset.to(cuda), None, None, order.to(cuda))
    assert torch.equal(idx.cpu(), ref_idx)
    assert torch.equal(dist.cpu(), oracle.sqrt_f32(ref_d2))
    # without an explicit order: a fresh permutation per call, same acceptance rule
    idx2, dist2 = pointops.random_ball_query(nsample, max_r, min_r, xyz.to(cuda), offset.to(cuda))
    assert torch.equal((idx2 >= 0).sum(1).cpu(), (ref_idx >= 0).sum(1))
    ok = idx2 >= 0
    d = (xyz.to(cuda)[idx2.clamp(min=0).long()] - xyz.to(cuda)[:, None, :]).norm(dim=-1)
    assert bool((d[ok] < max_r * (1 + 1e-5)).all())


def test_radius_queries_equal_reference_kernels(cuda, oracle):
    """The reference's own ball_query / random_ball_query kernels (compiled unmodified) on tie-free input."""
    import pointops
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libpointops_ref.so not built")
    ref = ctypes.CDLL(REF_SO)
    xyz, offset = cloud([4000, 1500], 8)
    xd, od = xyz.to(cuda), offset.to(cuda)
    m, ns = xyz.shape[0], 16
    F = ctypes.c_float
    for max_r, min_r in ((0.2, 0.0), (0.1, 0.04)):
        idx = torch.zeros((m, ns), dtype=torch.int32, device=cuda)
        d2 = torch.zeros((m, ns), dtype=torch.float32, device=cuda)
        torch.cuda.synchronize()
        ref.ball_query_cuda_launcher(ctypes.c_int(m), ctypes.c_int(ns), F(min_r), F(max_r), P(xd), P(xd), P(od), P(od), P(idx), P(d2))
        torch.cuda.synchronize()
        out_idx, out_dist = pointops.ball_query(ns, max_r, min_r, xd, od)
        assert torch.equal(out_idx, idx)
        assert torch.equal(out_dist, torch.sqrt(d2))
        g = torch.Generator().manual_seed(1)
        order = torch.cat([torch.randperm(4000, generator=g), torch.randperm(1500, generator=g) + 4000]).int().to(cuda)
        idx.zero_(); d2.zero_()
        torch.cuda.synchronize()
        ref.random_ball_query_cuda_launcher(ctypes.c_int(m), ctypes.c_int(ns), F(min_r), F(max_r), P(order), P(xd), P(xd),
                                            P(od), P(od), P(idx), P(d2))
        torch.cuda.synchronize()
        out_idx, out_dist = pointops.random_ball_query(ns, max_r, min_r, xd, od, None, None, order)
        assert torch.equal(out_idx, idx)
        assert torch.equal(out_dist, torch.sqrt(d2))


@pytest.mark.parametrize("n,g,c,m", [(500, 4, 16, 4000), (300, 1, 37, 1000), (64, 8, 64, 5000)])
def test_attention_steps_forward_backward(cuda, oracle, n, g, c, m):
    import pointops
    gen = torch.Generator().manual_seed(n + c)
    q = torch.randn(n, g, c, generator=gen)
    k = torch.randn(n, g, c, generator=gen)
    v = torch.randn(n, g, c, generator=gen)
    w = torch.randn(c, generator=gen)
    it = torch.randint(0, n, (m,), generator=gen)
    ir = torch.randint(0, n, (m,), generator=gen)
    go_rel = torch.randn(m, g, generator=gen)
    go_fus = torch.randn(n, g, c, generator=gen)
    wf = torch.randn(m, g, generator=gen)

    def rel_err(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp(min=1e-6))

    # oracle values and gradients through plain torch autograd on the restated expressions (f64)
    q64, k64, v64, wf64 = (t.double().requires_grad_(True) for t in (q, k, v, wf))
    ref_rel = oracle.attention_relation_step(q64, k64, w.double(), it, ir)
    ref_rel.backward(go_rel.double())
    ref_fus = oracle.attention_fusion_step(wf64, v64, it, ir)
    ref_fus.backward(go_fus.double())

    qd, kd, vd, wfd = (t.to(cuda).requires_grad_(True) for t in (q, k, v, wf))
    wd = w.to(cuda).requires_grad_(True)
    out = pointops.attention_relation_step(qd, kd, wd, it.int().to(cuda), ir.to(cuda))   # int32 and int64 indices
    assert out.shape == (m, g)
    assert rel_err(out.detach().cpu().double(), ref_rel.detach()) <= 1e-5
    out.backward(go_rel.to(cuda))
    assert rel_err(qd.grad.cpu().double(), q64.grad) <= 1e-5 and rel_err(kd.grad.cpu().double(), k64.grad) <= 1e-5
    assert wd.grad is None     # the reference returns None for weight (functions/attention.py:62)

    fus = pointops.attention_fusion_step(wfd, vd, it.to(cuda), ir.int().to(cuda))
    assert fus.shape == (n, g, c)
    assert rel_err(fus.detach().cpu().double(), ref_fus.detach()) <= 1e-5
    fus.backward(go_fus.to(cuda))
    assert rel_err(wfd.grad.cpu().double(), wf64.grad) <= 1e-5 and rel_err(vd.grad.cpu().double(), v64.grad) <= 1e-5
