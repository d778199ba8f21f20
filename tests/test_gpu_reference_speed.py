"""The reference's own CUDA kernels (oracle/_ref, compiled unmodified for sm_100a) timed beside the
product's on the same B200 and the same inputs -- the true "before" for every kernel of the path.
Writes gpurun_out/reference_kernel_times.json when that directory exists (profiles/ keeps a copy)."""
import ctypes
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpointops_ref.so")


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def test_reference_kernels_vs_product(cuda):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libpointops_ref.so not built")
    import pointops
    from pointcloudpdf_b200 import synthetic as S
    ref = ctypes.CDLL(REF_SO)
    I = ctypes.c_int
    rows = {}
    for n, k in ((24000, 16), (80000, 8), (80000, 16)):
        b = S.s3dis_batch([n], seed=2025)
        xyz, off = b["coord"].to(cuda), b["offset"].to(cuda)
        idx = torch.zeros((n, k), dtype=torch.int32, device=cuda)
        d2 = torch.zeros((n, k), dtype=torch.float32, device=cuda)
        t_ref = timed(lambda: ref.knn_query_cuda_launcher(I(n), I(k), P(xyz), P(xyz), P(off), P(off), P(idx), P(d2)))

        def mine():
            pointops.clear_caches()
            return pointops.knn_query(k, xyz, off)
        t_mine = timed(mine)
        assert torch.equal(mine()[0], idx)
        rows[f"knn n={n} k={k}"] = (t_ref, t_mine)
    for n, m in ((80000, 20000), (20000, 5000), (5000, 1250)):
        b = S.s3dis_batch([n], seed=2026)
        xyz, off = b["coord"].to(cuda), b["offset"].to(cuda)
        noff = torch.tensor([m], dtype=torch.int32, device=cuda)
        out = torch.zeros(m, dtype=torch.int32, device=cuda)
        tmp = torch.empty(n, device=cuda)

        def run_ref():
            tmp.fill_(1e10)
            ref.farthest_point_sampling_cuda_launcher(I(1), I(n), P(xyz), P(off), P(noff), P(tmp), P(out))
        t_ref = timed(run_ref, reps=2)

        def mine():
            pointops.clear_caches()
            return pointops.farthest_point_sampling(xyz, off, noff)
        t_mine = timed(mine, reps=2)
        # exact f32 ties between two running minima do occur at 80k points (12 of 20000 positions in
        # this cloud): the reference resolves them by block mechanics, the contract by lowest index
        # (the product equals the contract oracle there, tests/test_gpu_parity.py).  The two orders
        # pick the tied points in swapped order, so the selected SET is the same.
        got = mine()
        assert torch.equal(torch.sort(got)[0], torch.sort(out)[0])
        assert (got != out).float().mean() < 2e-3
        rows[f"fps {n}->{m}"] = (t_ref, t_mine)
    n, ns, c, w_c = 80000, 8, 32, 4
    g = torch.Generator(device=cuda).manual_seed(0)
    b = S.s3dis_batch([n], seed=2025)
    idx, _ = pointops.knn_query(ns, b["coord"].to(cuda), b["offset"].to(cuda))
    f = torch.randn(n, c, device=cuda, generator=g)
    f2 = torch.randn(n, c, device=cuda, generator=g)
    pos = torch.randn(n, ns, c, device=cuda, generator=g)
    w = torch.randn(n, ns, w_c, device=cuda, generator=g)
    out3 = torch.empty(n, ns, c, device=cuda)
    out2 = torch.zeros(n, c, device=cuda)
    rows["grouping fwd 80k x8 x32"] = (timed(lambda: ref.grouping_forward_cuda_launcher(I(n), I(ns), I(c), P(f), P(idx), P(out3))),
                                       timed(lambda: pointops.grouping2(f, idx)))
    rows["subtraction fwd 80k x8 x32"] = (timed(lambda: ref.subtraction_forward_cuda_launcher(I(n), I(ns), I(c), P(f), P(f2), P(idx), P(out3))),
                                          timed(lambda: pointops.subtraction(f, f2, idx)))

    def agg_ref():
        out2.zero_()
        ref.aggregation_forward_cuda_launcher(I(n), I(ns), I(c), I(w_c), P(f), P(pos), P(w), P(idx), P(out2))
    rows["aggregation fwd 80k x8 x32"] = (timed(agg_ref), timed(lambda: pointops.aggregation(f, pos, w, idx)))
    gout = torch.randn(n, c, device=cuda, generator=g)
    gi, gp, gw = torch.zeros_like(f), torch.zeros_like(pos), torch.zeros_like(w)

    def aggb_ref():
        gi.zero_(); gw.zero_()
        ref.aggregation_backward_cuda_launcher(I(n), I(ns), I(c), I(w_c), P(f), P(pos), P(w), P(idx), P(gout), P(gi), P(gp), P(gw))
    fr, pr, wr = f.clone().requires_grad_(True), pos.clone().requires_grad_(True), w.clone().requires_grad_(True)
    o = pointops.aggregation(fr, pr, wr, idx)

    def aggb_mine():
        fr.grad = pr.grad = wr.grad = None
        o.backward(gout, retain_graph=True)
    rows["aggregation bwd 80k x8 x32"] = (timed(aggb_ref), timed(aggb_mine))
    table = {k: dict(reference_ms=a, b200_ms=b_, speedup=a / b_) for k, (a, b_) in rows.items()}
    print(json.dumps(table, indent=1))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(table, open(os.path.join(out_dir, "reference_kernel_times.json"), "w"), indent=1)
    assert table["fps 80000->20000"]["speedup"] > 2 and table["knn n=80000 k=8"]["speedup"] > 2
