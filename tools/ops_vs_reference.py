"""Every gather operator, forward AND backward, timed kernel-to-kernel against the reference's own CUDA kernels
(oracle/_ref/libpointops_ref.so, compiled unmodified for sm_100a) on the same B200 and the same inputs, at
BASELINE configs[0]'s shape (24 000 points, k = 16, C = 32) and at PTv1 stage-1 size (80 000 points, k = 8, C = 32).

Both sides are called through their C entry points (no autograd, no allocator inside the timed region).
  * product: 24 launches captured in one CUDA graph over 8 rotating argument sets, CUDA events around a replay
    (median of 5) -- the method of bench.py's ops_cfg1;
  * reference: its launchers hard-code the legacy default stream (not capturable), so 24 launches are issued back to
    back on that stream between two events (the kernels take 20-700 us each, far longer than a ctypes launch, so the
    queue never runs dry); median of 5.
GB/s = SURVEY.md 8(d) algorithmic bytes / time; frac = GB/s / MEASURED_PEAKS.json hbm_gbs.

    python tools/ops_vs_reference.py [out.json]
"""
import ctypes
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pointcloudpdf_b200 import _lib, synthetic as S  # noqa: E402
import pointcloudpdf_b200.pointops as pointops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load() if torch.cuda.is_available() else None
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpointops_ref.so")
ref = ctypes.CDLL(REF_SO) if os.path.exists(REF_SO) else None
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, KeyError, ValueError):
    HBM = 6650.0
P, I, L = _lib.ptr, ctypes.c_int, ctypes.c_int64
REPS, NSETS = 24, 8


def time_graph(fn):
    for a in range(2):
        fn(a)
    torch.cuda.synchronize()
    st = torch.cuda.Stream(device=dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
        for r in range(REPS):
            fn(r % NSETS)
    ts = []
    with torch.cuda.stream(st):
        g.replay()
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) / REPS * 1e3   # us


def time_default_stream(fn):
    for a in range(2):
        fn(a)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(REPS):
            fn(r % NSETS)
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) / REPS * 1e3


def run_shape(tag, n, k, c, wc):
    b = S.s3dis_batch([n], seed=2025)
    xyz, off = b["coord"].to(dev), b["offset"].to(dev)
    idx, _ = pointops.knn_query(k, xyz, off)
    g = torch.Generator(device=dev).manual_seed(0)
    R = lambda *s: torch.randn(*s, device=dev, generator=g)
    sets = [dict(f=R(n, c), f2=R(n, c), pos=R(n, k, c), w=R(n, k, wc), o3=torch.empty(n, k, c, device=dev), o2=torch.empty(n, c, device=dev),
                 ox=torch.empty(n, k, 3 + c, device=dev), g3=R(n, k, c), g2=R(n, c), gx=R(n, k, 3 + c),
                 gi=torch.zeros(n, c, device=dev), gi2=torch.zeros(n, c, device=dev), gp=torch.zeros(n, k, c, device=dev),
                 gw=torch.zeros(n, k, wc, device=dev)) for _ in range(NSETS)]
    st = lambda: _lib.current_stream(dev)
    B = 4
    rows = {}

    def copy_floor(nbytes):
        """A plain device copy moving the same number of bytes (half read, half written), timed the same way: what a
        launch of this size can reach at all -- the fixed cost of a launch is 2-3 us, a third of the small rows."""
        half = max(4, nbytes // 8 * 4)
        bufs = [(torch.empty(half // 4, device=dev), torch.empty(half // 4, device=dev)) for _ in range(NSETS)]
        return time_graph(lambda a: bufs[a][1].copy_(bufs[a][0]))

    def row(name, nbytes, mine, theirs=None):
        us = time_graph(mine)
        cu = copy_floor(nbytes)
        r = {"b200_us": us, "alg_MB": nbytes / 1e6, "b200_GBps": nbytes / us / 1e3, "b200_frac_of_hbm_peak": nbytes / us / 1e3 / HBM,
             "same_bytes_copy_us": cu, "b200_frac_of_copy_speed": cu / us}
        if theirs is not None and ref is not None:
            ru = time_default_stream(theirs)
            r.update({"reference_us": ru, "reference_GBps": nbytes / ru / 1e3, "speedup": ru / us})
        rows[name] = r
        print(f"{tag:8s} {name:28s} b200 {us:8.2f} us {r['b200_GBps']:7.0f} GB/s ({r['b200_frac_of_hbm_peak']:.2f} of peak, {r['b200_frac_of_copy_speed']:.2f} of a same-bytes copy at {cu:.2f} us)"
              + (f"   reference {r['reference_us']:9.2f} us   x{r['speedup']:.1f}" if "speedup" in r else ""), flush=True)

    row("grouping2 fwd", B * (n * c + n * k + n * k * c),
        lambda a: _lib.check(lib.pob_grouping_forward(L(n), I(k), I(c), P(sets[a]["f"]), P(idx), P(sets[a]["o3"]), st()), "g"),
        lambda a: ref.grouping_forward_cuda_launcher(I(n), I(k), I(c), P(sets[a]["f"]), P(idx), P(sets[a]["o3"])))
    row("grouping2 bwd", B * (n * c + n * k + n * k * c),
        lambda a: _lib.check(lib.pob_grouping_backward(L(n), I(k), I(c), P(sets[a]["g3"]), P(idx), P(sets[a]["gi"]), st()), "g"),
        lambda a: ref.grouping_backward_cuda_launcher(I(n), I(k), I(c), P(sets[a]["g3"]), P(idx), P(sets[a]["gi"])))
    row("group with_xyz fwd", B * (n * c + 3 * n + 3 * n + n * k + n * k * (3 + c)),
        lambda a: _lib.check(lib.pob_group_xyz_forward(L(n), I(k), I(c), I(1), P(sets[a]["f"]), I(0), P(xyz), P(xyz), P(idx), P(sets[a]["ox"]), st()), "g"))
    row("group with_xyz bwd", B * (n * k * (3 + c) + n * k + n * c),
        lambda a: _lib.check(lib.pob_group_xyz_backward(L(n), I(k), I(c), I(1), P(sets[a]["gx"]), P(idx), P(sets[a]["gi"]), st()), "g"))
    row("subtraction fwd", B * (2 * n * c + n * k + n * k * c),
        lambda a: _lib.check(lib.pob_subtraction_forward(L(n), I(k), I(c), P(sets[a]["f"]), P(sets[a]["f2"]), P(idx), P(sets[a]["o3"]), st()), "s"),
        lambda a: ref.subtraction_forward_cuda_launcher(I(n), I(k), I(c), P(sets[a]["f"]), P(sets[a]["f2"]), P(idx), P(sets[a]["o3"])))
    row("subtraction bwd", B * (n * k * c + n * k + 2 * n * c),
        lambda a: _lib.check(lib.pob_subtraction_backward(L(n), I(k), I(c), P(idx), P(sets[a]["g3"]), P(sets[a]["gi"]), P(sets[a]["gi2"]), st()), "s"),
        lambda a: ref.subtraction_backward_cuda_launcher(I(n), I(k), I(c), P(idx), P(sets[a]["g3"]), P(sets[a]["gi"]), P(sets[a]["gi2"])))
    fwd_bytes = B * (n * c + n * k * c + n * k * wc + n * k + n * c)
    row("aggregation fwd", fwd_bytes,
        lambda a: _lib.check(lib.pob_aggregation_forward(L(n), I(k), I(c), I(wc), P(sets[a]["f"]), P(sets[a]["pos"]), P(sets[a]["w"]), P(idx), P(sets[a]["o2"]), st()), "a"),
        lambda a: ref.aggregation_forward_cuda_launcher(I(n), I(k), I(c), I(wc), P(sets[a]["f"]), P(sets[a]["pos"]), P(sets[a]["w"]), P(idx), P(sets[a]["o2"])))
    row("aggregation bwd", fwd_bytes + B * (n * c + n * k * c + n * k * wc),
        lambda a: _lib.check(lib.pob_aggregation_backward(L(n), I(k), I(c), I(wc), P(sets[a]["f"]), P(sets[a]["pos"]), P(sets[a]["w"]), P(idx), P(sets[a]["g2"]),
                                                          P(sets[a]["gi"]), P(sets[a]["gp"]), P(sets[a]["gw"]), st()), "a"),
        lambda a: ref.aggregation_backward_cuda_launcher(I(n), I(k), I(c), I(wc), P(sets[a]["f"]), P(sets[a]["pos"]), P(sets[a]["w"]), P(idx), P(sets[a]["g2"]),
                                                         P(sets[a]["gi"]), P(sets[a]["gp"]), P(sets[a]["gw"])))
    i3 = idx[:, :3].contiguous()
    w3 = torch.rand(n, 3, device=dev, generator=g)
    row("interpolation fwd (k=3)", B * (n * c + 2 * n * 3 + n * c),
        lambda a: _lib.check(lib.pob_interpolation_forward(L(n), I(c), I(3), P(sets[a]["f"]), P(i3), P(w3), P(sets[a]["o2"]), st()), "i"),
        lambda a: ref.interpolation_forward_cuda_launcher(I(n), I(c), I(3), P(sets[a]["f"]), P(i3), P(w3), P(sets[a]["o2"])))
    row("interpolation bwd (k=3)", B * (n * c + 2 * n * 3 + n * c),
        lambda a: _lib.check(lib.pob_interpolation_backward(L(n), I(c), I(3), P(sets[a]["g2"]), P(i3), P(w3), P(sets[a]["gi"]), st()), "i"),
        lambda a: ref.interpolation_backward_cuda_launcher(I(n), I(c), I(3), P(sets[a]["g2"]), P(i3), P(w3), P(sets[a]["gi"])))
    # kNN: grid build + query vs the reference's brute-force kernel (one launch each; the reference takes milliseconds)
    from pointcloudpdf_b200.pointops import _common as C
    ki = torch.empty(n, k, dtype=torch.int32, device=dev)
    kd = torch.empty(n, k, dtype=torch.float32, device=dev)

    def knn_mine(a):
        C.clear_caches()
        C.get_grid(xyz, off).query(k, xyz, off, True, False)
    us = time_graph(knn_mine)
    r = {"b200_us": us, "bruteforce_equivalent_TFLOPs": 8 * n * n / us / 1e6}
    if ref is not None:
        ru = time_default_stream(lambda a: ref.knn_query_cuda_launcher(I(n), I(k), P(xyz), P(xyz), P(off), P(off), P(ki), P(kd)))
        r.update({"reference_us": ru, "speedup": ru / us})
    rows[f"knn_query k={k} (build + query)"] = r
    print(f"{tag:8s} knn k={k}: b200 {us:.1f} us" + (f"  reference {r['reference_us']:.1f} us  x{r['speedup']:.1f}" if ref else ""), flush=True)
    C.clear_caches()
    return rows


def main():
    out = {"hbm_peak_GBps": HBM, "timing": __doc__.split("Both sides")[1].split("GB/s =")[0].strip(),
           "cfg1 (n=24000, k=16, C=32, w_c=4)": run_shape("cfg1", 24000, 16, 32, 4),
           "stage1 (n=80000, k=8, C=32, w_c=4)": run_shape("stage1", 80000, 8, 32, 4)}
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    with torch.no_grad():
        main()
