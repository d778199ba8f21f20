"""Per-shape timing of pob_linear_forward (every tile configuration) against the cuBLAS route it replaces,
on the linear shapes of PTv1-Seg50 over an 80 000-point room.  CUDA graph of 20 back-to-back launches over 4
rotating operand sets, CUDA events around a replay, median of 7.  python tools/linear_time.py"""
import os, sys, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import _lib
from pointcloudpdf_b200.pointops import fused as FZ

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
lib = _lib.load()
SHAPES = [(80000, 6, 32, "br"), (80000, 32, 32, "br"), (80000, 32, 96, "p"), (80000, 32, 32, "brr"), (80000, 32, 64, "p"),
          (80000, 32, 13, "b"), (20000, 64, 64, "br"), (20000, 64, 192, "p"), (20000, 64, 64, "brr"), (20000, 64, 128, "p"),
          (5000, 128, 128, "br"), (5000, 128, 384, "p"), (5000, 128, 128, "brr"), (5000, 128, 256, "p"),
          (1250, 256, 256, "br"), (1250, 256, 768, "p"), (1250, 256, 256, "brr"), (1250, 256, 512, "p"),
          (312, 512, 512, "br"), (312, 512, 1536, "p"), (312, 512, 512, "brr"), (312, 512, 256, "br")]
REPS = 20


def timed(fn, sets):
    for s in sets[:2]:
        fn(s)
    torch.cuda.synchronize()
    st = torch.cuda.Stream(device=dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
        for r in range(REPS):
            fn(sets[r % len(sets)])
    ts = []
    with torch.cuda.stream(st):
        g.replay()
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) / REPS * 1e3   # us


rows = []
gen = torch.Generator(device=dev).manual_seed(0)
for (m, k, n, ep) in SHAPES:
    sets = []
    for _ in range(4):
        w = torch.randn(n, k, device=dev, generator=gen) / k ** 0.5
        sets.append(dict(x=torch.randn(m, k, device=dev, generator=gen), w=w, wt=w.t().contiguous(),
                         b=torch.randn(n, device=dev, generator=gen), r=torch.randn(m, n, device=dev, generator=gen)))
    bias = lambda s: s["b"] if "b" in ep else None
    res = lambda s: s["r"] if ep == "brr" else None
    relu = ep in ("br", "brr")

    def cublas(s):
        wt = s["w"].t()
        if ep == "p":
            return torch.mm(s["x"], wt)
        if ep == "b":
            return torch.addmm(s["b"], s["x"], wt)
        if ep == "br":
            return torch._addmm_activation(s["b"], s["x"], wt)
        z = torch.addmm(s["r"], s["x"], wt)
        return FZ.affine_act(z, None, s["b"], None, relu=True, inplace=True)

    CFG = [0]

    def pob(s):
        return FZ.linear(s["x"], s["wt"], bias(s), res(s), relu, config=CFG[0])

    with torch.no_grad():
        rec = dict(shape=[m, k, n], epilogue=ep, cublas_us=timed(cublas, sets))
        aligned = k % 4 == 0 and n % 4 == 0
        for cfg in ([0, 21] + ([17, 18, 19, 20] if aligned else [])):
            CFG[0] = cfg
            rec[f"pob_cfg{cfg}_us"] = timed(pob, sets)
        CFG[0] = 0
    best = min([c for c in (17, 18, 19, 20, 21) if f"pob_cfg{c}_us" in rec], key=lambda c: rec[f"pob_cfg{c}_us"])
    rec["best_cfg"] = best
    rec["gflop"] = 2 * m * k * n / 1e9
    rec["auto_TFLOPs"] = rec["gflop"] / rec["pob_cfg0_us"] * 1e3
    rows.append(rec)
    print(f"{m:6d} x {k:3d} x {n:4d} {ep:>3}  cublas {rec['cublas_us']:6.1f}  auto(mma) {rec['pob_cfg0_us']:6.1f}  ffma {rec['pob_cfg21_us']:6.1f}  " +
          " ".join(f"c{c}={rec[f'pob_cfg{c}_us']:.1f}" for c in (17, 18, 19, 20) if f"pob_cfg{c}_us" in rec) + f"  best c{best}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/r02_linear_time.json", "w"), indent=1)
print("sum over shapes: cublas %.0f us, mma auto %.0f us, ffma %.0f us" % (sum(r["cublas_us"] for r in rows), sum(r["pob_cfg0_us"] for r in rows), sum(r["pob_cfg21_us"] for r in rows)))
