"""How much do the gather-type kernels care about the ORDER of the points in memory?  Same 80 000-point room, once in
its shuffled order (ShufflePoint) and once sorted along a Hilbert curve; CUDA-graph replay timing of the kNN search
(build + query), the cross queries of TransitionDown / interpolation, and the neighbour gathers.
    python tools/order_effect.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S
import pointcloudpdf_b200.pointops as pointops
from pointcloudpdf_b200.pointops import _common as C

dev = torch.device("cuda:0")
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fps_order_experiment.py")).read()
ns = {}
exec("import torch\n" + src[src.index("def part1by2"):src.index("for n in (80000")], ns)


def graph_time(fn, reps=12):
    fn(); torch.cuda.synchronize()
    st = torch.cuda.Stream(device=dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
        for _ in range(reps):
            fn()
    ts = []
    with torch.cuda.stream(st):
        g.replay()
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) / reps * 1e3


def hilbert_perm(c):
    lo = c.min(0)[0]
    ext = float((c.max(0)[0] - lo).max())
    q = ((c - lo) / ext * 1024.0).long().clamp(0, 1023)
    return torch.sort(ns["hilbert"](q, 10), stable=True)[1]


with torch.no_grad():
    n = 80000
    b = S.s3dis_batch([n], seed=2026)
    for tag in ("shuffled", "hilbert"):
        xyz = b["coord"].to(dev)
        if tag == "hilbert":
            xyz = xyz[hilbert_perm(xyz)].contiguous()
        off = torch.tensor([n], dtype=torch.int32, device=dev)
        m = n // 4
        noff = torch.tensor([m], dtype=torch.int32, device=dev)
        fidx = pointops.farthest_point_sampling(xyz, off, noff)
        new_xyz = xyz[fidx.long()].contiguous()          # FPS order (scattered), as in the model
        new_sorted = new_xyz[hilbert_perm(new_xyz)].contiguous()
        feat = torch.randn(n, 32, device=dev)
        res = {}

        def self_knn(k):
            C.clear_caches()
            return C.get_grid(xyz, off).query(k, xyz, off, True, False)
        res["self kNN k=8 (build+query)"] = graph_time(lambda: self_knn(8))
        res["self kNN k=16 (build+query)"] = graph_time(lambda: self_knn(16))
        grid = C.get_grid(xyz, off)
        res["cross kNN k=16: 20k FPS-ordered queries vs 80k"] = graph_time(lambda: grid.query(16, new_xyz, noff, True, False))
        res["cross kNN k=16: 20k curve-ordered queries vs 80k"] = graph_time(lambda: grid.query(16, new_sorted, noff, True, False))
        cgrid = C.get_grid(new_xyz, noff)
        res["interp kNN k=3: 80k queries vs 20k"] = graph_time(lambda: cgrid.query(3, xyz, off, True, False))
        idx = grid.query(8, xyz, off, True, False)[0]
        w = torch.randn(n, 8, 4, device=dev); pos = torch.randn(n, 8, 32, device=dev)
        res["grouping2 (n,8,32)"] = graph_time(lambda: pointops.grouping2(feat, idx))
        res["aggregation (n,8,32)"] = graph_time(lambda: pointops.aggregation(feat, pos, w, idx))
        for k, v in res.items():
            print(f"{tag:9s} {k:52s} {v:8.1f} us", flush=True)
        C.clear_caches()
