"""Time pob_farthest_point_sampling variants on S3DIS-shaped scenes (one GPU): ms, ns per sample, samples per
exchange, ns per round.   python tools/fps_time.py [--sizes 80000,20000,...] [--variants merge,chain]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import _common as C
from pointcloudpdf_b200.pointops.sampling import VARIANTS

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="80000,20000,5000,1250")
ap.add_argument("--variants", default="merge,chain")
ap.add_argument("--clusters", default="0")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--json", default="")
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = _lib.load()
rows = []
for n in [int(v) for v in args.sizes.split(",")]:
    m = n // 4
    b = S.s3dis_batch([n], seed=2026)
    xyz, off = b["coord"].to(dev), b["offset"].to(dev)
    noff = torch.tensor([m], dtype=torch.int32, device=dev)
    grid = C.NeighbourGrid(xyz, off) if n > 2048 else None
    tmp = torch.empty(n, dtype=torch.float32, device=dev)
    ref = None
    for v in args.variants.split(","):
        for cl in [int(c) for c in args.clusters.split(",")]:
            out = torch.empty(m, dtype=torch.int32, device=dev)
            stats = torch.zeros(4, dtype=torch.int64, device=dev)
            best = 1e9
            for r in range(args.reps):
                stats.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.pob_farthest_point_sampling(1, n, _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(noff), _lib.ptr(tmp), _lib.ptr(out), cl,
                                                     _lib.ptr(grid.workspace if grid else None), n, grid.cell_pts if grid else 0.0,
                                                     VARIANTS[v], _lib.ptr(stats), _lib.current_stream(dev))
                e1.record(); torch.cuda.synchronize()
                assert rc == 0, rc
                best = min(best, e0.elapsed_time(e1))
            if ref is None:
                ref = out.clone()
            st = stats.tolist()
            chain = st[1] / max(st[0], 1)
            row = dict(n=n, m=m, variant=v, cluster=cl, ms=best, ns_per_sample=best * 1e6 / m, rounds=st[0], samples_per_round=chain,
                       ns_per_round=best * 1e6 / max(st[0], 1), d2_evals=st[2], same_as_first=bool(torch.equal(out, ref)))
            rows.append(row)
            print(f"n={n:7d} m={m:6d} {v:6s} C={cl:2d}: {best:8.3f} ms {row['ns_per_sample']:7.1f} ns/sample rounds={st[0]:6d} "
                  f"chain={chain:5.2f} {row['ns_per_round']:7.1f} ns/round evals/sample={st[2] / max(st[1], 1):7.0f} same={row['same_as_first']}", flush=True)
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
