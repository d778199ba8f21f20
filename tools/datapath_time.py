"""f-4 timing: device GridSample / SphereCrop / scatter_mean against the reference's numpy formulation on the box's host
(stable-sort restatement of GridSample; the same arithmetic the dataloader workers run).   python tools/datapath_time.py [out.json]"""
import json, os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloudpdf_b200.datapath import grid_sample, sphere_crop, scatter_mean


class DO:
    """The reference's numpy formulation, spelled out here (pointcept/datasets/transform.py:813-857, 911-925; torch_scatter's
    scatter_mean) so that this timing tool does not reach into oracle/ (test infrastructure)."""

    @staticmethod
    def grid_sample_stable(coord, grid_size):
        g = np.floor(coord / np.array(grid_size)).astype(int)
        g -= g.min(0)
        a = g.astype(np.uint64)
        key = np.uint64(14695981039346656037) * np.ones(a.shape[0], dtype=np.uint64)
        for j in range(3):
            key *= np.uint64(1099511628211)
            key = np.bitwise_xor(key, a[:, j])
        idx_sort = np.argsort(key, kind="stable")
        _, inverse, count = np.unique(key[idx_sort], return_inverse=True, return_counts=True)
        inv = np.zeros_like(inverse)
        inv[idx_sort] = inverse
        return dict(inverse=inv, count=count)

    @staticmethod
    def scatter_mean(src, index, dim_size):
        out = np.zeros(dim_size)
        cnt = np.zeros(dim_size)
        np.add.at(out, index, src.astype(np.float64))
        np.add.at(cnt, index, 1)
        return out / np.maximum(cnt, 1)

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
out = []
for n in (250_000, 1_000_000, 4_000_000):
    coord = (rng.random((n, 3)) * np.array([8.0, 6.0, 3.0])).astype(np.float32)
    t0 = time.perf_counter(); ref = DO.grid_sample_stable(coord, 0.04); cpu = time.perf_counter() - t0
    c = torch.from_numpy(coord).to(dev)
    grid_sample(c, 0.04); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = grid_sample(c, 0.04); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    same = bool(torch.equal(r["inverse"].cpu(), torch.from_numpy(ref["inverse"])))
    out.append(dict(op="GridSample(0.04, train)", n=n, voxels=int(r["count"].numel()), cpu_numpy_ms=cpu * 1e3, b200_ms=statistics.median(ts),
                    speedup=cpu * 1e3 / statistics.median(ts), inverse_equal=same))
    print(out[-1], flush=True)
    t0 = time.perf_counter(); d2 = np.sum(np.square(coord - coord[n // 2]), 1); refc = np.argsort(d2, kind="stable")[:80000]; cpu = time.perf_counter() - t0
    sphere_crop(c, 80000, "center"); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx = sphere_crop(c, 80000, "center"); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    out.append(dict(op="SphereCrop(80000, center)", n=n, cpu_numpy_ms=cpu * 1e3, b200_ms=statistics.median(ts), speedup=cpu * 1e3 / statistics.median(ts),
                    equal=bool(torch.equal(idx.cpu(), torch.from_numpy(refc)))))
    print(out[-1], flush=True)
rows, dim = 4_000_000, 1_000_000
src = rng.random(rows).astype(np.float32); index = rng.integers(0, dim, rows)
t0 = time.perf_counter(); ref = DO.scatter_mean(src, index, dim); cpu = time.perf_counter() - t0
s, i = torch.from_numpy(src).to(dev), torch.from_numpy(index).to(dev)
scatter_mean(s, i, dim); torch.cuda.synchronize(); ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o = scatter_mean(s, i, dim); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
out.append(dict(op="scatter_mean (fragment scores)", rows=rows, dim_size=dim, cpu_numpy_ms=cpu * 1e3, b200_ms=statistics.median(ts),
                speedup=cpu * 1e3 / statistics.median(ts), max_abs_err=float(np.abs(o.cpu().numpy() - ref).max())))
print(out[-1], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
