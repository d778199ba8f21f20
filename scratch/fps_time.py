import sys, torch, time
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import _common as C
dev = torch.device('cuda:0')
lib = _lib.load()
def run(n, m, cluster, use_grid, reps=3):
    b = S.s3dis_batch([n], seed=2026)
    xyz = b['coord'].to(dev); off = b['offset'].to(dev); noff = torch.tensor([m], dtype=torch.int32, device=dev)
    out = torch.empty(m, dtype=torch.int32, device=dev)
    grid = C.NeighbourGrid(xyz, off) if use_grid else None
    best = 1e9
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    lib.pob_fps_set_stats(_lib.ptr(stats))
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.pob_farthest_point_sampling(1, n, _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(noff), None, _lib.ptr(out), cluster,
              _lib.ptr(grid.workspace if grid else None), n, grid.cell_pts if grid else 0.0, _lib.current_stream(dev))
        e1.record(); torch.cuda.synchronize()
        assert rc == 0, rc
        best = min(best, e0.elapsed_time(e1))
    lib.pob_fps_set_stats(None)
    st = stats.tolist()
    return best, out, (st[1] / max(st[0], 1))
for n, m in ((80000, 20000), (20000, 5000), (5000, 1250), (1250, 312)):
    ref = None
    for use_grid in (1,):
        for cl in (1, 4, 8, 16):
            if n / ((cl % 100) * (512 if cl >= 100 else 256)) > (16 if cl >= 100 else 32): continue
            ms, out, chain = run(n, m, cl, use_grid)
            if ref is None: ref = out.clone()
            same = torch.equal(out, ref)
            print(f"n={n:6d} m={m:6d} grid={use_grid} C={cl:2d}: {ms:8.3f} ms  {ms*1e6/m:7.1f} ns/iter  same={same} chain={chain:.2f} ns/round={ms*1e6/m*chain:7.1f}")
