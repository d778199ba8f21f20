"""FPS time per layout (wide: C CTAs x 256 threads; tall: C/2 CTAs x 512 threads, two groups each) on the
four stage sizes of an 80 000-point room.  python scratch/fps_layout_time.py"""
import sys, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import _common as C
dev = torch.device('cuda:0')
lib = _lib.load()
for n, m in ((80000, 20000), (20000, 5000), (5000, 1250)):
    b = S.s3dis_batch([n], seed=2026)
    xyz = b['coord'].to(dev); off = b['offset'].to(dev); noff = torch.tensor([m], dtype=torch.int32, device=dev)
    grid = C.NeighbourGrid(xyz, off)
    ref = None
    forms = [("wide/reg", 0, 0), ("wide/smem", 1, 0), ("tall/smem", 1, 1)]
    if "--fine" in sys.argv:          # experimental 32-group layout
        forms.append(("fine/reg", 0, 2))
    for name, pts, lay in forms:
        lib.pob_fps_set_points(pts); lib.pob_fps_set_layout(lay)
        out = torch.empty(m, dtype=torch.int32, device=dev)
        stats = torch.zeros(2, dtype=torch.int64, device=dev)
        lib.pob_fps_set_stats(_lib.ptr(stats))
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.pob_farthest_point_sampling(1, n, _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(noff), None, _lib.ptr(out), 0,
                                                 _lib.ptr(grid.workspace), n, grid.cell_pts, _lib.current_stream(dev))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, rc
            best = min(best, e0.elapsed_time(e1))
        lib.pob_fps_set_stats(None)
        st = stats.tolist()
        if ref is None: ref = out.clone()
        print(f"n={n:6d} m={m:6d} {name:10s} {best:8.3f} ms  {best*1e6/m:7.1f} ns/sample  chain={st[1]/max(st[0],1):.2f}  same={torch.equal(out, ref)}", flush=True)
lib.pob_fps_set_points(-1); lib.pob_fps_set_layout(-1)
