import sys, torch, ctypes
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import _common as C
dev = torch.device('cuda:0')
lib = ctypes.CDLL(_lib.LIB_PATH)
L = _lib.load()
for n, m, cl in ((80000, 20000, 16), (20000, 5000, 16), (5000, 1250, 1)):
    b = S.s3dis_batch([n], seed=2026)
    xyz = b['coord'].to(dev); off = b['offset'].to(dev); noff = torch.tensor([m], dtype=torch.int32, device=dev)
    out = torch.empty(m, dtype=torch.int32, device=dev)
    grid = C.NeighbourGrid(xyz, off)
    buf = (ctypes.c_ulonglong * 2)()
    torch.cuda.synchronize(); lib.pob_debug_touch(buf)
    rc = L.pob_farthest_point_sampling(1, n, _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(noff), None, _lib.ptr(out), cl,
          _lib.ptr(grid.workspace), n, grid.cell_pts, _lib.current_stream(dev))
    torch.cuda.synchronize(); lib.pob_debug_touch(buf)
    print(n, m, cl, "touched warps per iteration:", buf[0] / max(buf[1], 1), "of", cl * 8)
