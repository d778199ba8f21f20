"""Launch pob_farthest_point_sampling a few times on one S3DIS-shaped scene (target of ncu captures).
python tools/fps_one.py N [variant] [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import _common as C
from pointcloudpdf_b200.pointops.sampling import VARIANTS
n = int(sys.argv[1]); v = sys.argv[2] if len(sys.argv) > 2 else "merge"; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0"); lib = _lib.load()
b = S.s3dis_batch([n], seed=2026)
xyz, off = b["coord"].to(dev), b["offset"].to(dev)
m = n // 4
noff = torch.tensor([m], dtype=torch.int32, device=dev)
grid = C.NeighbourGrid(xyz, off) if n > 2048 else None
tmp = torch.empty(n, dtype=torch.float32, device=dev)
out = torch.empty(m, dtype=torch.int32, device=dev)
for _ in range(reps):
    rc = lib.pob_farthest_point_sampling(1, n, _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(noff), _lib.ptr(tmp), _lib.ptr(out), 0,
                                         _lib.ptr(grid.workspace if grid else None), n, grid.cell_pts if grid else 0.0,
                                         VARIANTS[v], None, _lib.current_stream(dev))
    assert rc == 0
torch.cuda.synchronize()
print("done", out[:8].tolist())
