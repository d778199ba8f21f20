"""The kNN launches of one 80k-point room's first level (self k=8, cross k=16 from the FPS subset, 3-NN back),
for `ncu --set full -k regex:knn_grid_kernel`."""
import sys, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S
import pointops
dev = torch.device('cuda:0')
b = S.s3dis_batch([80000], seed=2026)
xyz, off = b['coord'].to(dev), b['offset'].to(dev)
noff = torch.tensor([20000], dtype=torch.int32, device=dev)
sel = pointops.farthest_point_sampling(xyz, off, noff)
sub = xyz[sel.long()].contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for r in range(2):
    pointops.clear_caches()
    flush.zero_(); pointops.knn_query(8, xyz, off)
    flush.zero_(); pointops.knn_query(16, xyz, off, sub, noff)
    flush.zero_(); pointops.knn_query(3, sub, noff, xyz, off)
torch.cuda.synchronize()
