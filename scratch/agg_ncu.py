import sys, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S
import pointops
dev = torch.device('cuda:0')
N, ns, Cc = 80000, 8, 32
b = S.s3dis_batch([N], seed=2025)
xyz = b['coord'].to(dev); off = b['offset'].to(dev)
idx, _ = pointops.knn_query(ns, xyz, off)
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda *shape: torch.randn(*shape, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
f = mk(N, Cc); pos = mk(N, ns, Cc); w = mk(N, ns, Cc // 8)
for rep in range(3):
    flush.zero_(); o4 = pointops.aggregation(f, pos, w, idx)
torch.cuda.synchronize()
