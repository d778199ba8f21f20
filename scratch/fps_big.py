import sys, time, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S
import pointops
dev = torch.device('cuda:0')
for n in (150000, 250000, 500000, 1000000):
    b = S.s3dis_batch([n], seed=2029)
    xyz, off = b['coord'].to(dev), b['offset'].to(dev)
    m = n // 4 if n <= 500000 else 20000
    noff = torch.tensor([m], dtype=torch.int32, device=dev)
    pointops.clear_caches(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = pointops.farthest_point_sampling(xyz, off, noff); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    print(f"n={n} m={m}: {ms:.1f} ms  {ms*1e6/m:.0f} ns/sample  unique={out.unique().numel()}")
