"""Launch the standalone gather operators once each at PTv1 stage-1 size (target of ncu captures).
python tools/ops_one.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S
import pointops
dev = torch.device("cuda:0")
N, ns, Cc = 80000, 8, 32
b = S.s3dis_batch([N], seed=2025)
xyz, off = b["coord"].to(dev), b["offset"].to(dev)
idx, _ = pointops.knn_query(ns, xyz, off)
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda *shape: torch.randn(*shape, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
f, f2 = mk(N, Cc), mk(N, Cc)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    flush.zero_(); pointops.grouping(idx, f, xyz, xyz, True)
    flush.zero_(); pointops.grouping2(f, idx)
    flush.zero_(); o = pointops.subtraction(f.requires_grad_(True), f2.requires_grad_(True), idx)
    flush.zero_(); o.backward(mk(N, ns, Cc))
torch.cuda.synchronize()
print("done")
