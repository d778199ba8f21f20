"""Launch pob_linear_forward on one PTv1 shape a few times (target of ncu captures).  python tools/linear_one.py M K N [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200.pointops import fused as FZ
m, k, n = (int(v) for v in sys.argv[1:4]); cfg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
x = torch.randn(m, k, device=dev); wt = (torch.randn(n, k, device=dev) / k ** 0.5).t().contiguous(); b = torch.randn(n, device=dev)
with torch.no_grad():
    for _ in range(3):
        y = FZ.linear(x, wt, b, None, True, config=cfg)
torch.cuda.synchronize()
print("done", float(y.abs().mean()))
