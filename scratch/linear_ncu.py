"""Launch pob_linear_forward once per PTv1-Seg50 linear shape (L2 evicted before each) for an ncu capture:
ncu --set full --import-source on --clock-control none -k regex:linear_tile -o gpurun_out/<tag>_linear python scratch/linear_ncu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200.pointops import fused as FZ
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for (m, k, n, ep) in [(80000, 32, 32, "br"), (80000, 32, 96, "p"), (80000, 32, 32, "brr"), (20000, 64, 64, "br"),
                      (20000, 64, 192, "p"), (5000, 128, 128, "br"), (5000, 128, 384, "p"), (1250, 256, 256, "brr"),
                      (1250, 256, 768, "p"), (312, 512, 512, "br"), (312, 512, 1536, "p")]:
    x = torch.randn(m, k, device=dev, generator=g)
    wt = torch.randn(k, n, device=dev, generator=g)
    b = torch.randn(n, device=dev, generator=g) if "b" in ep else None
    r = torch.randn(m, n, device=dev, generator=g) if ep == "brr" else None
    FZ.linear(x, wt, b, r, ep != "p")      # warm (module load)
    flush.zero_()
    FZ.linear(x, wt, b, r, ep != "p")
torch.cuda.synchronize()
