"""cProfile of the host side of OpenSegPTv1.infer_stream (where do the ~4.5 ms of python per room go)."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S
from pointcloudpdf_b200.ptv1 import OpenSegPTv1
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(2024)
net = OpenSegPTv1(in_channels=6, num_classes=13, method="msp").to(dev).eval()
rooms = [S.s3dis_batch([80000], seed=2026 + i) for i in range(4)]
host = [(r["coord"].pin_memory(), r["feat"].pin_memory(), r["offset"].pin_memory()) for r in rooms]
seq = [host[i % 4] for i in range(30)]
for _ in net.infer_stream(seq[:8], depth=3, device=dev):
    pass
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in net.infer_stream(seq, depth=3, device=dev):
    pass
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(40)
