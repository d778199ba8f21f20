"""Small launches of the kernels added late in round 2, meant to run under compute-sanitizer (memcheck / racecheck):
the 3xTF32 linear on ragged shapes, the Hilbert ordering pass + merged-list FPS on a ragged batch, the tiny-scene FPS kernel.
    compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S
import pointcloudpdf_b200.pointops as pointops
from pointcloudpdf_b200.pointops import fused as FZ
from pointcloudpdf_b200.pointops.sampling import fps_launch

dev = torch.device("cuda:0")
with torch.no_grad():
    for (m, k, n, cfg) in ((10007, 36, 96, 18), (10007, 32, 160, 19), (12001, 64, 32, 17), (333, 128, 64, 20), (20000, 64, 192, 0)):
        x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) / k ** 0.5
        wt = w.t().contiguous(); b = torch.randn(n, device=dev)
        y = FZ.linear(x, wt, b, None, True, config=cfg)
        ref = torch.relu(x.double() @ w.double().t() + b.double()).float()
        print("linear", (m, k, n, cfg), float((y - ref).abs().max()))
    for sizes in ([5000], [700, 9000, 120, 3000], [1250], [2048, 17, 600]):
        b = S.s3dis_batch(sizes, seed=7)
        xyz, off = b["coord"].to(dev), b["offset"].to(dev)
        noff_host, acc = [], 0
        for s in sizes:
            acc += max(1, s // 4); noff_host.append(acc)
        noff = torch.tensor(noff_host, dtype=torch.int32, device=dev)
        a = fps_launch(xyz, off, noff, b["offset"].tolist(), noff_host, variant="auto")
        c = fps_launch(xyz, off, noff, b["offset"].tolist(), noff_host, variant="chain")
        d = fps_launch(xyz, off, noff, b["offset"].tolist(), noff_host, variant="merge")
        print("fps", sizes, bool(torch.equal(a, c)), bool(torch.equal(d, c)))
torch.cuda.synchronize()
print("done")
