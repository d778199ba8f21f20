"""One launch of pob_pt_layer_forward per PTv1 stage shape (after a warm-up launch), L2 evicted first:
the target of `ncu --set full -k regex:pt_layer` (see profiles/)."""
import sys, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import fused as FZ
import pointops
dev = torch.device('cuda:0')
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
stages = ((80000, 8, 32), (20000, 16, 64), (5000, 16, 128), (1250, 16, 256), (312, 16, 512))
for (n, ns, c) in stages:
    b = S.s3dis_batch([n], seed=2025)
    xyz, off = b['coord'].to(dev), b['offset'].to(dev)
    idx, _ = pointops.knn_query(ns, xyz, off)
    wc = c // 8
    qkv = torch.randn(n, 3 * c, device=dev, generator=g)
    params = torch.randn(int(lib.pob_pt_layer_param_floats(c, wc)), device=dev, generator=g) * 0.1
    for r in range(2):
        flush.zero_()
        out = FZ.pt_layer_forward(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], xyz, idx, params, True)
    torch.cuda.synchronize()
