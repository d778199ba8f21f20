"""pt_layer_forward at the five PTv1-Seg50 stage shapes (80k-point room), per split variant.
L2 evicted by a 256 MiB memset before every launch; CUDA events around the C-ABI call."""
import sys, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import fused as FZ
import pointops
dev = torch.device('cuda:0')
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for (n, ns, c) in ((80000, 8, 32), (20000, 16, 64), (5000, 16, 128), (1250, 16, 256), (312, 16, 512)):
    b = S.s3dis_batch([n], seed=2025)
    xyz, off = b['coord'].to(dev), b['offset'].to(dev)
    idx, _ = pointops.knn_query(ns, xyz, off)
    wc = c // 8
    qkv = torch.randn(n, 3 * c, device=dev, generator=g)
    params = torch.randn(int(lib.pob_pt_layer_param_floats(c, wc)), device=dev, generator=g) * 0.1
    ref = None
    for split in (1, 0):
        lib.pob_pt_layer_set_split(split)
        ts = []
        for r in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = FZ.pt_layer_forward(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], xyz, idx, params, True)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        if ref is None: ref = out
        err = (out - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
        alg = 4 * (4 * n * c + 3 * n + n * ns) + 4 * params.numel()
        us = sorted(ts)[len(ts) // 2]
        print(f"n={n:6d} ns={ns:2d} C={c:3d} split={split:2d}: {us:8.1f} us (min {min(ts):7.1f})  alg {alg/1e6:6.2f} MB -> {alg/us/1e3:7.1f} GB/s  relerr vs split1 {err:.1e}")
lib.pob_pt_layer_set_split(0)
