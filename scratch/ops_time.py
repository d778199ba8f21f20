"""Per-kernel timing at BASELINE cfg1 (N=24000,k=16,C=32) and PTv1 stage-1 shapes; rotating buffer
sets larger than L2 so every launch streams from HBM."""
import sys, json, torch
sys.path.insert(0, '.')
from pointcloudpdf_b200 import synthetic as S, _lib
import pointops
from pointcloudpdf_b200.pointops import _common as C
from pointcloudpdf_b200.scoring import fused_scores
dev = torch.device('cuda:0')
PEAK = 6555.5
def timeit(make_args, fn, alg_bytes, name, reps=20):
    nsets = max(2, int(400e6 // max(alg_bytes, 1)) + 1)
    nsets = min(nsets, 24)
    sets = [make_args() for _ in range(nsets)]
    for a in sets[:2]: fn(*a)
    torch.cuda.synchronize()
    _lib.PROFILE = prof = _lib.OpProfile()
    for r in range(reps):
        fn(*sets[r % nsets])
    _lib.PROFILE = None
    summ = prof.summary()
    for k, d in summ.items():
        ms = d['ms'] / d['calls']
        gbs = d['alg_bytes'] / d['calls'] / ms / 1e6
        print(f"{name:34s} {k:28s} {ms*1e3:9.1f} us  {d['alg_bytes']/d['calls']/1e6:8.2f} MB  {gbs:8.1f} GB/s  {100*gbs/PEAK:5.1f}% of measured peak")
for (N, ns, Cc, tag) in ((24000, 16, 32, 'cfg1'), (80000, 8, 32, 'stage1'), (20000, 16, 64, 'stage2')):
    b = S.s3dis_batch([N], seed=2025)
    xyz = b['coord'].to(dev); off = b['offset'].to(dev)
    idx, _ = pointops.knn_query(ns, xyz, off)
    w_c = Cc // 8
    g = torch.Generator(device=dev).manual_seed(0)
    mk = lambda *shape: torch.randn(*shape, device=dev, generator=g)
    timeit(lambda: (idx, mk(N, Cc), xyz, xyz, True), lambda i, f, x, nx, w: pointops.grouping(i, f, x, nx, w), 0, f'{tag} grouping with_xyz')
    timeit(lambda: (idx, mk(N, Cc), xyz, xyz, False), lambda i, f, x, nx, w: pointops.grouping(i, f, x, nx, w), 0, f'{tag} grouping feat only')
    timeit(lambda: (mk(N, Cc), idx), lambda f, i: pointops.grouping2(f, i), 0, f'{tag} grouping2 fwd')
    timeit(lambda: (mk(N, Cc), mk(N, Cc), idx), lambda a, b2, i: pointops.subtraction(a, b2, i), 0, f'{tag} subtraction fwd')
    timeit(lambda: (mk(N, Cc), mk(N, ns, Cc), mk(N, ns, w_c), idx), lambda a, p, w, i: pointops.aggregation(a, p, w, i), 0, f'{tag} aggregation fwd')
    def agg_bwd(a, p, w, i, go):
        a.requires_grad_(True); p.requires_grad_(True); w.requires_grad_(True)
        a.grad = p.grad = w.grad = None
        pointops.aggregation(a, p, w, i).backward(go)
    timeit(lambda: (mk(N, Cc), mk(N, ns, Cc), mk(N, ns, w_c), idx, mk(N, Cc)), agg_bwd, 0, f'{tag} aggregation fwd+bwd')
    def g2_bwd(f, i, go):
        f.requires_grad_(True); f.grad = None
        pointops.grouping2(f, i).backward(go)
    timeit(lambda: (mk(N, Cc), idx, mk(N, ns, Cc)), g2_bwd, 0, f'{tag} grouping2 fwd+bwd')
    def sub_bwd(a, b2, i, go):
        a.requires_grad_(True); b2.requires_grad_(True); a.grad = b2.grad = None
        pointops.subtraction(a, b2, i).backward(go)
    timeit(lambda: (mk(N, Cc), mk(N, Cc), idx, mk(N, ns, Cc)), sub_bwd, 0, f'{tag} subtraction fwd+bwd')
    def gx_bwd(i, f, x, go):
        f.requires_grad_(True); f.grad = None
        pointops.grouping(i, f, x, x, True).backward(go)
    timeit(lambda: (idx, mk(N, Cc), xyz, mk(N, ns, Cc + 3)), gx_bwd, 0, f'{tag} grouping xyz fwd+bwd')
    # knn: fresh grid each time
    def knn(x, o):
        pointops.clear_caches(); pointops.knn_query(ns, x, o)
    timeit(lambda: (xyz.clone(), off), knn, 0, f'{tag} knn build+query k={ns}')
b = S.s3dis_batch([80000], seed=2025)
xyz = b['coord'].to(dev); off = b['offset'].to(dev)
noff = torch.tensor([20000], dtype=torch.int32, device=dev)
sel = pointops.farthest_point_sampling(xyz, off, noff)
coarse = xyz[sel.long()].contiguous()
g = torch.Generator(device=dev).manual_seed(0)
def interp(f):
    pointops.clear_caches(); pointops.interpolation(coarse, xyz, f, noff, off)
timeit(lambda: (torch.randn(20000, 32, device=dev, generator=g),), interp, 0, 'dec1 interpolation 20k->80k C=32')
timeit(lambda: (torch.randn(80000, 13, device=dev, generator=g),), lambda l: fused_scores(l, want=('msp_score',)), 0, 'msp score 80000x13')
timeit(lambda: (torch.randn(150000, 20, device=dev, generator=g), torch.randn(150000, device=dev, generator=g)), lambda l, c: fused_scores(l, c, want=('pdf_score',)), 0, 'pdf score 150000x20')
