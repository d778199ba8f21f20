"""Experiment: does the ORDER in which the FPS kernel receives its points matter?  The kernel reads the scene from the
kNN grid's cell-sorted array (cells in x-fastest scanline order: a warp's 640 consecutive points are a strip across the
room).  Here that array is re-permuted in place (Morton order at several quantisations) before the launch; results
must not change (ties go to the lower ORIGINAL index, carried in .w), only the time.   python tools/fps_order_experiment.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloudpdf_b200 import synthetic as S, _lib
from pointcloudpdf_b200.pointops import _common as C
from pointcloudpdf_b200.pointops.sampling import VARIANTS

dev = torch.device("cuda:0")
lib = _lib.load()


def align(v, a=256):
    return (v + a - 1) // a * a


def layout(n, b, cell_pts):
    cap = int(n / cell_pts) + b + 8
    o = 0
    off = {}
    for name, size in (("scene", 64 * b), ("bbox", 24 * b), ("cnt", 4 * (cap + 1)), ("start", 4 * (cap + 1)),
                       ("tiles", 4 * (cap // 2048 + 2)), ("pcell", 4 * n), ("sorted", 16 * n)):
        off[name] = o
        o = align(o + size)
    return off


def part1by2(v):
    v = v & 0x3ff
    v = (v | (v << 16)) & 0x30000ff
    v = (v | (v << 8)) & 0x300f00f
    v = (v | (v << 4)) & 0x30c30c3
    v = (v | (v << 2)) & 0x9249249
    return v


def hilbert(q, bits=10):
    """Skilling's axes -> transposed Hilbert index, vectorised; returns the interleaved 3*bits key."""
    X = [q[:, 0].clone(), q[:, 1].clone(), q[:, 2].clone()]
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:
        P = Q - 1
        for i in range(3):
            hit = (X[i] & Q) != 0
            t = (X[0] ^ X[i]) & P
            X0_inv = X[0] ^ P
            X0_ex = X[0] ^ t
            Xi_ex = X[i] ^ t
            if i == 0:
                X[0] = torch.where(hit, X0_inv, X[0])   # exchange with itself is the identity
            else:
                X[0] = torch.where(hit, X0_inv, X0_ex)
                X[i] = torch.where(hit, X[i], Xi_ex)
        Q >>= 1
    for i in range(1, 3):
        X[i] = X[i] ^ X[i - 1]
    t = torch.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = torch.where((X[2] & Q) != 0, t ^ (Q - 1), t)
        Q >>= 1
    for i in range(3):
        X[i] = X[i] ^ t
    return (part1by2(X[0]) << 2) | (part1by2(X[1]) << 1) | part1by2(X[2])


def morton(q):
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


for n in (80000, 20000, 150000):
    m = n // 4
    b = S.s3dis_batch([n], seed=2026)
    xyz, off = b["coord"].to(dev), b["offset"].to(dev)
    noff = torch.tensor([m], dtype=torch.int32, device=dev)
    tmp = torch.empty(max(n, 1 << 20), dtype=torch.float32, device=dev)
    ref = None
    lo = xyz.min(0)[0]
    ext = float((xyz.max(0)[0] - lo).max())
    for order in ("cells (as built)", "morton 10 bit", "morton 9 bit", "hilbert 10 bit", "hilbert 9 bit", "hilbert 8 bit", "library (morton 10 bit in the launcher)"):
        grid = C.NeighbourGrid(xyz, off)
        L = layout(n, 1, grid.cell_pts)
        ws = grid.workspace
        sc = ws[L["scene"]:L["scene"] + 64].view(torch.float32)
        h = float(sc[8])
        srt = ws[L["sorted"]:L["sorted"] + 16 * n].view(torch.float32).view(n, 4)
        variant = "merge" if order.startswith("library") else "merge_cells"
        if order != "cells (as built)" and variant == "merge_cells":
            pts = srt[:, :3]
            bits = int(order.split()[1])
            q = ((pts - lo) / ext * float(1 << bits)).long().clamp(0, (1 << bits) - 1)
            key = hilbert(q, bits) if order.startswith("hilbert") else morton(q)
            perm = torch.sort(key, stable=True)[1]
            srt.copy_(srt[perm].clone())
        out = torch.empty(m, dtype=torch.int32, device=dev)
        stats = torch.zeros(4, dtype=torch.int64, device=dev)
        best = 1e9
        for r in range(4):
            stats.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.pob_farthest_point_sampling(1, n, _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(noff), _lib.ptr(tmp), _lib.ptr(out), 0,
                                                 _lib.ptr(ws), n, grid.cell_pts, VARIANTS[variant], _lib.ptr(stats), _lib.current_stream(dev))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0, rc
            best = min(best, e0.elapsed_time(e1))
        if ref is None:
            ref = out.clone()
        st = stats.tolist()
        print(f"n={n:7d} {order:40s} {best:8.3f} ms rounds={st[0]:6d} samples/round={st[1] / max(st[0], 1):5.2f} ns/round={best * 1e6 / max(st[0], 1):7.1f} "
              f"evals/sample={st[2] / max(st[1], 1):7.0f} same={bool(torch.equal(out, ref))}", flush=True)
        del grid
        C.clear_caches()
