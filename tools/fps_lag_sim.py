"""CPU simulation of the LAGGED merged-list FPS protocol (design study): lists are published first (after popping what the
merge took and truncating at the first entry a foreign sample lowered), the accepted samples are applied and the lists
refilled while the merge of the same round is running -- so what a warp publishes lags one round behind its state.
Checks equality with plain FPS and reports samples per exchange.   python tools/fps_lag_sim.py n W D KC LMAX"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloudpdf_b200 import synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 112
D = int(sys.argv[3]) if len(sys.argv) > 3 else 3
KC = int(sys.argv[4]) if len(sys.argv) > 4 else 4
LMAX = int(sys.argv[5]) if len(sys.argv) > 5 else 32
WPC = int(sys.argv[6]) if len(sys.argv) > 6 else 7     # point warps per CTA
m = n // 4
xyz = S.s3dis_batch([n], seed=2026)["coord"].numpy().astype(np.float64)
Lb = xyz.max(0) - xyz.min(0)
h = max((Lb.prod() / (n / 2)) ** (1 / 3), 1e-3)
cell = np.floor((xyz - xyz.min(0)) / h).astype(np.int64)
dd = cell.max(0) + 1
key = (cell[:, 2] * dd[1] + cell[:, 1]) * dd[0] + cell[:, 0]
order = np.argsort(key, kind="stable")
pts = xyz[order]; gid = order
d2 = lambda a, b: ((a - b) ** 2).sum(-1)
tmp0 = np.full(n, 1e10); ref = [0]
for _ in range(m - 1):
    tmp0 = np.minimum(tmp0, d2(xyz, xyz[ref[-1]])); ref.append(int(np.argmax(tmp0)))
per = -(-n // W)
bounds = [(w * per, min(n, (w + 1) * per)) for w in range(W)]
lo = np.array([pts[a:b].min(0) if b > a else np.zeros(3) for a, b in bounds])
hi = np.array([pts[a:b].max(0) if b > a else np.zeros(3) for a, b in bounds])
tmp = np.full(n, 1e10)
spec = [None] * W; lists = [[] for _ in range(W)]; term = np.zeros(W); wmax = np.full(W, 1e10); rebuild = [True] * W

def local_step(w):
    a, b = bounds[w]
    if b <= a: term[w] = 0.0; return
    s = spec[w]; v = s.max(); c = np.where(s == v)[0]
    top = a + c[np.argmin(gid[a + c])]
    lists[w].append((v, int(gid[top]), top))
    np.minimum(s, d2(pts[a:b], pts[top]), out=s); term[w] = s.max()

def slow_path(w, samples):
    a, b = bounds[w]
    if b <= a: return
    for p, own in samples:
        np.minimum(tmp[a:b], d2(pts[a:b], pts[p]), out=tmp[a:b])
        if not own and not rebuild[w]: np.minimum(spec[w], d2(pts[a:b], pts[p]), out=spec[w])
    if rebuild[w]:
        spec[w] = tmp[a:b].copy()
        for (v, g, p) in lists[w]: np.minimum(spec[w], d2(pts[a:b], pts[p]), out=spec[w])
        term[w] = spec[w].max(); rebuild[w] = False
    while len(lists[w]) < D: local_step(w)
    wmax[w] = tmp[a:b].max()

out = [0]; first = np.where(gid == 0)[0][0]
prev = [(first, -1)]          # accepted in the previous round: (position, owner warp)
for w in range(W): slow_path(w, [(first, False)])
prev = []
rounds = 0; stops = {"term": 0, "conflict": 0, "lmax": 0}
while len(out) < m:
    rounds += 1
    # ---- fast path: pop / truncate against prev (already applied to tmp? no: prev is applied in THIS round's slow path) ----
    touched = [[] for _ in range(W)]
    for (p, ow) in prev:
        ex = np.maximum(np.maximum(lo - pts[p], pts[p] - hi), 0.0); b2 = (ex ** 2).sum(1)
        for w in np.where(b2 < wmax)[0]: touched[w].append((p, ow == w))
        if ow >= 0 and not any(q == p for q, _ in touched[ow]): touched[ow].append((p, True))
    for w in range(W):
        own = [p for p, o in touched[w] if o]
        for p in own:
            assert lists[w] and lists[w][0][2] == p; lists[w].pop(0)
        foreign = [p for p, o in touched[w] if not o]
        for j, (v, g, q) in enumerate(lists[w]):
            if any(d2(pts[q], pts[p]) < v for p in foreign):
                term[w] = max(v, 0.0) if True else term[w]
                term[w] = v; lists[w] = lists[w][:j]; rebuild[w] = True; break
    # ---- merge (CTA top KC, then cluster) ----
    ents = []
    for c in range(W // WPC):
        ce = []
        for w in range(c * WPC, c * WPC + WPC):
            ce += [(-v, g, p, w, 0) for (v, g, p) in lists[w]]; ce.append((-term[w], -1, -1, w, 1))
        ce.sort(); ents += ce[:KC]; ents.append((ce[KC][0], -1, -1, -c - 1, 1)) if len(ce) > KC else None
    ents.sort(); acc = []; reason = "lmax"
    for e in ents:
        if e[4] == 1: reason = "term"; break
        if len(acc) >= LMAX: break
        if any(a_[3] != e[3] and d2(pts[e[2]], pts[a_[2]]) < -e[0] for a_ in acc): reason = "conflict"; break
        acc.append(e)
    stops[reason] += 1
    # ---- slow path of this round: apply prev, refill (concurrent with the merge: does not see acc) ----
    for w in range(W): slow_path(w, touched[w]) if (touched[w] or len(lists[w]) < D or rebuild[w]) else None
    if not acc:            # every list is empty or lagging: nothing to commit this round (costs a round, stays exact)
        prev = []; continue
    acc = acc[: m - len(out)]; out += [e[1] for e in acc]
    prev = [(e[2], e[3]) for e in acc]
print(f"n={n} W={W} D={D} KC={KC} LMAX={LMAX}: identical={out == ref[:len(out)]} rounds={rounds} samples={len(out)-1} chain={(len(out)-1)/rounds:.2f} stops={stops}")
