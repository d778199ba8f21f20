"""CPU simulation of the MERGED-LIST speculative FPS protocol (round-2 design study for csrc/fps.cu).

Every group (a warp's compact run of points) keeps a LOCAL FPS continuation: a_1, a_2, ... a_D = what the
group's own argmax sequence would be if only its own candidates were applied.  One exchange merges all
lists by key (value desc, index asc); the merged prefix is exactly the global FPS sequence as long as no
entry is lowered by an earlier entry of ANOTHER group (pairwise point test) and no group's list has run
out (terminal bound).  Checks equality with plain FPS and reports samples per exchange.

python tools/fps_merge_sim.py n W D_restart D_max [Kc] [m_limit]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloudpdf_b200 import synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
D0 = int(sys.argv[3]) if len(sys.argv) > 3 else 2
DMAX = int(sys.argv[4]) if len(sys.argv) > 4 else 4
KC = int(sys.argv[5]) if len(sys.argv) > 5 else 0      # entries a CTA (8 warps) may publish; 0 = flat merge
MLIM = int(sys.argv[6]) if len(sys.argv) > 6 else 0
LMAX = int(sys.argv[7]) if len(sys.argv) > 7 else 32
LAZY = int(sys.argv[8]) if len(sys.argv) > 8 else 0   # 1: a clean warp refills only when its list is empty
KEEP = int(os.environ.get("KEEP", "0"))   # 1: a foreign sample is applied to spec too; restart only if it lowers a pending entry
m = n // 4
if MLIM:
    m = min(m, MLIM)
xyz = S.s3dis_batch([n], seed=2026)["coord"].numpy().astype(np.float64)
Lb = xyz.max(0) - xyz.min(0)
h = max((Lb.prod() / (n / 2)) ** (1 / 3), 1e-3)
cell = np.floor((xyz - xyz.min(0)) / h).astype(np.int64)
dd = cell.max(0) + 1
key = (cell[:, 2] * dd[1] + cell[:, 1]) * dd[0] + cell[:, 0]
ORDER = os.environ.get("ORDER", "row")
if ORDER.startswith("morton"):
    B = int(ORDER[6:] or 2)            # block edge in cells
    blk = cell // B
    def spread(v):
        out = np.zeros_like(v)
        for i in range(10):
            out |= ((v >> i) & 1) << (3 * i)
        return out
    code = spread(blk[:, 0]) | (spread(blk[:, 1]) << 1) | (spread(blk[:, 2]) << 2)
    key = code * (dd.prod() + 1) + key     # Morton over blocks, row-major (z, y, x) inside a block
if ORDER == "hilbert":                 # what the launcher does: 10-bit Hilbert index over the scene's extent
    ext = Lb.max()
    c = np.minimum((xyz - xyz.min(0)) / ext * 1024.0, 1023.0).astype(np.int64)
    X = [c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy()]
    Q = 512
    while Q > 1:
        P_ = Q - 1
        for i in range(3):
            hit = (X[i] & Q) != 0
            t = (X[0] ^ X[i]) & P_
            x0 = np.where(hit, X[0] ^ P_, X[0] ^ t)
            if i:
                X[i] = np.where(hit, X[i], X[i] ^ t)
            X[0] = x0 if i else np.where(hit, X[0] ^ P_, X[0])
        Q >>= 1
    X[1] ^= X[0]; X[2] ^= X[1]
    t = np.zeros_like(X[0]); Q = 512
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ (Q - 1), t); Q >>= 1
    X = [x ^ t for x in X]
    def spread(v):
        out = np.zeros_like(v)
        for i in range(10):
            out |= ((v >> i) & 1) << (3 * i)
        return out
    key = (spread(X[0]) << 2) | (spread(X[1]) << 1) | spread(X[2])
order = np.argsort(key, kind="stable")
pts = xyz[order]
gid = order


def d2(a, b):
    return ((a - b) ** 2).sum(-1)


def fps_ref():
    tmp = np.full(n, 1e10)
    out = [0]
    for _ in range(m - 1):
        tmp = np.minimum(tmp, d2(xyz, xyz[out[-1]]))
        out.append(int(np.argmax(tmp)))
    return out


ref = fps_ref()

per = -(-n // W)
bounds = [(w * per, min(n, (w + 1) * per)) for w in range(W)]
lo = np.array([pts[a:b].min(0) if b > a else np.zeros(3) for a, b in bounds])
hi = np.array([pts[a:b].max(0) if b > a else np.zeros(3) for a, b in bounds])
tmp = np.full(n, 1e10)
spec = [None] * W          # speculative tmp of the group (own list applied)
lists = [[] for _ in range(W)]   # entries (value, gid, pos)
term = np.zeros(W)         # terminal bound
wmax = np.full(W, 1e10)


steps_tot = 0


step_log = {}


def local_step(w):
    global steps_tot
    steps_tot += 1
    step_log[w] = step_log.get(w, 0) + 1
    a, b = bounds[w]
    if b <= a:
        term[w] = 0.0
        return
    s = spec[w]
    v = s.max()
    c = np.where(s == v)[0]
    top = a + c[np.argmin(gid[a + c])]
    lists[w].append((v, int(gid[top]), top))
    np.minimum(s, d2(pts[a:b], pts[top]), out=s)
    term[w] = s.max()


def restart(w, steps):
    a, b = bounds[w]
    spec[w] = tmp[a:b].copy()
    lists[w] = []
    for _ in range(steps):
        local_step(w)


out = [0]
first = np.where(gid == 0)[0][0]
tmp = np.minimum(tmp, d2(pts, pts[first]))
for w in range(W):
    restart(w, D0)
rounds = 0
crit_steps = 0
touched_tot = 0
stop_reason = {"term": 0, "conflict": 0, "lmax": 0, "kc": 0}
while len(out) < m:
    rounds += 1
    # ---- build the merged order ----
    ents = []
    if KC:
        for c in range(W // 8):
            ce = []
            for w in range(c * 8, c * 8 + 8):
                ce += [(-v, g, p, w, 0) for (v, g, p) in lists[w]]
                ce.append((-term[w], -1, -1, w, 1))
            ce.sort()
            # CTA-level prefix: up to KC real entries without intra-CTA conflict; terminal = next key
            acc = []
            tb = 0.0
            for e in ce:
                if e[4] == 1:
                    tb = -e[0]; break
                if len(acc) >= KC:
                    tb = -e[0]; break
                if any(a_[3] != e[3] and d2(pts[e[2]], pts[a_[2]]) < -e[0] for a_ in acc):
                    tb = -e[0]; break
                acc.append(e)
            ents += acc
            ents.append((-tb, -1, -1, -c - 1, 1))
    else:
        for w in range(W):
            ents += [(-v, g, p, w, 0) for (v, g, p) in lists[w]]
            ents.append((-term[w], -1, -1, w, 1))
    ents.sort()
    acc = []
    reason = "lmax"
    for e in ents:
        if e[4] == 1:
            reason = "term"; break
        if len(acc) >= LMAX:
            break
        grp = e[3] if not KC else e[3] // 8
        if any((a_[3] if not KC else a_[3] // 8) != grp and d2(pts[e[2]], pts[a_[2]]) < -e[0] for a_ in acc):
            reason = "conflict"; break
        acc.append(e)
    stop_reason[reason] += 1
    assert acc, "no progress"
    acc = acc[: m - len(out)]
    out += [e[1] for e in acc]
    # ---- apply ----
    dirty = set()
    napply = np.zeros(W, dtype=np.int64)
    nsteps0 = steps_tot
    for e in acc:
        p = pts[e[2]]
        ex = np.maximum(np.maximum(lo - p, p - hi), 0.0)
        b2 = (ex ** 2).sum(1)
        for w in np.where(b2 < wmax)[0]:
            a, b = bounds[w]
            if b <= a:
                continue
            np.minimum(tmp[a:b], d2(pts[a:b], p), out=tmp[a:b])
            if w != e[3]:
                if KEEP:
                    np.minimum(spec[w], d2(pts[a:b], p), out=spec[w])
                    if any(d2(pts[le[2]], p) < le[0] for le in lists[w]):
                        dirty.add(w)
                else:
                    dirty.add(w)
            touched_tot += 1
            napply[w] += 1
        w = e[3]
        assert lists[w][0][2] == e[2]
        lists[w].pop(0)
    for w in range(W):
        a, b = bounds[w]
        if b > a:
            wmax[w] = tmp[a:b].max()
    mx = 0
    restarts_tot = globals().get("restarts_tot", 0) + len(dirty)
    for w in range(W):
        if w in dirty:
            restart(w, D0); mx = max(mx, D0)
        else:
            if (len(lists[w]) < DMAX and not LAZY) or len(lists[w]) == 0:
                need = 1 if len(lists[w]) >= 1 else max(1, D0)
                for _ in range(need):
                    local_step(w)
                mx = max(mx, 1)
    crit_steps += mx
    stepw = np.zeros(W, dtype=np.int64)
    for w in range(W):
        stepw[w] = step_log.get(w, 0)
    step_log.clear()
    cost = napply * 210 + stepw * 480 + np.array([300 if w in dirty else 0 for w in range(W)])
    cost_max_sum = globals().get("cost_max_sum", 0) + cost.max()
    cost_mean_sum = globals().get("cost_mean_sum", 0) + cost.mean()
    apply_max_sum = globals().get("apply_max_sum", 0) + napply.max()
print(f"n={n} W={W} D0={D0} DMAX={DMAX} KC={KC} LMAX={LMAX}: identical={out == ref[:len(out)]} rounds={rounds} samples={len(out)-1} "
      f"cost max/mean per round={cost_max_sum / rounds:.0f}/{cost_mean_sum / rounds:.0f} cyc, max applies/warp={apply_max_sum / rounds:.1f} "
      f"mean chain={(len(out) - 1) / rounds:.2f} restarts/round={restarts_tot / rounds:.1f} steps/round={steps_tot / rounds:.1f} touched/round={touched_tot / rounds:.1f} stops={stop_reason}")
