"""CPU simulation of the speculative-chain FPS protocol of csrc/fps.cu (fps_chain_kernel): checks that
the emitted sequence equals plain FPS and reports the mean accepted chain length.
python tools/fps_chain_sim.py [n] [groups] [warps_per_group]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloudpdf_b200 import synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 16
NW = int(sys.argv[3]) if len(sys.argv) > 3 else 8
KMAX = int(sys.argv[4]) if len(sys.argv) > 4 else 8
lattice = len(sys.argv) > 5
m = n // 4
xyz = S.s3dis_batch([n], seed=2026)["coord"].numpy().astype(np.float64)
if lattice:
    xyz = np.round(xyz / 0.25) * 0.25   # heavy ties
# cell order like the kNN grid (z, y, x-fastest), ~2 points per cell budget
L = xyz.max(0) - xyz.min(0)
h = max((L.prod() / (n / 2)) ** (1 / 3), 1e-3)
cell = np.floor((xyz - xyz.min(0)) / h).astype(np.int64)
d = cell.max(0) + 1
key = (cell[:, 2] * d[1] + cell[:, 1]) * d[0] + cell[:, 0]
order = np.argsort(key, kind="stable")
pts = xyz[order]
gid = order  # original index


def d2(a, b):
    return ((a - b) ** 2).sum(-1)


# reference FPS on the original order, lowest index among maxima
def fps_ref():
    tmp = np.full(n, 1e10)
    out = [0]
    for _ in range(m - 1):
        tmp = np.minimum(tmp, d2(xyz, xyz[out[-1]]))
        out.append(int(np.argmax(tmp)))  # argmax returns first maximum = lowest index
    return out


ref = fps_ref()

# chain protocol
W = G * NW
per = -(-n // W)
warp_of = np.minimum(np.arange(n) // per, W - 1)
tmp = np.full(n, 1e10)
out = [0]
chain = [np.where(gid == 0)[0][0]]  # positions (in cell order) of accepted, not yet applied samples
rounds = 0
while True:
    for c in chain:
        tmp = np.minimum(tmp, d2(pts, pts[c]))
    if len(out) >= m:
        break
    rounds += 1
    # warp entries: max (value, lowest original idx), V = warp max after applying own top
    ents = []
    for w in range(W):
        sel = np.where(warp_of == w)[0]
        if len(sel) == 0:
            ents.append((0.0, 1 << 60, -1, 0.0)); continue
        v = tmp[sel].max()
        cands = sel[tmp[sel] == v]
        top = cands[np.argmin(gid[cands])]
        V = np.minimum(tmp[sel], d2(pts[sel], pts[top])).max()
        ents.append((v, int(gid[top]), top, V))
    groups = []
    for g in range(G):
        es = ents[g * NW:(g + 1) * NW]
        best = min(range(NW), key=lambda i: (-es[i][0], es[i][1]))
        second = max([es[i][0] for i in range(NW) if i != best], default=0.0)
        groups.append((es[best][0], es[best][1], es[best][2], max(es[best][3], second)))
    groups.sort(key=lambda e: (-e[0], e[1]))
    Ln = 1
    for j in range(1, min(KMAX, G)):
        okj = all((not (d2(pts[groups[j][2]], pts[groups[i][2]]) < groups[j][0])) and groups[i][3] < groups[j][0]
                  for i in range(j)) and groups[j][2] >= 0
        if not okj:
            break
        Ln = j + 1
    Ln = min(Ln, m - len(out))
    chain = [groups[j][2] for j in range(Ln)]
    out += [groups[j][1] for j in range(Ln)]
print(f"n={n} groups={G} warps/group={NW} kmax={KMAX} lattice={lattice}: identical={out == ref} "
      f"rounds={rounds} samples={m - 1} mean chain={(m - 1) / rounds:.2f}")
