"""Thread-by-thread CPU emulation of csrc/linear.cu's index arithmetic (no GPU in the build container):
transliterates fetch / stash / compute / split-K reduction / epilogue and compares with numpy."""
import numpy as np, itertools, sys

def run(BM, BN, BK, KS, M, K, N, lda=None, relu=True, use_bias=True, use_res=True, seed=0):
    rng = np.random.default_rng(seed)
    lda = lda or K
    Abuf = rng.standard_normal((M, lda)).astype(np.float32)
    A = Abuf.reshape(-1)
    Wt = rng.standard_normal((K, N)).astype(np.float32).reshape(-1)
    bias = rng.standard_normal(N).astype(np.float32) if use_bias else None
    res = rng.standard_normal((M, N)).astype(np.float32) if use_res else None
    out = np.full((M, N), np.nan, np.float32)
    TX, TY = BN // 4, BM // 4
    G = TX * TY; NT = G * KS; KPG = BK // KS; AS = BM + 4
    FA, FB = BM * BK // 4, BK * BN // 4
    LA, LB = -(-FA // NT), -(-FB // NT)
    TILE = 2 * BK * (AS + BN); RED = KS * BM * BN if KS > 1 else 0
    col_tiles = -(-N // BN); row_tiles = -(-M // BM)
    def load4(buf, off, valid):
        return [buf[off + j] if valid > j else np.float32(0) for j in range(4)]
    for bid in range(row_tiles * col_tiles):
        smem = np.full(max(TILE, RED), np.nan, np.float32)
        AsO, BsO = 0, 2 * BK * AS
        m0 = (bid // col_tiles) * BM; n0 = (bid % col_tiles) * BN
        ra = {}; rb = {}
        acc = np.zeros((NT, 4, 4), np.float32)
        def fetch(k0):
            for tid in range(NT):
                for i in range(LA):
                    f = tid + i * NT
                    if FA % NT == 0 or f < FA:
                        row, kq = f % BM, f // BM
                        gm, gk = m0 + row, k0 + kq * 4
                        ra[tid, i] = load4(A, (gm if gm < M else 0) * lda + gk, (K - gk) if gm < M else 0)
                for i in range(LB):
                    f = tid + i * NT
                    if FB % NT == 0 or f < FB:
                        kk, nq = f // (BN // 4), f % (BN // 4)
                        gk, gn = k0 + kk, n0 + nq * 4
                        rb[tid, i] = load4(Wt, (gk if gk < K else 0) * N + gn, (N - gn) if gk < K else 0)
        def stash(buf):
            a = AsO + buf * BK * AS; b = BsO + buf * BK * BN
            for tid in range(NT):
                for i in range(LA):
                    f = tid + i * NT
                    if FA % NT == 0 or f < FA:
                        row, kq = f % BM, f // BM
                        for j in range(4):
                            smem[a + (kq * 4 + j) * AS + row] = ra[tid, i][j]
                for i in range(LB):
                    f = tid + i * NT
                    if FB % NT == 0 or f < FB:
                        kk, nq = f // (BN // 4), f % (BN // 4)
                        for j in range(4):
                            smem[b + kk * BN + nq * 4 + j] = rb[tid, i][j]
        tiles = -(-K // BK)
        fetch(0); stash(0)
        for t in range(tiles):
            more = t + 1 < tiles
            if more: fetch((t + 1) * BK)
            for tid in range(NT):
                g, r = tid // G, tid % G; tx, ty = r % TX, r // TX
                a = AsO + (t & 1) * BK * AS + (g * KPG) * AS + ty * 4
                b = BsO + (t & 1) * BK * BN + (g * KPG) * BN + tx * 4
                for kk in range(KPG):
                    av = smem[a + kk * AS: a + kk * AS + 4]; bv = smem[b + kk * BN: b + kk * BN + 4]
                    assert not np.isnan(av).any() and not np.isnan(bv).any()
                    acc[tid] += np.outer(av, bv)
            if more: stash((t + 1) & 1)
        def finish(gm, gn, v):
            if gm >= M or gn >= N: return
            valid = N - gn
            v = list(v)
            for j in range(min(4, valid)):
                x = v[j]
                if bias is not None: x += bias[gn + j]
                if res is not None: x += res[gm, gn + j]
                if relu: x = max(x, 0)
                assert np.isnan(out[gm, gn + j]), "double write"
                out[gm, gn + j] = x
        if KS == 1:
            for tid in range(NT):
                r = tid % G; tx, ty = r % TX, r // TX
                for i in range(4): finish(m0 + ty * 4 + i, n0 + tx * 4, acc[tid, i])
        else:
            red = np.full(RED, np.nan, np.float32)
            for tid in range(NT):
                g, r = tid // G, tid % G; tx, ty = r % TX, r // TX
                for i in range(4):
                    o = ((g * BM) + ty * 4 + i) * BN + tx * 4
                    assert np.isnan(red[o:o + 4]).all()
                    red[o:o + 4] = acc[tid, i]
            for tid in range(NT):
                for f in range(tid, BM * BN // 4, NT):
                    row, nq = f // (BN // 4), f % (BN // 4)
                    s = red[row * BN + nq * 4: row * BN + nq * 4 + 4].copy()
                    for h in range(1, KS):
                        o = ((h * BM) + row) * BN + nq * 4
                        s += red[o:o + 4]
                    finish(m0 + row, n0 + nq * 4, s)
    ref = Abuf[:, :K].astype(np.float64) @ Wt.reshape(K, N).astype(np.float64)
    if bias is not None: ref += bias
    if res is not None: ref += res
    if relu: ref = np.maximum(ref, 0)
    assert not np.isnan(out).any(), "unwritten outputs"
    err = np.abs(out - ref).max() / max(1, np.abs(ref).max())
    return err

cfgs = [(64, 64, 16, 1), (32, 64, 32, 2), (16, 64, 32, 4)] if len(sys.argv) > 1 else [(128, 32, 16, 1), (64, 32, 16, 2), (32, 32, 32, 4), (16, 32, 32, 8)]
shapes = [(130, 32, 32), (70, 6, 13), (33, 72, 40), (17, 64, 96), (5, 20, 7)]
for cfg in cfgs:
    for (M, K, N) in shapes:
        e = run(*cfg, M, K, N, lda=K + (4 if K % 4 == 0 else 3))
        print(cfg, (M, K, N), f"{e:.2e}")
        assert e < 1e-5
print("ok")
