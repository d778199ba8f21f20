"""Thread-by-thread CPU emulation of csrc/linear.cu's index arithmetic (no GPU in the build container):
transliterates fetch / stash / compute / split-K tree / epilogue and compares with numpy."""
import numpy as np, sys

def run(BM, BN, BK, KS, TM, TN, M, K, N, lda=None, relu=True, use_bias=True, use_res=True, seed=0):
    rng = np.random.default_rng(seed)
    lda = lda or K
    Abuf = rng.standard_normal((M, lda)).astype(np.float32)
    A = Abuf.reshape(-1)
    Wt = rng.standard_normal((K, N)).astype(np.float32).reshape(-1)
    bias = rng.standard_normal(N).astype(np.float32) if use_bias else None
    res = rng.standard_normal((M, N)).astype(np.float32) if use_res else None
    out = np.full((M, N), np.nan, np.float32)
    TX, TY = BN // TN, BM // TM
    RH, CH = TM // 4, TN // 4
    G = TX * TY; NT = G * KS; KPG = BK // KS; AS = BM + 4
    FA, FB = BM * BK // 4, BK * BN // 4
    LA, LB = -(-FA // NT), -(-FB // NT)
    TILE = 2 * BK * (AS + BN); RED = (KS // 2) * BM * BN if KS > 1 else 0
    assert max(TILE, RED) * 4 <= 48 * 1024
    col_tiles = -(-N // BN); row_tiles = -(-M // BM)
    def load4(buf, off, valid):
        return [buf[off + j] if valid > j else np.float32(0) for j in range(4)]
    for bid in range(row_tiles * col_tiles):
        smem = np.full(max(TILE, RED), np.nan, np.float32)
        AsO, BsO = 0, 2 * BK * AS
        m0 = (bid // col_tiles) * BM; n0 = (bid % col_tiles) * BN
        ra = {}; rb = {}
        acc = np.zeros((NT, TM, TN), np.float32)
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
                    ar = np.concatenate([smem[a + kk * AS + h * (BM // RH): a + kk * AS + h * (BM // RH) + 4] for h in range(RH)])
                    br = np.concatenate([smem[b + kk * BN + h * (BN // CH): b + kk * BN + h * (BN // CH) + 4] for h in range(CH)])
                    assert not np.isnan(ar).any() and not np.isnan(br).any()
                    acc[tid] += np.outer(ar, br)
            if more: stash((t + 1) & 1)
        def finish(gm, gn, v):
            if gm >= M or gn >= N: return
            valid = N - gn
            for j in range(min(4, valid)):
                x = v[j]
                if bias is not None: x += bias[gn + j]
                if res is not None: x += res[gm, gn + j]
                if relu: x = max(x, 0)
                assert np.isnan(out[gm, gn + j]), "double write"
                out[gm, gn + j] = x
        if KS > 1:
            Q = TM * TN // 4
            half = KS // 2
            while half >= 1:
                red = np.full((RED // 4, 4), np.nan, np.float32)
                for tid in range(NT):
                    g, r = tid // G, tid % G
                    if half <= g < 2 * half:
                        for i in range(TM):
                            for h in range(CH):
                                o = ((g - half) * Q + i * CH + h) * G + r
                                assert np.isnan(red[o]).all()
                                red[o] = acc[tid, i, h * 4:h * 4 + 4]
                for tid in range(NT):
                    g, r = tid // G, tid % G
                    if g < half:
                        for i in range(TM):
                            for h in range(CH):
                                p = red[(g * Q + i * CH + h) * G + r]
                                assert not np.isnan(p).any()
                                acc[tid, i, h * 4:h * 4 + 4] += p
                half >>= 1
        for tid in range(NT):
            g, r = tid // G, tid % G; tx, ty = r % TX, r // TX
            if g != 0: continue
            for i in range(TM):
                for h in range(CH):
                    finish(m0 + (i // 4) * (BM // RH) + ty * 4 + (i % 4), n0 + h * (BN // CH) + tx * 4, acc[tid, i, h * 4:h * 4 + 4])
    ref = Abuf[:, :K].astype(np.float64) @ Wt.reshape(K, N).astype(np.float64)
    if bias is not None: ref += bias
    if res is not None: ref += res
    if relu: ref = np.maximum(ref, 0)
    assert not np.isnan(out).any(), "unwritten outputs"
    return np.abs(out - ref).max() / max(1, np.abs(ref).max())

cfgs = [(128, 32, 16, 1, 4, 4), (64, 32, 16, 2, 4, 4), (32, 32, 32, 4, 4, 4), (16, 32, 32, 8, 4, 4), (64, 64, 16, 1, 4, 4),
        (32, 64, 32, 2, 4, 4), (16, 64, 32, 4, 4, 4), (128, 32, 16, 1, 8, 4), (256, 32, 16, 1, 8, 4), (128, 64, 16, 1, 8, 8),
        (64, 64, 16, 2, 8, 8), (64, 64, 32, 4, 8, 8), (32, 64, 32, 8, 8, 8), (128, 64, 16, 2, 8, 8), (64, 128, 16, 2, 8, 8),
        (32, 128, 32, 4, 8, 8)]
shapes = [(130, 32, 32), (70, 8, 12), (33, 72, 40), (17, 64, 96), (5, 20, 136)]
for n, cfg in enumerate(cfgs, 1):
    for (M, K, N) in shapes:
        e = run(*cfg, M, K, N, lda=K + 4)
        assert e < 1e-5, (cfg, M, K, N, e)
    print("config", n, cfg, "ok", flush=True)
print("ok")
