// gather_fast.cuh -- compile-time-tiled versions of the gather / scatter kernels for the shapes
// PTv1 uses (C = 4*LPR with LPR in {1,2,4,8,16,32}; nsample 8 or 16 for the reductions).
//
// Why a second set of kernels: the ncu capture of the run-time-tiled ones
// (profiles/r01_gather_v1_ncu.md) showed ~43 thread-instructions per 16 bytes moved -- integer
// address arithmetic, constant-bank reloads of the tile descriptor and per-row bounds checks --
// and 36-110 % XU/ALU pipe utilisation, i.e. the kernels were instruction-bound, not HBM-bound.
// Here every warp works on a TILE of R = U * (32/LPR) consecutive rows: the tile's rows of the
// streamed operand are one contiguous span, so lane l's u-th 16-byte access is simply
// base + u*32 + l; the only per-row integer work left is the gathered row's index.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace pob {

constexpr int FAST_THREADS = 256;

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ldcs4(const float4* p) { return __ldcs(p); }   // streamed, evict-first
__device__ __forceinline__ void stcs4(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 scale4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 xor4(float4 v, int o) {
    return make_float4(__shfl_xor_sync(FULL, v.x, o), __shfl_xor_sync(FULL, v.y, o), __shfl_xor_sync(FULL, v.z, o),
                       __shfl_xor_sync(FULL, v.w, o));
}
__device__ __forceinline__ void red_add4(float4* p, float4 v) { atomicAdd(p, v); }  // RED.E.ADD.F32x4

// four consecutive channels of a feature row as f32, for f32 / f16 / bf16 storage
template <typename T> struct Vec4Load;
template <> struct Vec4Load<float> {
    static __device__ __forceinline__ float4 ld(const float* base, int64_t vec) { return __ldg(reinterpret_cast<const float4*>(base) + vec); }
};
template <> struct Vec4Load<__half> {
    static __device__ __forceinline__ float4 ld(const __half* base, int64_t vec) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(base) + vec);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
};
template <> struct Vec4Load<__nv_bfloat16> {
    static __device__ __forceinline__ float4 ld(const __nv_bfloat16* base, int64_t vec) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(base) + vec);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
};

// ---------------------------------------------------------------------------------------------
// out[r, :] = in[idx[r], :]                    (grouping2 forward; MASKED: idx < 0 -> zeros)
// out[r, :] = in1[r / ns, :] - in[idx[r], :]   (SUB: subtraction forward)
template <typename T, int LPR, int U, bool SUB, bool MASKED>
__global__ void __launch_bounds__(FAST_THREADS)
gather_rows_fast(int64_t rows, int ns, int ns_shift, const T* __restrict__ in, const float4* __restrict__ in1,
                 const int* __restrict__ idx, float4* __restrict__ out) {
    constexpr int SPAR = 32 / LPR, R = SPAR * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (rows + R - 1) / R;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t row0 = tile * R;
        const bool full = row0 + R <= rows;
        const int* ip = idx + row0 + sub;
        float4* op = out + row0 * LPR + lane;
        int src[U];
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) src[u] = (full || row0 + u * SPAR + sub < rows) ? __ldg(ip + u * SPAR) : -1;
#pragma unroll
        for (int u = 0; u < U; u++) {
            v[u] = (!MASKED && full) || src[u] >= 0 ? Vec4Load<T>::ld(in, (int64_t)src[u] * LPR + cl) : zero4();
            if (SUB) {
                const int64_t r = row0 + u * SPAR + sub;
                const int64_t p = ns_shift >= 0 ? (r >> ns_shift) : r / ns;
                if (full || r < rows) v[u] = sub4(ldg4(in1 + p * LPR + cl), v[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (full || row0 + u * SPAR + sub < rows) stcs4(op + u * 32, v[u]);
    }
}

// ---------------------------------------------------------------------------------------------
// out[n, :] = sum_s (in[idx[n,s], :] + pos[n,s,:]) * w[n,s, : % w_c]      (aggregation forward)
// Tile = PPT whole points; the position rows of a tile are one contiguous span.
template <int LPR, int NS>
__global__ void __launch_bounds__(FAST_THREADS)
aggregation_fwd_fast(int64_t n, int wvec, const float4* __restrict__ in, const float4* __restrict__ pos,
                     const float4* __restrict__ w, const int* __restrict__ idx, float4* __restrict__ out) {
    constexpr int SPAR = 32 / LPR;
    constexpr int U = (NS / SPAR > 8) ? NS / SPAR : 8;  // row-slots per lane
    constexpr int R = U * SPAR, PPT = R / NS, UPP = U / PPT;
    static_assert(NS % SPAR == 0 && R % NS == 0, "tile must hold whole points");
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int wv = cl & (wvec - 1);
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (n + PPT - 1) / PPT;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t p0 = tile * PPT, row0 = p0 * NS;
        const bool full = p0 + PPT <= n;
        const int* ip = idx + row0 + sub;
        const float4* pp = pos + row0 * LPR + lane;
        const float4* wp = w + (row0 + sub) * wvec + wv;
        int src[U];
        float4 a[U], b[U], ww[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = full || p0 + u / UPP < n;
            src[u] = ok ? __ldg(ip + u * SPAR) : -1;
            b[u] = ok ? ldcs4(pp + u * 32) : zero4();
            ww[u] = ok ? ldg4(wp + (int64_t)u * SPAR * wvec) : zero4();
        }
#pragma unroll
        for (int u = 0; u < U; u++) a[u] = src[u] >= 0 ? ldg4(in + (int64_t)src[u] * LPR + cl) : zero4();
        float4 acc[PPT];
#pragma unroll
        for (int q = 0; q < PPT; q++) acc[q] = zero4();
#pragma unroll
        for (int u = 0; u < U; u++) acc[u / UPP] = fma4(add4(a[u], b[u]), ww[u], acc[u / UPP]);
#pragma unroll
        for (int q = 0; q < PPT; q++) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) acc[q] = add4(acc[q], xor4(acc[q], o));
            if (sub == q % SPAR && (full || p0 + q < n)) out[(p0 + q) * LPR + cl] = acc[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// out[n, :] = sum_{i<k} in[idx[n,i], :] * w[n,i]                           (interpolation forward)
template <int LPR, int U>
__global__ void __launch_bounds__(FAST_THREADS)
interpolation_fwd_fast(int64_t n, int k, const float4* __restrict__ in, const int* __restrict__ idx,
                       const float* __restrict__ w, float4* __restrict__ out) {
    constexpr int SPAR = 32 / LPR, R = SPAR * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (n + R - 1) / R;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t p0 = tile * R;
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            acc[u] = zero4();
            const int64_t p = p0 + u * SPAR + sub;
            if (p < n) {
                for (int i = 0; i < k; i++) {
                    const int src = __ldg(idx + p * k + i);
                    const float wi = __ldg(w + p * k + i);
                    acc[u] = fma4(ldg4(in + (int64_t)src * LPR + cl), make_float4(wi, wi, wi, wi), acc[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (p0 + u * SPAR + sub < n) stcs4(out + p0 * LPR + u * 32 + lane, acc[u]);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused grouping-with-xyz forward: out[m, s, :] = [ (xyz[idx] - new_xyz[m]) , feat[idx] ], zero
// rows for idx < 0.  Rows are W = 3 + C floats, i.e. NOT 16-byte aligned on their own, but the
// ns rows of one query are one contiguous span of ns*W floats: each warp assembles that span in
// shared memory (128-bit gathers in, conflict-free scalar stores) and streams it out with
// 128-bit stores.  Dynamic shared memory: warps_per_cta * ns * W floats.
template <typename T, int LPR, int CH>
__global__ void __launch_bounds__(FAST_THREADS)
group_xyz_fwd_fast(int64_t m, int ns, const T* __restrict__ feat, const float* __restrict__ xyz,
                   const float* __restrict__ new_xyz, const int* __restrict__ idx, float* __restrict__ out) {
    constexpr int SPAR = 32 / LPR, C = 4 * LPR, W = C + 3;
    extern __shared__ __align__(16) float s_tile[];
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int per = ns * W;                       // floats per query
    float* tile = s_tile + (threadIdx.x >> 5) * ((per + 3) & ~3);
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    for (int64_t q = warp; q < m; q += nwarps) {
        const int* iq = idx + q * ns;
        // one index load per neighbour (lane s holds idx[q, s]); the feature passes fetch theirs by shuffle
        const int my_src = lane < ns ? __ldg(iq + lane) : -1;
        const float qx = __ldg(new_xyz + q * 3), qy = __ldg(new_xyz + q * 3 + 1), qz = __ldg(new_xyz + q * 3 + 2);
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (my_src >= 0) {
            dx = __fsub_rn(__ldg(xyz + (int64_t)my_src * 3), qx);
            dy = __fsub_rn(__ldg(xyz + (int64_t)my_src * 3 + 1), qy);
            dz = __fsub_rn(__ldg(xyz + (int64_t)my_src * 3 + 2), qz);
        }
        // the gathers of CH passes (CH * SPAR rows; the launcher picks CH so that ns = 8 / 16 take one chunk) are
        // issued before anything is stored: independent 16-byte loads per lane.  ns <= 32 (one index per lane).
        for (int s0 = 0; s0 < ns; s0 += CH * SPAR) {
            float4 v[CH];
#pragma unroll
            for (int ps = 0; ps < CH; ps++) {
                const int s = s0 + ps * SPAR + sub;
                const int src = __shfl_sync(FULL, my_src, s & 31);
                v[ps] = (s < ns && src >= 0) ? Vec4Load<T>::ld(feat, (int64_t)src * LPR + cl) : zero4();
            }
#pragma unroll
            for (int ps = 0; ps < CH; ps++) {
                const int s = s0 + ps * SPAR + sub;
                if (s < ns) {
                    float* t = tile + s * W + 3 + 4 * cl;
                    t[0] = v[ps].x; t[1] = v[ps].y; t[2] = v[ps].z; t[3] = v[ps].w;
                }
            }
        }
        if (lane < ns) { tile[lane * W] = dx; tile[lane * W + 1] = dy; tile[lane * W + 2] = dz; }
        __syncwarp();
        float* o = out + q * per;
        if ((per & 3) == 0) {   // q*per*4 bytes is then a multiple of 16
            const float4* t4 = reinterpret_cast<const float4*>(tile);
            float4* o4 = reinterpret_cast<float4*>(o);
            const int nv = per / 4;
#pragma unroll 4
            for (int vv = lane; vv < nv; vv += 32) stcs4(o4 + vv, t4[vv]);
        } else {
            for (int f = lane; f < per; f += 32) __stcs(o + f, tile[f]);
        }
        __syncwarp();
    }
}

// grad_feat[idx[m,s], :] += grad_out[m, s, 3:]   for idx >= 0  (same staging, other direction)
template <int LPR>
__global__ void __launch_bounds__(FAST_THREADS)
group_xyz_bwd_fast(int64_t m, int ns, const float* __restrict__ gout, const int* __restrict__ idx,
                   float4* __restrict__ gfeat) {
    constexpr int SPAR = 32 / LPR, C = 4 * LPR, W = C + 3;
    extern __shared__ __align__(16) float s_tile[];
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int per = ns * W;
    float* tile = s_tile + (threadIdx.x >> 5) * ((per + 3) & ~3);
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    for (int64_t q = warp; q < m; q += nwarps) {
        const float* g = gout + q * per;
        if ((per & 3) == 0) {
            const float4* g4 = reinterpret_cast<const float4*>(g);
            float4* t4 = reinterpret_cast<float4*>(tile);
            for (int v = lane; v < per / 4; v += 32) t4[v] = ldcs4(g4 + v);
        } else {
            for (int f = lane; f < per; f += 32) tile[f] = __ldcs(g + f);
        }
        __syncwarp();
        const int* iq = idx + q * ns;
        for (int s0 = 0; s0 < ns; s0 += SPAR) {
            const int s = s0 + sub;
            if (s < ns) {
                const int dst = __ldg(iq + s);
                if (dst >= 0) {
                    const float* t = tile + s * W + 3 + 4 * cl;
                    red_add4(gfeat + (int64_t)dst * LPR + cl, make_float4(t[0], t[1], t[2], t[3]));
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// gin[idx[r], :] += sign * gout[r, :]          (grouping2 backward, subtraction backward w.r.t. input2)
template <int LPR, int U>
__global__ void __launch_bounds__(FAST_THREADS)
scatter_rows_fast(int64_t rows, float sign, const float4* __restrict__ gout, const int* __restrict__ idx,
                  float4* __restrict__ gin) {
    constexpr int SPAR = 32 / LPR, R = SPAR * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (rows + R - 1) / R;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t row0 = tile * R;
        const bool full = row0 + R <= rows;
        int dst[U];
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = full || row0 + u * SPAR + sub < rows;
            dst[u] = ok ? __ldg(idx + row0 + u * SPAR + sub) : -1;
            v[u] = ok ? ldcs4(gout + row0 * LPR + u * 32 + lane) : zero4();
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (dst[u] >= 0) red_add4(gin + (int64_t)dst[u] * LPR + cl, scale4(v[u], sign));
    }
}

// out[n, s, :] = in1[n, :] - in2[idx[n, s], :]          (subtraction forward, whole points per tile)
// The row-tiled form above (SUB) reloads in1 for every row and holds two float4 per row slot: 75 registers,
// 3 CTAs per SM, 32 % of the warps active and twice the time of the plain gather (ncu, profiles/r02_ops_ncu.md).
// Here a tile is PPT whole points: in1 is read once per point, and the launch bound keeps 5 CTAs per SM.
template <int LPR, int NS>
__global__ void __launch_bounds__(FAST_THREADS, 5)
subtraction_fwd_fast(int64_t n, const float4* __restrict__ in1, const float4* __restrict__ in2, const int* __restrict__ idx,
                     float4* __restrict__ out) {
    constexpr int SPAR = 32 / LPR;
    constexpr int U = (NS / SPAR > 8) ? NS / SPAR : 8;
    constexpr int R = U * SPAR, PPT = R / NS, UPP = U / PPT;
    static_assert(NS % SPAR == 0 && R % NS == 0, "tile must hold whole points");
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (n + PPT - 1) / PPT;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t p0 = tile * PPT, row0 = p0 * NS;
        const bool full = p0 + PPT <= n;
        const int* ip = idx + row0 + sub;
        float4* op = out + row0 * LPR + lane;
        int src[U];
#pragma unroll
        for (int u = 0; u < U; u++) src[u] = (full || p0 + u / UPP < n) ? __ldg(ip + u * SPAR) : -1;
        float4 a1[PPT];
#pragma unroll
        for (int q = 0; q < PPT; q++) a1[q] = (full || p0 + q < n) ? ldg4(in1 + (p0 + q) * LPR + cl) : zero4();
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (full || p0 + u / UPP < n) {
                const float4 g = src[u] >= 0 ? ldg4(in2 + (int64_t)src[u] * LPR + cl) : zero4();
                stcs4(op + u * 32, sub4(a1[u / UPP], g));
            }
        }
    }
}

// g1[n, :] = sum_s gout[n, s, :] ; g2[idx[n, s], :] -= gout[n, s, :]     (subtraction backward, ONE pass over gout:
// the reduction for input1 and the vector-atomic scatter for input2 share the streamed tile)
template <int LPR, int NS>
__global__ void __launch_bounds__(FAST_THREADS)
subtraction_bwd_fast(int64_t n, const float4* __restrict__ gout, const int* __restrict__ idx, float4* __restrict__ g1,
                     float4* __restrict__ g2) {
    constexpr int SPAR = 32 / LPR;
    constexpr int U = (NS / SPAR > 8) ? NS / SPAR : 8;
    constexpr int R = U * SPAR, PPT = R / NS, UPP = U / PPT;
    static_assert(NS % SPAR == 0 && R % NS == 0, "tile must hold whole points");
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (n + PPT - 1) / PPT;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t p0 = tile * PPT, row0 = p0 * NS;
        const bool full = p0 + PPT <= n;
        const float4* gp = gout + row0 * LPR + lane;
        const int* ip = idx + row0 + sub;
        float4 b[U];
        int dst[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = full || p0 + u / UPP < n;
            dst[u] = ok ? __ldg(ip + u * SPAR) : -1;
            b[u] = ok ? ldcs4(gp + u * 32) : zero4();
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (dst[u] >= 0) red_add4(g2 + (int64_t)dst[u] * LPR + cl, scale4(b[u], -1.f));
        float4 acc[PPT];
#pragma unroll
        for (int q = 0; q < PPT; q++) acc[q] = zero4();
#pragma unroll
        for (int u = 0; u < U; u++) acc[u / UPP] = add4(acc[u / UPP], b[u]);
#pragma unroll
        for (int q = 0; q < PPT; q++) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) acc[q] = add4(acc[q], xor4(acc[q], o));
            if (sub == q % SPAR && (full || p0 + q < n)) g1[(p0 + q) * LPR + cl] = acc[q];
        }
    }
}

// g1[n, :] = sum_s gout[n, s, :]               (subtraction backward w.r.t. input1)
template <int LPR, int NS>
__global__ void __launch_bounds__(FAST_THREADS)
reduce_neighbours_fast(int64_t n, const float4* __restrict__ gout, float4* __restrict__ g1) {
    constexpr int SPAR = 32 / LPR;
    constexpr int U = (NS / SPAR > 8) ? NS / SPAR : 8;
    constexpr int R = U * SPAR, PPT = R / NS, UPP = U / PPT;
    static_assert(NS % SPAR == 0 && R % NS == 0, "tile must hold whole points");
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (n + PPT - 1) / PPT;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t p0 = tile * PPT;
        const bool full = p0 + PPT <= n;
        const float4* gp = gout + p0 * NS * LPR + lane;
        float4 b[U];
#pragma unroll
        for (int u = 0; u < U; u++) b[u] = (full || p0 + u / UPP < n) ? ldcs4(gp + u * 32) : zero4();
        float4 acc[PPT];
#pragma unroll
        for (int q = 0; q < PPT; q++) acc[q] = zero4();
#pragma unroll
        for (int u = 0; u < U; u++) acc[u / UPP] = add4(acc[u / UPP], b[u]);
#pragma unroll
        for (int q = 0; q < PPT; q++) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) acc[q] = add4(acc[q], xor4(acc[q], o));
            if (sub == q % SPAR && (full || p0 + q < n)) g1[(p0 + q) * LPR + cl] = acc[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// aggregation backward: grad_in (vector atomics), grad_pos (streamed), grad_w (shuffle reduction)
template <int LPR, int NS>
__global__ void __launch_bounds__(FAST_THREADS)
aggregation_bwd_fast(int64_t n, int wvec, const float4* __restrict__ in, const float4* __restrict__ pos,
                     const float4* __restrict__ w, const int* __restrict__ idx, const float4* __restrict__ gout,
                     float4* __restrict__ gin, float4* __restrict__ gpos, float4* __restrict__ gw) {
    constexpr int SPAR = 32 / LPR;
    constexpr int U = (NS / SPAR > 4) ? NS / SPAR : 4;
    constexpr int R = U * SPAR, PPT = R / NS, UPP = U / PPT;
    static_assert(NS % SPAR == 0 && R % NS == 0, "tile must hold whole points");
    const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
    const int wv = cl & (wvec - 1);
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;
    const int64_t ntiles = (n + PPT - 1) / PPT;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
        const int64_t p0 = tile * PPT, row0 = p0 * NS;
        const bool full = p0 + PPT <= n;
        float4 g[PPT];
#pragma unroll
        for (int q = 0; q < PPT; q++) g[q] = (full || p0 + q < n) ? ldg4(gout + (p0 + q) * LPR + cl) : zero4();
        int src[U];
        float4 a[U], b[U], ww[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = full || p0 + u / UPP < n;
            src[u] = ok ? __ldg(idx + row0 + u * SPAR + sub) : -1;
            b[u] = ok ? ldcs4(pos + row0 * LPR + u * 32 + lane) : zero4();
            ww[u] = ok ? ldg4(w + (row0 + u * SPAR + sub) * wvec + wv) : zero4();
        }
#pragma unroll
        for (int u = 0; u < U; u++) a[u] = src[u] >= 0 ? ldg4(in + (int64_t)src[u] * LPR + cl) : zero4();
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = full || p0 + u / UPP < n;
            const float4 gq = g[u / UPP];
            const float4 gwv = mul4(gq, ww[u]);
            if (ok) stcs4(gpos + row0 * LPR + u * 32 + lane, gwv);
            if (src[u] >= 0) red_add4(gin + (int64_t)src[u] * LPR + cl, gwv);
            float4 part = mul4(gq, add4(a[u], b[u]));
            // lanes of one row whose channel vectors agree modulo wvec share a weight vector
            for (int o = wvec; o < LPR; o <<= 1) part = add4(part, xor4(part, o));
            if (ok && cl < wvec) gw[(row0 + u * SPAR + sub) * wvec + cl] = part;
        }
    }
}


// =============================================================================================
// Bulk-async (TMA) pipelined variants for the kernels whose dominant operand is a contiguous
// stream (position / weight / index tiles of aggregation).  The compile-time-tiled kernels above
// are latency-bound (ncu: 18-25 long-scoreboard stalls per issue, DRAM at 36-50 %): a warp loads
// a tile, waits, gathers, waits, stores, and only occupancy overlaps those waits.  Here every
// warp runs its own 3-stage ring in shared memory: lane 0 issues cp.async.bulk copies of the next
// tiles' contiguous spans (one instruction per span, completion counted by an mbarrier), so
// ~13 KB per warp / ~200 KB per SM are in flight while the warp gathers and reduces the current
// tile.  Persistent grid: 2 CTAs per SM.
namespace pipe {

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int STAGES = 3;

}  // namespace pipe

// out[n, :] = sum_s (in[idx[n,s], :] + pos[n,s,:]) * w[n,s, : % w_c]  over `ntiles` FULL tiles
template <int LPR, int NS>
__global__ void __launch_bounds__(FAST_THREADS)
aggregation_fwd_pipe(int64_t ntiles, int wvec, const float4* __restrict__ in, const float4* __restrict__ pos,
                     const float4* __restrict__ w, const int* __restrict__ idx, float4* __restrict__ out) {
    using namespace pipe;
    constexpr int SPAR = 32 / LPR;
    constexpr int U = (NS / SPAR > 8) ? NS / SPAR : 8;
    constexpr int R = U * SPAR, PPT = R / NS, UPP = U / PPT;
    static_assert(NS % SPAR == 0 && R % NS == 0, "tile must hold whole points");
    constexpr unsigned PB = R * LPR * 16, IB = R * 4;
    extern __shared__ __align__(128) unsigned char s_raw[];
    const unsigned WB = (unsigned)(R * wvec * 16);
    const unsigned stage_bytes = PB + WB + ((IB + 15u) & ~15u);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, sub = lane / LPR, cl = lane % LPR;
    const int wv = cl & (wvec - 1);
    unsigned char* my = s_raw + (size_t)wid * (STAGES * stage_bytes + 64);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(my + STAGES * stage_bytes);
    const int64_t warp = ((int64_t)blockIdx.x * FAST_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * FAST_THREADS) >> 5;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) bar_init(s32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int s, int64_t tile) {  // lane 0 only
        const int64_t row0 = tile * R;
        unsigned char* st = my + s * stage_bytes;
        const unsigned bar = s32(bars + s);
        bar_expect(bar, PB + WB + IB);
        bulk_g2s(s32(st), pos + row0 * LPR, PB, bar);
        bulk_g2s(s32(st + PB), w + row0 * wvec, WB, bar);
        bulk_g2s(s32(st + PB + WB), idx + row0, IB, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int64_t t = warp + (int64_t)s * nwarps;
            if (t < ntiles) issue(s, t);
        }
    }
    // k-th tile of this warp lives in stage k % STAGES, completed phase parity (k / STAGES) & 1.
    // The gathers of tile k+1 are issued before tile k is reduced, so their L2/DRAM latency hides
    // behind the arithmetic and the stage hand-over of tile k.
    auto gather = [&](int k, float4 (&a)[U]) {
        const int s = k % STAGES;
        bar_wait(s32(bars + s), (unsigned)(k / STAGES) & 1u);
        const int* si = reinterpret_cast<const int*>(my + s * stage_bytes + PB + WB);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int src = si[u * SPAR + sub];
            a[u] = src >= 0 ? ldg4(in + (int64_t)src * LPR + cl) : zero4();
        }
    };
    float4 a_cur[U], a_next[U];
    if (warp < ntiles) gather(0, a_cur);
    int k = 0;
    for (int64_t tile = warp; tile < ntiles; tile += nwarps, k++) {
        const int s = k % STAGES;
        const bool more = tile + nwarps < ntiles;
        if (more) gather(k + 1, a_next);
        const unsigned char* st = my + s * stage_bytes;
        const float4* sp = reinterpret_cast<const float4*>(st);
        const float4* sw = reinterpret_cast<const float4*>(st + PB);
        float4 acc[PPT];
#pragma unroll
        for (int q = 0; q < PPT; q++) acc[q] = zero4();
#pragma unroll
        for (int u = 0; u < U; u++) {
            const float4 b = sp[u * 32 + lane];
            const float4 ww = sw[(u * SPAR + sub) * wvec + wv];
            acc[u / UPP] = fma4(add4(a_cur[u], b), ww, acc[u / UPP]);
        }
        __syncwarp();  // everyone is done reading this stage
        if (lane == 0) {
            const int64_t nt = tile + (int64_t)STAGES * nwarps;
            if (nt < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(s, nt);
            }
        }
        const int64_t p0 = tile * PPT;
#pragma unroll
        for (int q = 0; q < PPT; q++) {
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) acc[q] = add4(acc[q], xor4(acc[q], o));
            if (sub == q % SPAR) out[(p0 + q) * LPR + cl] = acc[q];
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < U; u++) a_cur[u] = a_next[u];
        }
    }
}

}  // namespace pob
