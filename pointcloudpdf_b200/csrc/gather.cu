// gather.cu -- the HBM-bound gather / scatter-add operators of the PTv1 path (sm_100a).
//
// Replaces the one-thread-per-scalar kernels of libs/pointops/src/{grouping,subtraction,
// aggregation,interpolation}/*_cuda_kernel.cu (3 integer div/mod per element, idx re-read per
// channel, 4-byte accesses, read-modify-write of the output in global memory, two scalar
// atomics per element in backward) by row-vector kernels:
//   * channels are moved as 128-bit vectors (float4 / red.global.add.v4.f32);
//   * a warp owns whole neighbour rows: LPR = min(C/4, 32) lanes cover one row, 32/LPR rows are
//     in flight per warp, so each 128-byte line of a gathered row is fetched by one request;
//   * reductions over the neighbour axis (aggregation fwd, subtraction bwd) and over the
//     share_planes axis (aggregation bwd, grad_weight) happen in registers + shuffles: no
//     atomics and no output read-modify-write; only true scatters (grad of gathered rows) use
//     vector atomics;
//   * all index arithmetic is int64 (quirk C2: the reference overflows int above 2^31 elements).
// The fast paths need C % 4 == 0 with C/4 a power of two (or a multiple of 32) -- every PTv1
// width (32..512) qualifies; other shapes take the scalar kernels at the bottom, which are
// still CUDA: there is no CPU fallback anywhere.
#include <cstdlib>
#include "common.cuh"
#include "gather_fast.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace pob {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4_stream(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red4(float* p, float4 v) {
    atomicAdd(reinterpret_cast<float4*>(p), v);  // result unused -> RED.E.ADD.F32x4 (sm_90+)
}
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_neg(float4 a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
__device__ __forceinline__ float4 f4_shfl_xor(float4 v, int o) {
    return make_float4(__shfl_xor_sync(FULL, v.x, o), __shfl_xor_sync(FULL, v.y, o), __shfl_xor_sync(FULL, v.z, o),
                       __shfl_xor_sync(FULL, v.w, o));
}

// how a warp tiles rows of `cvec` float4: lpr lanes per row, spar rows side by side, chunks per lane
struct RowTile {
    int cvec, lpr, lpr_shift, spar, chunks;
    bool ok;
};
static inline RowTile row_tile(int c) {
    RowTile t = {};
    t.ok = false;
    if (c <= 0 || c % 4) return t;
    t.cvec = c / 4;
    if (t.cvec <= 32) {
        if (t.cvec & (t.cvec - 1)) return t;
        t.lpr = t.cvec;
        t.chunks = 1;
    } else {
        if (t.cvec % 32) return t;
        t.lpr = 32;
        t.chunks = t.cvec / 32;
    }
    t.spar = 32 / t.lpr;
    t.lpr_shift = 0;
    while ((1 << t.lpr_shift) < t.lpr) t.lpr_shift++;
    t.ok = true;
    return t;
}

constexpr int GATHER_THREADS = 256;

// ----------------------------------------------------------------------- grouping --
// out[r, :] = in[idx[r], :]   r = m*ns + s      (grouping_cuda_kernel.cu:5-14)
// SUB: out[r, :] = in1[r / ns, :] - in2[idx[r], :]   (subtraction_cuda_kernel.cu:5-16)
template <bool SUB>
__global__ void __launch_bounds__(GATHER_THREADS)
gather_rows_kernel(int64_t rows, int ns, RowTile t, const float* __restrict__ in, const float* __restrict__ in1,
                   const int* __restrict__ idx, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int sub = lane >> t.lpr_shift, cl = lane & (t.lpr - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int U = 4;  // independent rows in flight per lane
    const int64_t step = (int64_t)t.spar * U;
    for (int64_t base = warp * step; base < rows; base += nwarps * step) {
        int src[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t r = base + u * t.spar + sub;
            src[u] = r < rows ? __ldg(idx + r) : -1;
        }
        for (int ch = 0; ch < t.chunks; ch++) {
            const int cv = ch * t.lpr + cl;
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t r = base + u * t.spar + sub;
                if (r < rows) {
                    v[u] = ld4(in + ((int64_t)src[u] * t.cvec + cv) * 4);
                    if (SUB) v[u] = f4_sub(ld4(in1 + ((r / ns) * t.cvec + cv) * 4), v[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t r = base + u * t.spar + sub;
                if (r < rows) st4_stream(out + (r * t.cvec + cv) * 4, v[u]);
            }
        }
    }
}

// grad_in[idx[r], :] += sign * grad_out[r, :]    (grouping_cuda_kernel.cu:16-25,
// subtraction_cuda_kernel.cu:28-29 with sign = -1)
__global__ void __launch_bounds__(GATHER_THREADS)
scatter_rows_kernel(int64_t rows, RowTile t, float sign, const float* __restrict__ gout, const int* __restrict__ idx,
                    float* __restrict__ gin) {
    const int lane = threadIdx.x & 31;
    const int sub = lane >> t.lpr_shift, cl = lane & (t.lpr - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int U = 4;
    const int64_t step = (int64_t)t.spar * U;
    for (int64_t base = warp * step; base < rows; base += nwarps * step) {
        for (int ch = 0; ch < t.chunks; ch++) {
            const int cv = ch * t.lpr + cl;
            float4 v[U];
            int dst[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t r = base + u * t.spar + sub;
                dst[u] = -1;
                if (r < rows) { dst[u] = __ldg(idx + r); v[u] = ld4_stream(gout + (r * t.cvec + cv) * 4); }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (dst[u] >= 0) {
                    float4 g = v[u];
                    if (sign < 0.f) g = f4_neg(g);
                    red4(gin + ((int64_t)dst[u] * t.cvec + cv) * 4, g);
                }
            }
        }
    }
}

// g1[n, :] = sum_s grad_out[n, s, :]   (subtraction_cuda_kernel.cu:28, without atomics)
__global__ void __launch_bounds__(GATHER_THREADS)
reduce_neighbours_kernel(int64_t n, int ns, RowTile t, const float* __restrict__ gout, float* __restrict__ g1) {
    const int lane = threadIdx.x & 31;
    const int sub = lane >> t.lpr_shift, cl = lane & (t.lpr - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n; p += nwarps) {
        for (int ch = 0; ch < t.chunks; ch++) {
            const int cv = ch * t.lpr + cl;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = sub; s < ns; s += t.spar) acc = f4_add(acc, ld4_stream(gout + ((p * ns + s) * t.cvec + cv) * 4));
            for (int o = t.lpr; o < 32; o <<= 1) acc = f4_add(acc, f4_shfl_xor(acc, o));
            if (sub == 0) st4(g1 + (p * t.cvec + cv) * 4, acc);
        }
    }
}

// -------------------------------------------------------------------- aggregation --
// out[n, c] = sum_s (in[idx[n,s], c] + pos[n,s,c]) * w[n,s,c % w_c]
// (aggregation_cuda_kernel.cu:5-20).  One warp per point; spar neighbour rows in flight, the
// partial sums meet in registers through shuffles; the output is written once.
__global__ void __launch_bounds__(GATHER_THREADS)
aggregation_fwd_kernel(int64_t n, int ns, RowTile t, int wvec, const float* __restrict__ in,
                       const float* __restrict__ pos, const float* __restrict__ w, const int* __restrict__ idx,
                       float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int sub = lane >> t.lpr_shift, cl = lane & (t.lpr - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n; p += nwarps) {
        for (int ch = 0; ch < t.chunks; ch++) {
            const int cv = ch * t.lpr + cl;
            const int wv = cv & (wvec - 1);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int s = sub; s < ns; s += t.spar) {
                const int64_t r = p * ns + s;
                const int src = __ldg(idx + r);
                // idx < 0 (kNN placeholder) contributes a zero row, like pointops.grouping does
                const float4 a = src >= 0 ? ld4(in + ((int64_t)src * t.cvec + cv) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 b = ld4_stream(pos + (r * t.cvec + cv) * 4);
                const float4 ww = ld4(w + (r * wvec + wv) * 4);
                acc = f4_fma(f4_add(a, b), ww, acc);
            }
            for (int o = t.lpr; o < 32; o <<= 1) acc = f4_add(acc, f4_shfl_xor(acc, o));
            if (sub == 0) st4(out + (p * t.cvec + cv) * 4, acc);
        }
    }
}

// grad_in[idx[n,s], c] += g[n,c] * w[n,s,c%w_c]        (atomic scatter, vector red)
// grad_pos[n,s,c]       = g[n,c] * w[n,s,c%w_c]        (streamed store)
// grad_w[n,s,j]         = sum_{c%w_c==j} g[n,c] * (in[idx[n,s],c] + pos[n,s,c])   (shuffles)
// (aggregation_cuda_kernel.cu:22-39)
__global__ void __launch_bounds__(GATHER_THREADS)
aggregation_bwd_kernel(int64_t n, int ns, RowTile t, int wvec, const float* __restrict__ in,
                       const float* __restrict__ pos, const float* __restrict__ w, const int* __restrict__ idx,
                       const float* __restrict__ gout, float* __restrict__ gin, float* __restrict__ gpos,
                       float* __restrict__ gw) {
    const int lane = threadIdx.x & 31;
    const int sub = lane >> t.lpr_shift, cl = lane & (t.lpr - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n; p += nwarps) {
        // all lanes take the same number of trips so the shuffles below stay convergent
        for (int s0 = 0; s0 < ns; s0 += t.spar) {
            const int s = s0 + sub;
            const bool live = s < ns;
            const int64_t r = p * ns + (live ? s : 0);
            const int src = live ? __ldg(idx + r) : 0;
            float4 gws = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int ch = 0; ch < t.chunks; ch++) {
                const int cv = ch * t.lpr + cl;
                const int wv = cv & (wvec - 1);
                if (live) {
                    const float4 g = ld4(gout + (p * t.cvec + cv) * 4);
                    const float4 ww = ld4(w + (r * wvec + wv) * 4);
                    const float4 a = src >= 0 ? ld4(in + ((int64_t)src * t.cvec + cv) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 b = ld4_stream(pos + (r * t.cvec + cv) * 4);
                    const float4 gwv = f4_mul(g, ww);
                    st4_stream(gpos + (r * t.cvec + cv) * 4, gwv);
                    if (src >= 0) red4(gin + ((int64_t)src * t.cvec + cv) * 4, gwv);
                    gws = f4_fma(g, f4_add(a, b), gws);
                }
            }
            // lanes of one row whose cv agree modulo wvec share a weight vector
            for (int o = wvec; o < t.lpr; o <<= 1) gws = f4_add(gws, f4_shfl_xor(gws, o));
            if (live && cl < wvec) st4(gw + (r * wvec + cl) * 4, gws);
        }
    }
}

// ------------------------------------------------------------------ interpolation --
// out[n, c] = sum_{i<k} in[idx[n,i], c] * w[n,i]     (interpolation_cuda_kernel.cu:5-18)
__global__ void __launch_bounds__(GATHER_THREADS)
interpolation_fwd_kernel(int64_t n, int k, RowTile t, const float* __restrict__ in, const int* __restrict__ idx,
                         const float* __restrict__ w, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int sub = lane >> t.lpr_shift, cl = lane & (t.lpr - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * t.spar; base < n; base += nwarps * t.spar) {
        const int64_t p = base + sub;
        if (p >= n) continue;
        for (int ch = 0; ch < t.chunks; ch++) {
            const int cv = ch * t.lpr + cl;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < k; i++) {
                const int src = __ldg(idx + p * k + i);
                const float wi = __ldg(w + p * k + i);
                const float4 a = ld4(in + ((int64_t)src * t.cvec + cv) * 4);
                acc = f4_fma(a, make_float4(wi, wi, wi, wi), acc);
            }
            st4_stream(out + (p * t.cvec + cv) * 4, acc);
        }
    }
}

// grad_in[idx[n,i], c] += g[n,c] * w[n,i]            (interpolation_cuda_kernel.cu:20-33)
__global__ void __launch_bounds__(GATHER_THREADS)
interpolation_bwd_kernel(int64_t n, int k, RowTile t, const float* __restrict__ gout, const int* __restrict__ idx,
                         const float* __restrict__ w, float* __restrict__ gin) {
    const int lane = threadIdx.x & 31;
    const int sub = lane >> t.lpr_shift, cl = lane & (t.lpr - 1);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * t.spar; base < n; base += nwarps * t.spar) {
        const int64_t p = base + sub;
        if (p >= n) continue;
        for (int ch = 0; ch < t.chunks; ch++) {
            const int cv = ch * t.lpr + cl;
            const float4 g = ld4_stream(gout + (p * t.cvec + cv) * 4);
            for (int i = 0; i < k; i++) {
                const int dst = __ldg(idx + p * k + i);
                const float wi = __ldg(w + p * k + i);
                red4(gin + ((int64_t)dst * t.cvec + cv) * 4, make_float4(g.x * wi, g.y * wi, g.z * wi, g.w * wi));
            }
        }
    }
}

// ------------------------------------------------ fused group-with-xyz (a3 semantics) --
// out[m, s, :] = cat( (xyz[idx[m,s]] - new_xyz[m]) * [idx>=0] , feat[idx[m,s]] * [idx>=0] )
// (functions/grouping.py:36-60: five torch passes there, one here).  feat may be f32/f16/bf16;
// the output is f32 (the reference's cat with an f32 zero row promotes).  One warp per query m
// walks the contiguous ns*(3+C) output floats, so stores are fully coalesced.
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(GATHER_THREADS)
group_xyz_fwd_kernel(int64_t m, int ns, int c, int with_xyz, const T* __restrict__ feat, const float* __restrict__ xyz,
                     const float* __restrict__ new_xyz, const int* __restrict__ idx, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int x = with_xyz ? 3 : 0;
    const int W = x + c;
    const int per = ns * W;
    for (int64_t q = warp; q < m; q += nwarps) {
        float nq[3] = {0.f, 0.f, 0.f};
        if (with_xyz) { nq[0] = __ldg(new_xyz + q * 3); nq[1] = __ldg(new_xyz + q * 3 + 1); nq[2] = __ldg(new_xyz + q * 3 + 2); }
        float* o = out + q * per;
        const int* iq = idx + q * ns;
        int s = lane / W, chn = lane % W;
        for (int f = lane; f < per; f += 32) {
            const int src = __ldg(iq + s);
            float v = 0.f;
            if (src >= 0) {
                if (chn < x) v = __fsub_rn(__ldg(xyz + (int64_t)src * 3 + chn), chn == 0 ? nq[0] : (chn == 1 ? nq[1] : nq[2]));
                else v = to_f32<T>(feat[(int64_t)src * c + (chn - x)]);
            }
            __stcs(o + f, v);
            chn += 32;
            while (chn >= W) { chn -= W; s++; }
        }
    }
}

// grad_feat[idx[m,s], c] += grad_out[m, s, coff + c]  for idx >= 0   (autograd of the torch
// indexing in functions/grouping.py:43-45; xyz gets no gradient)
__global__ void __launch_bounds__(GATHER_THREADS)
group_xyz_bwd_kernel(int64_t m, int ns, int c, int coff, const float* __restrict__ gout, const int* __restrict__ idx,
                     float* __restrict__ gfeat) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int W = coff + c;
    const int per = ns * W;
    for (int64_t q = warp; q < m; q += nwarps) {
        const float* g = gout + q * per;
        const int* iq = idx + q * ns;
        int s = lane / W, chn = lane % W;
        for (int f = lane; f < per; f += 32) {
            const int dst = __ldg(iq + s);
            if (dst >= 0 && chn >= coff) atomicAdd(gfeat + (int64_t)dst * c + (chn - coff), __ldcs(g + f));
            chn += 32;
            while (chn >= W) { chn -= W; s++; }
        }
    }
}

// out[m, s, :] = (xyz[idx[m,s]] - new_xyz[m]) * [idx >= 0]   (the coordinate half of
// pointops.grouping(with_xyz=True), functions/grouping.py:49-57, on its own)
__global__ void __launch_bounds__(256)
group_relxyz_kernel(int64_t rows, int ns, int ns_shift, const float* __restrict__ xyz,
                    const float* __restrict__ new_xyz, const int* __restrict__ idx, float* __restrict__ out) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int src = __ldg(idx + r);
        const int64_t q = ns_shift >= 0 ? (r >> ns_shift) : r / ns;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (src >= 0) {
            dx = __fsub_rn(__ldg(xyz + (int64_t)src * 3), __ldg(new_xyz + q * 3));
            dy = __fsub_rn(__ldg(xyz + (int64_t)src * 3 + 1), __ldg(new_xyz + q * 3 + 1));
            dz = __fsub_rn(__ldg(xyz + (int64_t)src * 3 + 2), __ldg(new_xyz + q * 3 + 2));
        }
        out[r * 3] = dx; out[r * 3 + 1] = dy; out[r * 3 + 2] = dz;
    }
}

// ------------------------------------------------------- scalar kernels (any shape) --
__global__ void gather_rows_scalar_kernel(int64_t total, int ns, int c, int sub, const float* __restrict__ in,
                                          const float* __restrict__ in1, const int* __restrict__ idx,
                                          float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / c;
        const int ch = (int)(e - r * c);
        float v = __ldg(in + (int64_t)__ldg(idx + r) * c + ch);
        if (sub) v = __ldg(in1 + (r / ns) * c + ch) - v;
        out[e] = v;
    }
}

__global__ void scatter_rows_scalar_kernel(int64_t total, int c, float sign, const float* __restrict__ gout,
                                           const int* __restrict__ idx, float* __restrict__ gin) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / c;
        const int ch = (int)(e - r * c);
        atomicAdd(gin + (int64_t)__ldg(idx + r) * c + ch, sign * gout[e]);
    }
}

__global__ void reduce_neighbours_scalar_kernel(int64_t n, int ns, int c, const float* __restrict__ gout,
                                                float* __restrict__ g1) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * c; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / c;
        const int ch = (int)(e - p * c);
        float acc = 0.f;
        for (int s = 0; s < ns; s++) acc += gout[(p * ns + s) * c + ch];
        g1[e] = acc;
    }
}

__global__ void aggregation_fwd_scalar_kernel(int64_t n, int ns, int c, int w_c, const float* __restrict__ in,
                                              const float* __restrict__ pos, const float* __restrict__ w,
                                              const int* __restrict__ idx, float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * c; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / c;
        const int ch = (int)(e - p * c);
        float acc = 0.f;
        for (int s = 0; s < ns; s++) {
            const int64_t r = p * ns + s;
            const int src = __ldg(idx + r);
            const float a = src >= 0 ? __ldg(in + (int64_t)src * c + ch) : 0.f;
            acc = fmaf(a + pos[r * c + ch], __ldg(w + r * w_c + ch % w_c), acc);
        }
        out[e] = acc;
    }
}

// grad_weight must be zero on entry for this path (atomics over the share_planes channels)
__global__ void aggregation_bwd_scalar_kernel(int64_t n, int ns, int c, int w_c, const float* __restrict__ in,
                                              const float* __restrict__ pos, const float* __restrict__ w,
                                              const int* __restrict__ idx, const float* __restrict__ gout,
                                              float* __restrict__ gin, float* __restrict__ gpos,
                                              float* __restrict__ gw) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * c; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / c;
        const int ch = (int)(e - p * c);
        const float g = gout[e];
        for (int s = 0; s < ns; s++) {
            const int64_t r = p * ns + s;
            const int src = __ldg(idx + r);
            const int64_t ii = (int64_t)src * c + ch;
            const float wt = __ldg(w + r * w_c + ch % w_c);
            if (src >= 0) atomicAdd(gin + ii, g * wt);
            gpos[r * c + ch] = g * wt;
            atomicAdd(gw + r * w_c + ch % w_c, g * ((src >= 0 ? __ldg(in + ii) : 0.f) + pos[r * c + ch]));
        }
    }
}

__global__ void interpolation_fwd_scalar_kernel(int64_t n, int c, int k, const float* __restrict__ in,
                                                const int* __restrict__ idx, const float* __restrict__ w,
                                                float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * c; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / c;
        const int ch = (int)(e - p * c);
        float acc = 0.f;
        for (int i = 0; i < k; i++) acc = fmaf(__ldg(in + (int64_t)__ldg(idx + p * k + i) * c + ch), __ldg(w + p * k + i), acc);
        out[e] = acc;
    }
}

__global__ void interpolation_bwd_scalar_kernel(int64_t n, int c, int k, const float* __restrict__ gout,
                                                const int* __restrict__ idx, const float* __restrict__ w,
                                                float* __restrict__ gin) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * c; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / c;
        const int ch = (int)(e - p * c);
        for (int i = 0; i < k; i++) atomicAdd(gin + (int64_t)__ldg(idx + p * k + i) * c + ch, gout[e] * __ldg(w + p * k + i));
    }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// warps needed -> CTAs, capped at a few waves of the machine
static inline unsigned warp_grid(int64_t warps_needed) {
    return grid_for(warps_needed * 32, GATHER_THREADS, 8, 4);
}


// ------------------------------------------------------------ fast-path dispatch ----------
// lanes per row of the compile-time-tiled kernels: C = 4 * LPR, LPR a power of two <= 32
static inline int lpr_of(int c) {
    if (c <= 0 || c % 4) return 0;
    const int v = c / 4;
    return (v <= 32 && (v & (v - 1)) == 0) ? v : 0;
}
static inline int shift_of(int v) {  // log2 for powers of two, else -1
    if (v <= 0 || (v & (v - 1))) return -1;
    int s = 0;
    while ((1 << s) < v) s++;
    return s;
}
static const bool USE_PIPE = getenv("POINTOPS_B200_NO_PIPE") == nullptr;

static inline unsigned fast_grid(int64_t warps_needed) {
    const int64_t ctas = ceil_div(warps_needed, FAST_THREADS / 32);
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}
#define POB_LPR_SWITCH(lpr, M) \
    switch (lpr) { case 1: M(1); break; case 2: M(2); break; case 4: M(4); break; case 8: M(8); break; \
                   case 16: M(16); break; case 32: M(32); break; default: break; }

template <typename T, bool SUB, bool MASKED>
static bool launch_gather_fast(int lpr, int64_t rows, int ns, const T* in, const float* in1, const int* idx, float* out,
                               cudaStream_t stream) {
    if (!lpr) return false;
#define M(L) { constexpr int U = (L >= 8) ? 8 : 4; constexpr int R = (32 / L) * U; \
        gather_rows_fast<T, L, U, SUB, MASKED><<<fast_grid(ceil_div(rows, R)), FAST_THREADS, 0, stream>>>( \
            rows, ns, shift_of(ns), in, (const float4*)in1, idx, (float4*)out); }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

static bool launch_scatter_fast(int lpr, int64_t rows, float sign, const float* gout, const int* idx, float* gin,
                                cudaStream_t stream) {
    if (!lpr) return false;
#define M(L) { constexpr int U = (L >= 8) ? 8 : 4; constexpr int R = (32 / L) * U; \
        scatter_rows_fast<L, U><<<fast_grid(ceil_div(rows, R)), FAST_THREADS, 0, stream>>>( \
            rows, sign, (const float4*)gout, idx, (float4*)gin); }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

// whole-point tiles exist when nsample is 8 or 16 and a multiple of the rows a warp holds side by side
static inline bool ns_tiled(int lpr, int ns) { return lpr && (ns == 8 || ns == 16) && ns % (32 / lpr) == 0; }

template <int L, int NS> static constexpr bool tile_ok() { return NS % (32 / L) == 0; }

static bool launch_reduce_fast(int lpr, int ns, int64_t n, const float* gout, float* g1, cudaStream_t stream) {
    if (!ns_tiled(lpr, ns)) return false;
#define M(L) { if (ns == 8) { if constexpr (tile_ok<L, 8>()) { constexpr int U = ((8 / (32 / L)) > 8) ? 8 / (32 / L) : 8; \
            reduce_neighbours_fast<L, 8><<<fast_grid(ceil_div(n, (U * (32 / L)) / 8)), FAST_THREADS, 0, stream>>>(n, (const float4*)gout, (float4*)g1); } } \
        else { if constexpr (tile_ok<L, 16>()) { constexpr int U = ((16 / (32 / L)) > 8) ? 16 / (32 / L) : 8; \
            reduce_neighbours_fast<L, 16><<<fast_grid(ceil_div(n, (U * (32 / L)) / 16)), FAST_THREADS, 0, stream>>>(n, (const float4*)gout, (float4*)g1); } } }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

static bool launch_sub_fwd_fast(int lpr, int ns, int64_t n, const float* in1, const float* in2, const int* idx, float* out,
                                cudaStream_t stream) {
    if (!ns_tiled(lpr, ns)) return false;
#define M(L) { if (ns == 8) { if constexpr (tile_ok<L, 8>()) { constexpr int U = ((8 / (32 / L)) > 8) ? 8 / (32 / L) : 8; \
            subtraction_fwd_fast<L, 8><<<fast_grid(ceil_div(n, (U * (32 / L)) / 8)), FAST_THREADS, 0, stream>>>(n, (const float4*)in1, (const float4*)in2, idx, (float4*)out); } } \
        else { if constexpr (tile_ok<L, 16>()) { constexpr int U = ((16 / (32 / L)) > 8) ? 16 / (32 / L) : 8; \
            subtraction_fwd_fast<L, 16><<<fast_grid(ceil_div(n, (U * (32 / L)) / 16)), FAST_THREADS, 0, stream>>>(n, (const float4*)in1, (const float4*)in2, idx, (float4*)out); } } }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

static bool launch_sub_bwd_fast(int lpr, int ns, int64_t n, const float* gout, const int* idx, float* g1, float* g2,
                                cudaStream_t stream) {
    if (!ns_tiled(lpr, ns)) return false;
#define M(L) { if (ns == 8) { if constexpr (tile_ok<L, 8>()) { constexpr int U = ((8 / (32 / L)) > 8) ? 8 / (32 / L) : 8; \
            subtraction_bwd_fast<L, 8><<<fast_grid(ceil_div(n, (U * (32 / L)) / 8)), FAST_THREADS, 0, stream>>>(n, (const float4*)gout, idx, (float4*)g1, (float4*)g2); } } \
        else { if constexpr (tile_ok<L, 16>()) { constexpr int U = ((16 / (32 / L)) > 8) ? 16 / (32 / L) : 8; \
            subtraction_bwd_fast<L, 16><<<fast_grid(ceil_div(n, (U * (32 / L)) / 16)), FAST_THREADS, 0, stream>>>(n, (const float4*)gout, idx, (float4*)g1, (float4*)g2); } } }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

// pipelined (bulk-async) aggregation over the full tiles; returns the number of points it covered
template <int L, int NS>
static int64_t launch_agg_fwd_pipe(int64_t n, int wvec, const float* in, const float* pos, const float* w,
                                   const int* idx, float* out, cudaStream_t stream, int* err) {
    constexpr int SPAR = 32 / L;
    constexpr int U = (NS / SPAR > 8) ? NS / SPAR : 8;
    constexpr int R = U * SPAR, PPT = R / NS;
    const int64_t ntiles = n / PPT;
    const int64_t ctas = (int64_t)sm_count() * 2;
    if (ntiles < ctas * (FAST_THREADS / 32) * 3) return 0;  // too little work to fill the rings
    const size_t stage = (size_t)R * L * 16 + (size_t)R * wvec * 16 + (((size_t)R * 4 + 15) & ~(size_t)15);
    const size_t smem = (FAST_THREADS / 32) * (pipe::STAGES * stage + 64);
    if (smem > 115000) return 0;  // two CTAs per SM must fit in 227 KB
    auto kern = aggregation_fwd_pipe<L, NS>;
    {   // the opt-in is per DEVICE and cheap: set it on every launch (a process may drive several GPUs)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { *err = (int)e; return 0; }
    }
    kern<<<(unsigned)ctas, FAST_THREADS, smem, stream>>>(ntiles, wvec, (const float4*)in, (const float4*)pos,
                                                         (const float4*)w, idx, (float4*)out);
    pob_count_launches(1);
    return ntiles * PPT;
}

static bool launch_agg_fwd_fast(int lpr, int ns, int64_t n, int wvec, const float* in, const float* pos, const float* w,
                                const int* idx, float* out, cudaStream_t stream) {
    if (!ns_tiled(lpr, ns)) return false;
    if (USE_PIPE) {
        int err = 0;
        int64_t done = 0;
#define M(L) { if (ns == 8) { if constexpr (tile_ok<L, 8>()) done = launch_agg_fwd_pipe<L, 8>(n, wvec, in, pos, w, idx, out, stream, &err); } \
               else { if constexpr (tile_ok<L, 16>()) done = launch_agg_fwd_pipe<L, 16>(n, wvec, in, pos, w, idx, out, stream, &err); } }
        POB_LPR_SWITCH(lpr, M)
#undef M
        if (done >= n) return true;
        if (done > 0) {  // remainder (< one tile of points) through the plain tiled kernel
            in = in; pos += done * ns * (int64_t)lpr * 4; w += done * ns * (int64_t)wvec * 4; idx += done * ns;
            out += done * (int64_t)lpr * 4; n -= done;
        }
    }
#define M(L) { if (ns == 8) { if constexpr (tile_ok<L, 8>()) { constexpr int U = ((8 / (32 / L)) > 8) ? 8 / (32 / L) : 8; \
            aggregation_fwd_fast<L, 8><<<fast_grid(ceil_div(n, (U * (32 / L)) / 8)), FAST_THREADS, 0, stream>>>( \
                n, wvec, (const float4*)in, (const float4*)pos, (const float4*)w, idx, (float4*)out); } } \
        else { if constexpr (tile_ok<L, 16>()) { constexpr int U = ((16 / (32 / L)) > 8) ? 16 / (32 / L) : 8; \
            aggregation_fwd_fast<L, 16><<<fast_grid(ceil_div(n, (U * (32 / L)) / 16)), FAST_THREADS, 0, stream>>>( \
                n, wvec, (const float4*)in, (const float4*)pos, (const float4*)w, idx, (float4*)out); } } }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

static bool launch_agg_bwd_fast(int lpr, int ns, int64_t n, int wvec, const float* in, const float* pos, const float* w,
                                const int* idx, const float* gout, float* gin, float* gpos, float* gw,
                                cudaStream_t stream) {
    if (!ns_tiled(lpr, ns)) return false;
#define M(L) { if (ns == 8) { if constexpr (tile_ok<L, 8>()) { constexpr int U = ((8 / (32 / L)) > 4) ? 8 / (32 / L) : 4; \
            aggregation_bwd_fast<L, 8><<<fast_grid(ceil_div(n, (U * (32 / L)) / 8)), FAST_THREADS, 0, stream>>>( \
                n, wvec, (const float4*)in, (const float4*)pos, (const float4*)w, idx, (const float4*)gout, \
                (float4*)gin, (float4*)gpos, (float4*)gw); } } \
        else { if constexpr (tile_ok<L, 16>()) { constexpr int U = ((16 / (32 / L)) > 4) ? 16 / (32 / L) : 4; \
            aggregation_bwd_fast<L, 16><<<fast_grid(ceil_div(n, (U * (32 / L)) / 16)), FAST_THREADS, 0, stream>>>( \
                n, wvec, (const float4*)in, (const float4*)pos, (const float4*)w, idx, (const float4*)gout, \
                (float4*)gin, (float4*)gpos, (float4*)gw); } } }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

static bool launch_interp_fwd_fast(int lpr, int64_t n, int k, const float* in, const int* idx, const float* w, float* out,
                                   cudaStream_t stream) {
    if (!lpr) return false;
#define M(L) { constexpr int U = 4; constexpr int R = (32 / L) * U; \
        interpolation_fwd_fast<L, U><<<fast_grid(ceil_div(n, R)), FAST_THREADS, 0, stream>>>( \
            n, k, (const float4*)in, idx, w, (float4*)out); }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return true;
}

template <typename T>
static int launch_group_xyz_fwd_fast(int lpr, int64_t m, int ns, const T* feat, const float* xyz, const float* new_xyz,
                                     const int* idx, float* out, cudaStream_t stream) {
    const int per = ns * (4 * lpr + 3);
    const size_t smem = sizeof(float) * (size_t)((per + 3) & ~3) * (FAST_THREADS / 32);
    if (smem > 200 * 1024 || ns > 32) return -1;   // the kernel keeps one neighbour index per lane
    // passes per chunk: all of them for the PTv1 neighbourhood sizes (ns = 8 / 16), at most 8 gathers in flight per lane
#define K(L, CHV) { auto kern = group_xyz_fwd_fast<T, L, CHV>; \
        if (smem > 48 * 1024) POB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<fast_grid(m), FAST_THREADS, smem, stream>>>(m, ns, feat, xyz, new_xyz, idx, out); }
#define M(L) { constexpr int SP_ = 32 / L; const int need = (ns + SP_ - 1) / SP_; \
        if (need <= 1) K(L, 1) else if (need <= 2) K(L, 2) else if (need <= 4) K(L, 4) else K(L, 8) }
    POB_LPR_SWITCH(lpr, M)
#undef M
#undef K
    return 0;
}

static int launch_group_xyz_bwd_fast(int lpr, int64_t m, int ns, const float* gout, const int* idx, float* gfeat,
                                     cudaStream_t stream) {
    const int per = ns * (4 * lpr + 3);
    const size_t smem = sizeof(float) * (size_t)((per + 3) & ~3) * (FAST_THREADS / 32);
    if (smem > 200 * 1024) return -1;
#define M(L) { auto kern = group_xyz_bwd_fast<L>; \
        if (smem > 48 * 1024) POB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<fast_grid(m), FAST_THREADS, smem, stream>>>(m, ns, gout, idx, (float4*)gfeat); }
    POB_LPR_SWITCH(lpr, M)
#undef M
    return 0;
}

}  // namespace pob

using namespace pob;

// ===================================================================== C ABI ==

// grouping_forward_cuda_launcher(m, nsample, c, input, idx, output)  (grouping_cuda_kernel.h:14)
POB_API int pob_grouping_forward(int64_t m, int nsample, int c, const float* input, const int* idx, float* output,
                                 cudaStream_t stream) {
    if (m < 0 || nsample < 1 || c < 1) return POB_ERR_BAD_ARG;
    const int64_t rows = m * nsample;
    if (rows == 0) return 0;
    if (!input || !idx || !output) return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    if (aligned16(input) && aligned16(output) &&
        launch_gather_fast<float, false, false>(lpr_of(c), rows, nsample, input, nullptr, idx, output, stream)) {
    } else if (t.ok && aligned16(input) && aligned16(output)) {
        gather_rows_kernel<false><<<warp_grid(ceil_div(rows, t.spar * 4)), GATHER_THREADS, 0, stream>>>(
            rows, nsample, t, input, nullptr, idx, output);
    } else {
        gather_rows_scalar_kernel<<<grid_for(rows * c, 256, 8), 256, 0, stream>>>(rows * c, nsample, c, 0, input,
                                                                                  nullptr, idx, output);
    }
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// grouping_backward_cuda_launcher(m, nsample, c, grad_output, idx, grad_input)  (grouping_cuda_kernel.h:15)
// grad_input is accumulated into (the caller zeroes it, as functions/grouping.py:31 does).
POB_API int pob_grouping_backward(int64_t m, int nsample, int c, const float* grad_output, const int* idx,
                                  float* grad_input, cudaStream_t stream) {
    if (m < 0 || nsample < 1 || c < 1) return POB_ERR_BAD_ARG;
    const int64_t rows = m * nsample;
    if (rows == 0) return 0;
    if (!grad_output || !idx || !grad_input) return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    if (aligned16(grad_output) && aligned16(grad_input) &&
        launch_scatter_fast(lpr_of(c), rows, 1.f, grad_output, idx, grad_input, stream)) {
    } else if (t.ok && aligned16(grad_output) && aligned16(grad_input)) {
        scatter_rows_kernel<<<warp_grid(ceil_div(rows, t.spar * 4)), GATHER_THREADS, 0, stream>>>(
            rows, t, 1.f, grad_output, idx, grad_input);
    } else {
        scatter_rows_scalar_kernel<<<grid_for(rows * c, 256, 8), 256, 0, stream>>>(rows * c, c, 1.f, grad_output, idx,
                                                                                   grad_input);
    }
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// subtraction_forward_cuda_launcher(n, nsample, c, input1, input2, idx, output)  (subtraction_cuda_kernel.h:14)
POB_API int pob_subtraction_forward(int64_t n, int nsample, int c, const float* input1, const float* input2,
                                    const int* idx, float* output, cudaStream_t stream) {
    if (n < 0 || nsample < 1 || c < 1) return POB_ERR_BAD_ARG;
    const int64_t rows = n * nsample;
    if (rows == 0) return 0;
    if (!input1 || !input2 || !idx || !output) return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    if (aligned16(input1) && aligned16(input2) && aligned16(output) &&
        (launch_sub_fwd_fast(lpr_of(c), nsample, n, input1, input2, idx, output, stream) ||
         launch_gather_fast<float, true, false>(lpr_of(c), rows, nsample, input2, input1, idx, output, stream))) {
    } else if (t.ok && aligned16(input1) && aligned16(input2) && aligned16(output)) {
        gather_rows_kernel<true><<<warp_grid(ceil_div(rows, t.spar * 4)), GATHER_THREADS, 0, stream>>>(
            rows, nsample, t, input2, input1, idx, output);
    } else {
        gather_rows_scalar_kernel<<<grid_for(rows * c, 256, 8), 256, 0, stream>>>(rows * c, nsample, c, 1, input2,
                                                                                  input1, idx, output);
    }
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// subtraction_backward_cuda_launcher(n, nsample, c, idx, grad_output, grad_input1, grad_input2)
// (subtraction_cuda_kernel.h:15).  grad_input1 is overwritten; grad_input2 is accumulated into.
POB_API int pob_subtraction_backward(int64_t n, int nsample, int c, const int* idx, const float* grad_output,
                                     float* grad_input1, float* grad_input2, cudaStream_t stream) {
    if (n < 0 || nsample < 1 || c < 1) return POB_ERR_BAD_ARG;
    const int64_t rows = n * nsample;
    if (rows == 0) return 0;
    if (!idx || !grad_output || !grad_input1 || !grad_input2) return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    if (t.ok && aligned16(grad_output) && aligned16(grad_input1) && aligned16(grad_input2) &&
        launch_sub_bwd_fast(lpr_of(c), nsample, n, grad_output, idx, grad_input1, grad_input2, stream)) {
        pob_count_launches(1);
        POB_RETURN_LAST_ERROR();
    }
    if (t.ok && aligned16(grad_output) && aligned16(grad_input1) && aligned16(grad_input2)) {
        if (!launch_reduce_fast(lpr_of(c), nsample, n, grad_output, grad_input1, stream))
            reduce_neighbours_kernel<<<warp_grid(n), GATHER_THREADS, 0, stream>>>(n, nsample, t, grad_output, grad_input1);
        if (!launch_scatter_fast(lpr_of(c), rows, -1.f, grad_output, idx, grad_input2, stream))
            scatter_rows_kernel<<<warp_grid(ceil_div(rows, t.spar * 4)), GATHER_THREADS, 0, stream>>>(
                rows, t, -1.f, grad_output, idx, grad_input2);
    } else {
        reduce_neighbours_scalar_kernel<<<grid_for(n * c, 256, 8), 256, 0, stream>>>(n, nsample, c, grad_output,
                                                                                     grad_input1);
        scatter_rows_scalar_kernel<<<grid_for(rows * c, 256, 8), 256, 0, stream>>>(rows * c, c, -1.f, grad_output, idx,
                                                                                   grad_input2);
    }
    pob_count_launches(2);
    POB_RETURN_LAST_ERROR();
}

static inline bool agg_fast(const RowTile& t, int c, int w_c, int* wvec) {
    if (!t.ok || w_c < 4 || w_c % 4 || c % w_c) return false;
    const int v = w_c / 4;
    if (v & (v - 1)) return false;
    if (v > t.lpr) return false;
    *wvec = v;
    return true;
}

// aggregation_forward_cuda_launcher(n, nsample, c, w_c, input, position, weight, idx, output)
// (aggregation_cuda_kernel.h:14).  output is overwritten (the reference accumulates into a
// zeroed buffer, functions/aggregation.py:21).
POB_API int pob_aggregation_forward(int64_t n, int nsample, int c, int w_c, const float* input, const float* position,
                                    const float* weight, const int* idx, float* output, cudaStream_t stream) {
    if (n < 0 || nsample < 1 || c < 1 || w_c < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!input || !position || !weight || !idx || !output) return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    int wvec = 0;
    if (agg_fast(t, c, w_c, &wvec) && aligned16(input) && aligned16(position) && aligned16(weight) && aligned16(output)) {
        if (!launch_agg_fwd_fast(lpr_of(c), nsample, n, wvec, input, position, weight, idx, output, stream))
            aggregation_fwd_kernel<<<warp_grid(n), GATHER_THREADS, 0, stream>>>(n, nsample, t, wvec, input, position,
                                                                                weight, idx, output);
    } else {
        aggregation_fwd_scalar_kernel<<<grid_for(n * c, 256, 8), 256, 0, stream>>>(n, nsample, c, w_c, input, position,
                                                                                   weight, idx, output);
    }
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// aggregation_backward_cuda_launcher(...)  (aggregation_cuda_kernel.h:15).  grad_input is
// accumulated into (zeroed by the caller); grad_position and grad_weight are overwritten
// (grad_weight is zeroed here first on the scalar path, which still needs atomics).
POB_API int pob_aggregation_backward(int64_t n, int nsample, int c, int w_c, const float* input,
                                     const float* position, const float* weight, const int* idx,
                                     const float* grad_output, float* grad_input, float* grad_position,
                                     float* grad_weight, cudaStream_t stream) {
    if (n < 0 || nsample < 1 || c < 1 || w_c < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!input || !position || !weight || !idx || !grad_output || !grad_input || !grad_position || !grad_weight)
        return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    int wvec = 0;
    if (agg_fast(t, c, w_c, &wvec) && aligned16(input) && aligned16(position) && aligned16(weight) &&
        aligned16(grad_output) && aligned16(grad_input) && aligned16(grad_position) && aligned16(grad_weight)) {
        if (!launch_agg_bwd_fast(lpr_of(c), nsample, n, wvec, input, position, weight, idx, grad_output, grad_input,
                                 grad_position, grad_weight, stream))
            aggregation_bwd_kernel<<<warp_grid(n), GATHER_THREADS, 0, stream>>>(
                n, nsample, t, wvec, input, position, weight, idx, grad_output, grad_input, grad_position, grad_weight);
    } else {
        POB_CHECK(cudaMemsetAsync(grad_weight, 0, sizeof(float) * (size_t)n * nsample * w_c, stream));
        aggregation_bwd_scalar_kernel<<<grid_for(n * c, 256, 8), 256, 0, stream>>>(
            n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input, grad_position, grad_weight);
    }
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// interpolation_forward_cuda_launcher(n, c, k, input, idx, weight, output)  (interpolation_cuda_kernel.h:14)
// output is overwritten.
POB_API int pob_interpolation_forward(int64_t n, int c, int k, const float* input, const int* idx, const float* weight,
                                      float* output, cudaStream_t stream) {
    if (n < 0 || c < 1 || k < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!input || !idx || !weight || !output) return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    if (aligned16(input) && aligned16(output) && launch_interp_fwd_fast(lpr_of(c), n, k, input, idx, weight, output, stream)) {
    } else if (t.ok && aligned16(input) && aligned16(output)) {
        interpolation_fwd_kernel<<<warp_grid(ceil_div(n, t.spar)), GATHER_THREADS, 0, stream>>>(n, k, t, input, idx,
                                                                                                weight, output);
    } else {
        interpolation_fwd_scalar_kernel<<<grid_for(n * c, 256, 8), 256, 0, stream>>>(n, c, k, input, idx, weight, output);
    }
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// interpolation_backward_cuda_launcher(n, c, k, grad_output, idx, weight, grad_input)
// (interpolation_cuda_kernel.h:15).  grad_input is accumulated into.
POB_API int pob_interpolation_backward(int64_t n, int c, int k, const float* grad_output, const int* idx,
                                       const float* weight, float* grad_input, cudaStream_t stream) {
    if (n < 0 || c < 1 || k < 1) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!grad_output || !idx || !weight || !grad_input) return POB_ERR_BAD_ARG;
    const RowTile t = row_tile(c);
    if (t.ok && aligned16(grad_output) && aligned16(grad_input)) {
        interpolation_bwd_kernel<<<warp_grid(ceil_div(n, t.spar)), GATHER_THREADS, 0, stream>>>(n, k, t, grad_output,
                                                                                                idx, weight, grad_input);
    } else {
        interpolation_bwd_scalar_kernel<<<grid_for(n * c, 256, 8), 256, 0, stream>>>(n, c, k, grad_output, idx, weight,
                                                                                     grad_input);
    }
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// Fused replacement for pointops.grouping (functions/grouping.py:36-60).  feat_dtype: 0 f32,
// 1 f16, 2 bf16.  output is (m, nsample, 3 + c) f32 when with_xyz, else (m, nsample, c).
POB_API int pob_group_xyz_forward(int64_t m, int nsample, int c, int with_xyz, const void* feat, int feat_dtype,
                                  const float* xyz, const float* new_xyz, const int* idx, float* output,
                                  cudaStream_t stream) {
    if (m < 0 || nsample < 1 || c < 1 || feat_dtype < 0 || feat_dtype > 2) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!feat || !idx || !output || (with_xyz && (!xyz || !new_xyz))) return POB_ERR_BAD_ARG;
    const int lpr = lpr_of(c);
    const bool vec_ok = lpr && ((uintptr_t)feat & 15) == 0 && aligned16(output);
    if (vec_ok && with_xyz) {
        int rc = -1;
        if (feat_dtype == 0) rc = launch_group_xyz_fwd_fast<float>(lpr, m, nsample, (const float*)feat, xyz, new_xyz, idx, output, stream);
        else if (feat_dtype == 1) rc = launch_group_xyz_fwd_fast<__half>(lpr, m, nsample, (const __half*)feat, xyz, new_xyz, idx, output, stream);
        else rc = launch_group_xyz_fwd_fast<__nv_bfloat16>(lpr, m, nsample, (const __nv_bfloat16*)feat, xyz, new_xyz, idx, output, stream);
        if (rc > 0) return rc;
        if (rc == 0) { pob_count_launches(1); POB_RETURN_LAST_ERROR(); }
    } else if (vec_ok) {  // feature-only: a masked row gather
        const int64_t rows = m * nsample;
        bool done;
        if (feat_dtype == 0) done = launch_gather_fast<float, false, true>(lpr, rows, nsample, (const float*)feat, nullptr, idx, output, stream);
        else if (feat_dtype == 1) done = launch_gather_fast<__half, false, true>(lpr, rows, nsample, (const __half*)feat, nullptr, idx, output, stream);
        else done = launch_gather_fast<__nv_bfloat16, false, true>(lpr, rows, nsample, (const __nv_bfloat16*)feat, nullptr, idx, output, stream);
        if (done) { pob_count_launches(1); POB_RETURN_LAST_ERROR(); }
    }
    const unsigned grid = warp_grid(m);
    if (feat_dtype == 0)
        group_xyz_fwd_kernel<float><<<grid, GATHER_THREADS, 0, stream>>>(m, nsample, c, with_xyz, (const float*)feat, xyz,
                                                                        new_xyz, idx, output);
    else if (feat_dtype == 1)
        group_xyz_fwd_kernel<__half><<<grid, GATHER_THREADS, 0, stream>>>(m, nsample, c, with_xyz, (const __half*)feat,
                                                                         xyz, new_xyz, idx, output);
    else
        group_xyz_fwd_kernel<__nv_bfloat16><<<grid, GATHER_THREADS, 0, stream>>>(
            m, nsample, c, with_xyz, (const __nv_bfloat16*)feat, xyz, new_xyz, idx, output);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// Backward of the above w.r.t. feat: grad_feat (n, c) f32, accumulated into (caller zeroes).
// grad_output is (m, nsample, coff + c) with coff = 3 when the forward used with_xyz.
POB_API int pob_group_xyz_backward(int64_t m, int nsample, int c, int with_xyz, const float* grad_output,
                                   const int* idx, float* grad_feat, cudaStream_t stream) {
    if (m < 0 || nsample < 1 || c < 1) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!grad_output || !idx || !grad_feat) return POB_ERR_BAD_ARG;
    const int lpr = lpr_of(c);
    if (lpr && aligned16(grad_feat) && aligned16(grad_output)) {
        if (with_xyz) {
            const int rc = launch_group_xyz_bwd_fast(lpr, m, nsample, grad_output, idx, grad_feat, stream);
            if (rc > 0) return rc;
            if (rc == 0) { pob_count_launches(1); POB_RETURN_LAST_ERROR(); }
        } else if (launch_scatter_fast(lpr, m * nsample, 1.f, grad_output, idx, grad_feat, stream)) {
            pob_count_launches(1);
            POB_RETURN_LAST_ERROR();
        }
    }
    group_xyz_bwd_kernel<<<warp_grid(m), GATHER_THREADS, 0, stream>>>(m, nsample, c, with_xyz ? 3 : 0, grad_output, idx,
                                                                      grad_feat);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// Coordinate half of pointops.grouping(with_xyz=True) alone: output (m, nsample, 3) f32 =
// xyz[idx] - new_xyz[m], zero rows for idx < 0.  Lets a caller keep the gathered features in
// their own 16-byte-aligned tensor instead of the 3+C interleaved one.
POB_API int pob_group_relxyz_forward(int64_t m, int nsample, const float* xyz, const float* new_xyz, const int* idx,
                                     float* output, cudaStream_t stream) {
    if (m < 0 || nsample < 1) return POB_ERR_BAD_ARG;
    if (m == 0) return 0;
    if (!xyz || !new_xyz || !idx || !output) return POB_ERR_BAD_ARG;
    const int64_t rows = m * nsample;
    group_relxyz_kernel<<<grid_for(rows, 256, 8), 256, 0, stream>>>(rows, nsample, shift_of(nsample), xyz, new_xyz, idx,
                                                                   output);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}
