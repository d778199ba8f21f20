// datapath.cu -- device pieces of the data path either side of the model.
//
//   pob_grid_hash     voxel coordinates + FNV key of GridSample (pointcept/datasets/transform.py:813-823, 911-925):
//                     grid_coord = floor(coord / grid_size) - min, key = FNV64 over the three coordinates exactly as
//                     fnv_hash_vec spells it (offset basis 14695981039346656037, per coordinate: multiply by the
//                     prime 1099511628211, then xor), in wrapping unsigned 64-bit arithmetic.  The key is returned
//                     with its sign bit flipped so that a SIGNED 64-bit sort (torch.sort) orders it like numpy's
//                     unsigned argsort does.
//   pob_scatter_mean  out[index[r], :] = mean of the rows r that map there, 0 where none do: torch_scatter.scatter_mean
//                     as the tester uses it to average fragment scores (pointcept/engines/test.py:243-248).
// Both are HBM-bound integer / byte work: one coalesced pass, 128-bit accesses where rows allow it.
#include "common.cuh"

namespace pob {

__global__ void __launch_bounds__(256)
grid_hash_kernel(int64_t n, const float* __restrict__ coord, double gx, double gy, double gz, const long long* __restrict__ min_cell,
                 int* __restrict__ grid_coord, long long* __restrict__ key) {
    const unsigned long long basis = 14695981039346656037ull, prime = 1099511628211ull;
    const long long m0 = min_cell[0], m1 = min_cell[1], m2 = min_cell[2];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // numpy: coord (f32) / np.array(grid_size) is a float64 division, np.floor, astype(int)
        const long long c0 = (long long)floor((double)__ldg(coord + i * 3) / gx) - m0;
        const long long c1 = (long long)floor((double)__ldg(coord + i * 3 + 1) / gy) - m1;
        const long long c2 = (long long)floor((double)__ldg(coord + i * 3 + 2) / gz) - m2;
        if (grid_coord) { grid_coord[i * 3] = (int)c0; grid_coord[i * 3 + 1] = (int)c1; grid_coord[i * 3 + 2] = (int)c2; }
        unsigned long long h = basis;
        h *= prime; h ^= (unsigned long long)c0;
        h *= prime; h ^= (unsigned long long)c1;
        h *= prime; h ^= (unsigned long long)c2;
        key[i] = (long long)(h ^ 0x8000000000000000ull);
    }
}

// sums[index[r], :] += src[r, :] ; cnt[index[r]] += 1   (one thread per element; vector atomics when c % 4 == 0)
__global__ void __launch_bounds__(256)
scatter_sum_kernel(int64_t rows, int c, const float* __restrict__ src, const long long* __restrict__ index, int64_t dim_size,
                   float* __restrict__ sums, float* __restrict__ cnt) {
    const int vec = (c % 4 == 0) ? c / 4 : 0;
    if (vec) {
        const int64_t total = rows * vec;
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = t / vec;
            const int v = (int)(t % vec);
            const long long d = index[r];
            if (d < 0 || d >= dim_size) continue;
            const float4 x = __ldcs(reinterpret_cast<const float4*>(src) + t);
            atomicAdd(reinterpret_cast<float4*>(sums) + d * vec + v, x);
            if (v == 0) atomicAdd(cnt + d, 1.f);
        }
    } else {
        const int64_t total = rows * c;
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = t / c;
            const int j = (int)(t % c);
            const long long d = index[r];
            if (d < 0 || d >= dim_size) continue;
            atomicAdd(sums + d * c + j, __ldcs(src + t));
            if (j == 0) atomicAdd(cnt + d, 1.f);
        }
    }
}

__global__ void __launch_bounds__(256)
scatter_div_kernel(int64_t dim_size, int c, float* __restrict__ sums, const float* __restrict__ cnt) {
    const int64_t total = dim_size * c;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const float k = cnt[t / c];
        if (k > 1.f) sums[t] = __fdiv_rn(sums[t], k);   // torch_scatter: sum / clamp(count, 1)
    }
}

}  // namespace pob

using namespace pob;

// grid_size: the three voxel edge lengths (GridSample takes a scalar or a triple); min_cell: device int64[3] =
// floor(coord / grid_size).min(0), computed by the caller (one torch reduction); grid_coord (n, 3) i32 may be NULL.
POB_API int pob_grid_hash(int64_t n, const float* coord, double gx, double gy, double gz, const long long* min_cell,
                          int* grid_coord, long long* key, cudaStream_t stream) {
    if (n < 0 || !(gx > 0) || !(gy > 0) || !(gz > 0)) return POB_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (!coord || !min_cell || !key) return POB_ERR_BAD_ARG;
    grid_hash_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(n, coord, gx, gy, gz, min_cell, grid_coord, key);
    pob_count_launches(1);
    POB_RETURN_LAST_ERROR();
}

// out (dim_size, c) and count (dim_size) must be zero on entry; out receives the means.
POB_API int pob_scatter_mean(int64_t rows, int c, const float* src, const long long* index, int64_t dim_size, float* out,
                             float* count, cudaStream_t stream) {
    if (rows < 0 || c < 1 || dim_size < 0) return POB_ERR_BAD_ARG;
    if (rows == 0 || dim_size == 0) return 0;
    if (!src || !index || !out || !count) return POB_ERR_BAD_ARG;
    if (c % 4 == 0 && (((uintptr_t)src | (uintptr_t)out) & 15)) return POB_ERR_BAD_ARG;
    const int64_t work = c % 4 == 0 ? rows * (c / 4) : rows * c;
    scatter_sum_kernel<<<grid_for(work, 256, 8), 256, 0, stream>>>(rows, c, src, index, dim_size, out, count);
    scatter_div_kernel<<<grid_for(dim_size * c, 256, 8), 256, 0, stream>>>(dim_size, c, out, count);
    pob_count_launches(2);
    POB_RETURN_LAST_ERROR();
}
