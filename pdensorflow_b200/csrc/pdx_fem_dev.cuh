// pdx_fem_dev.cuh - device helpers shared by the FEM translation units (pdx_fem.cu, pdx_fem_part.cu).
#pragma once
#include "pdx_internal.h"

namespace pdx {

// G lanes cooperate on one row (G chosen from the mean row length).
template <int G>
__device__ __forceinline__ float row_dot(const int32_t* __restrict__ col, const float* __restrict__ val,
                                         const float* x, int beg, int end, int sub, bool cg_loads) {
    float acc = 0.0f;
    for (int k = beg + sub; k < end; k += G) {
        const float xv = cg_loads ? __ldcg(x + col[k]) : x[col[k]];
        acc = fmaf(val[k], xv, acc);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
    return acc;
}

inline int pick_group(int n, int64_t nnz) {
    const double mean = n > 0 ? (double)nnz / n : 1.0;
    if (mean <= 6) return 4;
    if (mean <= 12) return 8;
    if (mean <= 24) return 16;
    return 32;
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float t = 0.0f;
    const int nw = blockDim.x >> 5;
    if (w == 0) {
        t = lane < nw ? sh[lane] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// Two block sums through one pair of block barriers (sh: 66 floats).  Same shuffle trees and the same order of the
// per-warp partials as two block_sum calls, so the results are bitwise those of block_sum.
__device__ __forceinline__ void block_sum2(float a, float b, float* sh, float& ra, float& rb) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sh[w] = a; sh[33 + w] = b; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    if (w == 0) {
        float t = lane < nw ? sh[lane] : 0.0f;
        float u = lane < nw ? sh[33 + lane] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_xor_sync(0xffffffffu, t, o);
            u += __shfl_xor_sync(0xffffffffu, u, o);
        }
        if (lane == 0) { sh[32] = t; sh[65] = u; }
    }
    __syncthreads();
    ra = sh[32];
    rb = sh[65];
}

// grid_total of two partial arrays in one pass
__device__ __forceinline__ void grid_total2(const float* pa, const float* pb, float* sh, float& ra, float& rb) {
    float v = 0.0f, w = 0.0f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { v += __ldcg(pa + i); w += __ldcg(pb + i); }
    block_sum2(v, w, sh, ra, rb);
}

// Sum of the per-block partials, same order in every block => every block gets the same bits.
__device__ __forceinline__ float grid_total(const float* part, float* sh) {
    float v = 0.0f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += __ldcg(part + i);
    return block_sum(v, sh);
}

}  // namespace pdx
