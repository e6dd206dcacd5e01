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

// ---- SELL-32 row products with the matrix stream on the copy engine ----------------------------------------------
// A warp owns slices gw, gw + nw, ... ; lane 0 keeps SELL_STAGE_NB of them in flight as cp.async.bulk copies (the col
// and val words of a slice are contiguous: two copies on one mbarrier) into the warp's ring in dynamic shared memory;
// the lanes read their columns from shared memory, so a whole row of gathers is in flight per lane.  Row products are
// accumulated in row order (k = 0, 1, ...), i.e. bitwise those of a per-thread loop over the slice.
// Dynamic shared memory layout: [warps][NB] slots of slot_bytes | [warps][NB] mbarriers | [warps][NB] slice widths.
constexpr int SELL_STAGE_NB = 2;

__host__ __device__ inline size_t sell_stage_smem_bytes(int threads, int slot_bytes) {
    return (size_t)(threads / 32) * SELL_STAGE_NB * ((size_t)slot_bytes + 8 + 4) + 128;
}
__device__ __forceinline__ uint32_t sell_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// once per kernel, by every warp, before the first sell_rows_staged call
__device__ __forceinline__ void sell_stage_init(unsigned char* dsm, int slot_bytes) {
    const int nwb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsm + (size_t)nwb * SELL_STAGE_NB * slot_bytes) + wib * SELL_STAGE_NB;
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < SELL_STAGE_NB; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sell_smem_addr(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}

// uses: slots this warp has consumed so far (carries the mbarrier phase from call to call).  ldx(c) returns x[c];
// f(row, A_row . x) is called by the lane that owns the row.  policy: L2 cache policy of the matrix stream.
// pre(slice) / post(slice): called by every lane of the warp right before the gathers of a slice and right after its
// rows are done (hooks for the neighbour flags of a partitioned step).
// smap(v): the slice visited v-th (a permutation of 0 .. nsl-1; identity unless the caller wants an order).
template <class LDX, class F, class PRE, class POST, class SMAP>
__device__ __forceinline__ void sell_rows_staged(unsigned char* dsm, int slot_bytes, int n, const int32_t* slice_ptr,
                                                 const int32_t* col, const float* vals, unsigned long long policy,
                                                 unsigned int& uses, LDX&& ldx, F&& f, PRE&& pre, POST&& post, SMAP&& smap) {
    constexpr int NB = SELL_STAGE_NB;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    unsigned char* ring = dsm + (size_t)wib * NB * slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsm + (size_t)nwb * NB * slot_bytes) + wib * NB;
    int* widths = reinterpret_cast<int*>(dsm + (size_t)nwb * NB * (slot_bytes + 8)) + wib * NB;
    const int gw = blockIdx.x * nwb + wib, nw = gridDim.x * nwb;
    const int nsl = (n + 31) >> 5;
    const int cnt = gw < nsl ? (nsl - gw + nw - 1) / nw : 0;

    auto issue = [&](int t, int beg, int end) {          // lane 0
        const int sl = (int)((uses + t) % NB);
        const uint32_t bytes = (uint32_t)(end - beg) * 4u;
        if (2u * bytes > (uint32_t)slot_bytes) __trap();      // a slice wider than the registered width: fail loudly
        const uint32_t bar = sell_smem_addr(&bars[sl]);
        widths[sl] = (end - beg) >> 5;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
        const uint32_t dst = sell_smem_addr(ring + (size_t)sl * slot_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst), "l"(col + beg), "r"(bytes), "r"(bar), "l"(policy) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst + bytes), "l"(vals + beg), "r"(bytes), "r"(bar), "l"(policy) : "memory");
    };
    if (lane == 0) {
        for (int t = 0; t < NB && t < cnt; ++t) {
            const int s0 = smap(gw + t * nw);
            issue(t, slice_ptr[s0], slice_ptr[s0 + 1]);
        }
    }
    for (int t = 0; t < cnt; ++t) {
        int nbeg = 0, nend = 0;
        const bool refill = lane == 0 && t + NB < cnt;
        if (refill) {                                    // slice_ptr of the refill: loaded now, used after the products
            const int s1 = smap(gw + (t + NB) * nw);
            nbeg = slice_ptr[s1]; nend = slice_ptr[s1 + 1];
        }
        const unsigned int use = uses + t;
        const int sl = (int)(use % NB);
        const uint32_t bar = sell_smem_addr(&bars[sl]), parity = (use / NB) & 1u;
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        const int slice = smap(gw + t * nw);
        pre(slice);
        const int w = widths[sl];
        const int32_t* scol = reinterpret_cast<const int32_t*>(ring + (size_t)sl * slot_bytes) + lane;
        const float* sval = reinterpret_cast<const float*>(ring + (size_t)sl * slot_bytes + (size_t)w * 128) + lane;
        float acc = 0.0f;
        int k = 0;
        for (; k + 4 <= w; k += 4) {
            const int c0 = scol[k * 32], c1 = scol[k * 32 + 32], c2 = scol[k * 32 + 64], c3 = scol[k * 32 + 96];
            const float x0 = ldx(c0), x1 = ldx(c1), x2 = ldx(c2), x3 = ldx(c3);
            acc = fmaf(sval[k * 32], x0, acc);
            acc = fmaf(sval[k * 32 + 32], x1, acc);
            acc = fmaf(sval[k * 32 + 64], x2, acc);
            acc = fmaf(sval[k * 32 + 96], x3, acc);
        }
        for (; k < w; ++k) acc = fmaf(sval[k * 32], ldx(scol[k * 32]), acc);
        const int row = slice * 32 + lane;
        if (row < n) f(row, acc);
        post(slice);
        __syncwarp();                                    // every lane is done with the slot
        if (refill) issue(t + NB, nbeg, nend);
    }
    uses += (unsigned int)cnt;
}

template <class LDX, class F>
__device__ __forceinline__ void sell_rows_staged(unsigned char* dsm, int slot_bytes, int n, const int32_t* slice_ptr,
                                                 const int32_t* col, const float* vals, unsigned long long policy,
                                                 unsigned int& uses, LDX&& ldx, F&& f) {
    sell_rows_staged(dsm, slot_bytes, n, slice_ptr, col, vals, policy, uses, ldx, f, [](int) {}, [](int) {},
                     [](int v) { return v; });
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
