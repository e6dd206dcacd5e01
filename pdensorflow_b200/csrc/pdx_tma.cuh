// pdx_tma.cuh - device helpers shared by the TMA tile kernels (pdx_fd_tma.cu, pdx_fd_aniso.cu): mbarrier and
// cp.async.bulk.tensor wrappers, 128-bit register vectors.
#pragma once
#include "pdx_internal.h"

namespace pdx {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA: 3-D tiled bulk tensor load global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

struct F4 {
    float v[4];
};
__device__ __forceinline__ F4 lds4(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    return F4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ F4 ldg4(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    return F4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void stg4(float* p, const F4& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
// z-edge patch of the four centre columns of a row (index clamp to [zclo, zchi])
__device__ __forceinline__ void zfix(F4& a, bool lo2, bool hi2) {
    if (lo2) a.v[0] = a.v[1];
    if (hi2) a.v[3] = a.v[2];
}

}  // namespace
}  // namespace pdx
