// pdx_fd_aniso.cu - fused FD step with the 19-point fibre-tensor operator on TMA-staged tiles (sm_100a).
//
// Reference: gpuSolve/diffop3D/laplace_heterogeneous_anisotropic_diffusion.py:3-95 (operator),
// Tests/FD/Fenton_atria/fenton.py:106-112 (the solve() it would be used in).
//
// The reference forms the conductivity tensor C = sigma_t I + (sigma_l - sigma_t) a (x) a from DIFF [.,2] and
// AVEC [.,6] at every call (:40-45).  The fields are constants of a run, so the plan forms the six components ONCE
// (same fp32 operations, bit-identical) into six [nx][ny][nz] arrays (aniso_tensor_kernel): a step then reads 6 floats
// per node instead of 8 channel-last ones, and every component is a plain 3-D array a tensor map can tile.
//
// Kernel: the K1 scheme (pdx_fd_tma.cu) with seven staged fields instead of one.  A CTA (8 warps) owns an 8 x 128 tile
// of the (y,z) plane and marches along x; U and the six tensor components of a plane arrive as seven
// cp.async.bulk.tensor boxes of 1 x 10 x 136 floats (halo included, zero filled outside the array) in one slot of a
// 5-deep shared-memory ring (3 planes in use, 2 in flight; 188 KB, one CTA per SM).  A lane evaluates four consecutive z
// nodes from 128-bit shared-memory reads:
//   flux(axis a, side s) = s 0.5 ( T_a + sum_{b != a} T_b ),
//   T_a = (C_aa[s e_a] + C_aa[0]) (X[s e_a] - X[0]) / h_a,
//   T_b = 0.5 (C_ab[s e_a + e_b] + C_ab[s e_a - e_b]) (X[s e_a + e_b] - X[s e_a - e_b]) / h_b,
//   lap = (e - w)/h_x + (n - s)/h_y + (u - d)/h_z,
// in the reference's order of operations (EXACT: separately rounded, IEEE divisions - bitwise the generic kernel and
// the oracle; FAST: reciprocal multiplies, FMA contraction).  X is U0 = enforce_boundary(U) (indices clamped to
// [1, n-2]), the tensor is symmetric padded (indices clamped to the array): both clamps are applied when addressing
// shared memory, z edges are patched in registers.  The ionic update, the Euler update and the fused halo stores are
// those of K1; the gates are prefetched one plane ahead in registers.
//
// Algorithmic bytes per node-step with Fenton-Karma: 6 x 4 (tensor) + 32 (U,V,W,S read + written) = 56.
#include "pdx_internal.h"
#include "pdx_tma.cuh"

namespace pdx {

namespace {

constexpr int TZ = 128;
constexpr int ZH = 4;
constexpr int COLS = TZ + 2 * ZH;        // 136
constexpr int NWARP = 8;
constexpr int NTHREADS = NWARP * 32;
constexpr int TY = NWARP;
constexpr int ROWS = TY + 2;
constexpr int NF = 7;                    // U, C00, C01, C02, C11, C12, C22
constexpr int STAGES = 5;
constexpr int FIELD_FLOATS = (ROWS * COLS * 4 + 127) / 128 * 128 / 4;   // 1376
constexpr int FIELD_BYTES = FIELD_FLOATS * 4;
constexpr int SLOT_BYTES = NF * FIELD_BYTES;
constexpr int TX_BYTES = NF * ROWS * COLS * 4;
constexpr int BAR_OFF = STAGES * SLOT_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + STAGES * 8;

enum { F_U = 0, F_C00, F_C01, F_C02, F_C11, F_C12, F_C22 };

struct AnisoMaps {
    CUtensorMap m[NF];
};

// six values z-1 .. z+4 of a row
struct W6 {
    float v[6];
};
__device__ __forceinline__ W6 ldw6(const float* row) {      // row points at column z of the shared-memory row
    const float4 t = *reinterpret_cast<const float4*>(row);
    return W6{{row[-1], t.x, t.y, t.z, t.w, row[4]}};
}

template <int MATH> __device__ __forceinline__ float mdivh(float x, float h, float rh) {
    return MATH == PDX_MATH_EXACT ? xdiv(x, h) : x * rh;
}
template <int MATH> __device__ __forceinline__ float msub(float a, float b) {
    return MATH == PDX_MATH_EXACT ? xsub(a, b) : a - b;
}
// normal term  (Cn + C0) * (Xn - X0) / h
template <int MATH>
__device__ __forceinline__ float t_main(float cn, float c0, float xn, float x0, float h, float rh) {
    return mdivh<MATH>(mmul<MATH>(madd<MATH>(cn, c0), msub<MATH>(xn, x0)), h, rh);
}
// transverse term  (0.5 * (Cp + Cm)) * (Xp - Xm) / h
template <int MATH>
__device__ __forceinline__ float t_cross(float cp, float cm, float xp, float xm, float h, float rh) {
    return mdivh<MATH>(mmul<MATH>(mmul<MATH>(0.5f, madd<MATH>(cp, cm)), msub<MATH>(xp, xm)), h, rh);
}

template <int MODEL, int MATH>
__global__ void __launch_bounds__(NTHREADS, 1)
fd_aniso_kernel(const __grid_constant__ AnisoMaps maps, const __grid_constant__ FdArgs a) {
    constexpr int NS = NStates<MODEL>::value;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int b = blockIdx.x;
    const int tz = b % a.ntz;  b /= a.ntz;
    const int ty = b % a.nty;  b /= a.nty;
    const int chunk = b;
    const int z0 = tz * TZ, y0 = ty * TY;
    const int xb = a.xb + chunk * a.x_chunk;
    const int xe = min(xb + a.x_chunk, a.xe);
    const int np = xe - xb;
    const int nlog = np + 2;                         // logical planes q = xb-1 .. xe

    auto slot = [&](int j, int f) { return reinterpret_cast<float*>(smem + (j % STAGES) * SLOT_BYTES + f * FIELD_BYTES); };
    auto issue = [&](int j) {   // thread 0 only
        const int q = xb - 1 + j;
        uint64_t* bar = &full[j % STAGES];
        mbar_expect_tx(bar, TX_BYTES);
        tma_load_3d(slot(j, F_U), &maps.m[F_U], z0 - ZH, y0 - 1, clampi(q, a.xclo, a.xchi), bar);
        const int qc = clampi(q, a.dxlo, a.dxhi);    // tensor: symmetric pad only
#pragma unroll
        for (int f = 1; f < NF; ++f) tma_load_3d(slot(j, f), &maps.m[f], z0 - ZH, y0 - 1, qc, bar);
    };
    auto wait_plane = [&](int j) { mbar_wait(&full[j % STAGES], (uint32_t)((j / STAGES) & 1)); };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int n0 = min(STAGES, nlog);
        for (int j = 0; j < n0; ++j) issue(j);
    }

    // ---- per-thread geometry (constant along the march) ----
    const int  z   = z0 + 4 * lane;
    const int  col = ZH + 4 * lane;
    const int  gy  = y0 + warp;
    const bool act = z < a.nz && gy < a.ny;
    const int  g   = min(gy, a.ny - 1);
    // X = U0: rows clamped to [yclo, ychi]; tensor: rows clamped to the array
    const int rc = (clampi(g, a.yclo, a.ychi) - (y0 - 1)) * COLS + col;
    const int rm = (clampi(g - 1, a.yclo, a.ychi) - (y0 - 1)) * COLS + col;
    const int rp = (clampi(g + 1, a.yclo, a.ychi) - (y0 - 1)) * COLS + col;
    const int cc = (g - (y0 - 1)) * COLS + col;
    const int cm = (max(g - 1, 0) - (y0 - 1)) * COLS + col;
    const int cp = (min(g + 1, a.ny - 1) - (y0 - 1)) * COLS + col;
    const bool zb_tile = (z0 - 1 < a.zclo) || (z0 + TZ > a.zchi);        // CTA uniform
    const bool yb_tile = (y0 - 1 < a.yclo) || (y0 + TY > a.ychi);
    const bool xlo2 = z < a.zclo, xlo1 = z - 1 < a.zclo, xhi2 = z + 3 > a.zchi, xhi1 = z + 4 > a.zchi;
    const bool clo1 = z - 1 < 0, chi1 = z + 4 > a.nz - 1;
    // z-edge patches: X clamps to [zclo, zchi], the tensor to [0, nz-1]
    auto fixX6 = [&](W6& w) {
        if (xlo2) w.v[1] = w.v[2];
        if (xhi2) w.v[4] = w.v[3];
        if (xlo1) w.v[0] = w.v[1];
        if (xhi1) w.v[5] = w.v[4];
    };
    auto fixX4 = [&](F4& w) {
        if (xlo2) w.v[0] = w.v[1];
        if (xhi2) w.v[3] = w.v[2];
    };
    auto fixC6 = [&](W6& w) {
        if (clo1) w.v[0] = w.v[1];
        if (chi1) w.v[5] = w.v[4];
    };

    const float hx = a.hx, hy = a.hy, hz = a.hz;
    const float rhx = 1.0f / hx, rhy = 1.0f / hy, rhz = 1.0f / hz;      // FAST only

    F4 stn[NS > 0 ? NS : 1];
    auto load_states = [&](int p) {
        if (!act) return;
        const size_t gi = ((size_t)p * a.ny + gy) * a.nz + z;
#pragma unroll
        for (int s = 0; s < NS; ++s) stn[s] = ldg4(a.st[s] + gi);
    };
    if (NS > 0 && np > 0) load_states(xb);

    for (int i = 0; i < np; ++i) {
        const int p = xb + i;
        const int jc = i + 1;
        if (i == 0) { wait_plane(0); wait_plane(1); }
        wait_plane(i + 2);
        F4 stc[NS > 0 ? NS : 1];
#pragma unroll
        for (int s = 0; s < NS; ++s) stc[s] = stn[s];
        if (NS > 0 && i + 1 < np) load_states(p + 1);

        if (act) {
            const size_t gi = ((size_t)p * a.ny + gy) * a.nz + z;
            const float* Um = slot(jc - 1, F_U);
            const float* Uc = slot(jc, F_U);
            const float* Up = slot(jc + 1, F_U);
            // ---- potential ----
            W6 x_mc = ldw6(Um + rc), x_pc = ldw6(Up + rc);                 // planes i-1 / i+1, row j
            W6 x_cm = ldw6(Uc + rm), x_cc = ldw6(Uc + rc), x_cp = ldw6(Uc + rp);   // plane i, rows j-1, j, j+1
            F4 x_mm = lds4(Um + rm), x_mp = lds4(Um + rp), x_pm = lds4(Up + rm), x_pp = lds4(Up + rp);
            F4 raw = F4{{x_cc.v[1], x_cc.v[2], x_cc.v[3], x_cc.v[4]}};
            const bool bnd = yb_tile || zb_tile || p < a.xclo || p > a.xchi;     // CTA uniform
            if (bnd && MODEL != PDX_MODEL_NONE) raw = ldg4(a.Uin + gi);
            if (zb_tile) {
                fixX6(x_mc); fixX6(x_pc); fixX6(x_cm); fixX6(x_cc); fixX6(x_cp);
                fixX4(x_mm); fixX4(x_mp); fixX4(x_pm); fixX4(x_pp);
            }
            // ---- tensor ----
            const F4 c00m = lds4(slot(jc - 1, F_C00) + cc), c00c = lds4(slot(jc, F_C00) + cc), c00p = lds4(slot(jc + 1, F_C00) + cc);
            const F4 c11m = lds4(slot(jc, F_C11) + cm), c11c = lds4(slot(jc, F_C11) + cc), c11p = lds4(slot(jc, F_C11) + cp);
            W6 c22 = ldw6(slot(jc, F_C22) + cc);
            const F4 c01mm = lds4(slot(jc - 1, F_C01) + cm), c01mp = lds4(slot(jc - 1, F_C01) + cp);
            const F4 c01pm = lds4(slot(jc + 1, F_C01) + cm), c01pp = lds4(slot(jc + 1, F_C01) + cp);
            W6 c02m = ldw6(slot(jc - 1, F_C02) + cc), c02p = ldw6(slot(jc + 1, F_C02) + cc);
            W6 c12m = ldw6(slot(jc, F_C12) + cm), c12p = ldw6(slot(jc, F_C12) + cp);
            if (zb_tile) { fixC6(c22); fixC6(c02m); fixC6(c02p); fixC6(c12m); fixC6(c12p); }

            F4 out;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = e, k1 = e + 1, k2 = e + 2;          // window indices of z-1, z, z+1
                const float x0 = x_cc.v[k1];
                // axis x, side +/-: normal, transverse y, transverse z     (:50-62)
                float tp = t_main<MATH>(c00p.v[e], c00c.v[e], x_pc.v[k1], x0, hx, rhx);
                tp = madd<MATH>(tp, t_cross<MATH>(c01pp.v[e], c01pm.v[e], x_pp.v[e], x_pm.v[e], hy, rhy));
                tp = madd<MATH>(tp, t_cross<MATH>(c02p.v[k2], c02p.v[k0], x_pc.v[k2], x_pc.v[k0], hz, rhz));
                float tm = t_main<MATH>(c00m.v[e], c00c.v[e], x_mc.v[k1], x0, hx, rhx);
                tm = madd<MATH>(tm, t_cross<MATH>(c01mp.v[e], c01mm.v[e], x_mp.v[e], x_mm.v[e], hy, rhy));
                tm = madd<MATH>(tm, t_cross<MATH>(c02m.v[k2], c02m.v[k0], x_mc.v[k2], x_mc.v[k0], hz, rhz));
                const float gx = mdivh<MATH>(msub<MATH>(mmul<MATH>(0.5f, tp), mmul<MATH>(-0.5f, tm)), hx, rhx);
                // axis y: transverse x, normal, transverse z               (:64-76)
                float up = t_cross<MATH>(c01pp.v[e], c01mp.v[e], x_pp.v[e], x_mp.v[e], hx, rhx);
                up = madd<MATH>(up, t_main<MATH>(c11p.v[e], c11c.v[e], x_cp.v[k1], x0, hy, rhy));
                up = madd<MATH>(up, t_cross<MATH>(c12p.v[k2], c12p.v[k0], x_cp.v[k2], x_cp.v[k0], hz, rhz));
                float um = t_cross<MATH>(c01pm.v[e], c01mm.v[e], x_pm.v[e], x_mm.v[e], hx, rhx);
                um = madd<MATH>(um, t_main<MATH>(c11m.v[e], c11c.v[e], x_cm.v[k1], x0, hy, rhy));
                um = madd<MATH>(um, t_cross<MATH>(c12m.v[k2], c12m.v[k0], x_cm.v[k2], x_cm.v[k0], hz, rhz));
                const float gyv = mdivh<MATH>(msub<MATH>(mmul<MATH>(0.5f, up), mmul<MATH>(-0.5f, um)), hy, rhy);
                // axis z: transverse x, transverse y, normal               (:78-90)
                float wp = t_cross<MATH>(c02p.v[k2], c02m.v[k2], x_pc.v[k2], x_mc.v[k2], hx, rhx);
                wp = madd<MATH>(wp, t_cross<MATH>(c12p.v[k2], c12m.v[k2], x_cp.v[k2], x_cm.v[k2], hy, rhy));
                wp = madd<MATH>(wp, t_main<MATH>(c22.v[k2], c22.v[k1], x_cc.v[k2], x0, hz, rhz));
                float wm = t_cross<MATH>(c02p.v[k0], c02m.v[k0], x_pc.v[k0], x_mc.v[k0], hx, rhx);
                wm = madd<MATH>(wm, t_cross<MATH>(c12p.v[k0], c12m.v[k0], x_cp.v[k0], x_cm.v[k0], hy, rhy));
                wm = madd<MATH>(wm, t_main<MATH>(c22.v[k0], c22.v[k1], x_cc.v[k0], x0, hz, rhz));
                const float gz = mdivh<MATH>(msub<MATH>(mmul<MATH>(0.5f, wp), mmul<MATH>(-0.5f, wm)), hz, rhz);
                const float lap = madd<MATH>(madd<MATH>(gx, gyv), gz);   // :93

                float u1 = x0;
                if (MODEL != PDX_MODEL_NONE) {
                    float st[PDX_MAX_STATES];
#pragma unroll
                    for (int s = 0; s < NS; ++s) st[s] = stc[s].v[e];
                    const float dU = ionic_update<MODEL, MATH>(raw.v[e], st, a.mp);
#pragma unroll
                    for (int s = 0; s < NS; ++s) stc[s].v[e] = st[s];
                    u1 = madd<MATH>(u1, mmul<MATH>(a.dt, dU));
                }
                out.v[e] = madd<MATH>(u1, mmul<MATH>(a.coef, lap));
            }
            stg4(a.Uout + gi, out);
            if (a.mirror_lo && p == a.mirror_lo_plane) stg4(a.mirror_lo + (size_t)gy * a.nz + z, out);
            if (a.mirror_hi && p == a.mirror_hi_plane) stg4(a.mirror_hi + (size_t)gy * a.nz + z, out);
#pragma unroll
            for (int s = 0; s < NS; ++s) stg4(a.st[s] + gi, stc[s]);
        }
        __syncthreads();                              // everyone is done with logical plane i
        if (tid == 0 && i + STAGES < nlog) issue(i + STAGES);
    }
}

// C = sigma_t I + (sigma_l - sigma_t) a (x) a, the six components as six [n] arrays      (:40-45)
__global__ void __launch_bounds__(256) aniso_tensor_kernel(const float* __restrict__ DIFF, const float* __restrict__ AVEC,
                                                           float* __restrict__ C, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < n; id += stride) {
        const float st = DIFF[2 * id], dl = DIFF[2 * id + 1];
        const float* A = AVEC + 6 * id;
        // AVEC order (a1a1, a1a2, a1a3, a2a2, a2a3, a3a3) -> C00, C01, C02, C11, C12, C22
        C[0 * n + id] = xadd(st, xmul(A[0], dl));
        C[1 * n + id] = xmul(A[1], dl);
        C[2 * n + id] = xmul(A[2], dl);
        C[3 * n + id] = xadd(st, xmul(A[3], dl));
        C[4 * n + id] = xmul(A[4], dl);
        C[5 * n + id] = xadd(st, xmul(A[5], dl));
    }
}

template <int MODEL, int MATH>
int launch_inst(const FdArgs& a0, const AnisoMaps& m, cudaStream_t s) {
    auto kern = fd_aniso_kernel<MODEL, MATH>;
    static bool attr_done[PDX_MAX_DEVICES] = {};
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (dev < 0 || dev >= PDX_MAX_DEVICES || !attr_done[dev]) {
        if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
                       "cudaFuncSetAttribute(smem, aniso)"))
            return 1;
        if (dev >= 0 && dev < PDX_MAX_DEVICES) attr_done[dev] = true;
    }
    const int nplanes = a0.xe - a0.xb;
    if (nplanes <= 0) return 0;
    FdArgs a = a0;
    a.nchunk = (nplanes + a.x_chunk - 1) / a.x_chunk;
    const long long nb = (long long)a.ntz * a.nty * a.nchunk;
    if (nb > 0x7fffffffLL) { set_error("fd_aniso: grid too large"); return 1; }
    kern<<<(unsigned)nb, NTHREADS, SMEM_BYTES, s>>>(m, a);
    count_launch();
    return check_cuda(cudaGetLastError(), "fd_aniso_kernel launch");
}

template <int MODEL>
int launch_math(int math, const FdArgs& a, const AnisoMaps& m, cudaStream_t s) {
    return math == PDX_MATH_EXACT ? launch_inst<MODEL, PDX_MATH_EXACT>(a, m, s) : launch_inst<MODEL, PDX_MATH_FAST>(a, m, s);
}

}  // namespace

bool fd_aniso_tile_supported(const pdx_fd_desc& d) {
    return d.ndim == 3 && d.lap_kind == PDX_LAP_ANISO && d.nz % 4 == 0 && d.nz >= 4 && d.ny >= 3 && !d.dirichlet &&
           (d.model == PDX_MODEL_NONE || d.model == PDX_MODEL_FENTON4V || d.model == PDX_MODEL_MMS2V || d.model == PDX_MODEL_MS2V);
}

int fd_aniso_auto_chunk(const pdx_fd_desc& d) {
    // One CTA per SM: long marches keep the two halo planes every chunk re-reads (seven fields each) small against the
    // planes it advances; shorten only to keep two waves of CTAs.
    const int nplanes = d.x_end - d.x_begin;
    if (nplanes <= 1) return 1;
    const long long tiles = (long long)((d.ny + TY - 1) / TY) * ((d.nz + TZ - 1) / TZ);
    int chunk = 32;
    while (chunk > 4 && tiles * ((nplanes + chunk - 1) / chunk) < 2LL * sm_count()) chunk /= 2;
    return chunk < nplanes ? chunk : nplanes;
}

int launch_aniso_tensor(const float* DIFF, const float* AVEC, float* C, int64_t n, cudaStream_t s) {
    int64_t nb = (n + 255) / 256;
    if (nb > sm_count() * 16) nb = sm_count() * 16;
    if (nb < 1) nb = 1;
    aniso_tensor_kernel<<<(unsigned)nb, 256, 0, s>>>(DIFF, AVEC, C, n);
    count_launch();
    return check_cuda(cudaGetLastError(), "aniso_tensor_kernel launch");
}

int launch_fd_aniso_tile(const pdx_fd_desc& d, const FdArgs& a, const CUtensorMap& mapU, const CUtensorMap* mapC,
                         cudaStream_t s) {
    AnisoMaps m;
    m.m[F_U] = mapU;
    for (int f = 1; f < NF; ++f) m.m[f] = mapC[f - 1];
    switch (d.model) {
        case PDX_MODEL_NONE:     return launch_math<PDX_MODEL_NONE>(d.math_mode, a, m, s);
        case PDX_MODEL_FENTON4V: return launch_math<PDX_MODEL_FENTON4V>(d.math_mode, a, m, s);
        case PDX_MODEL_MMS2V:    return launch_math<PDX_MODEL_MMS2V>(d.math_mode, a, m, s);
        case PDX_MODEL_MS2V:     return launch_math<PDX_MODEL_MS2V>(d.math_mode, a, m, s);
    }
    set_error("fd_aniso: unknown model");
    return 1;
}

}  // namespace pdx
