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
// Kernel: the K1 scheme (pdx_fd_tma.cu) with seven staged fields instead of one.  A CTA owns an 8 x 128 tile
// of the (y,z) plane and marches along x; U and the six tensor components of a plane arrive as seven
// cp.async.bulk.tensor boxes of 1 x 10 x 136 floats (halo included, zero filled outside the array) in one slot of a
// 5-deep shared-memory ring (188 KB + 36 KB of gates, one CTA per SM; slots are refilled two planes ahead - running the
// ring three planes ahead, which the register rotation below would allow, measured slower: 89.5 vs 97.5 Gnode-steps/s).
// The rows of the three planes in use live in registers and rotate with the march (see the loop).  A thread evaluates N = 4 (8 warps
// per CTA; FAST) or 2 (16 warps; EXACT) consecutive z nodes from 64 / 128-bit shared-memory reads:
//   flux(axis a, side s) = s 0.5 ( T_a + sum_{b != a} T_b ),
//   T_a = (C_aa[s e_a] + C_aa[0]) (X[s e_a] - X[0]) / h_a,
//   T_b = 0.5 (C_ab[s e_a + e_b] + C_ab[s e_a - e_b]) (X[s e_a + e_b] - X[s e_a - e_b]) / h_b,
//   lap = (e - w)/h_x + (n - s)/h_y + (u - d)/h_z,
// in the reference's order of operations (EXACT: separately rounded, IEEE divisions - bitwise the generic kernel and
// the oracle; FAST: reciprocal multiplies, FMA contraction).  X is U0 = enforce_boundary(U) (indices clamped to
// [1, n-2]), the tensor is symmetric padded (indices clamped to the array): both clamps are applied when addressing
// shared memory, z edges are patched in registers.  The ionic update, the Euler update and the fused halo stores are
// those of K1; the gates arrive by TMA too (boxes without halo, three planes ahead).
//
// Algorithmic bytes per node-step with Fenton-Karma: 6 x 4 (tensor) + 32 (U,V,W,S read + written) = 56.
#include <cstdlib>
#include <type_traits>
#include "pdx_internal.h"
#include "pdx_tma.cuh"

namespace pdx {

namespace {

constexpr int TZ = 128;
constexpr int ZH = 4;
constexpr int COLS = TZ + 2 * ZH;        // 136
constexpr int TY = 8;
constexpr int ROWS = TY + 2;
constexpr int NF = 7;                    // U, C00, C01, C02, C11, C12, C22
constexpr int STAGES = 5;
constexpr int FIELD_FLOATS = (ROWS * COLS * 4 + 127) / 128 * 128 / 4;   // 1376
constexpr int FIELD_BYTES = FIELD_FLOATS * 4;
constexpr int SLOT_BYTES = NF * FIELD_BYTES;
constexpr int TX_BYTES = NF * ROWS * COLS * 4;
constexpr int SSTAGES = 3;               // gates: planes in flight (a one-plane register prefetch left 22 % of the
                                         // samples waiting for the loads - ncu r02)
constexpr int ST_BYTES = TY * TZ * 4;
constexpr int SRING_OFF = STAGES * SLOT_BYTES;
constexpr int smem_bytes(int ns) { return SRING_OFF + SSTAGES * ns * ST_BYTES + (STAGES + SSTAGES) * 8; }

enum { F_U = 0, F_C00, F_C01, F_C02, F_C11, F_C12, F_C22 };

struct AnisoMaps {
    CUtensorMap m[NF];
    CUtensorMap st[PDX_MAX_STATES];      // the gates: boxes of 1 x TY x TZ, no halo
};

// N consecutive z values of a row / the N + 2 values z-1 .. z+N (N = nodes per thread: 4 or 2)
template <int N> struct Vn { float v[N]; };
template <int N> __device__ __forceinline__ Vn<N> ldv_s(const float* p);
template <> __device__ __forceinline__ Vn<4> ldv_s<4>(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    return Vn<4>{{t.x, t.y, t.z, t.w}};
}
template <> __device__ __forceinline__ Vn<2> ldv_s<2>(const float* p) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    return Vn<2>{{t.x, t.y}};
}
template <int N> __device__ __forceinline__ void stv_g(float* p, const Vn<N>& a);
template <> __device__ __forceinline__ void stv_g<4>(float* p, const Vn<4>& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
template <> __device__ __forceinline__ void stv_g<2>(float* p, const Vn<2>& a) {
    *reinterpret_cast<float2*>(p) = make_float2(a.v[0], a.v[1]);
}
template <int N> __device__ __forceinline__ Vn<N + 2> ldw(const float* row) {   // row points at column z of the row
    const Vn<N> t = ldv_s<N>(row);
    Vn<N + 2> w;
    w.v[0] = row[-1];
#pragma unroll
    for (int e = 0; e < N; ++e) w.v[e + 1] = t.v[e];
    w.v[N + 1] = row[N];
    return w;
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

// N nodes per thread: 4 (8 warps per CTA, 128-bit accesses) or 2 (16 warps, 64-bit accesses: a tenth more instructions
// per node, twice the warps to hide the shared-memory and issue latencies of a one-CTA-per-SM kernel)
template <int MODEL, int MATH, int N>
__global__ void __launch_bounds__(TY * TZ / N, 1)
fd_aniso_kernel(const __grid_constant__ AnisoMaps maps, const __grid_constant__ FdArgs a) {
    constexpr int NS = NStates<MODEL>::value;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sring = smem + SRING_OFF;
    uint64_t* full = reinterpret_cast<uint64_t*>(sring + SSTAGES * NS * ST_BYTES);
    uint64_t* sfull = full + STAGES;
    constexpr int TPR = TZ / N;                      // threads per tile row
    typedef Vn<N> V;
    typedef Vn<N + 2> W;
    const int tid = threadIdx.x, zl = tid % TPR, trow = tid / TPR;

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
    auto sslot = [&](int i, int s) { return reinterpret_cast<float*>(sring + ((i % SSTAGES) * (NS > 0 ? NS : 1) + s) * ST_BYTES); };
    auto issue_states = [&](int i) {   // thread 0 only; plane xb + i
        uint64_t* bar = &sfull[i % SSTAGES];
        mbar_expect_tx(bar, NS * ST_BYTES);
#pragma unroll
        for (int s = 0; s < NS; ++s) tma_load_3d(sslot(i, s), &maps.st[s], z0, y0, xb + i, bar);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
#pragma unroll
        for (int s = 0; s < SSTAGES; ++s) mbar_init(&sfull[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int n0 = min(STAGES, nlog);
        for (int j = 0; j < n0; ++j) {
            issue(j);
            if (NS > 0 && j < SSTAGES && j < np) issue_states(j);
        }
    }

    // ---- per-thread geometry (constant along the march) ----
    const int  z   = z0 + N * zl;
    const int  col = ZH + N * zl;
    const int  gy  = y0 + trow;
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
    const bool xlo2 = z < a.zclo, xlo1 = z - 1 < a.zclo, xhi2 = z + N - 1 > a.zchi, xhi1 = z + N > a.zchi;
    const bool clo1 = z - 1 < 0, chi1 = z + N > a.nz - 1;
    // the ionic update takes the raw potential: the centre row at its unclamped index, unless the PLANE was clamped
    const bool row_clamped = g < a.yclo || g > a.ychi;
    // z-edge patches: X clamps to [zclo, zchi], the tensor to [0, nz-1]
    auto fixXw = [&](W& w) {
        if (xlo2) w.v[1] = w.v[2];
        if (xhi2) w.v[N] = w.v[N - 1];
        if (xlo1) w.v[0] = w.v[1];
        if (xhi1) w.v[N + 1] = w.v[N];
    };
    auto fixXv = [&](V& w) {
        if (xlo2) w.v[0] = w.v[1];
        if (xhi2) w.v[N - 1] = w.v[N - 2];
    };
    auto fixCw = [&](W& w) {
        if (clo1) w.v[0] = w.v[1];
        if (chi1) w.v[N + 1] = w.v[N];
    };

    const float hx = a.hx, hy = a.hy, hz = a.hz;
    const float rhx = 1.0f / hx, rhy = 1.0f / hy, rhz = 1.0f / hz;      // FAST only

    const int srow = trow * TZ + N * zl;

    // ---- the march.  The rows of the three planes in use live in REGISTERS and rotate with the march (r02b): a plane's
    // rows j-1, j, j+1 of U (with their z neighbours), C00, C01 (rows j-1, j+1) and C02 are read from shared memory ONCE,
    // when the plane arrives, and serve as plane i+1, then i, then i-1 - 27 shared-memory loads per thread and plane
    // instead of 44 (ncu of the first version: short_scoreboard / mio_throttle led at one CTA per SM).  The loop is
    // unrolled by three so that the rotation is a renaming of registers, not a copy.
    W xw[3][3];            // [plane slot][row j-1, j, j+1]: U0, z-edge patched
    V c00[3];              // [plane slot] row j
    V c01[3][2];           // [plane slot][row j-1, j+1]
    W c02[3];              // [plane slot] row j with z neighbours
    auto load_plane = [&](auto slot_c, int j) {       // logical plane j (already waited for) -> register slot
        constexpr int SL = decltype(slot_c)::value;
        const float* Uj = slot(j, F_U);
        xw[SL][0] = ldw<N>(Uj + rm); xw[SL][1] = ldw<N>(Uj + rc); xw[SL][2] = ldw<N>(Uj + rp);
        c00[SL] = ldv_s<N>(slot(j, F_C00) + cc);
        c01[SL][0] = ldv_s<N>(slot(j, F_C01) + cm); c01[SL][1] = ldv_s<N>(slot(j, F_C01) + cp);
        c02[SL] = ldw<N>(slot(j, F_C02) + cc);
        if (zb_tile) { fixXw(xw[SL][0]); fixXw(xw[SL][1]); fixXw(xw[SL][2]); fixCw(c02[SL]); }
    };
    auto body = [&](auto role, int i) {
        constexpr int SM_ = decltype(role)::value, SC_ = (SM_ + 1) % 3, SP_ = (SM_ + 2) % 3;   // slots of planes i-1, i, i+1
        const int p = xb + i;
        const int jc = i + 1;
        wait_plane(i + 2);
        V stc[NS > 0 ? NS : 1];
        if (NS > 0) {
            mbar_wait(&sfull[i % SSTAGES], (uint32_t)((i / SSTAGES) & 1));
#pragma unroll
            for (int s = 0; s < NS; ++s) stc[s] = ldv_s<N>(sslot(i, s) + srow);
        }
        if (act) {
            const size_t gi = ((size_t)p * a.ny + gy) * a.nz + z;
            load_plane(std::integral_constant<int, SP_>{}, jc + 1);
            const W& x_mc = xw[SM_][1]; const W& x_pc = xw[SP_][1];
            const W& x_cm = xw[SC_][0]; const W& x_cc = xw[SC_][1]; const W& x_cp = xw[SC_][2];
            const W& x_mm = xw[SM_][0]; const W& x_mp = xw[SM_][2]; const W& x_pm = xw[SP_][0]; const W& x_pp = xw[SP_][2];
            const float* Uc = slot(jc, F_U);
            V raw;
#pragma unroll
            for (int e = 0; e < N; ++e) raw.v[e] = x_cc.v[e + 1];
            if (MODEL != PDX_MODEL_NONE) {
                if (p < a.xclo || p > a.xchi) raw = ldv_s<N>(a.Uin + gi);     // CTA uniform, two planes of the grid
                else if (row_clamped || zb_tile) raw = ldv_s<N>(Uc + cc);     // the unclamped row, z edges unpatched
            }
            // ---- tensor components that are only read in the centre plane ----
            const V c11m = ldv_s<N>(slot(jc, F_C11) + cm), c11c = ldv_s<N>(slot(jc, F_C11) + cc), c11p = ldv_s<N>(slot(jc, F_C11) + cp);
            W c22 = ldw<N>(slot(jc, F_C22) + cc);
            W c12m = ldw<N>(slot(jc, F_C12) + cm), c12p = ldw<N>(slot(jc, F_C12) + cp);
            if (zb_tile) { fixCw(c22); fixCw(c12m); fixCw(c12p); }
            const V& c00m = c00[SM_]; const V& c00c = c00[SC_]; const V& c00p = c00[SP_];
            const V& c01mm = c01[SM_][0]; const V& c01mp = c01[SM_][1]; const V& c01pm = c01[SP_][0]; const V& c01pp = c01[SP_][1];
            const W& c02m = c02[SM_]; const W& c02p = c02[SP_];

            V out;
#pragma unroll
            for (int e = 0; e < N; ++e) {
                const int k0 = e, k1 = e + 1, k2 = e + 2;          // window indices of z-1, z, z+1
                const float x0 = x_cc.v[k1];
                // axis x, side +/-: normal, transverse y, transverse z     (:50-62)
                float tp = t_main<MATH>(c00p.v[e], c00c.v[e], x_pc.v[k1], x0, hx, rhx);
                tp = madd<MATH>(tp, t_cross<MATH>(c01pp.v[e], c01pm.v[e], x_pp.v[k1], x_pm.v[k1], hy, rhy));
                tp = madd<MATH>(tp, t_cross<MATH>(c02p.v[k2], c02p.v[k0], x_pc.v[k2], x_pc.v[k0], hz, rhz));
                float tm = t_main<MATH>(c00m.v[e], c00c.v[e], x_mc.v[k1], x0, hx, rhx);
                tm = madd<MATH>(tm, t_cross<MATH>(c01mp.v[e], c01mm.v[e], x_mp.v[k1], x_mm.v[k1], hy, rhy));
                tm = madd<MATH>(tm, t_cross<MATH>(c02m.v[k2], c02m.v[k0], x_mc.v[k2], x_mc.v[k0], hz, rhz));
                const float gx = mdivh<MATH>(msub<MATH>(mmul<MATH>(0.5f, tp), mmul<MATH>(-0.5f, tm)), hx, rhx);
                // axis y: transverse x, normal, transverse z               (:64-76)
                float up = t_cross<MATH>(c01pp.v[e], c01mp.v[e], x_pp.v[k1], x_mp.v[k1], hx, rhx);
                up = madd<MATH>(up, t_main<MATH>(c11p.v[e], c11c.v[e], x_cp.v[k1], x0, hy, rhy));
                up = madd<MATH>(up, t_cross<MATH>(c12p.v[k2], c12p.v[k0], x_cp.v[k2], x_cp.v[k0], hz, rhz));
                float um = t_cross<MATH>(c01pm.v[e], c01mm.v[e], x_pm.v[k1], x_mm.v[k1], hx, rhx);
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
            stv_g<N>(a.Uout + gi, out);
            if (a.mirror_lo && p == a.mirror_lo_plane) stv_g<N>(a.mirror_lo + (size_t)gy * a.nz + z, out);
            if (a.mirror_hi && p == a.mirror_hi_plane) stv_g<N>(a.mirror_hi + (size_t)gy * a.nz + z, out);
#pragma unroll
            for (int s = 0; s < NS; ++s) stv_g<N>(a.st[s] + gi, stc[s]);
        }
        __syncthreads();                              // everyone is done with logical plane i
        if (tid == 0) {
            if (i + STAGES < nlog) issue(i + STAGES);
            if (NS > 0 && i + SSTAGES < np) issue_states(i + SSTAGES);
        }
    };
    if (np > 0) {
        wait_plane(0);
        wait_plane(1);
        if (act) {
            load_plane(std::integral_constant<int, 0>{}, 0);
            load_plane(std::integral_constant<int, 1>{}, 1);
        }
    }
    for (int i = 0; i < np; i += 3) {
        body(std::integral_constant<int, 0>{}, i);
        if (i + 1 < np) body(std::integral_constant<int, 1>{}, i + 1);
        if (i + 2 < np) body(std::integral_constant<int, 2>{}, i + 2);
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

template <int MODEL, int MATH, int N>
int launch_inst(const FdArgs& a0, const AnisoMaps& m, cudaStream_t s) {
    auto kern = fd_aniso_kernel<MODEL, MATH, N>;
    static bool attr_done[PDX_MAX_DEVICES] = {};
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (dev < 0 || dev >= PDX_MAX_DEVICES || !attr_done[dev]) {
        if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(NStates<MODEL>::value)),
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
    kern<<<(unsigned)nb, TY * TZ / N, smem_bytes(NStates<MODEL>::value), s>>>(m, a);
    count_launch();
    return check_cuda(cudaGetLastError(), "fd_aniso_kernel launch");
}

template <int MODEL>
int launch_math(int math, const FdArgs& a, const AnisoMaps& m, cudaStream_t s) {
    // nodes per thread (PDX_ANISO_NPT=4|2 overrides).  Measured on 256 x 512^2 FK (profiles/r02_aniso_throughput.jsonl):
    // FAST 97.5 Gnode-steps/s with 4 and with 2 alike; EXACT (IEEE divisions: issue bound) 8.3 against 13.6 - more warps
    // pay only there.
    static const int npt_env = [] { const char* e = getenv("PDX_ANISO_NPT"); return e ? atoi(e) : 0; }();
    if (math == PDX_MATH_EXACT)
        return npt_env == 4 ? launch_inst<MODEL, PDX_MATH_EXACT, 4>(a, m, s) : launch_inst<MODEL, PDX_MATH_EXACT, 2>(a, m, s);
    return npt_env == 2 ? launch_inst<MODEL, PDX_MATH_FAST, 2>(a, m, s) : launch_inst<MODEL, PDX_MATH_FAST, 4>(a, m, s);
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
    int chunk = 64;             // tools/chunk_variants2.py, 256 x 512^2 FK: 85.2 / 92.7 / 97.0 / 99.4 / 89.7 Gnode-steps/s at 8 / 16 / 32 / 64 / 128
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
                         const CUtensorMap* const* mapS, cudaStream_t s) {
    AnisoMaps m;
    m.m[F_U] = mapU;
    for (int f = 1; f < NF; ++f) m.m[f] = mapC[f - 1];
    for (int i = 0; i < PDX_MAX_STATES; ++i) m.st[i] = (mapS && mapS[i]) ? *mapS[i] : mapU;
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
