// pdx_ionic_lut.cu - the reference's two lookup-table ionic models on sm_100a:
//   Courtemanche-Ramirez-Nattel 1998 (gpuSolve/ionic/courtemanche_ramirez_nattel.py:423-514)
//   ten Tusscher-Panfilov 2006      (gpuSolve/ionic/ten_tusscher_panfilov.py:405-523)
// as (a) the pointwise IonicModel.differentiate step (+ the FEM RHS0) and (b) fused with the
// finite-difference stencil and the Euler update (one pass over U and the 19/22 state arrays).
//
// TILE kernels (lut_fd_tile_kernel / lut_ionic_tile_kernel; nz % 4 == 0, 16-byte aligned arrays): a CTA owns 512
// consecutive nodes.  All 19 / 22 state rows of the tile are fetched by the copy engine (1-D cp.async.bulk into
// shared memory on an mbarrier: 38-44 KB in flight per CTA, four CTAs per SM, no register or address-arithmetic
// cost), updated IN shared memory (conflict-free 4-byte accesses, a thread takes nodes t, t+128, t+256, t+384) and
// written back by bulk stores.  While the state rows are in flight every thread evaluates the stencil of four
// consecutive nodes with 128-bit loads and parks centre value / coef*lap / raw potential in shared memory.  Table
// rows are read as 128-bit words (a V row is one 128-byte line) and interpolated once per node.
// SCALAR kernels (any shape): one thread per node, states through 4-byte global accesses.
// The three tables (0.5-12 MB) stay L2/L1 resident - neighbouring nodes hit the same two rows.
//
// THIS FILE IS COMPILED TWICE (csrc/build.py), once per arithmetic policy (PDX_LUT_TU_MATH = PDX_MATH_*):
//   EXACT  -fmad=false: every expression is evaluated op-for-op in fp32 in the order the reference writes
//          it (IEEE division, no contraction) and exp/log/expm1 return the correctly rounded fp32 value
//          (fp64 evaluation, one rounding) - the oracle's convention, results bit-comparable with
//          oracle/ionic_lut.py.  ~1400 instructions per node: the kernel is issue/latency bound.
//   FAST   same formulas, FMA contraction allowed, approximate division (-prec-div=false, 2 ulp) and
//          CUDA's fp32 expf/logf/expm1f.  rel-L2 <= 1e-5 against the oracle (tests/test_ref_golden_gpu.py).
#include <cstdlib>
#include "pdx_stencil.cuh"

#ifndef PDX_LUT_TU_MATH
#define PDX_LUT_TU_MATH PDX_MATH_EXACT
#endif

namespace pdx {
namespace {

constexpr int TU_MATH = PDX_LUT_TU_MATH;

template <int MATH> __device__ __forceinline__ float exp_m(float x) {
    return MATH == PDX_MATH_EXACT ? (float)exp((double)x) : expf(x);
}
template <int MATH> __device__ __forceinline__ float log_m(float x) {
    return MATH == PDX_MATH_EXACT ? (float)log((double)x) : logf(x);
}
template <int MATH> __device__ __forceinline__ float expm1_m(float x) {
    return MATH == PDX_MATH_EXACT ? (float)expm1((double)x) : expm1f(x);
}

// Linear interpolation in a uniform table (courtemanche_ramirez_nattel.py:241-252):
// idx = int((clip(X) - mn)*step), clamped to [0, mx_idx-1]; w = (X - lower)/res.
struct Row {
    const float* lo;
    const float* hi;
    float w, omw;
    __device__ __forceinline__ float operator()(int col) const { return omw * __ldg(lo + col) + w * __ldg(hi + col); }
    // Rush-Larsen update A + B*y with (A,B) in columns col, col+1 (col even)
    __device__ __forceinline__ float rl(int col, float y) const {
        const float2 l = __ldg(reinterpret_cast<const float2*>(lo + col));
        const float2 h = __ldg(reinterpret_cast<const float2*>(hi + col));
        const float A = omw * l.x + w * h.x;
        const float B = omw * l.y + w * h.y;
        return A + B * y;
    }
};

__device__ __forceinline__ Row lookup(const LutTab& T, float X) {
    const float Xc = fminf(fmaxf(X, T.mn), T.mx);
    const int idx = (int)((Xc - T.mn) * T.step);
    const int lo = min(max(idx, 0), T.mx_idx - 1);
    const float pos = (float)lo * T.res + T.mn;
    const float w = (Xc - pos) / T.res;
    const float* p = T.t + (size_t)lo * T.stride;
    return Row{p, p + T.stride, w, 1.0f - w};
}

// The same interpolation with the whole row pair fetched up front in the widest aligned words the row stride allows
// (W = 4: 128-bit, 2: 64-bit, 1: 32-bit) and every column interpolated once: v[c] = omw*lo[c] + w*hi[c].
template <int STRIDE, int W>
struct IRow {
    float v[STRIDE];
    __device__ __forceinline__ float operator()(int col) const { return v[col]; }
    __device__ __forceinline__ float rl(int col, float y) const { return v[col] + v[col + 1] * y; }
};
template <int STRIDE, int W>
__device__ __forceinline__ IRow<STRIDE, W> lookup_row(const LutTab& T, float X) {
    const float Xc = fminf(fmaxf(X, T.mn), T.mx);
    const int idx = (int)((Xc - T.mn) * T.step);
    const int lo = min(max(idx, 0), T.mx_idx - 1);
    const float pos = (float)lo * T.res + T.mn;
    const float w = (Xc - pos) / T.res;
    const float omw = 1.0f - w;
    const float* p = T.t + (size_t)lo * STRIDE;
    IRow<STRIDE, W> r;
    if (W == 4) {
#pragma unroll
        for (int q = 0; q < STRIDE / 4; ++q) {
            const float4 l = __ldg(reinterpret_cast<const float4*>(p) + q);
            const float4 h = __ldg(reinterpret_cast<const float4*>(p + STRIDE) + q);
            r.v[4 * q + 0] = omw * l.x + w * h.x;
            r.v[4 * q + 1] = omw * l.y + w * h.y;
            r.v[4 * q + 2] = omw * l.z + w * h.z;
            r.v[4 * q + 3] = omw * l.w + w * h.w;
        }
    } else if (W == 2) {
#pragma unroll
        for (int q = 0; q < STRIDE / 2; ++q) {
            const float2 l = __ldg(reinterpret_cast<const float2*>(p) + q);
            const float2 h = __ldg(reinterpret_cast<const float2*>(p + STRIDE) + q);
            r.v[2 * q + 0] = omw * l.x + w * h.x;
            r.v[2 * q + 1] = omw * l.y + w * h.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < STRIDE; ++q) r.v[q] = omw * __ldg(p + q) + w * __ldg(p + STRIDE + q);
    }
    return r;
}

// how a kernel reads the tables: column by column (any stride) or as whole rows (canonical strides of _lut.py)
struct ScalarRows {
    template <int MODEL, int TAB> static __device__ __forceinline__ Row get(const LutTab& T, float X) { return lookup(T, X); }
};
struct VectorRows {
    // CRN: V 32 floats (128-bit words), Cai 6 (64-bit), fn 4 (128-bit); TTP: V 28 (128-bit), CaSS 2 (64-bit), V-Ek 1
    template <int MODEL, int TAB> static __device__ __forceinline__ auto get(const LutTab& T, float X) {
        if constexpr (MODEL == PDX_MODEL_CRN) {
            if constexpr (TAB == 0) return lookup_row<32, 4>(T, X);
            else if constexpr (TAB == 1) return lookup_row<6, 2>(T, X);
            else return lookup_row<4, 4>(T, X);
        } else {
            if constexpr (TAB == 0) return lookup_row<28, 4>(T, X);
            else if constexpr (TAB == 1) return lookup_row<2, 2>(T, X);
            else return lookup_row<1, 1>(T, X);
        }
    }
};

// where a kernel keeps the state of a node: SoA global arrays, or the CTA's shared-memory tile
struct GlobalStates {
    float* const* st;
    size_t n;
    __device__ __forceinline__ float ld(int s) const { return st[s][n]; }
    __device__ __forceinline__ void sto(int s, float v) const { st[s][n] = v; }
};
template <int TILE>
struct TileStates {
    float* base;                       // &tile[0][node]
    __device__ __forceinline__ float ld(int s) const { return base[s * TILE]; }
    __device__ __forceinline__ void sto(int s, float v) const { base[s * TILE] = v; }
};

// ---------------------------------------------------------------------------- CRN
// constants (host-folded in float64, include/pdx.h)
enum { C_RTF = 0, C_KO, C_KACH, C_GTO, C_FGKUR, C_GKR, C_GKS, C_GK1, C_FGTR, C_TAU_TR, C_FGREL, C_KREL, C_IUPR,
       C_DCAREL, C_KMCSQN, C_FVOLI, C_DCAUP, C_B1D, C_B1E, C_FN1, C_FN2, C_BFCA, C_BU, C_NCONST };
// V-table columns
enum { CV_m = 0, CV_h = 2, CV_j = 4, CV_oa = 6, CV_oi = 8, CV_ua = 10, CV_ui = 12, CV_xr = 14, CV_xs = 16, CV_d = 18,
       CV_f = 20, CV_w = 22, CV_GKur = 24, CV_INaK, CV_IbNa, CV_ICaL, CV_NaCa_a, CV_NaCa_b, CV_IbCa, CV_INa };
enum { CC_Iup = 0, CC_buf, CC_IpCa, CC_conCa, CC_fCa_A };
enum { CF_u_A = 0, CF_v = 1 };   // v_A, v_B in columns 1,2 (odd start: read as scalars)
// state order
enum { S_Ca_rel = 0, S_Ca_up, S_Cai, S_Ki, S_d, S_f, S_f_Ca, S_h, S_j, S_m, S_oa, S_oi, S_u, S_ua, S_ui, S_v, S_w,
       S_xr, S_xs, CRN_NS };

template <int MATH, class ROWS, class ST>
__device__ __forceinline__ float crn_update(const LutArgs& L, const ST S, float V) {
    const float* c = L.c;
    const float dt = L.dt;
    float y[CRN_NS];
#pragma unroll
    for (int s = 0; s < CRN_NS; ++s) y[s] = S.ld(s);
    const auto vr = ROWS::template get<PDX_MODEL_CRN, 0>(L.tab[0], V);
    const auto cr = ROWS::template get<PDX_MODEL_CRN, 1>(L.tab[1], y[S_Cai]);

    const float E_K = c[C_RTF] * log_m<MATH>(c[C_KO] / y[S_Ki]);
    const float VmEK = V - E_K;
    const float ICaL = vr(CV_ICaL) * y[S_d] * y[S_f] * y[S_f_Ca];
    const float IKACh = (c[C_KACH] * (0.0517f + 0.4516f / (1.0f + exp_m<MATH>((V + 59.53f) / 17.18f)))) * VmEK;
    const float INa = vr(CV_INa) * y[S_m] * y[S_m] * y[S_m] * y[S_h] * y[S_j];
    const float INaCa = vr(CV_NaCa_a) - y[S_Cai] * vr(CV_NaCa_b);
    const float IK1 = (c[C_GK1] * VmEK) / (1.0f + exp_m<MATH>(0.07f * (V + 80.0f)));
    const float IKr = ((c[C_GKR] * VmEK) / (1.0f + exp_m<MATH>((V + 15.0f) / 22.4f))) * y[S_xr];
    const float IKs = (c[C_GKS] * VmEK) * y[S_xs] * y[S_xs];
    const float IKur = ((c[C_FGKUR] * vr(CV_GKur)) * VmEK) * y[S_ua] * y[S_ua] * y[S_ua] * y[S_ui];
    const float IpCa = cr(CC_IpCa) + vr(CV_IbCa);
    const float Ito = (c[C_GTO] * VmEK) * y[S_oa] * y[S_oa] * y[S_oa] * y[S_oi];
    const float INaK = vr(CV_INaK);
    const float Iion = INa + IK1 + Ito + IKur + IKr + IKs + ICaL + IpCa + INaCa + vr(CV_IbNa) + INaK + IKACh;

    const float Itr = (c[C_FGTR] * (y[S_Ca_up] - y[S_Ca_rel])) / c[C_TAU_TR];
    const float Irel = c[C_FGREL] * y[S_u] * y[S_u] * y[S_v] * c[C_KREL] * y[S_w] * (y[S_Ca_rel] - cr(CC_conCa));
    const float dIups = cr(CC_Iup) - c[C_IUPR] * y[S_Ca_up];
    const float csq = y[S_Ca_rel] + c[C_KMCSQN];
    const float d_Ca_rel = (Itr - Irel) / (1.0f + (c[C_DCAREL] / csq) / csq);
    const float d_Ki = (-(((Ito + IKr + IKur + IKs + IK1 + IKACh) - 2.0f * INaK))) / c[C_FVOLI];
    const float d_Ca_up = dIups - Itr * c[C_DCAUP];
    const float d_Cai = ((c[C_B1D] * ((INaCa + INaCa) - IpCa - ICaL)) - (c[C_B1E] * dIups) + Irel) / cr(CC_buf);
    const float fn = (c[C_FN1] * Irel) - (c[C_FN2] * (ICaL - 0.4f * INaCa));
    const auto fr = ROWS::template get<PDX_MODEL_CRN, 2>(L.tab[2], fn);

    S.sto(S_Ca_rel, y[S_Ca_rel] + d_Ca_rel * dt);
    S.sto(S_Ca_up, y[S_Ca_up] + d_Ca_up * dt);
    S.sto(S_Cai, y[S_Cai] + d_Cai * dt);
    S.sto(S_Ki, y[S_Ki] + d_Ki * dt);
    S.sto(S_d, vr.rl(CV_d, y[S_d]));
    S.sto(S_f, vr.rl(CV_f, y[S_f]));
    S.sto(S_f_Ca, cr(CC_fCa_A) + c[C_BFCA] * y[S_f_Ca]);
    S.sto(S_h, vr.rl(CV_h, y[S_h]));
    S.sto(S_j, vr.rl(CV_j, y[S_j]));
    S.sto(S_m, vr.rl(CV_m, y[S_m]));
    S.sto(S_oa, vr.rl(CV_oa, y[S_oa]));
    S.sto(S_oi, vr.rl(CV_oi, y[S_oi]));
    S.sto(S_u, fr(CF_u_A) + c[C_BU] * y[S_u]);
    S.sto(S_ua, vr.rl(CV_ua, y[S_ua]));
    S.sto(S_ui, vr.rl(CV_ui, y[S_ui]));
    S.sto(S_v, fr(CF_v) + fr(CF_v + 1) * y[S_v]);
    S.sto(S_w, vr.rl(CV_w, y[S_w]));
    S.sto(S_xr, vr.rl(CV_xr, y[S_xr]));
    S.sto(S_xs, vr.rl(CV_xs, y[S_xs]));
    return -Iion;
}

// ---------------------------------------------------------------------------- TTP
enum { T_RTONF = 0, T_HRTONF, T_CAO, T_KO, T_KOPK, T_PKNA, T_NAO, T_GPCA, T_KPCA, T_FCONST, T_FRT, T_HFRT, T_SQRTKO,
       T_GNA, T_PMFNAK, T_KMNA, T_GBCA, T_GBNA, T_GPK, T_GK1, T_VLEAK, T_VMAXUP, T_KUP2, T_VXFER, T_INVVCFCM, T_MAXSR,
       T_DSR, T_EC, T_INVVCF2, T_CAP, T_VSRVC, T_BUFCK, T_KBUFC, T_K4, T_K2, T_K1, T_K3, T_VREL, T_BUFSRK, T_KBUFSR,
       T_VCVSS, T_VSRVSS, T_INVVSSF2, T_BUFSSK, T_KBUFSS, T_NCONST };
enum { TV_D = 0, TV_F2 = 2, TV_F = 4, TV_H = 6, TV_J = 8, TV_M = 10, TV_R = 12, TV_S = 14, TV_Xr1 = 16, TV_Xr2 = 18,
       TV_Xs = 20, TV_a2 = 22, TV_NaCa_A, TV_NaCa_B, TV_rec_iNaK, TV_rec_ipK };
enum { Q_GCaL = 0, Q_GKr, Q_GKs, Q_Gto, Q_CaSR, Q_CaSS, Q_Cai, Q_F, Q_F2, Q_FCaSS, Q_H, Q_J, Q_Ki, Q_M, Q_Nai, Q_R,
       Q_R_bar, Q_S, Q_Xr1, Q_Xr2, Q_Xs, Q_D, TTP_NS };

template <int MATH, class ROWS, class ST>
__device__ __forceinline__ float ttp_update(const LutArgs& L, const ST S, float V) {
    const float* c = L.c;
    const float dt = L.dt;
    float y[TTP_NS];
#pragma unroll
    for (int s = 0; s < TTP_NS; ++s) y[s] = S.ld(s);
    const float Cai = y[Q_Cai], CaSS = y[Q_CaSS], CaSR = y[Q_CaSR], Nai = y[Q_Nai], Ki = y[Q_Ki], R_bar = y[Q_R_bar];
    const auto vr = ROWS::template get<PDX_MODEL_TTP, 0>(L.tab[0], V);
    const auto cr = ROWS::template get<PDX_MODEL_TTP, 1>(L.tab[1], CaSS);

    const float Eca = c[T_HRTONF] * log_m<MATH>(c[T_CAO] / Cai);
    const float Ek = c[T_RTONF] * log_m<MATH>(c[T_KO] / Ki);
    const float Eks = c[T_RTONF] * log_m<MATH>(c[T_KOPK] / (Ki + c[T_PKNA] * Nai));
    const float Ena = c[T_RTONF] * log_m<MATH>(c[T_NAO] / Nai);
    const float IpCa = (c[T_GPCA] * Cai) / (c[T_KPCA] + Cai);
    const float v15 = V - 15.0f;
    const float a1 = (y[Q_GCaL] * c[T_FCONST] * c[T_FRT] * 4.0f) *
                     (fabsf(v15) < 1e-10f ? c[T_HFRT] : v15 / expm1_m<MATH>(2.0f * v15 * c[T_FRT]));
    const float ICaL_A = a1 * vr(TV_a2);
    const float ICaL_B = a1 * c[T_CAO];
    const float IKr = y[Q_GKr] * c[T_SQRTKO] * y[Q_Xr1] * y[Q_Xr2] * (V - Ek);
    const float IKs = y[Q_GKs] * y[Q_Xs] * y[Q_Xs] * (V - Eks);
    const float INa = c[T_GNA] * y[Q_M] * y[Q_M] * y[Q_M] * y[Q_H] * y[Q_J] * (V - Ena);
    const float INaK = c[T_PMFNAK] * (Nai / (Nai + c[T_KMNA])) * vr(TV_rec_iNaK);
    const float IbCa = c[T_GBCA] * (V - Eca);
    const float IbNa = c[T_GBNA] * (V - Ena);
    const float IpK = c[T_GPK] * vr(TV_rec_ipK) * (V - Ek);
    const float Ito = y[Q_Gto] * y[Q_R] * y[Q_S] * (V - Ek);
    const auto er = ROWS::template get<PDX_MODEL_TTP, 2>(L.tab[2], V - Ek);
    const float ICaL = y[Q_D] * y[Q_F] * y[Q_F2] * y[Q_FCaSS] * (ICaL_A * CaSS - ICaL_B);
    const float INaCa = vr(TV_NaCa_A) * Nai * Nai * Nai - vr(TV_NaCa_B) * Cai;
    const float IK1 = c[T_GK1] * er(0) * (V - Ek);
    const float Iion = IKr + IKs + IK1 + Ito + INa + IbNa + ICaL + IbCa + INaK + INaCa + IpCa + IpK;

    const float Ileak = c[T_VLEAK] * (CaSR - Cai);
    const float Iup = c[T_VMAXUP] / (1.0f + c[T_KUP2] / (Cai * Cai));
    const float Ixfer = c[T_VXFER] * (CaSS - Cai);
    const float d_Ki = (-(IK1 + Ito + IKr + IKs - 2.0f * INaK + IpK)) * c[T_INVVCFCM];
    const float d_Nai = (-(INa + IbNa + 3.0f * INaK + 3.0f * INaCa)) * c[T_INVVCFCM];
    const float ecr = c[T_EC] / CaSR;
    const float kCaSR = c[T_MAXSR] - c[T_DSR] / (1.0f + ecr * ecr);
    const float kbc = c[T_KBUFC] + Cai;
    const float d_Cai = ((Ixfer - ((IbCa + IpCa - 2.0f * INaCa) * c[T_INVVCF2] * c[T_CAP])) -
                         ((Iup - Ileak) * c[T_VSRVC])) / (1.0f + (c[T_BUFCK] / kbc) / kbc);
    const float d_R_bar = c[T_K4] * (1.0f - R_bar) - c[T_K2] * kCaSR * CaSS * R_bar;
    const float k1 = c[T_K1] / kCaSR;
    const float O = (k1 * CaSS * CaSS * R_bar) / (c[T_K3] + k1 * CaSS * CaSS);
    const float Irel = c[T_VREL] * O * (CaSR - CaSS);
    const float ksr = CaSR + c[T_KBUFSR];
    const float d_CaSR = (Iup - Irel - Ileak) / (1.0f + (c[T_BUFSRK] / ksr) / ksr);
    const float kss = CaSS + c[T_KBUFSS];
    const float d_CaSS = (-Ixfer * c[T_VCVSS] + Irel * c[T_VSRVSS] + (-ICaL * c[T_INVVSSF2] * c[T_CAP])) /
                         (1.0f + (c[T_BUFSSK] / kss) / kss);

    S.sto(Q_CaSR, CaSR + d_CaSR * dt);
    S.sto(Q_CaSS, CaSS + d_CaSS * dt);
    S.sto(Q_Cai, Cai + d_Cai * dt);
    S.sto(Q_Ki, Ki + d_Ki * dt);
    S.sto(Q_Nai, Nai + d_Nai * dt);
    S.sto(Q_R_bar, R_bar + d_R_bar * dt);
    S.sto(Q_D, vr.rl(TV_D, y[Q_D]));
    S.sto(Q_F, vr.rl(TV_F, y[Q_F]));
    S.sto(Q_F2, vr.rl(TV_F2, y[Q_F2]));
    S.sto(Q_FCaSS, cr.rl(0, y[Q_FCaSS]));
    S.sto(Q_H, vr.rl(TV_H, y[Q_H]));
    S.sto(Q_J, vr.rl(TV_J, y[Q_J]));
    S.sto(Q_M, vr.rl(TV_M, y[Q_M]));
    S.sto(Q_R, vr.rl(TV_R, y[Q_R]));
    S.sto(Q_S, vr.rl(TV_S, y[Q_S]));
    S.sto(Q_Xr1, vr.rl(TV_Xr1, y[Q_Xr1]));
    S.sto(Q_Xr2, vr.rl(TV_Xr2, y[Q_Xr2]));
    S.sto(Q_Xs, vr.rl(TV_Xs, y[Q_Xs]));
    return -Iion;
}

template <int MODEL, int MATH, class ROWS, class ST>
__device__ __forceinline__ float lut_update(const LutArgs& L, const ST S, float V) {
    return MODEL == PDX_MODEL_CRN ? crn_update<MATH, ROWS>(L, S, V) : ttp_update<MATH, ROWS>(L, S, V);
}

// ---------------------------------------------------------------------------- kernels
// IonicModel.differentiate (+ RHS0 = U + dt*(dU + I0) of MonodomainSolver.solve_step)
template <int MODEL, int MATH>
__global__ void __launch_bounds__(128, 8) lut_ionic_kernel(const __grid_constant__ LutArgs L, int64_t n, const float* U,
                                                        float* dU, const float* I0, float* rhs0) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float V = U[i];
        const float d = lut_update<MODEL, MATH, ScalarRows>(L, GlobalStates{L.st, (size_t)i}, V);
        if (dU) dU[i] = d;
        if (rhs0) rhs0[i] = V + L.dt * (d + (I0 ? I0[i] : 0.0f));
    }
}

// fused explicit FD step: U1 = (U0 + dt*dU(U)) + coef*lap(U0)  (Tests/FD/Fenton/fenton.py:90-99)
template <int MODEL, int LAP, int MATH>
__global__ void __launch_bounds__(128, 8) lut_fd_kernel(const __grid_constant__ FdArgs a,
                                                     const __grid_constant__ LutArgs L) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = a.xb + blockIdx.z;
    if (k >= a.nz || j >= a.ny || i >= a.xe) return;
    const size_t idx = ((size_t)i * a.ny + j) * a.nz + k;
    const Field f = make_field(a);
    float c;
    const float lap = node_laplacian<LAP, LAP == PDX_LAP_ANISO ? PDX_MATH_EXACT : MATH>(a, f, i, j, k, c);
    const float dU = lut_update<MODEL, MATH, ScalarRows>(L, GlobalStates{L.st, idx}, a.Uin[idx]);      // raw U, not U0
    const float u_new = (c + a.dt * dU) + a.coef * lap;
    a.Uout[idx] = u_new;
    if (a.mirror_lo && i == a.mirror_lo_plane) a.mirror_lo[(size_t)j * a.nz + k] = u_new;
    if (a.mirror_hi && i == a.mirror_hi_plane) a.mirror_hi[(size_t)j * a.nz + k] = u_new;
}

// ---------------------------------------------------------------------------- tile kernels
constexpr int TT = 128;            // threads per CTA
constexpr int NPT = 4;             // nodes per thread
constexpr int TILE = TT * NPT;     // nodes per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// copy engine: 1-D bulk copies global <-> shared (sizes and addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int MODEL> struct LutNS { static constexpr int value = MODEL == PDX_MODEL_CRN ? (int)CRN_NS : (int)TTP_NS; };
// TTP's first four "states" are the per-node conductances GCaL, GKr, GKs, Gto: read, never written
template <int MODEL> struct LutRO { static constexpr int value = MODEL == PDX_MODEL_CRN ? 0 : 4; };

template <int MODEL> constexpr int tile_smem_bytes(int extra_rows) { return (LutNS<MODEL>::value + extra_rows) * TILE * 4 + 16; }

// warp 0 starts the bulk loads of every state row of the tile (one lane per row)
template <int MODEL>
__device__ __forceinline__ void tile_issue_loads(const LutArgs& L, float* sst, uint64_t* bar, size_t g0, int cnt) {
    constexpr int NS = LutNS<MODEL>::value;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)(NS * cnt * 4));
        __syncwarp();
        if (tid < NS) bulk_load(sst + tid * TILE, L.st[tid] + g0, (uint32_t)(cnt * 4), bar);
    }
}
template <int MODEL>
__device__ __forceinline__ void tile_issue_stores(const LutArgs& L, float* sst, size_t g0, int cnt, float* extra_dst,
                                                  const float* extra_src) {
    constexpr int NS = LutNS<MODEL>::value;
    fence_async_smem();                 // this thread's shared-memory writes -> visible to the copy engine
    __syncthreads();
    const int tid = threadIdx.x;
    if (tid < 32) {
        bool any = false;
        if (tid >= LutRO<MODEL>::value && tid < NS) { bulk_store(L.st[tid] + g0, sst + tid * TILE, (uint32_t)(cnt * 4)); any = true; }
        if (tid == NS && extra_dst) { bulk_store(extra_dst, extra_src, (uint32_t)(cnt * 4)); any = true; }
        if (any) bulk_commit_wait();    // shared memory must outlive the reads of the copy engine
    }
}

// fused explicit FD step on a tile (see the header of this file)
template <int MODEL, int LAP, int MATH>
__global__ void __launch_bounds__(TT, 4) lut_fd_tile_kernel(const __grid_constant__ FdArgs a,
                                                            const __grid_constant__ LutArgs L) {
    constexpr int NS = LutNS<MODEL>::value;
    extern __shared__ __align__(128) float sm[];
    float* sst = sm;                    // [NS][TILE] state rows
    float* sc  = sm + NS * TILE;        // centre value of U0 = enforce_boundary(U)
    float* scl = sc + TILE;             // coef * lap(U0)
    float* sv  = scl + TILE;            // raw U, later the new potential
    uint64_t* bar = reinterpret_cast<uint64_t*>(sv + TILE);
    const int tid = threadIdx.x;
    const long long plane = (long long)a.ny * a.nz;
    const int i = a.xb + blockIdx.y;
    const long long t0 = (long long)blockIdx.x * TILE;
    const int cnt = (int)min((long long)TILE, plane - t0);          // multiple of 4 (nz % 4 == 0)
    const size_t g0 = (size_t)i * plane + t0;
    tile_issue_loads<MODEL>(L, sst, bar, g0, cnt);

    // ---- stencil of four consecutive nodes while the state rows are in flight ----
    const int q = 4 * tid;
    if (q < cnt) {
        const long long pq = t0 + q;
        const int j = (int)(pq / a.nz), k = (int)(pq - (long long)j * a.nz);
        const float4 raw = __ldg(reinterpret_cast<const float4*>(a.Uin + g0 + q));
        float cc[4], ll[4];
        if (LAP == PDX_LAP_HOMOG && !a.dirichlet) {
            const int ic = clampi_(i, a.xclo, a.xchi), jc = clampi_(j, a.yclo, a.ychi);
            auto row = [&](int ii, int jj) { return a.Uin + ((size_t)ii * a.ny + jj) * a.nz; };
            const float* rc = row(ic, jc);
            auto ld4 = [&](const float* r, float* o) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(r + k));
                o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
                if (k < a.zclo) o[0] = o[1];                       // index clamp to [zclo, zchi] (k, nz multiples of 4)
                if (k + 3 > a.zchi) o[3] = o[2];
            };
            float c4[4], ym[4], yp[4], xm[4], xp[4];
            ld4(rc, c4);
            ld4(row(ic, clampi_(j - 1, a.yclo, a.ychi)), ym);
            ld4(row(ic, clampi_(j + 1, a.yclo, a.ychi)), yp);
            if (a.has_x) {
                ld4(row(clampi_(i - 1, a.xclo, a.xchi), jc), xm);
                ld4(row(clampi_(i + 1, a.xclo, a.xchi), jc), xp);
            }
            const float zl = k == 0 ? c4[0] : __ldg(rc + k - 1);
            const float zr = k + 4 >= a.nz ? c4[3] : __ldg(rc + k + 4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float c = c4[e];
                float lap = lap_axis<MATH>(ym[e], c, yp[e], a.hysq, a.rhysq);
                if (a.has_x) lap = madd<MATH>(lap_axis<MATH>(xm[e], c, xp[e], a.hxsq, a.rhxsq), lap);
                lap = madd<MATH>(lap, lap_axis<MATH>(e == 0 ? zl : c4[e - 1], c, e == 3 ? zr : c4[e + 1], a.hzsq, a.rhzsq));
                cc[e] = c;
                ll[e] = lap;
            }
        } else {
            const Field f = make_field(a);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                ll[e] = node_laplacian<LAP, LAP == PDX_LAP_ANISO ? PDX_MATH_EXACT : MATH>(a, f, i, j, k + e, cc[e]);
        }
        *reinterpret_cast<float4*>(sc + q) = make_float4(cc[0], cc[1], cc[2], cc[3]);
        *reinterpret_cast<float4*>(scl + q) = make_float4(a.coef * ll[0], a.coef * ll[1], a.coef * ll[2], a.coef * ll[3]);
        *reinterpret_cast<float4*>(sv + q) = raw;
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // ---- ionic update + Euler step, nodes t, t+128, ... of the tile, everything in shared memory ----
#pragma unroll 1
    for (int e = 0; e < NPT; ++e) {
        const int n = tid + e * TT;
        if (n < cnt) {
            const float dU = lut_update<MODEL, MATH, VectorRows>(L, TileStates<TILE>{sst + n}, sv[n]);   // raw U, not U0
            const float u_new = (sc[n] + a.dt * dU) + scl[n];
            sv[n] = u_new;
            if (a.mirror_lo && i == a.mirror_lo_plane) a.mirror_lo[t0 + n] = u_new;
            if (a.mirror_hi && i == a.mirror_hi_plane) a.mirror_hi[t0 + n] = u_new;
        }
    }
    tile_issue_stores<MODEL>(L, sst, g0, cnt, a.Uout + g0, sv);
}

// IonicModel.differentiate (+ RHS0) on a tile
template <int MODEL, int MATH>
__global__ void __launch_bounds__(TT, 4) lut_ionic_tile_kernel(const __grid_constant__ LutArgs L, int64_t n_total,
                                                               const float* U, float* dU, const float* I0, float* rhs0) {
    extern __shared__ __align__(128) float sm[];
    float* sst = sm;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + LutNS<MODEL>::value * TILE);
    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * TILE;
    const int cnt = (int)min((int64_t)TILE, n_total - t0);
    tile_issue_loads<MODEL>(L, sst, bar, (size_t)t0, cnt);
    float V[NPT], f0[NPT];
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
        const int n = tid + e * TT;
        V[e] = n < cnt ? U[t0 + n] : 0.0f;
        f0[e] = (n < cnt && I0) ? I0[t0 + n] : 0.0f;
    }
    mbar_wait(bar, 0);
#pragma unroll 1
    for (int e = 0; e < NPT; ++e) {
        const int n = tid + e * TT;
        if (n < cnt) {
            const float d = lut_update<MODEL, MATH, VectorRows>(L, TileStates<TILE>{sst + n}, V[e]);
            if (dU) dU[t0 + n] = d;
            if (rhs0) rhs0[t0 + n] = V[e] + L.dt * (d + f0[e]);
        }
    }
    tile_issue_stores<MODEL>(L, sst, (size_t)t0, cnt, nullptr, nullptr);
}

// the tile kernels need: canonical table strides (VectorRows), 16-byte aligned tables and arrays, 4 | n
template <int MODEL>
bool tile_tables_ok(const LutArgs& L) {
    static const int strides[2][3] = {{32, 6, 4}, {28, 2, 1}};
    const char* e = std::getenv("PDX_LUT_TILE");                   // A/B switch: PDX_LUT_TILE=0 keeps the scalar kernels
    if (e && e[0] == '0') return false;
    for (int t = 0; t < 3; ++t)
        if (L.tab[t].stride != strides[MODEL == PDX_MODEL_TTP][t] || ((uintptr_t)L.tab[t].t % 16) != 0) return false;
    for (int s = 0; s < LutNS<MODEL>::value; ++s)
        if (((uintptr_t)L.st[s] % 16) != 0) return false;
    return true;
}

template <class K>
int tile_smem_attr(K kern, int bytes) {
    if (bytes <= 48 * 1024) return 0;
    return check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes),
                      "cudaFuncSetAttribute(lut tile smem)");
}

template <int MODEL, int LAP, int MATH>
int launch_fd_tile(const FdArgs& a, const LutArgs& L, cudaStream_t s) {
    constexpr int bytes = tile_smem_bytes<MODEL>(3);
    auto kern = lut_fd_tile_kernel<MODEL, LAP, MATH>;
    if (tile_smem_attr(kern, bytes)) return 1;
    const long long plane = (long long)a.ny * a.nz;
    dim3 grid((unsigned)((plane + TILE - 1) / TILE), (unsigned)(a.xe - a.xb), 1);
    kern<<<grid, TT, bytes, s>>>(a, L);
    count_launch();
    return check_cuda(cudaGetLastError(), "lut_fd_tile_kernel launch");
}

template <int MODEL, int MATH>
int launch_fd_lap(int lap, const FdArgs& a, const LutArgs& L, cudaStream_t s) {
    if (a.xe <= a.xb) return 0;
    const bool tile = a.nz % 4 == 0 && a.nz >= 4 && (a.xe - a.xb) <= 65535 && tile_tables_ok<MODEL>(L) &&
                      ((uintptr_t)a.Uin % 16) == 0 && ((uintptr_t)a.Uout % 16) == 0 &&
                      (((uintptr_t)a.mirror_lo | (uintptr_t)a.mirror_hi) % 4) == 0;
    if (tile) {
        switch (lap) {
            case PDX_LAP_HOMOG:   return launch_fd_tile<MODEL, PDX_LAP_HOMOG, MATH>(a, L, s);
            case PDX_LAP_CONV:    return launch_fd_tile<MODEL, PDX_LAP_CONV, MATH>(a, L, s);
            case PDX_LAP_HETEROG: return launch_fd_tile<MODEL, PDX_LAP_HETEROG, MATH>(a, L, s);
            case PDX_LAP_ANISO:   return launch_fd_tile<MODEL, PDX_LAP_ANISO, MATH>(a, L, s);
            default: set_error("unknown lap_kind"); return 1;
        }
    }
    dim3 block(32, 4, 1);
    dim3 grid((a.nz + 31) / 32, (a.ny + 3) / 4, a.xe - a.xb);
    if (grid.z > 65535u || grid.y > 65535u) { set_error("LUT-model FD kernel: grid too large"); return 1; }
    switch (lap) {
        case PDX_LAP_HOMOG:   lut_fd_kernel<MODEL, PDX_LAP_HOMOG, MATH><<<grid, block, 0, s>>>(a, L); break;
        case PDX_LAP_CONV:    lut_fd_kernel<MODEL, PDX_LAP_CONV, MATH><<<grid, block, 0, s>>>(a, L); break;
        case PDX_LAP_HETEROG: lut_fd_kernel<MODEL, PDX_LAP_HETEROG, MATH><<<grid, block, 0, s>>>(a, L); break;
        case PDX_LAP_ANISO:   lut_fd_kernel<MODEL, PDX_LAP_ANISO, MATH><<<grid, block, 0, s>>>(a, L); break;
        default: set_error("unknown lap_kind"); return 1;
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "lut_fd_kernel launch");
}

}  // namespace

#if PDX_LUT_TU_MATH == 0
int lut_model_num_consts(int model) {
    return model == PDX_MODEL_CRN ? (int)C_NCONST : model == PDX_MODEL_TTP ? (int)T_NCONST : -1;
}
#define PDX_LUT_FN(name) name##_exact
#else
#define PDX_LUT_FN(name) name##_fast
#endif

int PDX_LUT_FN(launch_lut_ionic)(int model, const LutArgs& L, int64_t n, const float* U, float* dU, const float* I0,
                                 float* rhs0, cudaStream_t s) {
    const bool aligned = n % 4 == 0 && ((uintptr_t)U % 4) == 0;
    if (aligned && model == PDX_MODEL_CRN && tile_tables_ok<PDX_MODEL_CRN>(L)) {
        auto kern = lut_ionic_tile_kernel<PDX_MODEL_CRN, TU_MATH>;
        constexpr int bytes = tile_smem_bytes<PDX_MODEL_CRN>(0);
        if (tile_smem_attr(kern, bytes)) return 1;
        kern<<<(unsigned)((n + TILE - 1) / TILE), TT, bytes, s>>>(L, n, U, dU, I0, rhs0);
        count_launch();
        return check_cuda(cudaGetLastError(), "lut_ionic_tile_kernel launch");
    }
    if (aligned && model == PDX_MODEL_TTP && tile_tables_ok<PDX_MODEL_TTP>(L)) {
        auto kern = lut_ionic_tile_kernel<PDX_MODEL_TTP, TU_MATH>;
        constexpr int bytes = tile_smem_bytes<PDX_MODEL_TTP>(0);
        if (tile_smem_attr(kern, bytes)) return 1;
        kern<<<(unsigned)((n + TILE - 1) / TILE), TT, bytes, s>>>(L, n, U, dU, I0, rhs0);
        count_launch();
        return check_cuda(cudaGetLastError(), "lut_ionic_tile_kernel launch");
    }
    int64_t nb = (n + 127) / 128;
    if (nb > sm_count() * 32) nb = sm_count() * 32;
    if (nb < 1) nb = 1;
    if (model == PDX_MODEL_CRN) lut_ionic_kernel<PDX_MODEL_CRN, TU_MATH><<<(unsigned)nb, 128, 0, s>>>(L, n, U, dU, I0, rhs0);
    else if (model == PDX_MODEL_TTP) lut_ionic_kernel<PDX_MODEL_TTP, TU_MATH><<<(unsigned)nb, 128, 0, s>>>(L, n, U, dU, I0, rhs0);
    else { set_error("not a lookup-table model"); return 1; }
    count_launch();
    return check_cuda(cudaGetLastError(), "lut_ionic_kernel launch");
}

int PDX_LUT_FN(launch_fd_lut)(const pdx_fd_desc& d, const FdArgs& a, const LutArgs& L, cudaStream_t s) {
    if (d.model == PDX_MODEL_CRN) return launch_fd_lap<PDX_MODEL_CRN, TU_MATH>(d.lap_kind, a, L, s);
    if (d.model == PDX_MODEL_TTP) return launch_fd_lap<PDX_MODEL_TTP, TU_MATH>(d.lap_kind, a, L, s);
    set_error("not a lookup-table model");
    return 1;
}

}  // namespace pdx
