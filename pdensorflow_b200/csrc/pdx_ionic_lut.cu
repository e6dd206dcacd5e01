// pdx_ionic_lut.cu - the reference's two lookup-table ionic models on sm_100a:
//   Courtemanche-Ramirez-Nattel 1998 (gpuSolve/ionic/courtemanche_ramirez_nattel.py:423-514)
//   ten Tusscher-Panfilov 2006      (gpuSolve/ionic/ten_tusscher_panfilov.py:405-523)
// as (a) the pointwise IonicModel.differentiate step (+ the FEM RHS0) and (b) fused with the
// finite-difference stencil and the Euler update (one pass over U and the 19/22 state arrays).
//
// One thread per node.  State arrays are SoA and cross HBM once per step (coalesced 4-byte
// accesses); the three tables (0.5-12 MB) stay L2/L1 resident - neighbouring nodes hit the same
// two rows - and are read through the read-only path, Rush-Larsen (A,B) pairs as 8-byte loads.
//
// THIS FILE IS COMPILED TWICE (csrc/build.py), once per arithmetic policy (PDX_LUT_TU_MATH = PDX_MATH_*):
//   EXACT  -fmad=false: every expression is evaluated op-for-op in fp32 in the order the reference writes
//          it (IEEE division, no contraction) and exp/log/expm1 return the correctly rounded fp32 value
//          (fp64 evaluation, one rounding) - the oracle's convention, results bit-comparable with
//          oracle/ionic_lut.py.  ~1400 instructions per node: the kernel is issue/latency bound.
//   FAST   same formulas, FMA contraction allowed, approximate division (-prec-div=false, 2 ulp) and
//          CUDA's fp32 expf/logf/expm1f.  rel-L2 <= 1e-5 against the oracle (tests/test_ref_golden_gpu.py).
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

template <int MATH>
__device__ __forceinline__ float crn_update(const LutArgs& L, size_t n, float V) {
    const float* c = L.c;
    const float dt = L.dt;
    float y[CRN_NS];
#pragma unroll
    for (int s = 0; s < CRN_NS; ++s) y[s] = L.st[s][n];
    const Row vr = lookup(L.tab[0], V);
    const Row cr = lookup(L.tab[1], y[S_Cai]);

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
    const Row fr = lookup(L.tab[2], fn);

    L.st[S_Ca_rel][n] = y[S_Ca_rel] + d_Ca_rel * dt;
    L.st[S_Ca_up][n] = y[S_Ca_up] + d_Ca_up * dt;
    L.st[S_Cai][n] = y[S_Cai] + d_Cai * dt;
    L.st[S_Ki][n] = y[S_Ki] + d_Ki * dt;
    L.st[S_d][n] = vr.rl(CV_d, y[S_d]);
    L.st[S_f][n] = vr.rl(CV_f, y[S_f]);
    L.st[S_f_Ca][n] = cr(CC_fCa_A) + c[C_BFCA] * y[S_f_Ca];
    L.st[S_h][n] = vr.rl(CV_h, y[S_h]);
    L.st[S_j][n] = vr.rl(CV_j, y[S_j]);
    L.st[S_m][n] = vr.rl(CV_m, y[S_m]);
    L.st[S_oa][n] = vr.rl(CV_oa, y[S_oa]);
    L.st[S_oi][n] = vr.rl(CV_oi, y[S_oi]);
    L.st[S_u][n] = fr(CF_u_A) + c[C_BU] * y[S_u];
    L.st[S_ua][n] = vr.rl(CV_ua, y[S_ua]);
    L.st[S_ui][n] = vr.rl(CV_ui, y[S_ui]);
    L.st[S_v][n] = fr(CF_v) + fr(CF_v + 1) * y[S_v];
    L.st[S_w][n] = vr.rl(CV_w, y[S_w]);
    L.st[S_xr][n] = vr.rl(CV_xr, y[S_xr]);
    L.st[S_xs][n] = vr.rl(CV_xs, y[S_xs]);
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

template <int MATH>
__device__ __forceinline__ float ttp_update(const LutArgs& L, size_t n, float V) {
    const float* c = L.c;
    const float dt = L.dt;
    float y[TTP_NS];
#pragma unroll
    for (int s = 0; s < TTP_NS; ++s) y[s] = L.st[s][n];
    const float Cai = y[Q_Cai], CaSS = y[Q_CaSS], CaSR = y[Q_CaSR], Nai = y[Q_Nai], Ki = y[Q_Ki], R_bar = y[Q_R_bar];
    const Row vr = lookup(L.tab[0], V);
    const Row cr = lookup(L.tab[1], CaSS);

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
    const Row er = lookup(L.tab[2], V - Ek);
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

    L.st[Q_CaSR][n] = CaSR + d_CaSR * dt;
    L.st[Q_CaSS][n] = CaSS + d_CaSS * dt;
    L.st[Q_Cai][n] = Cai + d_Cai * dt;
    L.st[Q_Ki][n] = Ki + d_Ki * dt;
    L.st[Q_Nai][n] = Nai + d_Nai * dt;
    L.st[Q_R_bar][n] = R_bar + d_R_bar * dt;
    L.st[Q_D][n] = vr.rl(TV_D, y[Q_D]);
    L.st[Q_F][n] = vr.rl(TV_F, y[Q_F]);
    L.st[Q_F2][n] = vr.rl(TV_F2, y[Q_F2]);
    L.st[Q_FCaSS][n] = cr.rl(0, y[Q_FCaSS]);
    L.st[Q_H][n] = vr.rl(TV_H, y[Q_H]);
    L.st[Q_J][n] = vr.rl(TV_J, y[Q_J]);
    L.st[Q_M][n] = vr.rl(TV_M, y[Q_M]);
    L.st[Q_R][n] = vr.rl(TV_R, y[Q_R]);
    L.st[Q_S][n] = vr.rl(TV_S, y[Q_S]);
    L.st[Q_Xr1][n] = vr.rl(TV_Xr1, y[Q_Xr1]);
    L.st[Q_Xr2][n] = vr.rl(TV_Xr2, y[Q_Xr2]);
    L.st[Q_Xs][n] = vr.rl(TV_Xs, y[Q_Xs]);
    return -Iion;
}

template <int MODEL, int MATH>
__device__ __forceinline__ float lut_update(const LutArgs& L, size_t n, float V) {
    return MODEL == PDX_MODEL_CRN ? crn_update<MATH>(L, n, V) : ttp_update<MATH>(L, n, V);
}

// ---------------------------------------------------------------------------- kernels
// IonicModel.differentiate (+ RHS0 = U + dt*(dU + I0) of MonodomainSolver.solve_step)
template <int MODEL, int MATH>
__global__ void __launch_bounds__(128, 8) lut_ionic_kernel(const __grid_constant__ LutArgs L, int64_t n, const float* U,
                                                        float* dU, const float* I0, float* rhs0) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float V = U[i];
        const float d = lut_update<MODEL, MATH>(L, (size_t)i, V);
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
    const float dU = lut_update<MODEL, MATH>(L, idx, a.Uin[idx]);      // raw U, not U0
    const float u_new = (c + a.dt * dU) + a.coef * lap;
    a.Uout[idx] = u_new;
    if (a.mirror_lo && i == a.mirror_lo_plane) a.mirror_lo[(size_t)j * a.nz + k] = u_new;
    if (a.mirror_hi && i == a.mirror_hi_plane) a.mirror_hi[(size_t)j * a.nz + k] = u_new;
}

template <int MODEL, int MATH>
int launch_fd_lap(int lap, const FdArgs& a, const LutArgs& L, cudaStream_t s) {
    if (a.xe <= a.xb) return 0;
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
