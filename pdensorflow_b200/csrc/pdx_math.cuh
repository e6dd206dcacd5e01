// pdx_math.cuh - fp32 arithmetic policies and the pointwise ionic models.
//
// Two arithmetic policies (see include/pdx.h):
//   EXACT: every operation is a separately rounded IEEE fp32 op (__f*_rn intrinsics are
//          never contracted into FMAs), divisions are IEEE divisions, tanh is the
//          correctly rounded fp32 value.  Evaluation order follows the reference source
//          token by token (gpuSolve/ionic/fenton4v.py:168-185, mms2v.py:95-105,
//          ms2v.py:90-100), so results are bit-comparable with the NumPy oracle.
//   FAST:  same formulas; divisions by parameters become multiplications by host-side
//          reciprocals, FMA contraction is allowed and 0.5*(1+tanh(x)) is evaluated as
//          1/(1+2^(-2x log2 e)) with ex2.approx / rcp.approx (2 MUFU ops).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pdx.h"

namespace pdx {

// ---- parameter block shared by every kernel (passed by value / __grid_constant__) ----
struct ModelParams {
    float p[PDX_MAX_PARAMS];   // raw parameters, layout in pdx.h
    float r[PDX_MAX_PARAMS];   // 1/p[i] (FAST)
    float DV;                  // vmax - vmin (fp32)
    float rDV;
    float d0, d1, d2, d3;      // model-specific derived constants (see fill_model_params)
    float dt;                  // f32(dt) advancing the gates
};

// Fenton4v parameter indices
enum { FK_tau_vp = 0, FK_tau_vn, FK_tau_wp, FK_tau_wn1, FK_tau_wn2, FK_tau_d, FK_tau_si, FK_tau_so,
       FK_tau_a, FK_u_c, FK_u_w, FK_u_0, FK_u_m, FK_u_csi, FK_u_so, FK_r_sp, FK_r_sn, FK_k,
       FK_a_so, FK_b_so, FK_c_so, FK_vmin, FK_DV, FK_NPARAM };
// (modified) Mitchell-Schaeffer parameter indices
enum { MS_tau_in = 0, MS_tau_out, MS_tau_open, MS_tau_close, MS_u_gate, MS_u_crit, MS_vmin, MS_DV,
       MS_NPARAM };

// ---- EXACT ops: separately rounded, never contracted ----
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xtanh(float x) { return (float)tanh((double)x); }
// tf.sign semantics: sign(0) = 0  =>  H(0) = G(0) = 0.5
__device__ __forceinline__ float step_H(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? 0.0f : 0.5f); }
__device__ __forceinline__ float step_G(float x) { return x > 0.0f ? 0.0f : (x < 0.0f ? 1.0f : 0.5f); }

// ---- FAST helpers ----
__device__ __forceinline__ float fast_ex2(float x) {
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
// 0.5*(1+tanh(x)) with x*(-2 log2 e) already folded into arg
__device__ __forceinline__ float fast_sigm_pre(float arg) { return fast_rcp(1.0f + fast_ex2(arg)); }

// Scalar parameter accessor (uniform parameters, the common case)
struct ScalarP {
    const ModelParams& m;
    __device__ __forceinline__ ScalarP(const ModelParams& mm) : m(mm) {}
    __device__ __forceinline__ float operator()(int i) const { return m.p[i]; }
    __device__ __forceinline__ float rcp(int i) const { return m.r[i]; }
};

// Per-node parameter accessor: arr[i] != nullptr overrides p[i] with arr[i][node]
// (IonicModel.set_parameter with an [N,1] array, ionicmodel.py:56-62).
struct NodalP {
    const ModelParams& m;
    const float* const* arr;
    int64_t node;
    __device__ __forceinline__ NodalP(const ModelParams& mm, const float* const* a, int64_t n)
        : m(mm), arr(a), node(n) {}
    __device__ __forceinline__ float operator()(int i) const {
        const float* a = arr[i];
        return a ? a[node] : m.p[i];
    }
    __device__ __forceinline__ float rcp(int i) const { return 1.0f / (*this)(i); }
};

// =====================================================================================
// Fenton-Karma 4v (Cherry-Ehrlich-Nattel-Fenton), gpuSolve/ionic/fenton4v.py:168-185.
// Returns dU; advances V,W,S in place by dt.  Currents use the old gates.
// =====================================================================================
template <class PA>
__device__ __forceinline__ float fk_exact(float U, float& V, float& W, float& S, const PA& P, float DV,
                                          float dt) {
    const float u    = xdiv(xsub(U, P(FK_vmin)), DV);                 // to_dimensionless
    const float uc   = xsub(u, P(FK_u_c));
    const float Hc   = step_H(uc);
    // I_fi = -V * H(u-u_c) * (u-u_c) * (u_m-u) / tau_d
    const float I_fi = xdiv(xmul(xmul(xmul(-V, Hc), uc), xsub(P(FK_u_m), u)), P(FK_tau_d));
    // I_si = -W * S / tau_si
    const float I_si = xdiv(xmul(-W, S), P(FK_tau_si));
    // I_so = 0.5*(a_so-tau_a)*(1+tanh((u-b_so)/c_so)) + (u-u_0)*G(u-u_so)/tau_so + H(u-u_so)*tau_a
    const float uso  = xsub(u, P(FK_u_so));
    const float t1   = xtanh(xdiv(xsub(u, P(FK_b_so)), P(FK_c_so)));
    const float A    = xmul(xmul(0.5f, xsub(P(FK_a_so), P(FK_tau_a))), xadd(1.0f, t1));
    const float B    = xdiv(xmul(xsub(u, P(FK_u_0)), step_G(uso)), P(FK_tau_so));
    const float C    = xmul(step_H(uso), P(FK_tau_a));
    const float I_so = xadd(xadd(A, B), C);
    const float dU   = -xmul(DV, xadd(xadd(I_fi, I_si), I_so));
    const bool  gc   = u > P(FK_u_c);
    const float dV   = gc ? xdiv(-V, P(FK_tau_vp)) : xdiv(xsub(1.0f, V), P(FK_tau_vn));
    const float dW   = gc ? xdiv(-W, P(FK_tau_wp))
                          : (u > P(FK_u_w) ? xdiv(xsub(1.0f, W), P(FK_tau_wn2))
                                           : xdiv(xsub(1.0f, W), P(FK_tau_wn1)));
    const float r_s  = xadd(xmul(xsub(P(FK_r_sp), P(FK_r_sn)), Hc), P(FK_r_sn));
    const float t2   = xtanh(xmul(xsub(u, P(FK_u_csi)), P(FK_k)));
    const float dS   = xmul(r_s, xsub(xmul(0.5f, xadd(1.0f, t2)), S));
    V = xadd(V, xmul(dt, dV));
    W = xadd(W, xmul(dt, dW));
    S = xadd(S, xmul(dt, dS));
    return dU;
}

// FAST: derived constants in ModelParams: d0 = -2*log2(e)/c_so, d1 = -2*log2(e)*k_,
// d2 = a_so - tau_a, d3 = r_sp - r_sn.
__device__ __forceinline__ float fk_fast(float U, float& V, float& W, float& S, const ModelParams& m) {
    const float* p = m.p;
    const float* r = m.r;
    const float u   = (U - p[FK_vmin]) * m.rDV;
    const float uc  = u - p[FK_u_c];
    const bool  gc  = uc > 0.0f;
    const float Hc  = gc ? 1.0f : (uc < 0.0f ? 0.0f : 0.5f);
    const float I_fi = gc ? (-V * uc) * (p[FK_u_m] - u) * r[FK_tau_d] : 0.0f;
    const float I_si = (-W * S) * r[FK_tau_si];
    const float uso = u - p[FK_u_so];
    const float A   = m.d2 * fast_sigm_pre((u - p[FK_b_so]) * m.d0);
    const float Gs  = step_G(uso);
    const float BC  = (u - p[FK_u_0]) * Gs * r[FK_tau_so] + (1.0f - Gs) * p[FK_tau_a];
    const float dU  = -m.DV * (I_fi + I_si + (A + BC));
    const float dV  = gc ? -V * r[FK_tau_vp] : (1.0f - V) * r[FK_tau_vn];
    const float dW  = gc ? -W * r[FK_tau_wp]
                         : (1.0f - W) * (u > p[FK_u_w] ? r[FK_tau_wn2] : r[FK_tau_wn1]);
    const float r_s = m.d3 * Hc + p[FK_r_sn];
    const float dS  = r_s * (fast_sigm_pre((u - p[FK_u_csi]) * m.d1) - S);
    V = fmaf(m.dt, dV, V);
    W = fmaf(m.dt, dW, W);
    S = fmaf(m.dt, dS, S);
    return dU;
}

// =====================================================================================
// modified Mitchell-Schaeffer (mms2v.py:95-105) and Mitchell-Schaeffer (ms2v.py:90-100)
// =====================================================================================
template <bool MODIFIED, class PA>
__device__ __forceinline__ float ms_exact(float U, float& Hs, const PA& P, float DV, float dt) {
    const float u = xdiv(xsub(U, P(MS_vmin)), DV);
    float J_in, J_out;
    if (MODIFIED) {
        // J_in = -1.0*H*u*(u-u_crit)*(1-u)/tau_in ; J_out = (1-H)*u/tau_out
        J_in  = xdiv(xmul(xmul(xmul(xmul(-1.0f, Hs), u), xsub(u, P(MS_u_crit))), xsub(1.0f, u)), P(MS_tau_in));
        J_out = xdiv(xmul(xsub(1.0f, Hs), u), P(MS_tau_out));
    } else {
        // J_in = -1.0*H*u*u*(1-u)/tau_in ; J_out = u/tau_out
        J_in  = xdiv(xmul(xmul(xmul(xmul(-1.0f, Hs), u), u), xsub(1.0f, u)), P(MS_tau_in));
        J_out = xdiv(u, P(MS_tau_out));
    }
    const float dU = -xmul(DV, xadd(J_in, J_out));
    const float dH = u > P(MS_u_gate) ? xdiv(-Hs, P(MS_tau_close)) : xdiv(xsub(1.0f, Hs), P(MS_tau_open));
    Hs = xadd(Hs, xmul(dt, dH));
    return dU;
}

template <bool MODIFIED>
__device__ __forceinline__ float ms_fast(float U, float& Hs, const ModelParams& m) {
    const float* p = m.p;
    const float* r = m.r;
    const float u = (U - p[MS_vmin]) * m.rDV;
    float J;
    if (MODIFIED)
        J = (-Hs * u) * (u - p[MS_u_crit]) * (1.0f - u) * r[MS_tau_in] + (1.0f - Hs) * u * r[MS_tau_out];
    else
        J = (-Hs * u) * u * (1.0f - u) * r[MS_tau_in] + u * r[MS_tau_out];
    const float dU = -m.DV * J;
    const float dH = u > p[MS_u_gate] ? -Hs * r[MS_tau_close] : (1.0f - Hs) * r[MS_tau_open];
    Hs = fmaf(m.dt, dH, Hs);
    return dU;
}

// Uniform front end used by the fused kernels: st[] holds the model's states.
template <int MODEL, int MATH>
__device__ __forceinline__ float ionic_update(float U, float (&st)[PDX_MAX_STATES], const ModelParams& m) {
    if (MODEL == PDX_MODEL_FENTON4V) {
        if (MATH == PDX_MATH_EXACT) return fk_exact(U, st[0], st[1], st[2], ScalarP(m), m.DV, m.dt);
        return fk_fast(U, st[0], st[1], st[2], m);
    } else if (MODEL == PDX_MODEL_MMS2V) {
        if (MATH == PDX_MATH_EXACT) return ms_exact<true>(U, st[0], ScalarP(m), m.DV, m.dt);
        return ms_fast<true>(U, st[0], m);
    } else if (MODEL == PDX_MODEL_MS2V) {
        if (MATH == PDX_MATH_EXACT) return ms_exact<false>(U, st[0], ScalarP(m), m.DV, m.dt);
        return ms_fast<false>(U, st[0], m);
    }
    return 0.0f;
}

template <int MODEL> struct NStates { static constexpr int value = 0; };
template <> struct NStates<PDX_MODEL_FENTON4V> { static constexpr int value = 3; };
template <> struct NStates<PDX_MODEL_MMS2V>    { static constexpr int value = 1; };
template <> struct NStates<PDX_MODEL_MS2V>     { static constexpr int value = 1; };

// ---- Laplacian arithmetic ----
// homogeneous, one axis: ((m - 2.0*c) + p) / hsq        (laplace_homogeneous...py:29-31)
template <int MATH>
__device__ __forceinline__ float lap_axis(float m_, float c, float p_, float hsq, float rhsq) {
    if (MATH == PDX_MATH_EXACT) return xdiv(xadd(xsub(m_, xmul(2.0f, c)), p_), hsq);
    return ((m_ - 2.0f * c) + p_) * rhsq;
}
// heterogeneous, one axis: (e - w)/hsq with e = 0.5*(Dp+Dc)*(Xp-Xc), w = -0.5*(Dm+Dc)*(Xm-Xc)
// (laplace_heterogeneous_isotropic_diffusion.py:20-34)
template <int MATH>
__device__ __forceinline__ float lap_axis_het(float xm, float xc, float xp, float dm, float dc, float dp,
                                              float hsq, float rhsq) {
    if (MATH == PDX_MATH_EXACT) {
        const float e = xmul(xmul(0.5f, xadd(dp, dc)), xsub(xp, xc));
        const float w = xmul(xmul(-0.5f, xadd(dm, dc)), xsub(xm, xc));
        return xdiv(xsub(e, w), hsq);
    }
    const float e = (0.5f * (dp + dc)) * (xp - xc);
    const float w = (-0.5f * (dm + dc)) * (xm - xc);
    return (e - w) * rhsq;
}
template <int MATH> __device__ __forceinline__ float madd(float a, float b) {
    return MATH == PDX_MATH_EXACT ? xadd(a, b) : a + b;
}
template <int MATH> __device__ __forceinline__ float mmul(float a, float b) {
    return MATH == PDX_MATH_EXACT ? xmul(a, b) : a * b;
}

}  // namespace pdx
