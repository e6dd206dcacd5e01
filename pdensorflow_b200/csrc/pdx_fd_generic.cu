// pdx_fd_generic.cu - shape-agnostic fused FD step, stand-alone operators, pointwise ionic
// step and stimulus application.  One thread per node, neighbours through clamped global
// loads (L1/L2 provide the reuse).  This is the fallback for shapes the TMA tile kernel
// does not take (nz % 4 != 0, conv alias, Dirichlet) and the independent implementation
// the TMA kernel is cross-checked against.
#include "pdx_stencil.cuh"

namespace pdx {

namespace {

// MODE: 0 = fused step (writes U1, advances states), 1 = Laplacian only (writes lap)
template <int MODEL, int LAP, int MATH, int MODE>
__global__ void __launch_bounds__(256) fd_generic_kernel(const __grid_constant__ FdArgs a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = a.xb + blockIdx.z;
    if (k >= a.nz || j >= a.ny || i >= a.xe) return;
    const size_t idx = ((size_t)i * a.ny + j) * a.pz + k;

    const Field f = make_field(a);
    float c;
    const float lap = node_laplacian<LAP, MATH>(a, f, i, j, k, c);
    if (MODE == 1) {
        a.Uout[idx] = lap;
        return;
    }
    float u1 = c;
    if (MODEL != PDX_MODEL_NONE) {
        constexpr int NS = NStates<MODEL>::value;
        float st[PDX_MAX_STATES];
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = a.st[s][idx];
        const float dU = ionic_update<MODEL, MATH>(a.Uin[idx], st, a.mp);   // raw U, not U0
#pragma unroll
        for (int s = 0; s < NS; ++s) a.st[s][idx] = st[s];
        u1 = madd<MATH>(u1, mmul<MATH>(a.dt, dU));
    }
    const float u_new = madd<MATH>(u1, mmul<MATH>(a.coef, lap));
    a.Uout[idx] = u_new;
    if (a.mirror_lo && i == a.mirror_lo_plane) a.mirror_lo[(size_t)j * a.pz + k] = u_new;
    if (a.mirror_hi && i == a.mirror_hi_plane) a.mirror_hi[(size_t)j * a.pz + k] = u_new;
}

template <int MODEL, int LAP, int MATH, int MODE>
int launch_one(const FdArgs& a, cudaStream_t s) {
    dim3 block(64, 4, 1);
    dim3 grid((a.nz + 63) / 64, (a.ny + 3) / 4, a.xe - a.xb);
    if (a.xe <= a.xb) return 0;
    if (grid.z > 65535u || grid.y > 65535u) { set_error("generic FD kernel: grid too large"); return 1; }
    fd_generic_kernel<MODEL, LAP, MATH, MODE><<<grid, block, 0, s>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "fd_generic_kernel launch");
}

template <int MODEL, int LAP, int MODE>
int launch_math(int math, const FdArgs& a, cudaStream_t s) {
    return math == PDX_MATH_EXACT ? launch_one<MODEL, LAP, PDX_MATH_EXACT, MODE>(a, s)
                                  : launch_one<MODEL, LAP, PDX_MATH_FAST, MODE>(a, s);
}
template <int MODEL, int MODE>
int launch_lap(int lap, int math, const FdArgs& a, cudaStream_t s) {
    switch (lap) {
        case PDX_LAP_HOMOG:   return launch_math<MODEL, PDX_LAP_HOMOG, MODE>(math, a, s);
        case PDX_LAP_CONV:    return launch_math<MODEL, PDX_LAP_CONV, MODE>(math, a, s);
        case PDX_LAP_HETEROG: return launch_math<MODEL, PDX_LAP_HETEROG, MODE>(math, a, s);
        case PDX_LAP_ANISO:   return launch_one<MODEL, PDX_LAP_ANISO, PDX_MATH_EXACT, MODE>(a, s);
    }
    set_error("unknown lap_kind");
    return 1;
}

}  // namespace

int launch_fd_generic(const pdx_fd_desc& d, const FdArgs& a, cudaStream_t s) {
    switch (d.model) {
        case PDX_MODEL_NONE:     return launch_lap<PDX_MODEL_NONE, 0>(d.lap_kind, d.math_mode, a, s);
        case PDX_MODEL_FENTON4V: return launch_lap<PDX_MODEL_FENTON4V, 0>(d.lap_kind, d.math_mode, a, s);
        case PDX_MODEL_MMS2V:    return launch_lap<PDX_MODEL_MMS2V, 0>(d.lap_kind, d.math_mode, a, s);
        case PDX_MODEL_MS2V:     return launch_lap<PDX_MODEL_MS2V, 0>(d.lap_kind, d.math_mode, a, s);
    }
    set_error("unknown model");
    return 1;
}

int launch_fd_laplace_only(int lap_kind, int math_mode, const FdArgs& a, cudaStream_t s) {
    return launch_lap<PDX_MODEL_NONE, 1>(lap_kind, math_mode, a, s);
}

int launch_fd_laplace_aniso(const FdArgs& a, cudaStream_t s) {
    return launch_one<PDX_MODEL_NONE, PDX_LAP_ANISO, PDX_MATH_EXACT, 1>(a, s);
}

// ---------------------------------------------------------------- enforce_boundary
__global__ void __launch_bounds__(256) enforce_boundary_kernel(const __grid_constant__ FdArgs a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = blockIdx.z;
    if (k >= a.nz || j >= a.ny) return;
    const Field f = make_field(a);
    a.Uout[((size_t)i * a.ny + j) * a.pz + k] = f.at(i, j, k);
}

int launch_enforce_boundary(const FdArgs& a, cudaStream_t s) {
    dim3 block(64, 4, 1);
    dim3 grid((a.nz + 63) / 64, (a.ny + 3) / 4, a.nx);
    enforce_boundary_kernel<<<grid, block, 0, s>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "enforce_boundary_kernel launch");
}

// ---------------------------------------------------------------- pointwise ionic step
struct IonicArgs {
    const float* U;
    float* st[PDX_MAX_STATES];
    float* dU;
    const float* I0;
    float* rhs0;
    const float* parr[PDX_MAX_PARAMS];
    int64_t n;
    ModelParams mp;
};

template <int MODEL, int MATH, bool NODAL>
__global__ void __launch_bounds__(256) ionic_kernel(const __grid_constant__ IonicArgs a) {
    constexpr int NS = NStates<MODEL>::value;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < a.n; n += stride) {
        float st[PDX_MAX_STATES];
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = a.st[s][n];
        const float U = a.U[n];
        float dU;
        if (NODAL) {
            NodalP P(a.mp, a.parr, n);
            const float DV = P(MODEL == PDX_MODEL_FENTON4V ? (int)FK_DV : (int)MS_DV);
            if (MODEL == PDX_MODEL_FENTON4V) dU = fk_exact(U, st[0], st[1], st[2], P, DV, a.mp.dt);
            else if (MODEL == PDX_MODEL_MMS2V) dU = ms_exact<true>(U, st[0], P, DV, a.mp.dt);
            else dU = ms_exact<false>(U, st[0], P, DV, a.mp.dt);
        } else {
            dU = ionic_update<MODEL, MATH>(U, st, a.mp);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) a.st[s][n] = st[s];
        if (a.dU) a.dU[n] = dU;
        if (a.rhs0) {
            // RHS0 = U + dt*(dU + I0): monodomainSolver.py:100-103
            const float f = a.I0 ? a.I0[n] : 0.0f;
            a.rhs0[n] = MATH == PDX_MATH_EXACT ? xadd(U, xmul(a.mp.dt, xadd(dU, f))) : fmaf(a.mp.dt, dU + f, U);
        }
    }
}

template <int MODEL>
int launch_ionic_model(int math, bool nodal, const IonicArgs& a, cudaStream_t s) {
    const int block = 256;
    int64_t nb = (a.n + block - 1) / block;
    if (nb > sm_count() * 16) nb = sm_count() * 16;
    if (nb < 1) nb = 1;
    if (nodal) ionic_kernel<MODEL, PDX_MATH_EXACT, true><<<(unsigned)nb, block, 0, s>>>(a);
    else if (math == PDX_MATH_EXACT) ionic_kernel<MODEL, PDX_MATH_EXACT, false><<<(unsigned)nb, block, 0, s>>>(a);
    else ionic_kernel<MODEL, PDX_MATH_FAST, false><<<(unsigned)nb, block, 0, s>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "ionic_kernel launch");
}

int launch_ionic(int model, int math, const float* params, const float* const* parr, int64_t n, const float* U,
                 float* const* states, float* dU, float dt, const float* I0, float* rhs0, cudaStream_t s) {
    IonicArgs a{};
    a.U = U; a.dU = dU; a.I0 = I0; a.rhs0 = rhs0; a.n = n;
    const int ns = pdx_model_num_states(model);
    const int np = pdx_model_num_params(model);
    if (ns <= 0) { set_error("pdx_ionic_step: model has no states / unknown model"); return 1; }
    for (int i = 0; i < ns; ++i) a.st[i] = states[i];
    bool nodal = false;
    if (parr)
        for (int i = 0; i < np; ++i) { a.parr[i] = parr[i]; nodal |= parr[i] != nullptr; }
    fill_model_params(model, params, dt, &a.mp);
    switch (model) {
        case PDX_MODEL_FENTON4V: return launch_ionic_model<PDX_MODEL_FENTON4V>(math, nodal, a, s);
        case PDX_MODEL_MMS2V:    return launch_ionic_model<PDX_MODEL_MMS2V>(math, nodal, a, s);
        case PDX_MODEL_MS2V:     return launch_ionic_model<PDX_MODEL_MS2V>(math, nodal, a, s);
    }
    set_error("pdx_ionic_step: unknown model");
    return 1;
}

// ---------------------------------------------------------------- stimulus
__global__ void __launch_bounds__(256) stim_kernel(float* U, const float* mask, float intensity, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __fmul_rn(intensity, mask[i]);     // stim = intensity*mask ; bool(stim)
        if (v != 0.0f) U[i] = fmaxf(U[i], v);
    }
}

// heat examples: U = maximum(U, intensity*mask) on EVERY node (Tests/FD/HeatEquation/heat.py:143-144)
__global__ void __launch_bounds__(256) stim_max_kernel(float* U, const float* mask, float intensity, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        U[i] = fmaxf(U[i], __fmul_rn(intensity, mask[i]));
}

int launch_stim_max(float* U, const float* mask, float intensity, int64_t n, cudaStream_t s) {
    int64_t nb = (n + 255) / 256;
    if (nb > sm_count() * 16) nb = sm_count() * 16;
    if (nb < 1) nb = 1;
    stim_max_kernel<<<(unsigned)nb, 256, 0, s>>>(U, mask, intensity, n);
    count_launch();
    return check_cuda(cudaGetLastError(), "stim_max_kernel launch");
}

int launch_stim(float* U, const float* mask, float intensity, int64_t n, cudaStream_t s) {
    int64_t nb = (n + 255) / 256;
    if (nb > sm_count() * 16) nb = sm_count() * 16;
    if (nb < 1) nb = 1;
    stim_kernel<<<(unsigned)nb, 256, 0, s>>>(U, mask, intensity, n);
    count_launch();
    return check_cuda(cudaGetLastError(), "stim_kernel launch");
}

}  // namespace pdx
