// pdx_fd_resident.cu - SM-resident multi-step kernel for 2-D grids that fit on chip
// (BASELINE.json config 2: 1024 x 1024 Fenton-Karma, 10 000 steps).
//
// A 1024^2 grid is 4 MB per array: one time step moves 33 MB, which a kernel launch per step turns
// into ~10 us of launch latency + L2 traffic.  Here the WHOLE grid lives in the shared memory of
// the SMs for `nsteps` time steps: the rows are dealt out in contiguous strips, one CTA per SM, a
// thread per column; U (double buffered, with one halo row per side) and the gate variables never
// leave the SM.  Per step a CTA
//   1. advances its first and last row (they are what the neighbours wait for),
//   2. publishes them (2 x nz floats) to a global halo buffer and raises its step flag
//      (st.release.gpu),
//   3. advances half of its interior rows - which hides the flag latency of the neighbours,
//   4. acquires the two neighbour flags and issues the loads of their rows from L2,
//   5. advances the other half while those loads are in flight, and stores them as next step's halo rows.
// No grid-wide barrier, no HBM traffic between the first load and the final store.  The CTAs spin
// on each other, so the kernel is launched cooperatively (co-residency guaranteed).
//
// Arithmetic: the same node formulas as the per-step kernels (pdx_math.cuh, evaluation order of
// pdx_stencil.cuh node_laplacian), so EXACT results are bit-identical to them and to the oracle.
// Reference: Tests/FD/Fenton/fenton.py:46-53,90-99 transcribed to 2-D (SURVEY.md section 8a, rows
// B, L2, I-*, T-FD).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include "pdx_internal.h"

namespace pdx {
namespace {

constexpr int kFlagStride = PDX_RESIDENT_FLAG_STRIDE;   // one flag per 128-byte line

struct ResArgs {
    const float* Uin;
    float* Uout;
    float* st[PDX_MAX_STATES];
    int ny, nz;
    int rows_per_cta;
    int nsteps;
    unsigned int base;          // flag value of "nothing published yet" for this launch
    unsigned int* flags;        // [ncta * kFlagStride]
    float* halo;                // [2][ncta][2][nz]
    float dt, coef;
    float hysq, hzsq, rhysq, rhzsq;
    ModelParams mp;
};

__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag has reached `want` (flags only grow; compare modulo 2^32)
__device__ __forceinline__ void wait_flag(const unsigned int* flag, unsigned int want) {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while ((int)(ld_acquire_gpu_u32(flag) - want) < 0) {
        if ((++spins & 4095u) == 0) {
            const unsigned long long t = global_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 5000000000ull) {          // a co-resident CTA never takes 5 s: logic error
                printf("pdx fd_resident_kernel: CTA %d waited 5 s for a neighbour, aborting\n", (int)blockIdx.x);
                __trap();
            }
        }
    }
}

template <int MODEL, int MATH>
__global__ void __launch_bounds__(1024, 1) fd_resident_kernel(const __grid_constant__ ResArgs a) {
    extern __shared__ float smem[];
    constexpr int NS = NStates<MODEL>::value;
    const int nz = a.nz, ny = a.ny;
    const int k = threadIdx.x;
    const int waiter_hi = blockDim.x > 32 ? 32 : 0;
    const int c = blockIdx.x, ncta = gridDim.x;
    const int R = a.rows_per_cta;
    const int j0 = c * R;
    const int rows = min(R, ny - j0);
    // shared memory: two U buffers of (R+2) rows (row 0 / rows+1 are the halos), then NS state arrays of R rows
    const int ustride = (R + 2) * nz;
    const int sstride = R * nz;
    float* const Sk = smem + 2 * ustride + k;          // this thread's column of state 0
    const bool active = k < nz;
    const bool has_lo = c > 0, has_hi = c < ncta - 1;

    if (active) {
        for (int r = 0; r < rows; ++r) {
            const size_t g = (size_t)(j0 + r) * nz + k;
            smem[(r + 1) * nz + k] = a.Uin[g];
#pragma unroll
            for (int q = 0; q < NS; ++q) Sk[q * sstride + r * nz] = a.st[q][g];
        }
        if (has_lo) smem[k] = a.Uin[(size_t)(j0 - 1) * nz + k];
        if (has_hi) smem[(rows + 1) * nz + k] = a.Uin[(size_t)(j0 + rows) * nz + k];
    }
    __syncthreads();

    // column indices after symmetric pad + enforce_boundary (clamp to the array, then to [1, n-2])
    const int kc = min(max(k, 1), nz - 2);
    const int km = min(max(max(k - 1, 0), 1), nz - 2);
    const int kp = min(max(min(k + 1, nz - 1), 1), nz - 2);
    const int mid = 1 + max(rows - 2, 0) / 2;            // interior rows [1, mid) before the halo prefetch, [mid, rows-1) after

    for (int s = 0; s < a.nsteps; ++s) {
        const float* cur = smem + (s & 1) * ustride;
        float* nxt = smem + ((s & 1) ^ 1) * ustride;
        const bool more = s + 1 < a.nsteps;
        auto update = [&](int r) {
            const int jg = j0 + r;
            const int oraw = (r + 1) * nz;               // this node's own row in the buffer
            int oc = oraw, om = oraw - nz, op = oraw + nz;
            if (jg < 2 || jg > ny - 3) {                 // rows touched by the boundary rule
                oc = (min(max(jg, 1), ny - 2) - j0 + 1) * nz;
                om = (min(max(max(jg - 1, 0), 1), ny - 2) - j0 + 1) * nz;
                op = (min(max(min(jg + 1, ny - 1), 1), ny - 2) - j0 + 1) * nz;
            }
            const float cval = cur[oc + kc];
            float lap = lap_axis<MATH>(cur[om + kc], cval, cur[op + kc], a.hysq, a.rhysq);
            lap = madd<MATH>(lap, lap_axis<MATH>(cur[oc + km], cval, cur[oc + kp], a.hzsq, a.rhzsq));
            float u1 = cval;
            if (MODEL != PDX_MODEL_NONE) {
                float st[PDX_MAX_STATES];
                float* sp = Sk + r * nz;
#pragma unroll
                for (int q = 0; q < NS; ++q) st[q] = sp[q * sstride];
                const float dU = ionic_update<MODEL, MATH>(cur[oraw + k], st, a.mp);    // raw U, not U0
#pragma unroll
                for (int q = 0; q < NS; ++q) sp[q * sstride] = st[q];
                u1 = madd<MATH>(u1, mmul<MATH>(a.dt, dU));
            }
            const float un = madd<MATH>(u1, mmul<MATH>(a.coef, lap));
            nxt[oraw + k] = un;
            return un;
        };
        // 1. the rows the neighbours wait for
        float first = 0.0f, last = 0.0f;
        if (active) {
            first = update(0);
            last = rows > 1 ? update(rows - 1) : first;
        }
        // 2. publish them
        float* hb = a.halo + (size_t)(s & 1) * ncta * 2 * nz;
        if (more) {
            if (active) {
                __stcg(hb + (c * 2 + 0) * nz + k, first);
                __stcg(hb + (c * 2 + 1) * nz + k, last);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                st_release_gpu_u32(a.flags + c * kFlagStride, a.base + s + 1);
            }
        }
        // 3. first half of the interior rows: covers the flight time of the neighbours' flags
        if (active)
            for (int r = 1; r < mid; ++r) update(r);
        // 4. the neighbours' rows of this step: two threads acquire the flags (more pollers only queue up on the
        //    L2 slice that holds them), the loads stay in flight during 5.
        float h_lo = 0.0f, h_hi = 0.0f;
        if (more) {
            if (threadIdx.x == 0 && has_lo) wait_flag(a.flags + (c - 1) * kFlagStride, a.base + s + 1);
            if (threadIdx.x == waiter_hi && has_hi) wait_flag(a.flags + (c + 1) * kFlagStride, a.base + s + 1);
            __syncthreads();
            if (active) {
                if (has_lo) h_lo = __ldcg(hb + ((c - 1) * 2 + 1) * nz + k);
                if (has_hi) h_hi = __ldcg(hb + ((c + 1) * 2 + 0) * nz + k);
            }
        }
        // 5. second half of the interior rows
        if (active)
            for (int r = mid; r < rows - 1; ++r) update(r);
        // 6. halo rows of the next step's input
        if (more && active) {
            if (has_lo) nxt[k] = h_lo;
            if (has_hi) nxt[(rows + 1) * nz + k] = h_hi;
        }
        __syncthreads();
    }

    const float* fin = smem + (a.nsteps & 1) * ustride;
    if (active) {
        for (int r = 0; r < rows; ++r) {
            const size_t g = (size_t)(j0 + r) * nz + k;
            a.Uout[g] = fin[(r + 1) * nz + k];
#pragma unroll
            for (int q = 0; q < NS; ++q) a.st[q][g] = Sk[q * sstride + r * nz];
        }
    }
}

// Two nodes per thread (columns 2t, 2t+1; nz even): every shared-memory access of the centre / upper / lower
// rows and of the gates is a 64-bit access shared by both nodes, and so is the index arithmetic - the one-node
// kernel is issue bound at 153 warp instructions per node of which 60 are floating point.  Same protocol, same
// node formulas in the same order (EXACT stays bit-identical).
template <int MODEL, int MATH>
__global__ void __launch_bounds__(512, 1) fd_resident2_kernel(const __grid_constant__ ResArgs a) {
    extern __shared__ float smem[];
    constexpr int NS = NStates<MODEL>::value;
    const int nz = a.nz, ny = a.ny;
    const int k0 = 2 * threadIdx.x;
    const int waiter_hi = blockDim.x > 32 ? 32 : 0;
    const int c = blockIdx.x, ncta = gridDim.x;
    const int R = a.rows_per_cta;
    const int j0 = c * R;
    const int rows = min(R, ny - j0);
    const int ustride = (R + 2) * nz;
    const int sstride = R * nz;
    float* const Sk = smem + 2 * ustride + k0;
    const bool active = k0 < nz;
    const bool has_lo = c > 0, has_hi = c < ncta - 1;
    const bool first_pair = k0 == 0, last_pair = k0 + 2 == nz;
    auto ld2 = [](const float* p) { return *reinterpret_cast<const float2*>(p); };
    auto st2 = [](float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; };

    if (active) {
        for (int r = 0; r < rows; ++r) {
            const size_t g = (size_t)(j0 + r) * nz + k0;
            st2(smem + (r + 1) * nz + k0, ld2(a.Uin + g));
#pragma unroll
            for (int q = 0; q < NS; ++q) st2(Sk + q * sstride + r * nz, ld2(a.st[q] + g));
        }
        if (has_lo) st2(smem + k0, ld2(a.Uin + (size_t)(j0 - 1) * nz + k0));
        if (has_hi) st2(smem + (rows + 1) * nz + k0, ld2(a.Uin + (size_t)(j0 + rows) * nz + k0));
    }
    __syncthreads();
    const int kl = max(k0 - 1, 0), kr = min(k0 + 2, nz - 1);      // left / right neighbours of the pair
    const int mid = 1 + max(rows - 2, 0) / 2;

    for (int s = 0; s < a.nsteps; ++s) {
        const float* cur = smem + (s & 1) * ustride;
        float* nxt = smem + ((s & 1) ^ 1) * ustride;
        const bool more = s + 1 < a.nsteps;
        auto update = [&](int r) {
            const int jg = j0 + r;
            const int oraw = (r + 1) * nz;
            int oc = oraw, om = oraw - nz, op = oraw + nz;
            const bool edge_row = jg < 2 || jg > ny - 3;
            if (edge_row) {
                oc = (min(max(jg, 1), ny - 2) - j0 + 1) * nz;
                om = (min(max(max(jg - 1, 0), 1), ny - 2) - j0 + 1) * nz;
                op = (min(max(min(jg + 1, ny - 1), 1), ny - 2) - j0 + 1) * nz;
            }
            const float2 up = ld2(cur + om + k0), ce = ld2(cur + oc + k0), dn = ld2(cur + op + k0);
            const float2 raw = edge_row ? ld2(cur + oraw + k0) : ce;
            const float lft = cur[oc + kl], rgt = cur[oc + kr];
            // node A = column k0, node B = column k0 + 1: (upper, centre, lower, left, right) of U0
            float ua = up.x, ca = ce.x, da = dn.x, la = lft, ra = ce.y;
            float ub = up.y, cb = ce.y, db = dn.y, lb = ce.x, rb = rgt;
            if (first_pair) { ua = up.y; ca = ce.y; da = dn.y; la = ce.y; ra = ce.y; lb = ce.y; }   // column 0 reads column 1
            if (last_pair) { ub = up.x; cb = ce.x; db = dn.x; lb = ce.x; rb = ce.x; ra = ce.x; }    // column nz-1 reads nz-2
            float lapa = lap_axis<MATH>(ua, ca, da, a.hysq, a.rhysq);
            lapa = madd<MATH>(lapa, lap_axis<MATH>(la, ca, ra, a.hzsq, a.rhzsq));
            float lapb = lap_axis<MATH>(ub, cb, db, a.hysq, a.rhysq);
            lapb = madd<MATH>(lapb, lap_axis<MATH>(lb, cb, rb, a.hzsq, a.rhzsq));
            float u1a = ca, u1b = cb;
            if (MODEL != PDX_MODEL_NONE) {
                float sa[PDX_MAX_STATES], sb[PDX_MAX_STATES];
                float* sp = Sk + r * nz;
#pragma unroll
                for (int q = 0; q < NS; ++q) { const float2 v = ld2(sp + q * sstride); sa[q] = v.x; sb[q] = v.y; }
                const float dUa = ionic_update<MODEL, MATH>(raw.x, sa, a.mp);       // raw U, not U0
                const float dUb = ionic_update<MODEL, MATH>(raw.y, sb, a.mp);
#pragma unroll
                for (int q = 0; q < NS; ++q) st2(sp + q * sstride, make_float2(sa[q], sb[q]));
                u1a = madd<MATH>(u1a, mmul<MATH>(a.dt, dUa));
                u1b = madd<MATH>(u1b, mmul<MATH>(a.dt, dUb));
            }
            const float2 un = make_float2(madd<MATH>(u1a, mmul<MATH>(a.coef, lapa)), madd<MATH>(u1b, mmul<MATH>(a.coef, lapb)));
            st2(nxt + oraw + k0, un);
            return un;
        };
        float2 first = make_float2(0.0f, 0.0f), last = first;
        if (active) {
            first = update(0);
            last = rows > 1 ? update(rows - 1) : first;
        }
        float* hb = a.halo + (size_t)(s & 1) * ncta * 2 * nz;
        if (more) {
            if (active) {
                __stcg(reinterpret_cast<float2*>(hb + (c * 2 + 0) * nz + k0), first);
                __stcg(reinterpret_cast<float2*>(hb + (c * 2 + 1) * nz + k0), last);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                st_release_gpu_u32(a.flags + c * kFlagStride, a.base + s + 1);
            }
        }
        if (active)
            for (int r = 1; r < mid; ++r) update(r);
        float2 h_lo = make_float2(0.0f, 0.0f), h_hi = h_lo;
        if (more) {
            if (threadIdx.x == 0 && has_lo) wait_flag(a.flags + (c - 1) * kFlagStride, a.base + s + 1);
            if (threadIdx.x == waiter_hi && has_hi) wait_flag(a.flags + (c + 1) * kFlagStride, a.base + s + 1);
            __syncthreads();
            if (active) {
                if (has_lo) h_lo = __ldcg(reinterpret_cast<const float2*>(hb + ((c - 1) * 2 + 1) * nz + k0));
                if (has_hi) h_hi = __ldcg(reinterpret_cast<const float2*>(hb + ((c + 1) * 2 + 0) * nz + k0));
            }
        }
        if (active)
            for (int r = mid; r < rows - 1; ++r) update(r);
        if (more && active) {
            if (has_lo) st2(nxt + k0, h_lo);
            if (has_hi) st2(nxt + (rows + 1) * nz + k0, h_hi);
        }
        __syncthreads();
    }

    const float* fin = smem + (a.nsteps & 1) * ustride;
    if (active) {
        for (int r = 0; r < rows; ++r) {
            const size_t g = (size_t)(j0 + r) * nz + k0;
            st2(a.Uout + g, ld2(fin + (r + 1) * nz + k0));
#pragma unroll
            for (int q = 0; q < NS; ++q) st2(a.st[q] + g, ld2(Sk + q * sstride + r * nz));
        }
    }
}

template <int MODEL, int MATH>
const void* kernel_ptr(bool two) {
    return two ? (const void*)fd_resident2_kernel<MODEL, MATH> : (const void*)fd_resident_kernel<MODEL, MATH>;
}

const void* pick_kernel(int model, int math, bool two) {
    const bool ex = math == PDX_MATH_EXACT;
    switch (model) {
        case PDX_MODEL_NONE:     return ex ? kernel_ptr<PDX_MODEL_NONE, PDX_MATH_EXACT>(two) : kernel_ptr<PDX_MODEL_NONE, PDX_MATH_FAST>(two);
        case PDX_MODEL_FENTON4V: return ex ? kernel_ptr<PDX_MODEL_FENTON4V, PDX_MATH_EXACT>(two) : kernel_ptr<PDX_MODEL_FENTON4V, PDX_MATH_FAST>(two);
        case PDX_MODEL_MMS2V:    return ex ? kernel_ptr<PDX_MODEL_MMS2V, PDX_MATH_EXACT>(two) : kernel_ptr<PDX_MODEL_MMS2V, PDX_MATH_FAST>(two);
        case PDX_MODEL_MS2V:     return ex ? kernel_ptr<PDX_MODEL_MS2V, PDX_MATH_EXACT>(two) : kernel_ptr<PDX_MODEL_MS2V, PDX_MATH_FAST>(two);
    }
    return nullptr;
}

}  // namespace

// Geometry of the resident kernel for a 2-D grid, or false when the grid does not fit on chip /
// the configuration is not covered (heterogeneous operator, Dirichlet pad, lookup-table models).
bool fd_resident_plan(const pdx_fd_desc& d, ResidentGeom* g) {
    if (d.ndim != 2 || d.lap_kind != PDX_LAP_HOMOG || d.dirichlet) return false;
    if (d.pitch_z != 0 && d.pitch_z != d.nz) return false;          // contiguous rows only
    if (d.model != PDX_MODEL_NONE && d.model != PDX_MODEL_FENTON4V && d.model != PDX_MODEL_MMS2V && d.model != PDX_MODEL_MS2V)
        return false;
    if (d.nz > 1024 || d.nz < 3 || d.ny < 3) return false;
    int dev = 0, sms = 0, smem_max = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (sms < 1) return false;
    const int ns = pdx_model_num_states(d.model);
    const int rows = (d.ny + sms - 1) / sms;                         // strips as even as the SM count allows
    const size_t bytes = ((size_t)2 * (rows + 2) + (size_t)ns * rows) * d.nz * sizeof(float);
    if (bytes > (size_t)smem_max) return false;
    g->rows_per_cta = rows;
    g->ncta = (d.ny + rows - 1) / rows;
    g->threads = (d.nz + 31) / 32 * 32;
    g->smem_bytes = bytes;
    return true;
}

int launch_fd_resident(const pdx_fd_desc& d, const FdArgs& f, const ResidentGeom& g, int nsteps, unsigned int base,
                       unsigned int* flags, float* halo, cudaStream_t s) {
    // two nodes per thread when the rows can be moved as 64-bit words (PDX_RESIDENT_NPT=1 forces the one-node kernel)
    bool two = d.nz % 2 == 0 && d.nz >= 4 && ((uintptr_t)f.Uin % 8 == 0) && ((uintptr_t)f.Uout % 8 == 0);
    for (int i = 0; i < pdx_model_num_states(d.model); ++i) two = two && ((uintptr_t)f.st[i] % 8 == 0);
    {
        const char* e = std::getenv("PDX_RESIDENT_NPT");
        if (e && e[0] == '1') two = false;
    }
    const int threads = two ? (d.nz / 2 + 31) / 32 * 32 : g.threads;
    const void* fn = pick_kernel(d.model, d.math_mode, two);
    if (!fn) { set_error("resident FD kernel: unsupported model"); return 1; }
    if (check_cuda(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes),
                   "resident FD kernel: shared memory opt-in"))
        return 1;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, g.smem_bytes), "occupancy query") ||
        per_sm * sms < g.ncta) {
        if (per_sm * sms < g.ncta) set_error("resident FD kernel: the strips cannot all be resident");
        return 1;
    }
    ResArgs a{};
    a.Uin = f.Uin; a.Uout = f.Uout;
    for (int i = 0; i < PDX_MAX_STATES; ++i) a.st[i] = f.st[i];
    a.ny = d.ny; a.nz = d.nz;
    a.rows_per_cta = g.rows_per_cta;
    a.nsteps = nsteps;
    a.base = base; a.flags = flags; a.halo = halo;
    a.dt = f.dt; a.coef = f.coef;
    a.hysq = f.hysq; a.hzsq = f.hzsq; a.rhysq = f.rhysq; a.rhzsq = f.rhzsq;
    a.mp = f.mp;
    void* args[] = {&a};
    count_launch();
    return check_cuda(cudaLaunchCooperativeKernel(fn, dim3(g.ncta), dim3(threads), args, g.smem_bytes, s),
                      "fd_resident_kernel cooperative launch");
}

}  // namespace pdx
