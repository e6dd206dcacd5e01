// pdx_fem.cu - FEM side of the hot path: CSR SpMV, Jacobi-PCG as one persistent
// cooperative kernel, and the explicit mass-lumped step fused with the ionic update.
//
// Reference call sites: tf.raw_ops.SparseMatrixMatMul (linearsolvers/conjgrad.py:99-101,
// physics/monodomainSolver.py:104), ConjGrad._initialize/_iterate/solve
// (conjgrad.py:172-203,257-312), JacobiPrecond.solve_precond_system (jacobi_precond.py:35-41).
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <new>
#include "pdx_internal.h"
#include "pdx_fem_dev.cuh"

namespace cg = cooperative_groups;

namespace pdx {
namespace {

// ---------------------------------------------------------------- SpMV
template <int G>
__global__ void __launch_bounds__(256) spmv_kernel(int n, const int32_t* __restrict__ rowptr,
                                                   const int32_t* __restrict__ col, const float* __restrict__ val,
                                                   const float* __restrict__ x, float* __restrict__ y) {
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int sub = threadIdx.x % G;
    const int ngroups = gridDim.x * blockDim.x / G;
    const int nround = (n + ngroups - 1) / ngroups * ngroups;   // all lanes reach the shuffles
    for (int row = gid; row < nround; row += ngroups) {
        const bool valid = row < n;
        const float acc = row_dot<G>(col, val, x, valid ? rowptr[row] : 0, valid ? rowptr[row + 1] : 0, sub, false);
        if (valid && sub == 0) y[row] = acc;
    }
}

// ---------------------------------------------------------------- PCG (cooperative)
struct PcgArgs {
    int n;
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* minv;      // may be null
    const float* b;         // right-hand side, or null: b = M * rhs0 with M = mval on A's pattern
    const float* mval;
    const float* rhs0;
    float* x;
    float *r, *p, *q;       // work vectors
    float* p2;              // second direction buffer (the direction update is folded into the SpMV)
    float* part;            // 3 * gridDim.x partial sums
    int maxiter;
    int check_every;
    double toll_sq;
    int32_t* info;
};

template <int G>
__global__ void __launch_bounds__(256) pcg_kernel(const __grid_constant__ PcgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float sh[33];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int gid = tid / G, sub = threadIdx.x % G, ngroups = nthreads / G;
    const int nb = gridDim.x;
    float* part_pq = a.part;
    float* part_rr = a.part + nb;
    float* part_rz = a.part + 2 * nb;
    const int nround = (a.n + ngroups - 1) / ngroups * ngroups;   // keep groups converged for shuffles

    // r = b - A x ; z = Minv r ; p = z ; rz = r.z          (conjgrad.py:191-203)
    float loc_rz = 0.0f, loc_rr = 0.0f;
    for (int row = gid; row < nround; row += ngroups) {
        {
            const bool valid = row < a.n;
            const int beg = valid ? a.rowptr[row] : 0, end = valid ? a.rowptr[row + 1] : 0;
            const float ax = row_dot<G>(a.col, a.val, a.x, beg, end, sub, true);
            float bv = 0.0f;
            if (!a.b) bv = row_dot<G>(a.col, a.mval, a.rhs0, beg, end, sub, true);     // uniform branch: RHS = MASS*RHS0
            if (valid && sub == 0) {
                if (a.b) bv = a.b[row];
                const float r = __fsub_rn(bv, ax);
                const float z = a.minv ? __fmul_rn(a.minv[row], r) : r;
                __stcg(a.r + row, r);
                __stcg(a.p + row, z);
                loc_rz = fmaf(r, z, loc_rz);
                loc_rr = fmaf(r, r, loc_rr);
            }
        }
    }
    float v = block_sum(loc_rz, sh);
    float w = block_sum(loc_rr, sh);
    if (threadIdx.x == 0) { __stcg(part_rz + blockIdx.x, v); __stcg(part_rr + blockIdx.x, w); }
    grid.sync();
    float rz = grid_total(part_rz, sh);
    float res = grid_total(part_rr, sh);

    int niters = 0;
    int flags = 0;
    // TWO grid barriers per iteration.  The direction update p_k = z_{k-1} + beta_{k-1} p_{k-1} (conjgrad.py:184-187)
    // is not a pass of its own: the SpMV of iteration k forms it on the fly for every column it gathers
    // (p_old, r and Minv of the neighbours - all L2 resident at the sizes this kernel serves) with the very
    // fma the owner uses, and the row's owner stores p_k into the OTHER p buffer for the x update and the next
    // iteration.  Same values, same summation order as the three-pass loop; one pass and one barrier less.
    float* pbuf[2] = {a.p, a.p2};
    int cur = 0;                     // pbuf[cur] holds p_{k-1} (p_0 = z_0 for k = 1, used with beta = 0)
    float beta = 0.0f;
    for (int k = 1; k <= a.maxiter; ++k) {
        niters = k;
        const float* pold = pbuf[cur];
        float* pnew = pbuf[cur ^ 1];
        const bool first = k == 1;
        // q = A p ; pq = p.q
        float loc_pq = 0.0f;
        for (int row = gid; row < nround; row += ngroups) {
            const bool valid = row < a.n;
            const int beg = valid ? a.rowptr[row] : 0, end = valid ? a.rowptr[row + 1] : 0;
            float acc = 0.0f;
            for (int e = beg + sub; e < end; e += G) {
                const int c = a.col[e];
                float pj = __ldcg(pold + c);
                if (!first) {
                    const float rj = __ldcg(a.r + c);
                    pj = fmaf(beta, pj, a.minv ? a.minv[c] * rj : rj);
                }
                acc = fmaf(a.val[e], pj, acc);
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
            if (valid && sub == 0) {
                float pi = __ldcg(pold + row);
                if (!first) {
                    const float ri = __ldcg(a.r + row);
                    pi = fmaf(beta, pi, a.minv ? a.minv[row] * ri : ri);
                }
                __stcg(pnew + row, pi);
                __stcg(a.q + row, acc);
                loc_pq = fmaf(pi, acc, loc_pq);
            }
        }
        v = block_sum(loc_pq, sh);
        if (threadIdx.x == 0) __stcg(part_pq + blockIdx.x, v);
        grid.sync();
        const float pq = grid_total(part_pq, sh);
        const float alpha = rz / pq;
        // x += alpha p ; r -= alpha q ; res = r.r ; z = Minv r ; rz' = r.z   (conjgrad.py:172-183)
        loc_rz = 0.0f; loc_rr = 0.0f;
        for (int i = tid; i < a.n; i += nthreads) {
            const float pi = __ldcg(pnew + i);
            __stcg(a.x + i, fmaf(alpha, pi, __ldcg(a.x + i)));
            const float r = fmaf(-alpha, __ldcg(a.q + i), __ldcg(a.r + i));
            __stcg(a.r + i, r);
            const float z = a.minv ? a.minv[i] * r : r;
            loc_rr = fmaf(r, r, loc_rr);
            loc_rz = fmaf(r, z, loc_rz);
        }
        v = block_sum(loc_rz, sh);
        w = block_sum(loc_rr, sh);
        // the partials of the two phases live in different arrays and each array is rewritten two barriers later
        if (threadIdx.x == 0) { __stcg(part_rz + blockIdx.x, v); __stcg(part_rr + blockIdx.x, w); }
        grid.sync();
        const float rznew = grid_total(part_rz, sh);
        res = grid_total(part_rr, sh);
        // batched convergence test + NaN/Inf guard (conjgrad.py:293-300)
        if (k % a.check_every == 0) {
            const bool finite = isfinite(res);
            if (!finite || (double)res < a.toll_sq) {
                if (!finite) flags |= 2;
                break;
            }
        }
        beta = rznew / rz;
        rz = rznew;
        cur ^= 1;
    }
    if (niters >= a.maxiter) flags |= 1;
    if (tid == 0 && a.info) {
        a.info[0] = niters;
        a.info[1] = __float_as_int(res);
        a.info[2] = flags;
        a.info[3] = 0;
    }
}

// ---------------------------------------------------------------- PCG, register-resident (small systems)
// Config 1 (63 001 rows, 4.4e5 non-zeros) is 5 MB: everything sits in L2 and an iteration of pcg_kernel is a chain
// of dependent L2 round trips (rowptr -> col/val -> gather -> reduce, then the vector passes) between grid barriers.
// Here every lane group keeps ITS rows for the whole solve: the matrix entries (col, val, Minv of the column) and the
// row's x, r, Minv, p live in REGISTERS; an iteration is
//   phase A: gather p_old and z = Minv r of the columns (independent loads, one L2 round trip), form p_k on the fly,
//            q = A p_k, publish p_k of the own row, p.q partial             -> grid barrier
//   phase B: x, r of the own rows updated in registers, z published, r.r / r.z partials  -> grid barrier
// i.e. two barriers and ~one L2 round trip of latency per iteration; x is written once, at the end.  Rows are dealt out
// exactly as in pcg_kernel (row = group + k * #groups), entries of a row to the lanes in the same order, so every sum
// has the same order and the iterates are bitwise those of pcg_kernel with the same grid.
// Needs at most K rows per group and rows of at most E*G entries (checked by the host at pdx_pcg_create).
template <int G, int K, int E>
__global__ void __launch_bounds__(256, K <= 2 ? 4 : (K <= 4 ? 2 : 1)) pcg_small_kernel(const __grid_constant__ PcgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float sh[33];
    __shared__ float sh2[66];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int gid = tid / G, sub = threadIdx.x % G, ngroups = nthreads / G;
    const int nb = gridDim.x;
    float* part_pq = a.part;
    float* part_rr = a.part + nb;
    float* part_rz = a.part + 2 * nb;

    int   ec[K][E];          // column of my e-th entry of my k-th row
    float ev[K][E];          // its value (0 where the row is shorter)
    float xk[K], rk[K], mk[K], pk[K], qk[K];
    bool  own[K];            // this lane owns row k (lane 0 of the group, row < n)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int row = gid + k * ngroups;
        const bool valid = row < a.n;
        const int beg = valid ? a.rowptr[row] : 0, end = valid ? a.rowptr[row + 1] : 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int idx = beg + sub + e * G;
            const bool has = idx < end;
            ec[k][e] = has ? a.col[idx] : 0;
            ev[k][e] = has ? a.val[idx] : 0.0f;
        }
        own[k] = valid && sub == 0;
        xk[k] = own[k] ? a.x[row] : 0.0f;
        mk[k] = (own[k] && a.minv) ? a.minv[row] : 1.0f;
        rk[k] = pk[k] = qk[k] = 0.0f;
    }

    // r = b - A x ; z = Minv r ; p = z ; rz = r.z          (conjgrad.py:191-203)
    float loc_rz = 0.0f, loc_rr = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int row = gid + k * ngroups;
        const bool valid = row < a.n;
        float ax = 0.0f;
#pragma unroll
        for (int e = 0; e < E; ++e) ax = fmaf(ev[k][e], a.x[ec[k][e]], ax);     // padded entries add +0
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) ax += __shfl_xor_sync(0xffffffffu, ax, o, G);
        float bv = 0.0f;
        if (!a.b) bv = row_dot<G>(a.col, a.mval, a.rhs0, valid ? a.rowptr[row] : 0, valid ? a.rowptr[row + 1] : 0, sub, true);
        if (own[k]) {
            if (a.b) bv = a.b[row];
            const float r = __fsub_rn(bv, ax);
            const float z = a.minv ? __fmul_rn(mk[k], r) : r;
            rk[k] = r;
            pk[k] = z;
            __stcg(a.p + row, z);
            loc_rz = fmaf(r, z, loc_rz);
            loc_rr = fmaf(r, r, loc_rr);
        }
    }
    float v = block_sum(loc_rz, sh);
    float w = block_sum(loc_rr, sh);
    if (threadIdx.x == 0) { __stcg(part_rz + blockIdx.x, v); __stcg(part_rr + blockIdx.x, w); }
    grid.sync();
    float rz = grid_total(part_rz, sh);
    float res = grid_total(part_rr, sh);

    int niters = 0, flags = 0, cur = 0;
    float* pbuf[2] = {a.p, a.p2};
    float beta = 0.0f;
    for (int it = 1; it <= a.maxiter; ++it) {
        niters = it;
        const float* pold = pbuf[cur];
        float* pnew = pbuf[cur ^ 1];
        const bool first = it == 1;
        // ---- phase A: all gathers of all my rows first (independent loads), then the arithmetic
        float gp[K][E], gr[K][E];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int e = 0; e < E; ++e) {
                gp[k][e] = __ldcg(pold + ec[k][e]);
                gr[k][e] = first ? 0.0f : __ldcg(a.r + ec[k][e]);       // z = Minv r of the column, published by its owner
            }
        float loc_pq = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float acc = 0.0f;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                float pj = gp[k][e];
                if (!first) pj = fmaf(beta, pj, gr[k][e]);
                acc = fmaf(ev[k][e], pj, acc);
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
            if (own[k]) {
                if (!first) pk[k] = fmaf(beta, pk[k], a.minv ? mk[k] * rk[k] : rk[k]);
                __stcg(pnew + gid + k * ngroups, pk[k]);
                qk[k] = acc;
                loc_pq = fmaf(pk[k], acc, loc_pq);
            }
        }
        v = block_sum(loc_pq, sh);
        if (threadIdx.x == 0) __stcg(part_pq + blockIdx.x, v);
        grid.sync();
        const float pq = grid_total(part_pq, sh);
        const float alpha = rz / pq;
        // ---- phase B: x += alpha p ; r -= alpha q ; res = r.r ; z = Minv r ; rz' = r.z, in registers
        loc_rz = 0.0f; loc_rr = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (own[k]) {
                xk[k] = fmaf(alpha, pk[k], xk[k]);
                const float r = fmaf(-alpha, qk[k], rk[k]);
                rk[k] = r;
                const float z = a.minv ? mk[k] * r : r;
                __stcg(a.r + gid + k * ngroups, z);                 // the work vector "r" carries z for the gathers
                loc_rr = fmaf(r, r, loc_rr);
                loc_rz = fmaf(r, z, loc_rz);
            }
        }
        block_sum2(loc_rz, loc_rr, sh2, v, w);              // both sums through one pair of block barriers
        if (threadIdx.x == 0) { __stcg(part_rz + blockIdx.x, v); __stcg(part_rr + blockIdx.x, w); }
        grid.sync();
        float rznew;
        grid_total2(part_rz, part_rr, sh2, rznew, res);
        if (it % a.check_every == 0) {
            const bool finite = isfinite(res);
            if (!finite || (double)res < a.toll_sq) {
                if (!finite) flags |= 2;
                break;
            }
        }
        beta = rznew / rz;
        rz = rznew;
        cur ^= 1;
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (own[k]) a.x[gid + k * ngroups] = xk[k];
    if (niters >= a.maxiter) flags |= 1;
    if (tid == 0 && a.info) {
        a.info[0] = niters;
        a.info[1] = __float_as_int(res);
        a.info[2] = flags;
        a.info[3] = 0;
    }
}

__global__ void max_row_kernel(int n, const int32_t* __restrict__ rowptr, int* out) {
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, rowptr[i + 1] - rowptr[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// ---------------------------------------------------------------- explicit lumped step
// Rows are LOCAL (0..n-1) and belong to global rows row_lo..row_lo+n-1 of a row-block partition;
// columns, U_in and U_out are GLOBAL (every rank holds full-length U buffers of which only its own
// rows and its ghost columns are valid); states, mlinv and I0 are local.  The new values of the
// first send_lo / last send_hi local rows are also stored into the neighbouring ranks' U_out
// buffers (NVLink peer memory) at the same global index: the ghost exchange is fused into the step.
// A non-partitioned call is row_lo = 0 with no mirrors.
struct ExplArgs {
    int n;
    int row_lo;
    const int32_t* rowptr;
    const int32_t* col;
    const float* kval;
    const float* mlinv;
    const float* Uin;
    float* Uout;
    float* st[PDX_MAX_STATES];
    const float* I0;
    float* mirror_lo;
    float* mirror_hi;
    int send_lo, send_hi;
    // neighbour flags (pdx_fem_explicit_step_sync): this rank's control words and the two neighbours'
    unsigned int* ctrl;
    unsigned int* ctrl_lo;
    unsigned int* ctrl_hi;
    int stage_slot_bytes;          // staged kernel: bytes of a ring slot (col + val words of the widest slice)
    int n_lo_ctas, n_hi_ctas;      // CTAs that hold rows < send_lo / >= n - send_hi (set by the launcher)
    int block_shift;               // SELL: CTA b works on row block (b + block_shift) % gridDim.x - the high edge first
    ModelParams mp;
};

// Control words of a rank (symmetric memory, zeroed by the caller).  Only the CTAs at the two ends of the row block
// take part: they are the ones that read ghost rows and store into a neighbour (the matrix is structurally symmetric),
// so "the lower neighbour may go on" means "my low-edge CTAs are done".  Per side: steps published so far, CTAs of the
// running step that are done; and the steps the neighbour on that side has published (written by the neighbour).
// A low-edge CTA of step k waits until the lower neighbour has published k steps: its ghost rows of step k-1 are then
// in U_in, and it has finished reading the buffer this step's ghost stores overwrite.  Interior CTAs never wait and
// never count (39 000 CTAs arriving on one word cost more than the barrier they replace: measured 0.324 against
// 0.293 ms per step on two GPUs).  Replaces the host-side barrier between steps.
enum { CTRL_STEP_LO = 0, CTRL_DONE_LO = 1, CTRL_STEP_HI = 2, CTRL_DONE_HI = 3, CTRL_FROM_LO = 4, CTRL_FROM_HI = 5 };

__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// every thread of an edge CTA calls it before touching U_in (lo / hi: CTA uniform)
__device__ __forceinline__ void expl_wait_neighbours(const ExplArgs& a, bool lo, bool hi) {
    if (!(lo || hi)) return;
    if (threadIdx.x == 0) {
        if (lo) {
            const unsigned int k = *((volatile unsigned int*)(a.ctrl + CTRL_STEP_LO));
            while ((int)(ld_acquire_sys_u32(a.ctrl + CTRL_FROM_LO) - k) < 0) { }
        }
        if (hi) {
            const unsigned int k = *((volatile unsigned int*)(a.ctrl + CTRL_STEP_HI));
            while ((int)(ld_acquire_sys_u32(a.ctrl + CTRL_FROM_HI) - k) < 0) { }
        }
    }
    __syncthreads();
}
// every thread of an edge CTA calls it after its last store; pushed = this thread stored into a neighbour's buffer
__device__ __forceinline__ void expl_signal_neighbours(const ExplArgs& a, bool lo, bool hi, bool pushed) {
    if (!(lo || hi)) return;
    if (pushed) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (lo && atomicAdd(a.ctrl + CTRL_DONE_LO, 1u) == (unsigned int)a.n_lo_ctas - 1) {   // last low-edge CTA
            const unsigned int k = a.ctrl[CTRL_STEP_LO] + 1;
            a.ctrl[CTRL_DONE_LO] = 0;
            a.ctrl[CTRL_STEP_LO] = k;
            __threadfence_system();
            st_release_sys_u32(a.ctrl_lo + CTRL_FROM_HI, k);       // I am my lower neighbour's UPPER neighbour
        }
        if (hi && atomicAdd(a.ctrl + CTRL_DONE_HI, 1u) == (unsigned int)a.n_hi_ctas - 1) {
            const unsigned int k = a.ctrl[CTRL_STEP_HI] + 1;
            a.ctrl[CTRL_DONE_HI] = 0;
            a.ctrl[CTRL_STEP_HI] = k;
            __threadfence_system();
            st_release_sys_u32(a.ctrl_hi + CTRL_FROM_LO, k);
        }
    }
}

template <int MODEL, int MATH, int G>
__global__ void __launch_bounds__(256) fem_explicit_kernel(const __grid_constant__ ExplArgs a) {
    constexpr int NS = NStates<MODEL>::value;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int sub = threadIdx.x % G;
    const int ngroups = gridDim.x * blockDim.x / G;
    const int nround = (a.n + ngroups - 1) / ngroups * ngroups;
    // grid-stride rows: every CTA holds rows of both edges
    const bool elo = a.ctrl && a.ctrl_lo, ehi = a.ctrl && a.ctrl_hi;
    expl_wait_neighbours(a, elo, ehi);
    bool pushed = false;
    for (int row = gid; row < nround; row += ngroups) {
        const bool valid = row < a.n;
        const float ku = row_dot<G>(a.col, a.kval, a.Uin, valid ? a.rowptr[row] : 0, valid ? a.rowptr[row + 1] : 0,
                                    sub, false);
        if (valid && sub == 0) {
            const size_t grow = (size_t)a.row_lo + row;
            const float U = a.Uin[grow];
            float st[PDX_MAX_STATES];
#pragma unroll
            for (int s = 0; s < NS; ++s) st[s] = a.st[s][row];
            float dU = MODEL != PDX_MODEL_NONE ? ionic_update<MODEL, MATH>(U, st, a.mp) : 0.0f;
#pragma unroll
            for (int s = 0; s < NS; ++s) a.st[s][row] = st[s];
            if (a.I0) dU += a.I0[row];
            // U1 = U + dt*(dU + I0) - dt*mlinv*(K U)
            const float u1 = fmaf(-a.mp.dt * a.mlinv[row], ku, fmaf(a.mp.dt, dU, U));
            a.Uout[grow] = u1;
            if (a.mirror_lo && row < a.send_lo) { a.mirror_lo[grow] = u1; pushed = true; }
            if (a.mirror_hi && row >= a.n - a.send_hi) { a.mirror_hi[grow] = u1; pushed = true; }
        }
    }
    expl_signal_neighbours(a, elo, ehi, pushed);
}

// SELL-32 (sliced ELLPACK, slice height = warp size) form of the same step: ONE THREAD PER ROW.
// Slice s holds rows 32s..32s+31 column-major: entry k of lane l at slice_ptr[s] + 32k + l, rows padded
// to the slice's longest row with (col = own row, val = 0).  Every col/val load is a fully coalesced
// 128-byte warp access, each thread keeps a whole row of gathers in flight, and - unlike the
// lanes-per-row CSR kernel, where one lane in G does the ionic update - all 32 lanes run the ionic
// model.  Products are accumulated in row order (k = 0, 1, ...).
template <int MODEL, int MATH>
__global__ void __launch_bounds__(256) fem_explicit_sell_kernel(const __grid_constant__ ExplArgs a) {
    constexpr int NS = NStates<MODEL>::value;
    int vb = blockIdx.x + a.block_shift;
    if (vb >= (int)gridDim.x) vb -= gridDim.x;
    const int row = vb * blockDim.x + threadIdx.x;
    const int slice = row >> 5, lane = threadIdx.x & 31;
    const bool elo = a.ctrl && a.ctrl_lo && vb * (int)blockDim.x < a.send_lo;
    const bool ehi = a.ctrl && a.ctrl_hi && (vb + 1) * (int)blockDim.x > a.n - a.send_hi;
    expl_wait_neighbours(a, elo, ehi);
    bool pushed = false;
    if (slice * 32 < a.n) {                              // (a whole warp may be out of range)
        const int beg = a.rowptr[slice], end = a.rowptr[slice + 1];     // rowptr = slice_ptr here
        const int32_t* __restrict__ col = a.col;
        const float* __restrict__ val = a.kval;
        float ku = 0.0f;
#pragma unroll 4
        for (int k = beg + lane; k < end; k += 32) ku = fmaf(val[k], a.Uin[col[k]], ku);
        if (row < a.n) {
            const size_t grow = (size_t)a.row_lo + row;
            const float U = a.Uin[grow];
            float st[PDX_MAX_STATES];
#pragma unroll
            for (int s = 0; s < NS; ++s) st[s] = a.st[s][row];
            float dU = MODEL != PDX_MODEL_NONE ? ionic_update<MODEL, MATH>(U, st, a.mp) : 0.0f;
#pragma unroll
            for (int s = 0; s < NS; ++s) a.st[s][row] = st[s];
            if (a.I0) dU += a.I0[row];
            const float u1 = fmaf(-a.mp.dt * a.mlinv[row], ku, fmaf(a.mp.dt, dU, U));
            a.Uout[grow] = u1;
            if (a.mirror_lo && row < a.send_lo) { a.mirror_lo[grow] = u1; pushed = true; }
            if (a.mirror_hi && row >= a.n - a.send_hi) { a.mirror_hi[grow] = u1; pushed = true; }
        }
    }
    expl_signal_neighbours(a, elo, ehi, pushed);
}

// The same step for LARGE blocks of rows: a persistent grid whose warps walk their slices with the matrix stream on the
// copy engine (sell_rows_staged, pdx_fem_dev.cuh): the col / val words of the next slices are in flight while a warp
// gathers and runs the ionic model - the plain kernel above waits for slice_ptr, then col / val, then the gathers
// (0.88 of the HBM peak on config 5).
template <int MODEL, int MATH>
__global__ void __launch_bounds__(512, 2) fem_explicit_sell_staged_kernel(const __grid_constant__ ExplArgs a) {
    constexpr int NS = NStates<MODEL>::value;
    extern __shared__ __align__(128) unsigned char dsm[];
    sell_stage_init(dsm, a.stage_slot_bytes);
    // Neighbour flags per SLICE (opt-in, sync='flags'): a warp waits for a neighbour right before its first slice that
    // holds edge rows, every finished edge slice counts towards this rank's publication.  Slices are dealt out round robin,
    // so the low edge comes first and the high edge last for every warp.  Measured on 8 GPUs (config 5, 2.5 M rows each):
    // 0.083 ms per step, edge slices walked last 0.090, all CTAs waiting at the start 0.080 - against 0.076-0.077 with a
    // barrier between steps, which therefore stays the default of PartitionedExplicitFEM.
    const bool flo = a.ctrl && a.ctrl_lo, fhi = a.ctrl && a.ctrl_hi;
    const int lo_slices = (a.send_lo + 31) >> 5;                 // slices [0, lo_slices) hold rows < send_lo
    const int hi_first = (a.n - a.send_hi) >> 5;                 // slices [hi_first, nsl) hold rows >= n - send_hi
    const int nsl = (a.n + 31) >> 5;
    const int lane = threadIdx.x & 31;
    bool waited_lo = false, waited_hi = false, pushed = false;
    unsigned int uses = 0;
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    sell_rows_staged(dsm, a.stage_slot_bytes, a.n, a.rowptr, a.col, a.kval, policy, uses,
                     [&](int c) { return a.Uin[c]; },
                     [&](int row, float ku) {
        const size_t grow = (size_t)a.row_lo + row;
        const float U = a.Uin[grow];
        float st[PDX_MAX_STATES];
#pragma unroll
        for (int s = 0; s < NS; ++s) st[s] = a.st[s][row];
        float dU = MODEL != PDX_MODEL_NONE ? ionic_update<MODEL, MATH>(U, st, a.mp) : 0.0f;
#pragma unroll
        for (int s = 0; s < NS; ++s) a.st[s][row] = st[s];
        if (a.I0) dU += a.I0[row];
        const float u1 = fmaf(-a.mp.dt * a.mlinv[row], ku, fmaf(a.mp.dt, dU, U));
        a.Uout[grow] = u1;
        if (a.mirror_lo && row < a.send_lo) { a.mirror_lo[grow] = u1; pushed = true; }
        if (a.mirror_hi && row >= a.n - a.send_hi) { a.mirror_hi[grow] = u1; pushed = true; }
    },
                     [&](int sl) {          // before the gathers of slice sl
        if (flo && !waited_lo && sl < lo_slices) {
            if (lane == 0) {
                const unsigned int k = *((volatile unsigned int*)(a.ctrl + CTRL_STEP_LO));
                while ((int)(ld_acquire_sys_u32(a.ctrl + CTRL_FROM_LO) - k) < 0) { }
            }
            __syncwarp();
            waited_lo = true;
        }
        if (fhi && !waited_hi && sl >= hi_first) {
            if (lane == 0) {
                const unsigned int k = *((volatile unsigned int*)(a.ctrl + CTRL_STEP_HI));
                while ((int)(ld_acquire_sys_u32(a.ctrl + CTRL_FROM_HI) - k) < 0) { }
            }
            __syncwarp();
            waited_hi = true;
        }
    },
                     [&](int sl) {          // after the rows of slice sl
        const bool elo = flo && sl < lo_slices, ehi = fhi && sl >= hi_first;
        if (!(elo || ehi)) return;
        if (pushed) __threadfence_system();
        pushed = false;
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            if (elo && atomicAdd(a.ctrl + CTRL_DONE_LO, 1u) == (unsigned int)lo_slices - 1) {
                const unsigned int k = a.ctrl[CTRL_STEP_LO] + 1;
                a.ctrl[CTRL_DONE_LO] = 0;
                a.ctrl[CTRL_STEP_LO] = k;
                __threadfence_system();
                st_release_sys_u32(a.ctrl_lo + CTRL_FROM_HI, k);
            }
            if (ehi && atomicAdd(a.ctrl + CTRL_DONE_HI, 1u) == (unsigned int)(nsl - hi_first) - 1) {
                const unsigned int k = a.ctrl[CTRL_STEP_HI] + 1;
                a.ctrl[CTRL_DONE_HI] = 0;
                a.ctrl[CTRL_STEP_HI] = k;
                __threadfence_system();
                st_release_sys_u32(a.ctrl_hi + CTRL_FROM_LO, k);
            }
        }
    },
                     [](int v) { return v; });
}

// staged launch geometry for a slice width registered through pdx_sell_hint: threads, blocks, dynamic shared memory
struct ExplStageGeom {
    int threads, blocks;
    size_t smem;
};
template <int MODEL, int MATH>
bool expl_stage_geometry(int wmax, ExplStageGeom* g) {
    static ExplStageGeom cache[PDX_MAX_DEVICES][64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PDX_MAX_DEVICES || wmax < 1 || wmax >= 64) return false;
    ExplStageGeom& c = cache[dev][wmax];
    if (c.threads == 0) {
        c.threads = -1;
        auto kern = fem_explicit_sell_staged_kernel<MODEL, MATH>;
        const int slot = wmax * 256;
        int smem_max = 0;
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {
            const int max_dyn = smem_max - (int)fa.sharedSizeBytes;
            if (max_dyn > 0 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) == cudaSuccess) {
                int best_warps = 0;
                for (int cand : {512, 384, 256}) {
                    const size_t need = sell_stage_smem_bytes(cand, slot);
                    if (need > (size_t)max_dyn) continue;
                    int blk = 0;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blk, kern, cand, need) != cudaSuccess) continue;
                    if (blk * cand / 32 > best_warps) {
                        best_warps = blk * cand / 32;
                        c.threads = cand; c.blocks = blk * sm_count(); c.smem = need;
                    }
                }
                if (best_warps < 16) c.threads = -1;
            }
        }
        cudaGetLastError();
    }
    if (c.threads <= 0) return false;
    *g = c;
    return true;
}

template <int MODEL, int MATH>
int launch_expl_sell(const ExplArgs& a0, cudaStream_t s) {
    // large row blocks whose slice width is known (pdx_sell_hint): persistent staged kernel.  PDX_EXPL_STAGE=0 / 1
    // turn it off / force it for any size (A/B runs, tests).
    {
        const char* e = getenv("PDX_EXPL_STAGE");
        const int mode = e ? (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2)) : 2;
        const int wmax = sell_width_hint(a0.rowptr);
        ExplStageGeom g;
        if (getenv("PDX_DEBUG")) fprintf(stderr, "[pdx] explicit sell: n %d wmax %d mode %d\n", a0.n, wmax, mode);
        if (wmax > 0 && mode != 0 && (a0.n >= 1000000 || mode == 1) && ((uintptr_t)a0.col % 16) == 0 && ((uintptr_t)a0.kval % 16) == 0 &&
            expl_stage_geometry<MODEL, MATH>(wmax, &g)) {
            ExplArgs a = a0;
            a.stage_slot_bytes = wmax * 256;
            const int warps_needed = (a.n + 31) / 32;
            int nb = g.blocks;
            const int per_block = g.threads / 32;
            if ((long long)nb * per_block > warps_needed) nb = (warps_needed + per_block - 1) / per_block;
            a.n_lo_ctas = a.n_hi_ctas = nb;
            fem_explicit_sell_staged_kernel<MODEL, MATH><<<(unsigned)nb, g.threads, g.smem, s>>>(a);
            count_launch();
            return check_cuda(cudaGetLastError(), "fem_explicit_sell_staged_kernel launch");
        }
    }
    ExplArgs a = a0;
    const int nb = (a.n + 255) / 256;
    if (a.ctrl) {
        a.n_lo_ctas = a.ctrl_lo ? (a.send_lo + 255) / 256 : 0;
        const int hb = (a.n - a.send_hi) / 256;                  // first CTA that holds a high-edge row
        a.n_hi_ctas = a.ctrl_hi ? nb - hb : 0;
        a.block_shift = a.ctrl_hi ? hb : 0;                       // run order: high edge, low edge, interior
    }
    fem_explicit_sell_kernel<MODEL, MATH><<<(unsigned)nb, 256, 0, s>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "fem_explicit_sell_kernel launch");
}

template <int MODEL, int MATH>
int launch_expl(int G, const ExplArgs& a0, unsigned nb, cudaStream_t s) {
    ExplArgs a = a0;
    a.n_lo_ctas = a.n_hi_ctas = (int)nb;
    switch (G) {
        case 4:  fem_explicit_kernel<MODEL, MATH, 4><<<nb, 256, 0, s>>>(a); break;
        case 8:  fem_explicit_kernel<MODEL, MATH, 8><<<nb, 256, 0, s>>>(a); break;
        case 16: fem_explicit_kernel<MODEL, MATH, 16><<<nb, 256, 0, s>>>(a); break;
        default: fem_explicit_kernel<MODEL, MATH, 32><<<nb, 256, 0, s>>>(a); break;
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "fem_explicit_kernel launch");
}

}  // namespace
}  // namespace pdx

using namespace pdx;

constexpr int PCG_SMALL_E = 2;     // entries per lane and row

struct pdx_pcg {
    int n;
    int64_t nnz;
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* minv;
    float* work;      // r, p, q
    float* part;
    int G;
    int nblocks;
    int max_row;     // longest row (selects the register-resident kernel)
    bool small;      // rows per lane group <= 4 and rows <= 2 Gs entries: pcg_small_kernel<Gs, Ks, 2>
    int Gs, Ks;
};

extern "C" {

int pdx_csr_spmv(int32_t n, const int32_t* rowptr, const int32_t* col, const float* val, const float* x, float* y,
                 void* stream) {
    if (n < 0 || !rowptr || !col || !val || !x || !y) { set_error("pdx_csr_spmv: bad arguments"); return 1; }
    if (n == 0) return 0;
    // group size from a cheap upper bound: read nnz lazily is a device sync, so use 8 lanes
    // (P1 triangle/tet meshes have 7-15 entries per row)
    constexpr int G = 8;
    long long threads = (long long)n * G;
    long long nb = (threads + 255) / 256;
    if (nb > (long long)sm_count() * 32) nb = (long long)sm_count() * 32;
    // round rows per pass so that groups stay converged: handled by the kernel loop bound (row < n)
    spmv_kernel<G><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(n, rowptr, col, val, x, y);
    count_launch();
    return check_cuda(cudaGetLastError(), "spmv_kernel launch");
}

int pdx_pcg_create(int32_t n, const int32_t* rowptr, const int32_t* col, const float* val, const float* minv,
                   pdx_pcg** out) {
    if (n <= 0 || !rowptr || !col || !val || !out) { set_error("pdx_pcg_create: bad arguments"); return 1; }
    pdx_pcg* s = new (std::nothrow) pdx_pcg();
    if (!s) { set_error("out of memory"); return 1; }
    s->n = n; s->rowptr = rowptr; s->col = col; s->val = val; s->minv = minv;
    int32_t last = 0;
    if (check_cuda(cudaMemcpy(&last, rowptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost), "read nnz")) { delete s; return 1; }
    s->nnz = last;
    s->G = pick_group(n, s->nnz);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 0;
    cudaError_t e;
    switch (s->G) {
        case 4:  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel<4>, 256, 0); break;
        case 8:  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel<8>, 256, 0); break;
        case 16: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel<16>, 256, 0); break;
        default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel<32>, 256, 0); break;
    }
    if (check_cuda(e, "occupancy query") || per_sm < 1) { delete s; if (per_sm < 1) set_error("pcg kernel cannot be resident"); return 1; }
    // small systems are latency bound: one block per SM keeps the grid barrier cheap
    long long want = ((long long)n * s->G + 255) / 256;
    long long cap = (long long)sms * (per_sm > 4 ? 4 : per_sm);
    long long nb = want < cap ? want : cap;
    if (nb < 1) nb = 1;
    if (nb > sms && (long long)n * s->G <= (long long)sms * 256 * 8) nb = sms;
    {   // PDX_PCG_BLOCKS=<n>: grid size override for tuning runs (clamped to what can be co-resident)
        const char* e = std::getenv("PDX_PCG_BLOCKS");
        if (e && std::atoi(e) > 0) nb = std::min<long long>(std::atoi(e), (long long)sms * per_sm);
    }
    s->nblocks = (int)nb;
    {   // longest row -> can the register-resident kernel take this system?  (one-time set-up read, like nnz above)
        int* dmax = nullptr;
        int hmax = 0;
        if (check_cuda(cudaMalloc(&dmax, sizeof(int)), "cudaMalloc(max row)") ||
            check_cuda(cudaMemset(dmax, 0, sizeof(int)), "clear max row")) { delete s; return 1; }
        max_row_kernel<<<64, 256>>>(n, rowptr, dmax);
        const bool bad = check_cuda(cudaMemcpy(&hmax, dmax, sizeof(int), cudaMemcpyDeviceToHost), "read max row") != 0;
        cudaFree(dmax);
        if (bad) { delete s; return 1; }
        s->max_row = hmax;
        // lanes per row of the register-resident kernel: the fewest that hold a row in PCG_SMALL_E entries per lane
        // (more groups = fewer rows per group = fewer registers), then 2 or 4 rows per group
        s->Gs = hmax <= 4 * PCG_SMALL_E ? 4 : hmax <= 8 * PCG_SMALL_E ? 8 : hmax <= 16 * PCG_SMALL_E ? 16 : 32;
        const char* e = std::getenv("PDX_PCG_SMALL");           // A/B switch: PDX_PCG_SMALL=0 keeps the loop kernel
        const char* eb = std::getenv("PDX_PCG_BLOCKS");
        const bool allowed = !(e && e[0] == '0') && hmax <= PCG_SMALL_E * s->Gs;
        auto rows_per_group = [&](long long blocks) {
            const long long ngroups = blocks * 256 / s->Gs;
            return (n + ngroups - 1) / ngroups;
        };
        s->small = false;
        if (allowed) {
            // measured on config 1 (tools/tune_pcg_blocks.py, r02g): 148 blocks x 8 rows per group 0.080 ms per step,
            // 296 x 4: 0.083, 592 x 2: 0.097 - the grid barrier gets cheaper faster than the phases get longer.
            // Without an override take the fewest blocks in {1, 2, 4} x SMs whose rows per group fit the registers.
            const long long cand[3] = {1LL * sms, 2LL * sms, 4LL * sms};
            const long long cap[3] = {8, 4, 2};
            long long pick = 0;
            if (eb && std::atoi(eb) > 0) pick = s->nblocks;
            else for (int c = 0; c < 3; ++c) if (!pick && rows_per_group(cand[c]) <= cap[c]) pick = cand[c];
            if (pick) {
                const long long rpg = rows_per_group(pick);
                const int K = rpg <= 2 ? 2 : rpg <= 4 ? 4 : 8;
                const long long per = K == 2 ? 4 : K == 4 ? 2 : 1;          // launch bounds of pcg_small_kernel
                if (rpg <= 8 && pick <= per * sms) { s->small = true; s->Ks = K; s->nblocks = (int)pick; }
            }
        }
    }
    if (check_cuda(cudaMalloc(&s->work, sizeof(float) * 4 * (size_t)n), "cudaMalloc(pcg work)")) { delete s; return 1; }
    if (check_cuda(cudaMalloc(&s->part, sizeof(float) * 3 * (size_t)s->nblocks), "cudaMalloc(pcg partials)")) {
        cudaFree(s->work); delete s; return 1;
    }
    *out = s;
    return 0;
}

void pdx_pcg_destroy(pdx_pcg* s) {
    if (!s) return;
    cudaFree(s->work);
    cudaFree(s->part);
    delete s;
}

static int pcg_launch(pdx_pcg* s, const float* b, const float* mval, const float* rhs0, float* x, int maxiter, double toll,
                      int check_every, int32_t* info, void* stream);

int pdx_pcg_solve(pdx_pcg* s, const float* b, float* x, int maxiter, double toll, int check_every, int32_t* info,
                  void* stream) {
    if (!s || !b || !x || maxiter < 0) { set_error("pdx_pcg_solve: bad arguments"); return 1; }
    return pcg_launch(s, b, nullptr, nullptr, x, maxiter, toll, check_every, info, stream);
}

int pdx_pcg_solve_rhs0(pdx_pcg* s, const float* mval, const float* rhs0, float* x, int maxiter, double toll,
                       int check_every, int32_t* info, void* stream) {
    if (!s || !mval || !rhs0 || !x || maxiter < 0) { set_error("pdx_pcg_solve_rhs0: bad arguments"); return 1; }
    return pcg_launch(s, nullptr, mval, rhs0, x, maxiter, toll, check_every, info, stream);
}

static int pcg_launch(pdx_pcg* s, const float* b, const float* mval, const float* rhs0, float* x, int maxiter, double toll,
                      int check_every, int32_t* info, void* stream) {
    PcgArgs a{};
    a.n = s->n; a.rowptr = s->rowptr; a.col = s->col; a.val = s->val; a.minv = s->minv;
    a.b = b; a.mval = mval; a.rhs0 = rhs0; a.x = x;
    a.r = s->work; a.p = s->work + s->n; a.q = s->work + 2 * (size_t)s->n; a.p2 = s->work + 3 * (size_t)s->n;
    a.part = s->part;
    a.maxiter = maxiter;
    a.check_every = check_every > 0 ? check_every : 1;
    a.toll_sq = toll * toll;
    a.info = info;
    void* args[] = {&a};
    const void* fn;
    if (s->small) {
#define PDX_SMALL_PICK(GG) (s->Ks == 2 ? (const void*)pcg_small_kernel<GG, 2, PCG_SMALL_E> : s->Ks == 4 ? \
                           (const void*)pcg_small_kernel<GG, 4, PCG_SMALL_E> : (const void*)pcg_small_kernel<GG, 8, PCG_SMALL_E>)
        switch (s->Gs) {
            case 4:  fn = PDX_SMALL_PICK(4); break;
            case 8:  fn = PDX_SMALL_PICK(8); break;
            case 16: fn = PDX_SMALL_PICK(16); break;
            default: fn = PDX_SMALL_PICK(32); break;
        }
#undef PDX_SMALL_PICK
    } else {
        switch (s->G) {
            case 4:  fn = (const void*)pcg_kernel<4>; break;
            case 8:  fn = (const void*)pcg_kernel<8>; break;
            case 16: fn = (const void*)pcg_kernel<16>; break;
            default: fn = (const void*)pcg_kernel<32>; break;
        }
    }
    count_launch();
    return check_cuda(cudaLaunchCooperativeKernel(fn, dim3(s->nblocks), dim3(256), args, 0, (cudaStream_t)stream),
                      "pcg_kernel cooperative launch");
}

static int explicit_step_impl(bool sell, int model, int math_mode, const float* params, int32_t n, int32_t row_lo,
                              const int32_t* rowptr, const int32_t* col, const float* kval, const float* mlinv,
                              const float* U_in, float* U_out, float* const* states, const float* I0, float dt,
                              int32_t send_lo, float* mirror_lo, int32_t send_hi, float* mirror_hi, void* stream,
                              uint32_t* ctrl = nullptr, uint32_t* ctrl_lo = nullptr, uint32_t* ctrl_hi = nullptr) {
    if (n <= 0 || row_lo < 0 || !rowptr || !col || !kval || !mlinv || !U_in || !U_out || U_in == U_out) {
        set_error("pdx_fem_explicit_step: bad arguments");
        return 1;
    }
    if (send_lo < 0 || send_hi < 0 || send_lo > n || send_hi > n) { set_error("pdx_fem_explicit_step: bad send ranges"); return 1; }
    const int ns = pdx_model_num_states(model);
    if (ns < 0 || ns > PDX_MAX_STATES || (ns > 0 && (!states || !params))) {
        set_error("pdx_fem_explicit_step: bad model/states (closed-form models only)");
        return 1;
    }
    ExplArgs a{};
    a.n = n; a.row_lo = row_lo; a.rowptr = rowptr; a.col = col; a.kval = kval; a.mlinv = mlinv;
    a.Uin = U_in; a.Uout = U_out; a.I0 = I0;
    a.mirror_lo = send_lo > 0 ? mirror_lo : nullptr; a.mirror_hi = send_hi > 0 ? mirror_hi : nullptr;
    a.send_lo = send_lo; a.send_hi = send_hi;
    // a neighbour takes part in the flag protocol when rows are exchanged with it
    a.ctrl = ctrl; a.ctrl_lo = ctrl && send_lo > 0 ? ctrl_lo : nullptr; a.ctrl_hi = ctrl && send_hi > 0 ? ctrl_hi : nullptr;
    for (int i = 0; i < ns; ++i) a.st[i] = states[i];
    if (ns > 0) fill_model_params(model, params, dt, &a.mp); else a.mp.dt = dt;
    cudaStream_t s = (cudaStream_t)stream;
    const bool ex = math_mode == PDX_MATH_EXACT;
    if (sell) {
        switch (model) {
            case PDX_MODEL_NONE:     return launch_expl_sell<PDX_MODEL_NONE, PDX_MATH_FAST>(a, s);
            case PDX_MODEL_FENTON4V: return ex ? launch_expl_sell<PDX_MODEL_FENTON4V, PDX_MATH_EXACT>(a, s)
                                               : launch_expl_sell<PDX_MODEL_FENTON4V, PDX_MATH_FAST>(a, s);
            case PDX_MODEL_MMS2V:    return ex ? launch_expl_sell<PDX_MODEL_MMS2V, PDX_MATH_EXACT>(a, s)
                                               : launch_expl_sell<PDX_MODEL_MMS2V, PDX_MATH_FAST>(a, s);
            case PDX_MODEL_MS2V:     return ex ? launch_expl_sell<PDX_MODEL_MS2V, PDX_MATH_EXACT>(a, s)
                                               : launch_expl_sell<PDX_MODEL_MS2V, PDX_MATH_FAST>(a, s);
        }
        set_error("pdx_fem_explicit_step_sell: unknown model");
        return 1;
    }
    // lanes per row from the mean row length (P1 triangles ~7, tetrahedra ~15 entries per row)
    // (registered once per matrix through pdx_csr_hint; 8 lanes when the caller gave no hint - no device read here,
    // so the call never blocks and is safe inside a stream capture)
    int G = csr_group_hint(rowptr);
    if (G == 0) G = 8;
    long long nb = ((long long)n * G + 255) / 256;
    if (nb > (long long)sm_count() * 32) nb = (long long)sm_count() * 32;
    switch (model) {
        case PDX_MODEL_NONE:     return launch_expl<PDX_MODEL_NONE, PDX_MATH_FAST>(G, a, (unsigned)nb, s);
        case PDX_MODEL_FENTON4V: return ex ? launch_expl<PDX_MODEL_FENTON4V, PDX_MATH_EXACT>(G, a, (unsigned)nb, s)
                                           : launch_expl<PDX_MODEL_FENTON4V, PDX_MATH_FAST>(G, a, (unsigned)nb, s);
        case PDX_MODEL_MMS2V:    return ex ? launch_expl<PDX_MODEL_MMS2V, PDX_MATH_EXACT>(G, a, (unsigned)nb, s)
                                           : launch_expl<PDX_MODEL_MMS2V, PDX_MATH_FAST>(G, a, (unsigned)nb, s);
        case PDX_MODEL_MS2V:     return ex ? launch_expl<PDX_MODEL_MS2V, PDX_MATH_EXACT>(G, a, (unsigned)nb, s)
                                           : launch_expl<PDX_MODEL_MS2V, PDX_MATH_FAST>(G, a, (unsigned)nb, s);
    }
    set_error("pdx_fem_explicit_step: unknown model");
    return 1;
}

int pdx_fem_explicit_step(int model, int math_mode, const float* params, int32_t n, const int32_t* rowptr,
                          const int32_t* col, const float* kval, const float* mlinv, const float* U_in, float* U_out,
                          float* const* states, const float* I0, float dt, void* stream) {
    return explicit_step_impl(false, model, math_mode, params, n, 0, rowptr, col, kval, mlinv, U_in, U_out, states, I0, dt,
                              0, nullptr, 0, nullptr, stream);
}

int pdx_fem_explicit_step_part(int model, int math_mode, const float* params, int32_t n_local, int32_t row_lo,
                               const int32_t* rowptr, const int32_t* col, const float* kval, const float* mlinv,
                               const float* U_in, float* U_out, float* const* states, const float* I0, float dt,
                               int32_t send_lo, float* mirror_lo, int32_t send_hi, float* mirror_hi, void* stream) {
    return explicit_step_impl(false, model, math_mode, params, n_local, row_lo, rowptr, col, kval, mlinv, U_in, U_out,
                              states, I0, dt, send_lo, mirror_lo, send_hi, mirror_hi, stream);
}

int pdx_fem_explicit_step_sell(int model, int math_mode, const float* params, int32_t n_local, int32_t row_lo,
                               const int32_t* slice_ptr, const int32_t* col, const float* kval, const float* mlinv,
                               const float* U_in, float* U_out, float* const* states, const float* I0, float dt,
                               int32_t send_lo, float* mirror_lo, int32_t send_hi, float* mirror_hi, void* stream) {
    return explicit_step_impl(true, model, math_mode, params, n_local, row_lo, slice_ptr, col, kval, mlinv, U_in, U_out,
                              states, I0, dt, send_lo, mirror_lo, send_hi, mirror_hi, stream);
}

int pdx_fem_explicit_step_sync(int sell, int model, int math_mode, const float* params, int32_t n_local, int32_t row_lo,
                               const int32_t* rowptr, const int32_t* col, const float* kval, const float* mlinv,
                               const float* U_in, float* U_out, float* const* states, const float* I0, float dt,
                               int32_t send_lo, float* mirror_lo, int32_t send_hi, float* mirror_hi, uint32_t* ctrl,
                               uint32_t* ctrl_lo, uint32_t* ctrl_hi, void* stream) {
    if (!ctrl) { set_error("pdx_fem_explicit_step_sync: null control block"); return 1; }
    if ((send_lo > 0 && !ctrl_lo) || (send_hi > 0 && !ctrl_hi)) {
        set_error("pdx_fem_explicit_step_sync: a neighbour that receives rows needs its control block");
        return 1;
    }
    return explicit_step_impl(sell != 0, model, math_mode, params, n_local, row_lo, rowptr, col, kval, mlinv, U_in, U_out,
                              states, I0, dt, send_lo, mirror_lo, send_hi, mirror_hi, stream, ctrl, ctrl_lo, ctrl_hi);
}

}  // extern "C"
