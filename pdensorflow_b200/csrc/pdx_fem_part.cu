// pdx_fem_part.cu - row-block partitioned Jacobi-PCG: the semi-implicit FEM step across GPUs
// (FEM: "per SpMV one ghost exchange, per CG iteration two scalar all-reduces").
//
// Same mathematics as pcg_kernel (pdx_fem.cu; reference linearsolvers/conjgrad.py:172-203,257-312,
// jacobi_precond.py:35-41), one PERSISTENT COOPERATIVE KERNEL PER RANK per solve.  The ranks talk only
// through NVLink peer memory, from inside the kernel:
//
//  * Every rank owns n rows and keeps WINDOWS [ghost_lo | n | ghost_hi] of the vectors the matrix
//    multiplies (x, p, rhs0); column indices are window-local.  The matrix must be banded enough for the
//    ghosts to be the neighbours' edge rows (RCM / structured order, fem_dist.plan_windows checks it).
//  * Ghost values are never "exchanged" as such.  x and p ghosts are RECOMPUTED locally with the same
//    arithmetic the owner uses (x_g = fma(alpha, p_g, x_g); p_g = fma(beta, p_g, z_g)), which needs only
//    z = Minv*r of the neighbours' edge rows.  Those are stored by the owner straight into this rank's
//    ghost slots while it updates r, and they become visible with the all-reduce that follows anyway.
//  * The two dot products of an iteration are all-reduced through peer memory: block 0 publishes the
//    rank total into slot[rank] of EVERY rank as one 8-byte (epoch, value) word with release semantics,
//    then waits for the `world` words in its own slots, adds them in rank order (same bits on every
//    rank, so every rank takes the same convergence decision) and hands the result to the other blocks.
//    The release/acquire pair on the slot also orders the z stores of the same phase.
//  * So an iteration costs 2 exchanges of 8 bytes per peer and nothing else; no NCCL call, no host
//    involvement, no extra launch.  A rank that never shows up makes block 0 time out (timeout_ms) and
//    the solve return flag bit 2 instead of hanging the GPU.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>
#include "pdx_internal.h"
#include "pdx_fem_dev.cuh"

namespace cg = cooperative_groups;

namespace pdx {
namespace {

constexpr int MAXW = PDX_PART_MAX_WORLD;

// Communication block at the start of every rank's symmetric workspace.
struct Comm {
    unsigned long long slot[2][2][MAXW];   // [epoch parity][which scalar][source rank] = epoch << 32 | float bits
    unsigned long long bcast[2];           // block 0 -> other blocks of this rank
    unsigned int abort_flag;
    unsigned int epoch;                    // all-reduces completed so far (identical on every rank)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// ---- L2 residency control (r02).  A CG iteration streams the matrix once (8 B per non-zero) and touches each of the
// vectors p, r, q, x, Minv two to four times; at 4e6 rows the five vectors are 82 MB - they FIT the 126 MB L2, but the
// 480 MB matrix stream evicts them between phases, so every vector access of the update phases went to DRAM
// (52 of 170 B per row).  Matrix loads therefore carry an evict-first policy and vector accesses evict-last;
// vectors beyond the budget (PartArgs::keep) keep the normal policy.
typedef unsigned long long l2pol_t;
__device__ __forceinline__ l2pol_t l2_policy(int kind) {   // 0 normal, 1 evict_last, 2 evict_first
    l2pol_t p;
    if (kind == 1)      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else                asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// vector accesses: L2-coherent (.cg, other SMs write these between grid barriers) + policy
__device__ __forceinline__ float ldv(const float* a, l2pol_t pol) {
    float v;
    asm volatile("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void stv(float* a, float v, l2pol_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
}
// matrix stream: read-only, never re-used inside an iteration
__device__ __forceinline__ float ldm(const float* a, l2pol_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ int ldm(const int32_t* a, l2pol_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}
struct Pol {
    l2pol_t mat, p, r, q, x, minv;
};

__device__ __forceinline__ unsigned long long pack(unsigned int epoch, float v) {
    return ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
}

struct PartArgs {
    int n, glo, ghi;                 // owned rows, ghost rows below / above
    const int32_t* rowptr;
    const int32_t* col;              // window-local
    const float* aval;               // A = M + dt K
    const float* mval;               // optional: b = M * rhs0 (same pattern)
    const float* minv;               // n, may be null
    const float* b;                  // n (used when mval == null)
    const float* rhs0;               // window (used when mval != null)
    float* x;                        // window
    float* p;                        // window
    float* zg;                       // window; only the ghost parts are used
    float *r, *q;                    // n
    float* part;                     // 2 buffers x 2 scalars x gridDim.x
    Comm* comm;
    Comm* peer[MAXW];
    // where this rank's first send_lo / last send_hi rows live in the neighbours' windows
    float* peer_lo_p;  float* peer_lo_zg;     // indexed by own row i            (i < send_lo)
    float* peer_hi_p;  float* peer_hi_zg;     // indexed by i - (n - send_hi)    (i >= n - send_hi)
    int send_lo, send_hi;
    int rank, world;
    int maxiter, check_every;
    double toll_sq;
    unsigned long long timeout_ns;
    int32_t* info;
    unsigned long long* prof;        // optional: block 0 stamps %globaltimer at every phase boundary (diagnostics)
    int nprof;
    int stage_slot_bytes;            // > 0: SELL slices of q = A p are staged through shared memory by the copy engine
    int keep;                        // L2 policy bits: 1 p, 2 r, 4 q, 8 x, 16 minv evict-last; 32 matrix evict-first
};

// All-reduce of NV (1 or 2) scalars over the ranks.  Returns false when a peer timed out.
// Every thread of the grid calls it; loc[] are the thread's partial sums.
// `pushed`: this thread has stored into a neighbour's window since the last all-reduce.
template <int NV>
__device__ __forceinline__ bool all_reduce(const PartArgs& a, cg::grid_group& grid, unsigned int& epoch, float* sh,
                                           float loc0, float loc1, float* out0, float* out1, bool pushed) {
    const int nb = gridDim.x;
    float v0 = block_sum(loc0, sh);
    float v1 = NV > 1 ? block_sum(loc1, sh) : 0.0f;
    // the block partials alternate between two buffers: a block may already publish the partial of the NEXT
    // all-reduce while a slower block is still adding up this one's (they are two grid barriers apart per buffer)
    float* part = a.part + ((epoch + 1) & 1) * 2 * nb;
    if (threadIdx.x == 0) {
        __stcg(part + blockIdx.x, v0);
        if (NV > 1) __stcg(part + nb + blockIdx.x, v1);
    }
    if (pushed) __threadfence_system();   // this thread's peer stores (z of the edge rows) before the publication
    grid.sync();
    ++epoch;
    if (a.world == 1) {
        // one rank: nothing to exchange - every block adds the block partials itself, in the same order, so all
        // blocks hold the same bits; no publication by block 0, no second wait
        *out0 = grid_total(part, sh);
        if (NV > 1) *out1 = grid_total(part + nb, sh);
        return true;
    }
    const int par = epoch & 1;
    Comm* me = a.comm;
    if (blockIdx.x == 0) {
        const float t0 = grid_total(part, sh);
        const float t1 = NV > 1 ? grid_total(part + nb, sh) : 0.0f;
        __shared__ float got[2][MAXW];
        __shared__ int timed_out;
        if (threadIdx.x == 0) timed_out = 0;
        __syncthreads();
        if ((int)threadIdx.x < a.world) {
            const int peer = threadIdx.x;
            __threadfence_system();
            st_release_sys(&a.peer[peer]->slot[par][0][a.rank], pack(epoch, t0));
            if (NV > 1) st_release_sys(&a.peer[peer]->slot[par][1][a.rank], pack(epoch, t1));
            const unsigned long long t_start = global_ns();
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                unsigned long long w;
                unsigned int spins = 0;
                while ((unsigned int)((w = ld_acquire_sys(&me->slot[par][v][peer])) >> 32) != epoch) {
                    if ((++spins & 1023u) == 0 && global_ns() - t_start > a.timeout_ns) { timed_out = 1; break; }
                }
                got[v][peer] = __uint_as_float((unsigned int)w);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float s0 = 0.0f, s1 = 0.0f;
            for (int rk = 0; rk < a.world; ++rk) { s0 += got[0][rk]; if (NV > 1) s1 += got[1][rk]; }
            if (timed_out) { me->abort_flag = 1u; __threadfence(); }
            if (NV > 1) st_release_gpu(&me->bcast[1], pack(epoch, s1));
            st_release_gpu(&me->bcast[0], pack(epoch, s0));
        }
    }
    __shared__ unsigned long long res[2];
    __shared__ unsigned int aborted;
    if (threadIdx.x == 0) {
        unsigned long long w;
        while ((unsigned int)((w = ld_acquire_gpu(&me->bcast[0])) >> 32) != epoch) { }
        res[0] = w;
        if (NV > 1) res[1] = ld_acquire_gpu(&me->bcast[1]);   // written before bcast[0]
        aborted = *((volatile unsigned int*)&me->abort_flag);
    }
    __syncthreads();
    *out0 = __uint_as_float((unsigned int)res[0]);
    if (NV > 1) *out1 = __uint_as_float((unsigned int)res[1]);
    const bool ok = aborted == 0u;
    __syncthreads();                 // res[] is reused by the next call
    return ok;
}

// Row products of the kernel's SpMV phases: f(row, A_row . xa, M_row . xm) is called once per owned row
// by the lane that owns it (mv == nullptr: the second product is skipped and 0 is passed).
//  G > 0: CSR, G lanes per row.   G == 0: SELL-32 (a.rowptr = slice_ptr): a warp per 32-row slice, one
//  thread per row, column-major inside the slice so that every col/val access is a coalesced 128-byte load.
template <int G, class F>
__device__ __forceinline__ void rows_dot(const PartArgs& a, const Pol& pol, l2pol_t polx, const float* xa, const float* mv,
                                         const float* xm, F&& f) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    if constexpr (G == 0) {
        const int lane = threadIdx.x & 31;
        const int nsl = (a.n + 31) >> 5;
        for (int sl = tid >> 5; sl < nsl; sl += nthreads >> 5) {
            const int beg = a.rowptr[sl], end = a.rowptr[sl + 1];
            float acc = 0.0f, accm = 0.0f;
            int k = beg + lane;
            if (mv) {
                for (; k + 32 < end; k += 64) {
                    const int c0 = ldm(a.col + k, pol.mat), c1 = ldm(a.col + k + 32, pol.mat);
                    const float v0 = ldm(a.aval + k, pol.mat), v1 = ldm(a.aval + k + 32, pol.mat);
                    const float m0 = ldm(mv + k, pol.mat), m1 = ldm(mv + k + 32, pol.mat);
                    const float x0 = ldv(xa + c0, polx), x1 = ldv(xa + c1, polx);
                    const float y0 = __ldcg(xm + c0), y1 = __ldcg(xm + c1);
                    acc = fmaf(v0, x0, acc);   accm = fmaf(m0, y0, accm);
                    acc = fmaf(v1, x1, acc);   accm = fmaf(m1, y1, accm);
                }
                for (; k < end; k += 32) {
                    const int c = ldm(a.col + k, pol.mat);
                    acc = fmaf(ldm(a.aval + k, pol.mat), ldv(xa + c, polx), acc);
                    accm = fmaf(ldm(mv + k, pol.mat), __ldcg(xm + c), accm);
                }
            } else {
                // four slice columns at a time: the matrix words first (plain read-only loads the compiler may hoist),
                // then the four gathers they feed
                for (; k + 96 < end; k += 128) {
                    const int c0 = ldm(a.col + k, pol.mat), c1 = ldm(a.col + k + 32, pol.mat);
                    const int c2 = ldm(a.col + k + 64, pol.mat), c3 = ldm(a.col + k + 96, pol.mat);
                    const float v0 = ldm(a.aval + k, pol.mat), v1 = ldm(a.aval + k + 32, pol.mat);
                    const float v2 = ldm(a.aval + k + 64, pol.mat), v3 = ldm(a.aval + k + 96, pol.mat);
                    const float x0 = ldv(xa + c0, polx), x1 = ldv(xa + c1, polx);
                    const float x2 = ldv(xa + c2, polx), x3 = ldv(xa + c3, polx);
                    acc = fmaf(v0, x0, acc);
                    acc = fmaf(v1, x1, acc);
                    acc = fmaf(v2, x2, acc);
                    acc = fmaf(v3, x3, acc);
                }
                for (; k < end; k += 32) acc = fmaf(ldm(a.aval + k, pol.mat), ldv(xa + ldm(a.col + k, pol.mat), polx), acc);
            }
            const int row = sl * 32 + lane;
            if (row < a.n) f(row, acc, accm);
        }
    } else {
        const int gid = tid / G, sub = threadIdx.x % G, ngroups = nthreads / G;
        const int nround = (a.n + ngroups - 1) / ngroups * ngroups;     // keep groups converged for the shuffles
        for (int row = gid; row < nround; row += ngroups) {
            const bool valid = row < a.n;
            const int beg = valid ? a.rowptr[row] : 0, end = valid ? a.rowptr[row + 1] : 0;
            const float acc = row_dot<(G > 0 ? G : 1)>(a.col, a.aval, xa, beg, end, sub, true);
            float accm = 0.0f;
            if (mv) accm = row_dot<(G > 0 ? G : 1)>(a.col, mv, xm, beg, end, sub, true);   // uniform branch
            if (valid && sub == 0) f(row, acc, accm);
        }
    }
}

// ---- SELL product with the matrix stream on the copy engine (r02).  In rows_dot a warp reads slice_ptr, then the
// slice's col / val words, then gathers: three dependent latencies per slice, and only the middle one moves the 480 MB
// that bound the phase (ncu: 0.25 eligible warps per scheduler).  Here lane 0 of every warp keeps STAGE_NB slices
// ahead in flight as cp.async.bulk copies (col and val words of a slice are contiguous: two copies, one mbarrier)
// into a per-warp shared-memory ring; the lanes then read their columns from shared memory and have a whole row of
// gathers in flight at once.  Products are accumulated in the same order as in rows_dot: same bits.
struct StageState {
    unsigned int uses;       // slots consumed so far by this warp (phase parity of the ring's mbarriers)
};

// the walker of pdx_fem_dev.cuh on this kernel's matrix, vector loads with their L2 policy
template <class F>
__device__ __forceinline__ void rows_dot_staged(const PartArgs& a, const Pol& pol, l2pol_t polx, const float* vals,
                                                const float* xa, StageState& stg, F&& f) {
    extern __shared__ __align__(128) unsigned char dsm[];
    sell_rows_staged(dsm, a.stage_slot_bytes, a.n, a.rowptr, a.col, vals, pol.mat, stg.uses,
                     [&](int c) { return ldv(xa + c, polx); }, [&](int row, float acc) { f(row, acc, 0.0f); });
}

// STAGED (G == 0 only): blocks of at most 512 threads with the shared-memory ring, 64 registers per thread
template <int G, bool STAGED = false>
__global__ void __launch_bounds__(STAGED ? 512 : 1024, 2) pcg_part_kernel(const __grid_constant__ PartArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float sh[33];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int glo = a.glo;
    const int hi_first = a.n - a.send_hi;          // first row the upper neighbour holds as a ghost
    unsigned int epoch = *((volatile unsigned int*)&a.comm->epoch);
    constexpr int UNR = STAGED ? 4 : 1;      // rows per thread and trip of the vector passes (32-register kernels: 1)
    StageState stg{0u};
    if constexpr (STAGED) {
        extern __shared__ __align__(128) unsigned char dsm[];
        sell_stage_init(dsm, a.stage_slot_bytes);
    }
    Pol pol;
    pol.mat  = l2_policy(a.keep & 32 ? 2 : 0);
    pol.p    = l2_policy(a.keep & 1 ? 1 : 0);
    pol.r    = l2_policy(a.keep & 2 ? 1 : 0);
    pol.q    = l2_policy(a.keep & 4 ? 1 : 0);
    pol.x    = l2_policy(a.keep & 8 ? 1 : 0);
    pol.minv = l2_policy(a.keep & 16 ? 1 : 0);

    // b = M rhs0 (optional) ; r = b - A x ; z = Minv r ; p = z ; rz = r.z      (conjgrad.py:191-203)
    float loc_rz = 0.0f, loc_rr = 0.0f;
    bool pushed = false;
    auto take_r = [&](int row, float ax, float bm) {
        const float bv = a.mval ? bm : a.b[row];
        const float r = __fsub_rn(bv, ax);
        const float z = a.minv ? __fmul_rn(ldv(a.minv + row, pol.minv), r) : r;
        stv(a.r + row, r, pol.r);
        stv(a.p + glo + row, z, pol.p);
        if (row < a.send_lo) { a.peer_lo_p[row] = z; pushed = true; }          // the neighbours' ghost p = z
        if (row >= hi_first) { a.peer_hi_p[row - hi_first] = z; pushed = true; }
        loc_rz = fmaf(r, z, loc_rz);
        loc_rr = fmaf(r, r, loc_rr);
    };
    if constexpr (STAGED) {
        // two staged products (b = M rhs0 parked in q, then A x) instead of one pass over three arrays: the ring slot
        // stays two arrays wide; a row is handled by the same lane in both, so no synchronisation in between
        if (a.mval)
            rows_dot_staged(a, pol, pol.x, a.mval, a.rhs0, stg, [&](int row, float bm, float) { a.q[row] = bm; });
        rows_dot_staged(a, pol, pol.x, a.aval, a.x, stg, [&](int row, float ax, float) { take_r(row, ax, a.mval ? a.q[row] : 0.0f); });
    } else {
        rows_dot<G>(a, pol, pol.x, a.x, a.mval, a.rhs0, take_r);
    }
    int nstamp = 0;
    auto stamp = [&]() {
        if (a.prof && tid == 0 && nstamp < a.nprof) a.prof[nstamp] = global_ns();
        ++nstamp;
    };
    stamp();
    float rz, res;
    bool ok = all_reduce<2>(a, grid, epoch, sh, loc_rz, loc_rr, &rz, &res, pushed);
    stamp();

    int niters = 0, flags = 0;
    for (int k = 1; ok && k <= a.maxiter; ++k) {
        niters = k;
        // q = A p ; pq = p.q
        float loc_pq = 0.0f;
        auto take_q = [&](int row, float ap, float) {
            stv(a.q + row, ap, pol.q);
            loc_pq = fmaf(ldv(a.p + glo + row, pol.p), ap, loc_pq);
        };
        if constexpr (STAGED) rows_dot_staged(a, pol, pol.p, a.aval, a.p, stg, take_q);
        else rows_dot<G>(a, pol, pol.p, a.p, nullptr, nullptr, take_q);
        float pq, unused;
        stamp();
        ok = all_reduce<1>(a, grid, epoch, sh, loc_pq, 0.0f, &pq, &unused, false);
        stamp();
        if (!ok) break;
        const float alpha = rz / pq;
        // x += alpha p ; r -= alpha q ; res = r.r ; z = Minv r ; rz' = r.z        (conjgrad.py:172-183)
        loc_rz = 0.0f; loc_rr = 0.0f;
        pushed = false;
        // (STAGED: four rows per thread and trip, all twenty loads first - the pass is latency bound otherwise, the staged form
        //  runs it with a third of the threads; rows are taken in the same order as a plain grid-stride loop)
        {
            auto row_update = [&](int i, float pi, float xi, float qi, float ri, float mi) {
                stv(a.x + glo + i, fmaf(alpha, pi, xi), pol.x);
                const float r = fmaf(-alpha, qi, ri);
                stv(a.r + i, r, pol.r);
                const float z = a.minv ? __fmul_rn(mi, r) : r;
                if (i < a.send_lo) { a.peer_lo_zg[i] = z; pushed = true; }            // edge rows: z into the neighbours' ghost slots
                if (i >= hi_first) { a.peer_hi_zg[i - hi_first] = z; pushed = true; }
                loc_rr = fmaf(r, r, loc_rr);
                loc_rz = fmaf(r, z, loc_rz);
            };
            int i = tid;
            for (; UNR > 1 && i + (UNR - 1) * nthreads < a.n; i += UNR * nthreads) {
                float pi[UNR], xi[UNR], qi[UNR], ri[UNR], mi[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int j = i + u * nthreads;
                    pi[u] = ldv(a.p + glo + j, pol.p);
                    xi[u] = ldv(a.x + glo + j, pol.x);
                    qi[u] = ldv(a.q + j, pol.q);
                    ri[u] = ldv(a.r + j, pol.r);
                    mi[u] = a.minv ? ldv(a.minv + j, pol.minv) : 1.0f;
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) row_update(i + u * nthreads, pi[u], xi[u], qi[u], ri[u], mi[u]);
            }
            for (; i < a.n; i += nthreads)
                row_update(i, ldv(a.p + glo + i, pol.p), ldv(a.x + glo + i, pol.x), ldv(a.q + i, pol.q), ldv(a.r + i, pol.r),
                           a.minv ? ldv(a.minv + i, pol.minv) : 1.0f);
        }
        // ghost x, recomputed exactly as its owner computes it
        for (int g = tid; g < a.glo + a.ghi; g += nthreads) {
            const int w = g < a.glo ? g : a.n + g;
            __stcg(a.x + w, fmaf(alpha, __ldcg(a.p + w), __ldcg(a.x + w)));
        }
        float rznew;
        stamp();
        ok = all_reduce<2>(a, grid, epoch, sh, loc_rz, loc_rr, &rznew, &res, pushed);
        stamp();
        if (!ok) break;
        // batched convergence test + NaN/Inf guard (conjgrad.py:293-300); p is not needed after the last iteration
        if (k % a.check_every == 0) {
            const bool finite = isfinite(res);
            if (!finite || (double)res < a.toll_sq) {
                if (!finite) flags |= 2;
                break;
            }
        }
        if (k == a.maxiter) break;
        const float beta = rznew / rz;
        // p = z + beta p, owned rows and ghosts alike
        {
            int i = tid;
            for (; UNR > 1 && i + (UNR - 1) * nthreads < a.n; i += UNR * nthreads) {
                float ri[UNR], pi[UNR], mi[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int j = i + u * nthreads;
                    ri[u] = ldv(a.r + j, pol.r);
                    pi[u] = ldv(a.p + glo + j, pol.p);
                    mi[u] = a.minv ? ldv(a.minv + j, pol.minv) : 1.0f;
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const float z = a.minv ? __fmul_rn(mi[u], ri[u]) : ri[u];
                    stv(a.p + glo + i + u * nthreads, fmaf(beta, pi[u], z), pol.p);
                }
            }
            for (; i < a.n; i += nthreads) {
                const float r = ldv(a.r + i, pol.r);
                const float pi = ldv(a.p + glo + i, pol.p);
                const float z = a.minv ? __fmul_rn(ldv(a.minv + i, pol.minv), r) : r;
                stv(a.p + glo + i, fmaf(beta, pi, z), pol.p);
            }
        }
        for (int g = tid; g < a.glo + a.ghi; g += nthreads) {
            const int w = g < a.glo ? g : a.n + g;
            __stcg(a.p + w, fmaf(beta, __ldcg(a.p + w), __ldcg(a.zg + w)));
        }
        rz = rznew;
        stamp();
        grid.sync();
        stamp();
    }
    if (!ok) flags |= 4;
    if (niters >= a.maxiter) flags |= 1;
    if (tid == 0) {
        a.comm->epoch = epoch;
        if (a.info) {
            a.info[0] = niters;
            a.info[1] = __float_as_int(res);
            a.info[2] = flags;
            a.info[3] = 0;
        }
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr size_t COMM_BYTES = 4096;
static_assert(sizeof(Comm) <= COMM_BYTES, "Comm must fit its slot of the workspace");

}  // namespace
}  // namespace pdx

using namespace pdx;

struct pdx_pcg_part {
    pdx_pcg_part_desc d;
    int window;
    int64_t nnz;
    int G, nblocks, nthreads;
    int stage_slot_bytes;   // SELL slices staged by the copy engine (0 = off)
    size_t smem_bytes;
    char* ws;               // own workspace
    float *p, *zg;          // windows inside the workspace (written by the neighbours)
    float* work;            // r, q
    float* part;
    PartArgs base;
};

extern "C" {

int64_t pdx_pcg_part_workspace_bytes(int32_t max_window) {
    if (max_window <= 0) return 0;
    return (int64_t)(COMM_BYTES + 2 * align_up(sizeof(float) * (size_t)max_window, 256));
}

int pdx_pcg_part_create(const pdx_pcg_part_desc* d, void* const* peer_ws, int32_t max_window, pdx_pcg_part** out) {
    if (!d || !peer_ws || !out) { set_error("pdx_pcg_part_create: null argument"); return 1; }
    if (d->n_local <= 0 || d->ghost_lo < 0 || d->ghost_hi < 0 || !d->rowptr || !d->col || !d->aval) {
        set_error("pdx_pcg_part_create: bad matrix description");
        return 1;
    }
    if (d->world < 1 || d->world > MAXW || d->rank < 0 || d->rank >= d->world) {
        set_error("pdx_pcg_part_create: world must be 1.." + std::to_string(MAXW) + " and 0 <= rank < world");
        return 1;
    }
    const int window = d->ghost_lo + d->n_local + d->ghost_hi;
    if (window > max_window) { set_error("pdx_pcg_part_create: window exceeds max_window"); return 1; }
    if (d->send_lo < 0 || d->send_hi < 0 || d->send_lo > d->n_local || d->send_hi > d->n_local ||
        (d->rank == 0 && (d->send_lo || d->ghost_lo)) || (d->rank == d->world - 1 && (d->send_hi || d->ghost_hi))) {
        set_error("pdx_pcg_part_create: bad ghost / send counts");
        return 1;
    }
    if ((d->send_lo > 0 && d->lower_window_index < 0) || (d->send_hi > 0 && d->upper_window_index < 0)) {
        set_error("pdx_pcg_part_create: neighbour window indices missing");
        return 1;
    }
    for (int r = 0; r < d->world; ++r)
        if (!peer_ws[r]) { set_error("pdx_pcg_part_create: null peer workspace"); return 1; }
    pdx_pcg_part* s = new (std::nothrow) pdx_pcg_part();
    if (!s) { set_error("out of memory"); return 1; }
    s->d = *d;
    s->window = window;
    int32_t last = 0;
    const int nptr = d->sell ? (d->n_local + 31) / 32 : d->n_local;
    if (check_cuda(cudaMemcpy(&last, d->rowptr + nptr, sizeof(int32_t), cudaMemcpyDeviceToHost), "read nnz")) {
        delete s; return 1;
    }
    s->nnz = last;
    s->G = d->sell ? 0 : pick_group(d->n_local, s->nnz);
    int dev = 0, sms = 0, per_sm = 0;
    // threads per block: big systems run 1024 (a quarter of the arrivals at every grid barrier and of the block partials every
    // block adds up; same 32-register budget and occupancy as 8 x 256), small ones 256.  PDX_PCG_PART_THREADS overrides.
    int bt = (long long)d->n_local * (s->G > 0 ? s->G : 1) >= 1000000LL ? 1024 : 256;
    if (const char* e = getenv("PDX_PCG_PART_THREADS")) { const int v = atoi(e); if (v == 256 || v == 512 || v == 1024) bt = v; }
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e;
    s->stage_slot_bytes = 0;
    s->smem_bytes = 0;
    {   // SELL systems large enough to be bandwidth bound: the matrix stream of q = A p goes through the copy engine
        // (rows_dot_staged).  The ring slot holds the widest slice; the block size is the one that keeps most warps
        // resident next to STAGE_NB slots per warp.  PDX_PCG_STAGE=0 turns it off (A/B runs).
        const char* ev = getenv("PDX_PCG_STAGE");
        // PDX_PCG_STAGE=1 forces it for any SELL system (tests)
        if (s->G == 0 && (bt == 1024 || (ev && ev[0] == '1')) && !(ev && ev[0] == '0')) {
            const int nsl = (d->n_local + 31) / 32;
            std::vector<int32_t> sp((size_t)nsl + 1);
            if (check_cuda(cudaMemcpy(sp.data(), d->rowptr, sizeof(int32_t) * ((size_t)nsl + 1), cudaMemcpyDeviceToHost), "read slice_ptr")) {
                delete s; return 1;
            }
            int wmax = 0;
            for (int i = 0; i < nsl; ++i) wmax = std::max(wmax, (sp[i + 1] - sp[i]) / 32);
            const int slot = wmax * 256;                       // col + val words of the widest slice
            int smem_max = 0;
            cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            int best_bt = 0, best_warps = 0;
            if (getenv("PDX_DEBUG")) fprintf(stderr, "[pdx] stage wmax %d slot %d smem_max %d col %p aval %p\n", wmax, slot, smem_max, (void*)d->col, (void*)d->aval);
            size_t best_smem = 0;
            if (slot > 0 && ((uintptr_t)d->col % 16) == 0 && ((uintptr_t)d->aval % 16) == 0) {
                // the occupancy query counts against the kernel's CURRENT dynamic shared-memory limit: raise it first
                cudaFuncAttributes fa{};
                int max_dyn = 0;
                if (cudaFuncGetAttributes(&fa, pcg_part_kernel<0, true>) == cudaSuccess) max_dyn = smem_max - (int)fa.sharedSizeBytes;
                if (max_dyn > 0 && cudaFuncSetAttribute(pcg_part_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) == cudaSuccess) {
                    for (int cand : {512, 384, 256}) {
                        const size_t need = sell_stage_smem_bytes(cand, slot);
                        if (need > (size_t)max_dyn) continue;
                        int blk = 0;
                        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blk, pcg_part_kernel<0, true>, cand, need) != cudaSuccess) continue;
                        const int warps = blk * cand / 32;
                        if (getenv("PDX_DEBUG")) fprintf(stderr, "[pdx] stage cand %d need %zu blk %d wmax %d\n", cand, need, blk, wmax);
                        if (warps > best_warps) { best_warps = warps; best_bt = cand; best_smem = need; }
                    }
                }
                cudaGetLastError();
            }
            if (best_warps >= 16) {
                bt = best_bt;
                s->stage_slot_bytes = slot;
                s->smem_bytes = best_smem;
            }
        }
    }
    s->nthreads = bt;
    switch (s->G) {
        case 0:  e = s->stage_slot_bytes > 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_part_kernel<0, true>, bt, s->smem_bytes)
                                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_part_kernel<0>, bt, 0); break;
        case 4:  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_part_kernel<4>, bt, 0); break;
        case 8:  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_part_kernel<8>, bt, 0); break;
        case 16: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_part_kernel<16>, bt, 0); break;
        default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_part_kernel<32>, bt, 0); break;
    }
    if (check_cuda(e, "occupancy query") || per_sm < 1) {
        if (per_sm < 1) set_error("pcg_part kernel cannot be resident");
        delete s; return 1;
    }
    const int lanes = s->G > 0 ? s->G : 1;
    long long want = ((long long)d->n_local * lanes + bt - 1) / bt;
    // big systems are bandwidth bound: fill the SMs; small ones are latency bound: one block per SM keeps the
    // grid barrier cheap
    long long cap = (long long)sms * per_sm;
    long long nb = want < cap ? want : cap;
    if (nb > sms && (long long)d->n_local * lanes <= (long long)sms * 256 * 8) nb = sms;   // (bt == 256 here)
    if (d->max_blocks > 0 && nb > d->max_blocks) nb = d->max_blocks;
    if (nb < 1) nb = 1;
    s->nblocks = (int)nb;
    const size_t vec = align_up(sizeof(float) * (size_t)max_window, 256);
    s->ws = (char*)peer_ws[d->rank];
    s->p    = (float*)(s->ws + COMM_BYTES);
    s->zg   = (float*)(s->ws + COMM_BYTES + vec);
    if (check_cuda(cudaMalloc(&s->work, sizeof(float) * 2 * (size_t)d->n_local), "cudaMalloc(pcg_part work)")) { delete s; return 1; }
    if (check_cuda(cudaMalloc(&s->part, sizeof(float) * 4 * (size_t)s->nblocks), "cudaMalloc(pcg_part partials)")) {
        cudaFree(s->work); delete s; return 1;
    }
    if (check_cuda(cudaMemset(s->ws, 0, COMM_BYTES + 2 * vec), "clear comm block") ||     // comm, p, zg
        check_cuda(cudaDeviceSynchronize(), "clear comm block (sync)")) {
        cudaFree(s->work); cudaFree(s->part); delete s; return 1;
    }
    PartArgs& a = s->base;
    a = PartArgs{};
    a.n = d->n_local; a.glo = d->ghost_lo; a.ghi = d->ghost_hi;
    a.rowptr = d->rowptr; a.col = d->col; a.aval = d->aval; a.mval = d->mval; a.minv = d->minv;
    a.p = s->p; a.zg = s->zg;
    a.r = s->work; a.q = s->work + d->n_local;
    a.part = s->part;
    a.comm = (Comm*)s->ws;
    for (int r = 0; r < d->world; ++r) a.peer[r] = (Comm*)peer_ws[r];
    a.send_lo = d->send_lo; a.send_hi = d->send_hi;
    if (d->send_lo > 0) {
        char* w = (char*)peer_ws[d->rank - 1];
        a.peer_lo_p  = (float*)(w + COMM_BYTES) + d->lower_window_index;
        a.peer_lo_zg = (float*)(w + COMM_BYTES + vec) + d->lower_window_index;
    }
    if (d->send_hi > 0) {
        char* w = (char*)peer_ws[d->rank + 1];
        a.peer_hi_p  = (float*)(w + COMM_BYTES) + d->upper_window_index;
        a.peer_hi_zg = (float*)(w + COMM_BYTES + vec) + d->upper_window_index;
    }
    a.stage_slot_bytes = s->stage_slot_bytes;
    a.rank = d->rank; a.world = d->world;
    a.timeout_ns = (unsigned long long)(d->timeout_ms > 0 ? d->timeout_ms : 10000) * 1000000ull;
    // L2 residency: vectors that fit ~3/4 of the L2 (in order of re-use: p, r, q, x, Minv) are kept with evict-last,
    // the matrix streams through with evict-first.  PDX_PCG_L2KEEP=<bits> overrides (0 = plain loads, for A/B runs).
    {
        int l2bytes = 0;
        cudaDeviceGetAttribute(&l2bytes, cudaDevAttrL2CacheSize, dev);
        const double budget = 0.75 * (double)l2bytes, vec_bytes = 4.0 * (double)window;
        // evict-first only for a matrix that cannot stay in the L2 anyway: a 5e5-row block (60 MB of col / val words, what
        // a rank of a strong-scaled solve holds) is L2 resident across iterations and must not be pushed out
        int keep = 8.0 * (double)s->nnz > (double)l2bytes ? 32 : 0, nfit = vec_bytes > 0 ? (int)(budget / vec_bytes) : 5;
        static const int order[5] = {1, 2, 4, 8, 16};
        for (int i = 0; i < 5 && i < nfit; ++i) keep |= order[i];
        if (const char* e = getenv("PDX_PCG_L2KEEP")) keep = atoi(e);
        a.keep = keep;
    }
    *out = s;
    return 0;
}

void pdx_pcg_part_destroy(pdx_pcg_part* s) {
    if (!s) return;
    cudaFree(s->work);
    cudaFree(s->part);
    delete s;
}

int32_t pdx_pcg_part_blocks(const pdx_pcg_part* s) { return s ? s->nblocks : 0; }
int32_t pdx_pcg_part_threads(const pdx_pcg_part* s) { return s ? s->nthreads : 0; }
int32_t pdx_pcg_part_stage_bytes(const pdx_pcg_part* s) { return s ? s->stage_slot_bytes : 0; }

int pdx_pcg_part_set_profile(pdx_pcg_part* s, uint64_t* device_stamps, int32_t nstamps) {
    if (!s) { set_error("pdx_pcg_part_set_profile: null handle"); return 1; }
    s->base.prof = (unsigned long long*)device_stamps;
    s->base.nprof = device_stamps ? nstamps : 0;
    return 0;
}

int pdx_pcg_part_solve(pdx_pcg_part* s, const float* b, const float* rhs0, float* x, int maxiter, double toll,
                       int check_every, int32_t* info, void* stream) {
    if (!s || !x || maxiter < 0) { set_error("pdx_pcg_part_solve: bad arguments"); return 1; }
    if (!b && !(s->d.mval && rhs0)) {
        set_error("pdx_pcg_part_solve: without b the handle needs a mass matrix and the call an rhs0 window");
        return 1;
    }
    PartArgs a = s->base;
    a.b = b;
    a.x = x;
    a.rhs0 = rhs0;
    if (b) a.mval = nullptr;
    a.maxiter = maxiter;
    a.check_every = check_every > 0 ? check_every : 1;
    a.toll_sq = toll * toll;
    a.info = info;
    void* args[] = {&a};
    const void* fn;
    switch (s->G) {
        case 0:  fn = s->stage_slot_bytes > 0 ? (const void*)pcg_part_kernel<0, true> : (const void*)pcg_part_kernel<0>; break;
        case 4:  fn = (const void*)pcg_part_kernel<4>; break;
        case 8:  fn = (const void*)pcg_part_kernel<8>; break;
        case 16: fn = (const void*)pcg_part_kernel<16>; break;
        default: fn = (const void*)pcg_part_kernel<32>; break;
    }
    count_launch();
    return check_cuda(cudaLaunchCooperativeKernel(fn, dim3(s->nblocks), dim3(s->nthreads), args, s->smem_bytes, (cudaStream_t)stream),
                      "pcg_part_kernel cooperative launch");
}

}  // extern "C"
