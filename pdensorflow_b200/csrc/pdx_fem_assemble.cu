// pdx_fem_assemble.cu - P1 mass / stiffness assembly on the device.
//
// Reference: gpuSolve/matrices/globalMatrices.py:82-287 (contravariant basis, local matrices, COO
// accumulation), localMass.py, localStiffness.py.  The reference assembles on the host in NumPy
// (7.9 s at 1 M nodes by its own numbers; minutes at the 2e7 nodes of config 5).  Here one thread takes one
// element: it reads its 2/3/4 vertices, forms the contravariant basis and the element measure in fp64 with the
// closed-form 3x3 inverse, evaluates  M_ab = meas * T_ab  and  K_ab = meas * G_a . (Sigma G_b)  and adds the
// nv^2 entries into the CSR value arrays at positions found by binary search in the (sorted) row.
// Accumulation is fp64 atomicAdd, rounded to fp32 once at the end: the sum of a row entry's <= ~30
// contributions is then order independent to ~1e-16, i.e. reproducible after the rounding except for rare
// ties, and agrees with the reference's fp32 running sums to a few fp32 ulp (tests: 2e-6 of the largest entry).
#include "pdx_internal.h"

namespace pdx {
namespace {

struct AsmArgs {
    int64_t ne;
    const int32_t* elems;     // [ne][stride]: nv vertex ids (+ optional region id)
    int stride;
    const double* pts;        // [npt][3]
    const int32_t* rowid;     // vertex id -> matrix row (renumbering), or null
    const double* sigma;      // [ne][6] xx xy xz yy yz zz, or null (identity)
    const int32_t* rowptr;
    const int32_t* col;
    double* macc;             // may be null
    double* kacc;             // may be null
    int32_t* status;          // set to 1 when an entry is not in the pattern
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

template <int NV>
__global__ void __launch_bounds__(128) fem_assemble_kernel(const __grid_constant__ AsmArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.ne; e += stride) {
        int nd[NV];
        double P[NV][3];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            nd[v] = a.elems[e * a.stride + v];
#pragma unroll
            for (int d = 0; d < 3; ++d) P[v][d] = a.pts[(size_t)nd[v] * 3 + d];
        }
        double r0[3], r1[3] = {0, 0, 0}, r2[3] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < 3; ++d) r0[d] = P[1][d] - P[0][d];
        double meas;
        double g[3][3];          // contravariant vectors g[i] (i < NV-1)
        if (NV == 2) {
            const double l2 = dot3(r0, r0);
            meas = sqrt(l2);
#pragma unroll
            for (int d = 0; d < 3; ++d) g[0][d] = r0[d] / l2;
        } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) r1[d] = P[2][d] - P[0][d];
            if (NV == 3) {
                double nrm[3];
                cross3(r0, r1, nrm);
                const double a2 = sqrt(dot3(nrm, nrm));
#pragma unroll
                for (int d = 0; d < 3; ++d) r2[d] = nrm[d] / a2;
                meas = 0.5 * a2;
            } else {
#pragma unroll
                for (int d = 0; d < 3; ++d) r2[d] = P[NV - 1][d] - P[0][d];
                double c01[3];
                cross3(r0, r1, c01);
                meas = fabs(dot3(c01, r2)) / 6.0;
            }
            // inverse of the matrix with ROWS r0, r1, r2: its columns are (r1 x r2, r2 x r0, r0 x r1) / det
            double c12[3], c20[3], c01[3];
            cross3(r1, r2, c12);
            cross3(r2, r0, c20);
            cross3(r0, r1, c01);
            const double idet = 1.0 / dot3(r0, c12);
#pragma unroll
            for (int d = 0; d < 3; ++d) { g[0][d] = c12[d] * idet; g[1][d] = c20[d] * idet; g[2][d] = c01[d] * idet; }
        }
        // gradients of the hat functions: G_0 = -(g_1 + ...), G_a = g_a
        double G[NV][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double s = 0.0;
#pragma unroll
            for (int v = 1; v < NV; ++v) { G[v][d] = g[v - 1][d]; s += g[v - 1][d]; }
            G[0][d] = -s;
        }
        double S[6] = {1, 0, 0, 1, 0, 1};
        if (a.sigma) {
#pragma unroll
            for (int q = 0; q < 6; ++q) S[q] = a.sigma[e * 6 + q];
        }
        double SG[NV][3];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            SG[v][0] = S[0] * G[v][0] + S[1] * G[v][1] + S[2] * G[v][2];
            SG[v][1] = S[1] * G[v][0] + S[3] * G[v][1] + S[4] * G[v][2];
            SG[v][2] = S[2] * G[v][0] + S[4] * G[v][1] + S[5] * G[v][2];
        }
        // localMass.py: Edges meas*(1/3,1/6), Trias 2*meas*(1/12,1/24), Tetras 6*meas*(1/60,1/120)
        const double md = NV == 2 ? 1.0 / 3.0 : (NV == 3 ? 2.0 / 12.0 : 6.0 / 60.0);
        const double mo = NV == 2 ? 1.0 / 6.0 : (NV == 3 ? 2.0 / 24.0 : 6.0 / 120.0);
        int row[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) row[v] = a.rowid ? a.rowid[nd[v]] : nd[v];
#pragma unroll
        for (int va = 0; va < NV; ++va) {
            const int beg = a.rowptr[row[va]], end = a.rowptr[row[va] + 1];
#pragma unroll
            for (int vb = 0; vb < NV; ++vb) {
                int lo = beg, hi = end - 1, pos = -1;
                while (lo <= hi) {
                    const int mid = (lo + hi) >> 1;
                    const int c = a.col[mid];
                    if (c == row[vb]) { pos = mid; break; }
                    if (c < row[vb]) lo = mid + 1; else hi = mid - 1;
                }
                if (pos < 0) { *a.status = 1; continue; }
                if (a.macc) atomicAdd(a.macc + pos, meas * (va == vb ? md : mo));
                if (a.kacc) atomicAdd(a.kacc + pos, meas * dot3(G[va], SG[vb]));
            }
        }
    }
}

__global__ void __launch_bounds__(256) round_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i];
}

}  // namespace
}  // namespace pdx

using namespace pdx;

extern "C" {

int pdx_fem_assemble(int nv, int64_t ne, const int32_t* elems, int elem_stride, const double* pts, const int32_t* rowid,
                     const double* sigma, const int32_t* rowptr, const int32_t* col, double* mass_acc, double* stiff_acc,
                     int32_t* status, void* stream) {
    if (nv < 2 || nv > 4 || ne < 0 || elem_stride < nv || !elems || !pts || !rowptr || !col || !status ||
        (!mass_acc && !stiff_acc)) {
        set_error("pdx_fem_assemble: bad arguments (P1 Edges / Trias / Tetras only)");
        return 1;
    }
    if (ne == 0) return 0;
    AsmArgs a{ne, elems, elem_stride, pts, rowid, sigma, rowptr, col, mass_acc, stiff_acc, status};
    int64_t nb = (ne + 127) / 128;
    if (nb > (long long)sm_count() * 64) nb = (long long)sm_count() * 64;
    cudaStream_t s = (cudaStream_t)stream;
    if (nv == 2) fem_assemble_kernel<2><<<(unsigned)nb, 128, 0, s>>>(a);
    else if (nv == 3) fem_assemble_kernel<3><<<(unsigned)nb, 128, 0, s>>>(a);
    else fem_assemble_kernel<4><<<(unsigned)nb, 128, 0, s>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "fem_assemble_kernel launch");
}

int pdx_round_to_f32(const double* in, float* out, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!in || !out))) { set_error("pdx_round_to_f32: bad arguments"); return 1; }
    if (n == 0) return 0;
    int64_t nb = (n + 255) / 256;
    if (nb > (long long)sm_count() * 32) nb = (long long)sm_count() * 32;
    round_to_f32_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(in, out, n);
    count_launch();
    return check_cuda(cudaGetLastError(), "round_to_f32_kernel launch");
}

}  // extern "C"
