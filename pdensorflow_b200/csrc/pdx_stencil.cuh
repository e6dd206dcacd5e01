// pdx_stencil.cuh - one-node evaluation of the reference's diffusion operators through
// clamped global loads (L1/L2 provide the reuse).  Shared by the generic fused step, the
// stand-alone operators and the lookup-table-model step.
//
//   HOMOG   gpuSolve/diffop3D/laplace_homogeneous_isotropic_diffusion.py:29-31 (+ diffop2D)
//   CONV    gpuSolve/diffop3D/laplace_convolution_homogeneous_isotropic_diffusion.py:30-47
//   HETEROG gpuSolve/diffop3D/laplace_heterogeneous_isotropic_diffusion.py:20-34 (+ diffop2D)
//   ANISO   gpuSolve/diffop3D/laplace_heterogeneous_anisotropic_diffusion.py:40-93
//
// tf.pad(mode='symmetric') by one cell == clamping the neighbour index to the array; the
// examples' enforce_boundary (Tests/FD/Fenton/fenton.py:46-53) == clamping the index once more
// to [1, n-2].  Both clamps are applied here, in that order.
#pragma once
#include "pdx_internal.h"

namespace pdx {

__device__ __forceinline__ int clampi_(int v, int lo, int hi) { return min(max(v, lo), hi); }

struct Field {
    const float* U;
    int ny, nz, pz;
    int xclo, xchi, yclo, ychi, zclo, zchi;
    int dirichlet;
    float dval;
    // value of U0 = enforce_boundary(U) at an in-array index
    __device__ __forceinline__ float at(int i, int j, int k) const {
        if (dirichlet) {
            const bool inside = i >= xclo && i <= xchi && j >= yclo && j <= ychi && k >= zclo && k <= zchi;
            return inside ? U[((size_t)i * ny + j) * pz + k] : dval;
        }
        i = clampi_(i, xclo, xchi);
        j = clampi_(j, yclo, ychi);
        k = clampi_(k, zclo, zchi);
        return U[((size_t)i * ny + j) * pz + k];
    }
};

__device__ __forceinline__ Field make_field(const FdArgs& a) {
    return Field{a.Uin, a.ny, a.nz, a.pz, a.xclo, a.xchi, a.yclo, a.ychi, a.zclo, a.zchi, a.dirichlet, a.dirichlet_value};
}

// 19-point fibre-tensor operator.  DIFF is [nx][ny][nz][2] = (sigma_t, sigma_l - sigma_t),
// AVEC is [nx][ny][nz][6] = (a1a1, a1a2, a1a3, a2a2, a2a3, a3a3).  For axis A and side SG
//   flux = SG*0.5*( sum_b T_b ),  T_A = (C_AA[SG e_A] + C_AA[0]) * (X[SG e_A] - X[0]) / h_A,
//   T_b = 0.5*(C_Ab[SG e_A + e_b] + C_Ab[SG e_A - e_b]) * (X[SG e_A + e_b] - X[SG e_A - e_b]) / h_b
// and lap = (e - w)/h_x + (n - s)/h_y + (u - d)/h_z.  Always evaluated op-for-op (the operator
// is not on any benchmarked configuration).
struct AnisoCtx {
    const FdArgs& a;
    const Field& f;
    int i, j, k;
    __device__ __forceinline__ int ci(int d) const { return clampi_(i + d, a.dxlo, a.dxhi); }
    __device__ __forceinline__ int cj(int d) const { return clampi_(j + d, 0, a.ny - 1); }
    __device__ __forceinline__ int ck(int d) const { return clampi_(k + d, 0, a.nz - 1); }
    __device__ __forceinline__ float X(int di, int dj, int dk) const { return f.at(ci(di), cj(dj), ck(dk)); }
    // tensor component comp (0..5) at an offset cell
    __device__ __forceinline__ float C(int comp, int di, int dj, int dk) const {
        const size_t id = ((size_t)ci(di) * a.ny + cj(dj)) * a.nz + ck(dk);
        const float dl = a.DIFF[2 * id + 1];
        const float t = xmul(a.AVEC[6 * id + comp], dl);
        return (comp == 0 || comp == 3 || comp == 5) ? xadd(a.DIFF[2 * id], t) : t;
    }
};

__device__ __forceinline__ int aniso_comp(int p, int q) {
    const int lo = p < q ? p : q, hi = p < q ? q : p;
    return lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);      // (0,0)0 (0,1)1 (0,2)2 (1,1)3 (1,2)4 (2,2)5
}

template <int A, int SG>
__device__ __forceinline__ float aniso_flux(const AnisoCtx& c) {
    const float h[3] = {c.a.hx, c.a.hy, c.a.hz};
    float total = 0.0f;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        int o[3] = {0, 0, 0};
        o[A] = SG;
        float t;
        if (b == A) {
            const int cc = aniso_comp(A, A);
            t = xdiv(xmul(xadd(c.C(cc, o[0], o[1], o[2]), c.C(cc, 0, 0, 0)),
                          xsub(c.X(o[0], o[1], o[2]), c.X(0, 0, 0))), h[A]);
        } else {
            const int cc = aniso_comp(A, b);
            int p[3] = {o[0], o[1], o[2]}, m[3] = {o[0], o[1], o[2]};
            p[b] += 1;
            m[b] -= 1;
            t = xdiv(xmul(xmul(0.5f, xadd(c.C(cc, p[0], p[1], p[2]), c.C(cc, m[0], m[1], m[2]))),
                          xsub(c.X(p[0], p[1], p[2]), c.X(m[0], m[1], m[2]))), h[b]);
        }
        total = b == 0 ? t : xadd(total, t);
    }
    return xmul(SG > 0 ? 0.5f : -0.5f, total);
}

// Laplacian of U0 = enforce_boundary(U) at node (i,j,k); also returns the centre value c.
template <int LAP, int MATH>
__device__ __forceinline__ float node_laplacian(const FdArgs& a, const Field& f, int i, int j, int k, float& c) {
    // symmetric pad: neighbour indices are first clamped to the (global) array bounds
    const int im = max(i - 1, a.dxlo), ip = min(i + 1, a.dxhi);
    const int jm = max(j - 1, 0), jp = min(j + 1, a.ny - 1);
    const int km = max(k - 1, 0), kp = min(k + 1, a.nz - 1);
    c = f.at(i, j, k);
    float lap;
    if (LAP == PDX_LAP_ANISO) {
        const AnisoCtx ctx{a, f, i, j, k};
        const float gx = xdiv(xsub(aniso_flux<0, 1>(ctx), aniso_flux<0, -1>(ctx)), a.hx);
        const float gy = xdiv(xsub(aniso_flux<1, 1>(ctx), aniso_flux<1, -1>(ctx)), a.hy);
        const float gz = xdiv(xsub(aniso_flux<2, 1>(ctx), aniso_flux<2, -1>(ctx)), a.hz);
        lap = xadd(xadd(gx, gy), gz);
    } else if (LAP == PDX_LAP_HETEROG) {
        const float* D = a.DIFF;
        auto dat = [&](int ii, int jj, int kk) { return D[((size_t)ii * a.ny + jj) * a.pz + kk]; };
        const float dc = dat(i, j, k);
        lap = lap_axis_het<MATH>(f.at(i, jm, k), c, f.at(i, jp, k), dat(i, jm, k), dc, dat(i, jp, k), a.hysq, a.rhysq);
        if (a.has_x) {
            const float lx = lap_axis_het<MATH>(f.at(im, j, k), c, f.at(ip, j, k), dat(im, j, k), dc, dat(ip, j, k),
                                                a.hxsq, a.rhxsq);
            lap = madd<MATH>(lx, lap);
        }
        lap = madd<MATH>(lap, lap_axis_het<MATH>(f.at(i, j, km), c, f.at(i, j, kp), dat(i, j, km), dc, dat(i, j, kp),
                                                 a.hzsq, a.rhzsq));
    } else if (LAP == PDX_LAP_CONV) {
        // 3x3x3 VALID convolution, taps summed in kernel memory order
        float o = mmul<MATH>(a.kx, f.at(im, j, k));
        o = madd<MATH>(o, mmul<MATH>(a.ky, f.at(i, jm, k)));
        o = madd<MATH>(o, mmul<MATH>(a.kz, f.at(i, j, km)));
        o = madd<MATH>(o, mmul<MATH>(a.kc, c));
        o = madd<MATH>(o, mmul<MATH>(a.kz, f.at(i, j, kp)));
        o = madd<MATH>(o, mmul<MATH>(a.ky, f.at(i, jp, k)));
        o = madd<MATH>(o, mmul<MATH>(a.kx, f.at(ip, j, k)));
        lap = o;
    } else {
        lap = lap_axis<MATH>(f.at(i, jm, k), c, f.at(i, jp, k), a.hysq, a.rhysq);
        if (a.has_x) lap = madd<MATH>(lap_axis<MATH>(f.at(im, j, k), c, f.at(ip, j, k), a.hxsq, a.rhxsq), lap);
        lap = madd<MATH>(lap, lap_axis<MATH>(f.at(i, j, km), c, f.at(i, j, kp), a.hzsq, a.rhzsq));
    }
    return lap;
}

}  // namespace pdx
