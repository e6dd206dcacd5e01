// pdx_fd_tma.cu - fused FD step for sm_100a: TMA-staged shared-memory tiles, march along
// the slowest axis, stencil + ionic update + Euler update in one pass.
//
// Work decomposition
//   CTA  = one (TY x 128) tile of the (y,z) plane marched over x_chunk planes.
//   warp = R consecutive rows of the tile; lane = 4 consecutive z nodes (float4).
// Data movement per node-step (Fenton-Karma): U is read once through TMA
// (cp.async.bulk.tensor.3d, box = 1 x (TY+2) x (128+8) including the halo, zero-filled
// outside the array) into a ring of NSTAGE shared-memory plane slots guarded by mbarriers;
// V,W,S are read and written once with 128-bit global accesses (prefetched one plane
// ahead in registers); U' is written once with 128-bit stores.  Halo re-reads hit L2.
// The clamp-index boundary rule (enforce_boundary + symmetric pad) is applied
// when reading shared memory: plane/row indices are clamped before addressing, z edges are
// patched in registers (only in tiles that touch the boundary).
#include "pdx_internal.h"
#include "pdx_tma.cuh"

namespace pdx {

namespace {

constexpr int TZ = 128;              // tile extent along z (contiguous axis)
constexpr int ZH = 4;                // z halo per side in floats (keeps 16-byte alignment)
constexpr int COLS = TZ + 2 * ZH;    // 136 floats per smem row
constexpr int NWARP = 8;
constexpr int NTHREADS = NWARP * 32;
constexpr int STAGES = 6;            // ring depth (3 planes in use + 3 in flight)
constexpr int SSTAGES = 3;           // state ring depth (TMA-staged states: planes in flight per CTA)

constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <int R, bool HET, int NST = 0>
struct Geo {
    static constexpr int TY = NWARP * R;
    static constexpr int ROWS = TY + 2;
    static constexpr int FIELD_BYTES = round_up(ROWS * COLS * 4, 128);
    static constexpr int SLOT_BYTES = FIELD_BYTES * (HET ? 2 : 1);
    static constexpr int TX_BYTES = ROWS * COLS * 4 * (HET ? 2 : 1);
    // TMA-staged states: NST arrays, one (TY x TZ) box each per plane, SSTAGES planes deep
    static constexpr int ST_BYTES = TY * TZ * 4;
    static constexpr int SRING_BYTES = SSTAGES * NST * ST_BYTES;
    static constexpr int BAR_OFF = STAGES * SLOT_BYTES + SRING_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + (STAGES + SSTAGES) * 8;
};

struct StateMaps {
    CUtensorMap m[PDX_MAX_STATES];
};

// TST: the state arrays are staged through shared memory by TMA (SSTAGES planes ahead) instead of
// being prefetched one plane ahead in registers
template <int MODEL, bool HET, int MATH, bool HAS_X, int R, bool TST>
__global__ void __launch_bounds__(NTHREADS, (R == 1 && !HET) ? 3 : 2)
fd_tma_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapD,
              const __grid_constant__ StateMaps mapS, const __grid_constant__ FdArgs a) {
    constexpr int NS = NStates<MODEL>::value;
    using G = Geo<R, HET, TST ? NS : 0>;
    constexpr int TY = G::TY;
    constexpr int NSTAGE = STAGES;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::BAR_OFF);
    uint64_t* sfull = full + NSTAGE;
    unsigned char* sring = smem + NSTAGE * G::SLOT_BYTES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Rasterisation.  raster 0: tiles fastest, x-chunks slowest (an x-layer of CTAs at a time): right when a layer
    // is a wave or two of CTAs, because the two halo planes a chunk re-reads are then still in L2 (512^3: DRAM
    // reads 1.006 x algorithmic).  A 2048^2 layer is 9 waves and its halo planes are long evicted when the next
    // layer runs (ncu: reads 1.067 x).  raster G > 0 (3-D): groups of G consecutive tiles (z fastest, so a group
    // reads contiguous rows - DRAM page locality), inside a group the x-chunks come next and odd chunks march
    // DOWNWARDS: the resident set holds every chunk of a few tile groups, neighbouring chunks start together at
    // their shared planes and meet again at the other pair when they finish (reads 1.018 x).
    int b = blockIdx.x, chunk, tz, ty;
    if (HAS_X && a.raster > 0) {
        const int G = a.raster;
        const int ntile = a.ntz * a.nty;
        const int per_group = G * a.nchunk;
        const int grp = b / per_group;
        const int g0 = grp * G;
        const int gsz = min(G, ntile - g0);              // the last group may be short
        const int r = b - grp * per_group;
        chunk = r / gsz;
        const int t = g0 + (r - chunk * gsz);
        tz = t % a.ntz;
        ty = t / a.ntz;
    } else if (HAS_X) {
        tz = b % a.ntz;  b /= a.ntz;
        ty = b % a.nty;  b /= a.nty;
        chunk = b;
    } else {
        // 2-D fields: a CTA marches over x_chunk consecutive y-tiles of one z-strip, so that the TMA rings have
        // something to pipeline (a tile per CTA is launch-and-wait: 0.70 of the HBM peak on 4096^2)
        tz = b % a.ntz;
        ty = (b / a.ntz) * a.x_chunk;
        chunk = 0;
    }
    const int z0 = tz * TZ;
    const int y00 = ty * TY;                        // first (2-D: marched) y-tile of the CTA
    const int xb = HAS_X ? a.xb + chunk * a.x_chunk : a.xb;
    const int xe = HAS_X ? min(xb + a.x_chunk, a.xe) : a.xe;
    const int np = HAS_X ? xe - xb : min(a.x_chunk, a.nty - ty);    // march steps: planes (3-D) or y-tiles (2-D)
    const int nlog = HAS_X ? np + 2 : np;           // logical planes q = xb-1 .. xe (3-D)
    const bool down = HAS_X && a.raster > 0 && (chunk & 1);   // CTA uniform
    const int qs = down ? -1 : 1;                   // logical plane j is array plane q0 + qs*j
    const int q0 = HAS_X ? (down ? xe : xb - 1) : xb;
    const int p0 = down ? xe - 1 : xb;              // i-th plane advanced: p0 + qs*i
    auto y0_of = [&](int i) { return HAS_X ? y00 : y00 + i * TY; };   // y origin of march step i

    auto slotU = [&](int j) { return reinterpret_cast<float*>(smem + (j % NSTAGE) * G::SLOT_BYTES); };
    auto slotD = [&](int j) { return reinterpret_cast<float*>(smem + (j % NSTAGE) * G::SLOT_BYTES + G::FIELD_BYTES); };
    auto issue = [&](int j) {   // thread 0 only
        const int q = q0 + qs * j;                  // 3-D only
        uint64_t* bar = &full[j % NSTAGE];
        mbar_expect_tx(bar, G::TX_BYTES);
        tma_load_3d(slotU(j), &mapU, z0 - ZH, y0_of(j) - 1, HAS_X ? clampi(q, a.xclo, a.xchi) : xb, bar);
        if (HET) tma_load_3d(slotD(j), &mapD, z0 - ZH, y0_of(j) - 1, HAS_X ? clampi(q, a.dxlo, a.dxhi) : xb, bar);
    };
    auto wait_plane = [&](int j) { mbar_wait(&full[j % NSTAGE], (uint32_t)((j / NSTAGE) & 1)); };
    auto sslot = [&](int i, int s) {
        return reinterpret_cast<float*>(sring + ((i % SSTAGES) * (NS > 0 ? NS : 1) + s) * G::ST_BYTES);
    };
    auto issue_states = [&](int i) {   // thread 0 only; plane xb + i
        uint64_t* bar = &sfull[i % SSTAGES];
        mbar_expect_tx(bar, NS * G::ST_BYTES);
#pragma unroll
        for (int s = 0; s < NS; ++s) tma_load_3d(sslot(i, s), &mapS.m[s], z0, y0_of(i), HAS_X ? p0 + qs * i : xb, bar);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
#pragma unroll
        for (int s = 0; s < SSTAGES; ++s) mbar_init(&sfull[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int n0 = min(NSTAGE, nlog);
        for (int j = 0; j < n0; ++j) {
            issue(j);
            if (TST && NS > 0 && j < SSTAGES && j < np) issue_states(j);
        }
    }

    // ---- per-thread geometry (constant along the march) ----
    const int  z    = z0 + 4 * lane;
    const int  col  = ZH + 4 * lane;
    const bool zact = z < a.nz;
    const bool zb_tile = (z0 - 1 < a.zclo) || (z0 + TZ > a.zchi);            // CTA uniform
    const bool lo2 = z < a.zclo, lo1 = z - 1 < a.zclo;
    const bool hi2 = z + 3 > a.zchi, hi1 = z + 4 > a.zchi;
    const bool dlo1 = z - 1 < 0, dhi1 = z + 4 > a.nz - 1;                    // DIFF: symmetric pad only
    // Pitched rows (nz % 4 != 0, rows padded to a multiple of four: pdx_fd_desc.pitch_z): the clamp column may sit
    // anywhere among a lane's four columns.  General patch: columns beyond index h take the value at h (h < 0: the
    // column left of the four, read from shared memory).  zgen is CTA uniform.
    const bool zgen = (a.nz & 3) != 0;
    const int hiz = a.zchi - z, hizd = a.nz - 1 - z;
    auto fixhi = [&](F4& v, const float* rowp, int h) {
        if (h < 3) {
            const float tv = h < 0 ? rowp[-1] : (h == 0 ? v.v[0] : (h == 1 ? v.v[1] : v.v[2]));
            if (h < 0) v.v[0] = tv;
            if (h < 1) v.v[1] = tv;
            if (h < 2) v.v[2] = tv;
            v.v[3] = tv;
        }
    };

    // row geometry of a y-tile: constant along the march in 3-D, recomputed per marched tile in 2-D
    int  gy[R], rc[R], rm[R], rp[R], drm[R], drp[R], drc0;
    bool act[R], yb_tile;
    auto row_geometry = [&](int y0) {
        yb_tile = (y0 - 1 < a.yclo) || (y0 + TY > a.ychi);                   // CTA uniform
#pragma unroll
        for (int r = 0; r < R; ++r) {
            gy[r]  = y0 + warp * R + r;
            act[r] = zact && gy[r] < a.ny;
            const int g = min(gy[r], a.ny - 1);          // keep inactive rows addressable
            rc[r]  = (clampi(g, a.yclo, a.ychi) - (y0 - 1)) * COLS + col;
            rm[r]  = (clampi(g - 1, a.yclo, a.ychi) - (y0 - 1)) * COLS + col;
            rp[r]  = (clampi(g + 1, a.yclo, a.ychi) - (y0 - 1)) * COLS + col;
            drm[r] = (max(g - 1, 0) - (y0 - 1)) * COLS + col;
            drp[r] = (min(g + 1, a.ny - 1) - (y0 - 1)) * COLS + col;
        }
        drc0 = (0 - (y0 - 1)) * COLS + col;              // + g*COLS gives the unclamped DIFF row
    };
    row_geometry(y00);

    // ---- state prefetch registers ----
    F4 stn[R][NS > 0 ? NS : 1];
    auto load_states = [&](int p) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!act[r]) continue;
            const size_t gi = ((size_t)p * a.ny + gy[r]) * a.pz + z;
#pragma unroll
            for (int s = 0; s < NS; ++s) stn[r][s] = ldg4(a.st[s] + gi);
        }
    };
    if (!TST && NS > 0 && np > 0) load_states(p0);
    int srow[R];
#pragma unroll
    for (int r = 0; r < R; ++r) srow[r] = (warp * R + r) * TZ + 4 * lane;

    for (int i = 0; i < np; ++i) {
        const int p = HAS_X ? p0 + qs * i : xb;
        const int jc = HAS_X ? i + 1 : i;
        if (!HAS_X && i > 0) row_geometry(y0_of(i));        // 2-D: the next y-tile of the strip
        if (HAS_X) {
            if (i == 0) { wait_plane(0); wait_plane(1); }
            wait_plane(i + 2);
        } else {
            wait_plane(i);
        }
        const float* Sc = slotU(jc);
        const float* Sm = HAS_X ? slotU(jc - qs) : Sc;      // array plane p - 1
        const float* Sp = HAS_X ? slotU(jc + qs) : Sc;      // array plane p + 1
        // CTA uniform.  Dirichlet pad: the planes next to the clamp box read a replaced neighbour plane too
        const bool raw_diff = yb_tile || zb_tile || (HAS_X && (p < a.xclo || p > a.xchi)) ||
                              (HAS_X && a.dirichlet && (p - 1 < a.xclo || p + 1 > a.xchi));

        F4 stc[R][NS > 0 ? NS : 1];
        if (TST) {
            if (NS > 0) {
                mbar_wait(&sfull[i % SSTAGES], (uint32_t)((i / SSTAGES) & 1));
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int s = 0; s < NS; ++s) stc[r][s] = lds4(sslot(i, s) + srow[r]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int s = 0; s < NS; ++s) stc[r][s] = stn[r][s];
            if (HAS_X && NS > 0 && i + 1 < np) load_states(p + qs);   // (2-D marches stage their states by TMA)
        }

#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!act[r]) continue;
            const size_t gi = ((size_t)p * a.ny + gy[r]) * a.pz + z;
            F4 c = lds4(Sc + rc[r]);
            float zl = Sc[rc[r] - 1], zr = Sc[rc[r] + 4];
            F4 ym = lds4(Sc + rm[r]), yp = lds4(Sc + rp[r]);
            F4 xm, xp;
            if (HAS_X) { xm = lds4(Sm + rc[r]); xp = lds4(Sp + rc[r]); }
            F4 raw = c;
            if (raw_diff) {
                if (MODEL != PDX_MODEL_NONE) raw = ldg4(a.Uin + gi);
                if (a.dirichlet) {
                    // constant pad of LaplaceSolver (laplaceSolver.py:44-52): U0 is the boundary value wherever the
                    // (plane, row, column) of a stencil point lies outside the clamp box; shared memory was addressed
                    // with clamped (valid) indices, the values read there are replaced
                    const float dv = a.dirichlet_value;
                    const int g = gy[r];
                    const bool pin = !HAS_X || (p >= a.xclo && p <= a.xchi);
                    const bool pmi = HAS_X && p - 1 >= a.xclo && p - 1 <= a.xchi;
                    const bool ppi = HAS_X && p + 1 >= a.xclo && p + 1 <= a.xchi;
                    const bool yin = pin && g >= a.yclo && g <= a.ychi;
                    const bool ymi = pin && g - 1 >= a.yclo && g - 1 <= a.ychi;
                    const bool ypi = pin && g + 1 >= a.yclo && g + 1 <= a.ychi;
                    const bool yx = g >= a.yclo && g <= a.ychi;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool zin = z + e >= a.zclo && z + e <= a.zchi;
                        c.v[e]  = (yin && zin) ? c.v[e] : dv;
                        ym.v[e] = (ymi && zin) ? ym.v[e] : dv;
                        yp.v[e] = (ypi && zin) ? yp.v[e] : dv;
                        if (HAS_X) {
                            xm.v[e] = (pmi && yx && zin) ? xm.v[e] : dv;
                            xp.v[e] = (ppi && yx && zin) ? xp.v[e] : dv;
                        }
                    }
                    zl = (yin && z - 1 >= a.zclo && z - 1 <= a.zchi) ? zl : dv;
                    zr = (yin && z + 4 >= a.zclo && z + 4 <= a.zchi) ? zr : dv;
                } else if (zb_tile) {
                    if (zgen) {
                        zfix(c, lo2, false); zfix(ym, lo2, false); zfix(yp, lo2, false);
                        if (HAS_X) { zfix(xm, lo2, false); zfix(xp, lo2, false); }
                        zl = lo1 ? c.v[0] : zl;
                        fixhi(c, Sc + rc[r], hiz); fixhi(ym, Sc + rm[r], hiz); fixhi(yp, Sc + rp[r], hiz);
                        if (HAS_X) { fixhi(xm, Sm + rc[r], hiz); fixhi(xp, Sp + rc[r], hiz); }
                        zr = hi1 ? c.v[3] : zr;          // after fixhi c.v[3] holds the clamp column's value
                    } else {
                        zfix(c, lo2, hi2); zfix(ym, lo2, hi2); zfix(yp, lo2, hi2);
                        if (HAS_X) { zfix(xm, lo2, hi2); zfix(xp, lo2, hi2); }
                        zl = lo1 ? c.v[0] : zl;          // after zfix c.v[0] already holds the clamped value
                        zr = hi1 ? c.v[3] : zr;
                    }
                }
            }
            F4 dc, dym, dyp, dxm, dxp;
            float dzl = 0.f, dzr = 0.f;
            if (HET) {
                const float* Dc = slotD(jc);
                const int drc = drc0 + min(gy[r], a.ny - 1) * COLS;
                dc = lds4(Dc + drc);
                dzl = dlo1 ? dc.v[0] : Dc[drc - 1];
                if (zgen) fixhi(dc, Dc + drc, hizd);
                dzr = dhi1 ? dc.v[3] : Dc[drc + 4];
                dym = lds4(Dc + drm[r]); dyp = lds4(Dc + drp[r]);
                if (HAS_X) { dxm = lds4(slotD(jc - qs) + drc); dxp = lds4(slotD(jc + qs) + drc); }
            }
            F4 out;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float cc = c.v[e];
                const float zm_ = e == 0 ? zl : c.v[e - 1];
                const float zp_ = e == 3 ? zr : c.v[e + 1];
                float lap;
                if (HET) {
                    const float dcc = dc.v[e];
                    const float dzm_ = e == 0 ? dzl : dc.v[e - 1];
                    const float dzp_ = e == 3 ? dzr : dc.v[e + 1];
                    lap = lap_axis_het<MATH>(ym.v[e], cc, yp.v[e], dym.v[e], dcc, dyp.v[e], a.hysq, a.rhysq);
                    if (HAS_X)
                        lap = madd<MATH>(lap_axis_het<MATH>(xm.v[e], cc, xp.v[e], dxm.v[e], dcc, dxp.v[e], a.hxsq,
                                                            a.rhxsq), lap);
                    lap = madd<MATH>(lap, lap_axis_het<MATH>(zm_, cc, zp_, dzm_, dcc, dzp_, a.hzsq, a.rhzsq));
                } else {
                    lap = lap_axis<MATH>(ym.v[e], cc, yp.v[e], a.hysq, a.rhysq);
                    if (HAS_X) lap = madd<MATH>(lap_axis<MATH>(xm.v[e], cc, xp.v[e], a.hxsq, a.rhxsq), lap);
                    lap = madd<MATH>(lap, lap_axis<MATH>(zm_, cc, zp_, a.hzsq, a.rhzsq));
                }
                float u1 = cc;
                if (MODEL != PDX_MODEL_NONE) {
                    float st[PDX_MAX_STATES];
#pragma unroll
                    for (int s = 0; s < NS; ++s) st[s] = stc[r][s].v[e];
                    const float dU = ionic_update<MODEL, MATH>(raw.v[e], st, a.mp);
#pragma unroll
                    for (int s = 0; s < NS; ++s) stc[r][s].v[e] = st[s];
                    u1 = madd<MATH>(u1, mmul<MATH>(a.dt, dU));
                }
                out.v[e] = madd<MATH>(u1, mmul<MATH>(a.coef, lap));
            }
            stg4(a.Uout + gi, out);
            if (HAS_X) {   // fused halo store into the neighbours' ghost planes (peer memory over NVLink)
                if (a.mirror_lo && p == a.mirror_lo_plane) stg4(a.mirror_lo + (size_t)gy[r] * a.pz + z, out);
                if (a.mirror_hi && p == a.mirror_hi_plane) stg4(a.mirror_hi + (size_t)gy[r] * a.pz + z, out);
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) stg4(a.st[s] + gi, stc[r][s]);
        }
        __syncthreads();                              // everyone is done with logical plane i
        if (tid == 0) {
            if (i + NSTAGE < nlog) issue(i + NSTAGE);
            if (TST && NS > 0 && i + SSTAGES < np) issue_states(i + SSTAGES);
        }
    }
}

template <int MODEL, bool HET, int MATH, bool HAS_X, int R, bool TST>
int launch_inst(const FdArgs& a0, const CUtensorMap& mU, const CUtensorMap& mD, const StateMaps& mS, cudaStream_t s) {
    using G = Geo<R, HET, TST ? NStates<MODEL>::value : 0>;
    auto kern = fd_tma_kernel<MODEL, HET, MATH, HAS_X, R, TST>;
    // the opt-in to > 48 KB of dynamic shared memory is a PER-DEVICE function attribute: one flag per
    // instantiation and device ordinal (a process may drive several GPUs)
    static bool attr_done[PDX_MAX_DEVICES] = {};
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (dev < 0 || dev >= PDX_MAX_DEVICES || !attr_done[dev]) {
        if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES),
                       "cudaFuncSetAttribute(smem)"))
            return 1;
        if (dev >= 0 && dev < PDX_MAX_DEVICES) attr_done[dev] = true;
    }
    const int nplanes = a0.xe - a0.xb;
    if (nplanes <= 0) return 0;
    FdArgs a = a0;
    if (HAS_X) {
        a.nchunk = (nplanes + a.x_chunk - 1) / a.x_chunk;
        if (a.nchunk < 2) a.raster = 0;
    } else {
        // 2-D: x_chunk = y-tiles marched per CTA; without TMA-staged states there is nothing to look ahead with
        if (!TST && NStates<MODEL>::value > 0) a.x_chunk = 1;
        if (a.x_chunk < 1) a.x_chunk = 1;
        a.nchunk = (a.nty + a.x_chunk - 1) / a.x_chunk;
        a.raster = 0;
    }
    const long long nb = HAS_X ? (long long)a.ntz * a.nty * a.nchunk : (long long)a.ntz * a.nchunk;
    if (nb > 0x7fffffffLL) { set_error("fd_tma: grid too large"); return 1; }
    kern<<<(unsigned)nb, NTHREADS, G::SMEM_BYTES, s>>>(mU, mD, mS, a);
    count_launch();
    return check_cuda(cudaGetLastError(), "fd_tma_kernel launch");
}

template <int MODEL, bool HET, int MATH, bool TST>
int launch_hasx(const FdArgs& a, const CUtensorMap& mU, const CUtensorMap& mD, const StateMaps& mS, cudaStream_t s) {
    return a.has_x ? launch_inst<MODEL, HET, MATH, true, 1, TST>(a, mU, mD, mS, s)
                   : launch_inst<MODEL, HET, MATH, false, 1, TST>(a, mU, mD, mS, s);
}
template <int MODEL, bool HET, int MATH>
int launch_tst(bool tst, const FdArgs& a, const CUtensorMap& mU, const CUtensorMap& mD, const StateMaps& mS,
               cudaStream_t s) {
    if (NStates<MODEL>::value == 0) tst = false;
    return tst ? launch_hasx<MODEL, HET, MATH, true>(a, mU, mD, mS, s)
               : launch_hasx<MODEL, HET, MATH, false>(a, mU, mD, mS, s);
}
template <int MODEL, bool HET>
int launch_math(int math, bool tst, const FdArgs& a, const CUtensorMap& mU, const CUtensorMap& mD,
                const StateMaps& mS, cudaStream_t s) {
    return math == PDX_MATH_EXACT ? launch_tst<MODEL, HET, PDX_MATH_EXACT>(tst, a, mU, mD, mS, s)
                                  : launch_tst<MODEL, HET, PDX_MATH_FAST>(tst, a, mU, mD, mS, s);
}
template <int MODEL>
int launch_het(bool het, int math, bool tst, const FdArgs& a, const CUtensorMap& mU, const CUtensorMap& mD,
               const StateMaps& mS, cudaStream_t s) {
    return het ? launch_math<MODEL, true>(math, tst, a, mU, mD, mS, s)
               : launch_math<MODEL, false>(math, tst, a, mU, mD, mS, s);
}

}  // namespace

bool fd_tma_supported(const pdx_fd_desc& d) {
    const int pz = d.pitch_z > 0 ? d.pitch_z : d.nz;
    if (pz % 4 != 0 || d.nz < 3) return false;          // tensor maps need rows that are 16-byte multiples
    if (d.lap_kind == PDX_LAP_CONV && d.math_mode == PDX_MATH_EXACT) return false;   // different rounding order
    if (d.ny < 3) return false;
    return true;
}

void fd_tma_state_box(int* box_y, int* box_z) {
    *box_y = NWARP * 1;
    *box_z = TZ;
}

void fd_tma_tile_geometry(const pdx_fd_desc& d, int* tile_y, int* tile_z, int* box_y, int* box_z) {
    (void)d;
    *tile_y = NWARP * 1;
    *tile_z = TZ;
    *box_y  = NWARP * 1 + 2;
    *box_z  = COLS;
}

int fd_tma_auto_chunk(const pdx_fd_desc& d) {
    // Planes marched per CTA.  Measured on B200 (tools/chunk_sweep.py, 512^3 and 128x2048^2): the
    // optimum is flat at 6-8 planes - short chunks keep several x-layers of CTAs resident at once,
    // so the two halo planes each chunk re-reads are L2 hits, and the grid has thousands of CTAs to
    // balance over 148 SMs x 3.  Small grids shorten the chunk further to keep >= 2 waves of CTAs.
    const int nplanes = d.x_end - d.x_begin;
    if (d.ndim == 2) {          // y-tiles marched per CTA: as long as the grid still fills the SMs twice over
        const long long nty = (d.ny + NWARP - 1) / NWARP, ntz = (d.nz + TZ - 1) / TZ;
        int chunk = 4;          // tools/chunk_variants2.py: 4096^2 FK 188.1 Gnode-steps/s at 4 tiles, 184.0 at 8, 171.8 at 16
        while (chunk > 1 && ntz * ((nty + chunk - 1) / chunk) < 148LL * 3 * 2) --chunk;
        return chunk;
    }
    if (nplanes <= 1) return 1;
    const long long tiles = (long long)((d.ny + NWARP - 1) / NWARP) * ((d.nz + TZ - 1) / TZ);
    // The two halo planes a chunk re-reads and its pipeline fill weigh more the less a plane costs: the heat equation
    // (8 B per node, no ionic work) and the heterogeneous operator (two staged fields, 2 CTAs per SM) want longer marches
    // (tools/chunk_variants.py on 512^3, profiles/r02_chunk_variants.jsonl: heat 671 -> 744 Gnode-steps/s at 24 planes,
    // heterogeneous FK 163.8 -> 171.6; FK / mMS on the homogeneous operator are flat or lose beyond 8-12).
    int chunk = 8;
    if (d.model == PDX_MODEL_NONE || (d.lap_kind == PDX_LAP_HETEROG && d.model == PDX_MODEL_FENTON4V)) chunk = 24;
    else if (d.lap_kind == PDX_LAP_HETEROG) chunk = 16;
    while (chunk > 3 && tiles * ((nplanes + chunk - 1) / chunk) < 148LL * 3 * 2) --chunk;
    if (chunk > nplanes) chunk = nplanes;
    return chunk;
}

int launch_fd_tma(const pdx_fd_desc& d, const FdArgs& a, const CUtensorMap& mU, const CUtensorMap& mD,
                  const CUtensorMap* const* mS, cudaStream_t s) {
    const bool het = d.lap_kind == PDX_LAP_HETEROG;
    StateMaps sm;
    bool tst = mS != nullptr;
    for (int i = 0; i < PDX_MAX_STATES; ++i) {
        if (tst && mS[i]) sm.m[i] = *mS[i];
        else sm.m[i] = mU;
    }
    switch (d.model) {
        case PDX_MODEL_NONE:     return launch_het<PDX_MODEL_NONE>(het, d.math_mode, tst, a, mU, mD, sm, s);
        case PDX_MODEL_FENTON4V: return launch_het<PDX_MODEL_FENTON4V>(het, d.math_mode, tst, a, mU, mD, sm, s);
        case PDX_MODEL_MMS2V:    return launch_het<PDX_MODEL_MMS2V>(het, d.math_mode, tst, a, mU, mD, sm, s);
        case PDX_MODEL_MS2V:     return launch_het<PDX_MODEL_MS2V>(het, d.math_mode, tst, a, mU, mD, sm, s);
    }
    set_error("unknown model");
    return 1;
}

}  // namespace pdx
