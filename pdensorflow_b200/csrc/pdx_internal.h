// pdx_internal.h - host/device structs shared by the translation units of libpdx.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <string>
#include "../../include/pdx.h"
#include "pdx_math.cuh"

namespace pdx {

#define PDX_MAX_DEVICES 64

// Arguments of one fused FD step launch (by value, __grid_constant__).
struct FdArgs {
    const float* Uin;
    float*       Uout;
    float*       st[PDX_MAX_STATES];
    const float* DIFF;
    const float* AVEC;           // fibre dyadic [nx][ny][nz][6] (ANISO)
    // fused halo store (slab decomposition over NVLink peer memory): the output of plane
    // mirror_lo_plane / mirror_hi_plane is ALSO stored to these plane-sized peer buffers (the
    // neighbours' ghost planes); nullptr = off
    float* mirror_lo;
    float* mirror_hi;
    int mirror_lo_plane, mirror_hi_plane;
    int nx, ny, nz;
    int pz;                      // floats between consecutive rows (= nz unless the arrays are pitched)
    int xb, xe;                  // planes advanced
    int xclo, xchi;              // clamp of stencil plane indices for U (enforce_boundary + pad)
    int yclo, ychi, zclo, zchi;  // same for the y / z axes
    int dxlo, dxhi;              // clamp of plane indices for DIFF (symmetric pad only)
    int x_chunk;                 // planes per CTA (TMA kernel)
    int nchunk;                  // x-chunks of this launch (set by the launcher)
    int raster;                  // G > 0: tile groups of G, x-chunks next, odd chunks march downwards (pdx_fd_tma.cu)
    int ntz, nty;                // tile counts (TMA kernel)
    int has_x;                   // 0 for 2-D fields
    int dirichlet;
    float dirichlet_value;
    float dt, coef;
    float hx, hy, hz;            // spacings (ANISO divides by them)
    float hxsq, hysq, hzsq;      // dx*dx ... (fp32 products)
    float rhxsq, rhysq, rhzsq;   // reciprocals (FAST)
    float kx, ky, kz, kc;        // conv-kernel taps (CONV): see laplace_convolution...py:30-40
    ModelParams mp;
};

// lookup-table models (pdx_ionic_lut.cu)
struct LutTab {
    const float* t;
    int stride;
    float mn, mx, res, step;
    int mx_idx;
};
struct LutArgs {
    LutTab tab[3];
    float  c[PDX_LUT_MAX_CONSTS];
    float* st[PDX_LUT_MAX_STATES];
    float  dt;
};

void set_error(const std::string& msg);
int  check_cuda(cudaError_t e, const char* what);
void count_launch(int64_t n = 1);
int  sm_count();                       // multiprocessors of the current device (cached per ordinal)
int  sell_width_hint(const int32_t* slice_ptr);   // widest slice registered through pdx_sell_hint, or 0
int  csr_group_hint(const int32_t* rowptr);   // lanes per row registered through pdx_csr_hint, or 0

void fill_model_params(int model, const float* params, float dt, ModelParams* out);

// generic (any shape, 2-D/3-D, every operator/model) fused step
int launch_fd_generic(const pdx_fd_desc& d, const FdArgs& a, cudaStream_t s);
int launch_fd_laplace_only(int lap_kind, int math_mode, const FdArgs& a, cudaStream_t s);
int launch_enforce_boundary(const FdArgs& a, cudaStream_t s);
int launch_ionic(int model, int math, const float* params, const float* const* parr, int64_t n, const float* U,
                 float* const* states, float* dU, float dt, const float* I0, float* rhs0, cudaStream_t s);
int launch_stim(float* U, const float* mask, float intensity, int64_t n, cudaStream_t s);
int launch_stim_max(float* U, const float* mask, float intensity, int64_t n, cudaStream_t s);
int launch_fd_laplace_aniso(const FdArgs& a, cudaStream_t s);
int lut_model_num_consts(int model);
// one translation unit per arithmetic policy (pdx_ionic_lut.cu is compiled twice)
int launch_lut_ionic_exact(int model, const LutArgs& L, int64_t n, const float* U, float* dU, const float* I0,
                           float* rhs0, cudaStream_t s);
int launch_lut_ionic_fast(int model, const LutArgs& L, int64_t n, const float* U, float* dU, const float* I0,
                          float* rhs0, cudaStream_t s);
int launch_fd_lut_exact(const pdx_fd_desc& d, const FdArgs& a, const LutArgs& L, cudaStream_t s);
int launch_fd_lut_fast(const pdx_fd_desc& d, const FdArgs& a, const LutArgs& L, cudaStream_t s);
// TMA-staged tile kernel: requires nz % 4 == 0, 16-byte aligned arrays, HOMOG/HETEROG
bool fd_tma_supported(const pdx_fd_desc& d);
// mapS: per-state tensor maps (box = fd_tma_state_box) to stage the states by TMA, or nullptr to
// prefetch them in registers
int  launch_fd_tma(const pdx_fd_desc& d, const FdArgs& a, const CUtensorMap& mapU, const CUtensorMap& mapD,
                   const CUtensorMap* const* mapS, cudaStream_t s);
void fd_tma_state_box(int* box_y, int* box_z);
int  make_tensor_map_3d(CUtensorMap* map, const float* base, int nx, int ny, int nz, int pitch_z, int box_x, int box_y,
                        int box_z);
// geometry of the TMA kernel's tile, needed by the host to build the tensor maps
void fd_tma_tile_geometry(const pdx_fd_desc& d, int* tile_y, int* tile_z, int* box_y, int* box_z);
int  fd_tma_auto_chunk(const pdx_fd_desc& d);

// 19-point fibre-tensor operator on TMA tiles (pdx_fd_aniso.cu): the six tensor components as six [n] arrays C
bool fd_aniso_tile_supported(const pdx_fd_desc& d);
int  fd_aniso_auto_chunk(const pdx_fd_desc& d);
int  launch_aniso_tensor(const float* DIFF, const float* AVEC, float* C, int64_t n, cudaStream_t s);
int  launch_fd_aniso_tile(const pdx_fd_desc& d, const FdArgs& a, const CUtensorMap& mapU, const CUtensorMap* mapC /* 6 */,
                          const CUtensorMap* const* mapS /* per state, box = fd_tma_state_box */, cudaStream_t s);

// SM-resident multi-step kernel for 2-D grids (pdx_fd_resident.cu)
#define PDX_RESIDENT_FLAG_STRIDE 32
struct ResidentGeom {
    int rows_per_cta, ncta, threads;
    size_t smem_bytes;
};
bool fd_resident_plan(const pdx_fd_desc& d, ResidentGeom* g);
int  launch_fd_resident(const pdx_fd_desc& d, const FdArgs& f, const ResidentGeom& g, int nsteps, unsigned int base,
                        unsigned int* flags, float* halo, cudaStream_t s);

}  // namespace pdx
