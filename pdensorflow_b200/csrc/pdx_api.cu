// pdx_api.cu - the C ABI of libpdx.so (see include/pdx.h): plans, tensor maps, dispatch.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <unordered_map>
#include "pdx_internal.h"

namespace pdx {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int sm_count() {
    static std::atomic<int> cached[PDX_MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PDX_MAX_DEVICES) return 148;
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

static std::mutex g_hint_mu;
static std::unordered_map<const void*, int> g_hint;
static std::unordered_map<const void*, int> g_sell_hint;
int sell_width_hint(const int32_t* slice_ptr) {
    std::lock_guard<std::mutex> lk(g_hint_mu);
    auto it = g_sell_hint.find(slice_ptr);
    return it == g_sell_hint.end() ? 0 : it->second;
}
int csr_group_hint(const int32_t* rowptr) {
    std::lock_guard<std::mutex> lk(g_hint_mu);
    auto it = g_hint.find(rowptr);
    return it == g_hint.end() ? 0 : it->second;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return 1;
}

static const float kFkDefaults[FK_NPARAM] = {3.33f, 19.2f, 160.0f, 75.0f, 75.0f, 0.065f, 31.8364f, 31.8364f, 0.009f,
                                             0.23f, 0.146f, 0.0f, 1.0f, 0.8f, 0.3f, 0.02f, 1.2f, 3.0f,
                                             0.115f, 0.84f, 0.02f, -80.0f, 100.0f};
static const float kMmsDefaults[MS_NPARAM] = {0.1f, 9.0f, 100.0f, 120.0f, 0.13f, 0.13f, -80.0f, 100.0f};
static const float kMsDefaults[MS_NPARAM]  = {0.3f, 6.0f, 120.0f, 150.0f, 0.13f, 0.0f, -80.0f, 100.0f};

void fill_model_params(int model, const float* params, float dt, ModelParams* out) {
    std::memset(out, 0, sizeof(*out));
    out->dt = dt;
    const int np = pdx_model_num_params(model);
    for (int i = 0; i < np; ++i) {
        out->p[i] = params[i];
        out->r[i] = params[i] != 0.0f ? (float)(1.0 / (double)params[i]) : 0.0f;
    }
    const double log2e = 1.4426950408889634;
    if (model == PDX_MODEL_FENTON4V) {
        out->DV  = params[FK_DV];
        out->rDV = out->r[FK_DV];
        out->d0  = (float)(-2.0 * log2e / (double)params[FK_c_so]);
        out->d1  = (float)(-2.0 * log2e * (double)params[FK_k]);
        out->d2  = params[FK_a_so] - params[FK_tau_a];     // fp32, as the reference forms it
        out->d3  = params[FK_r_sp] - params[FK_r_sn];
    } else if (model == PDX_MODEL_MMS2V || model == PDX_MODEL_MS2V) {
        out->DV  = params[MS_DV];
        out->rDV = out->r[MS_DV];
    }
}

static CUresult encode_tiled(CUtensorMap* map, void* base, const cuuint64_t* gdim, const cuuint64_t* gstride,
                             const cuuint32_t* box) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p ||
            qres != cudaDriverEntryPointSuccess)
            return CUDA_ERROR_NOT_FOUND;
        fn = (EncodeFn)p;
    }
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int make_tensor_map_3d(CUtensorMap* map, const float* base, int nx, int ny, int nz, int pitch_z, int box_x, int box_y,
                       int box_z) {
    // rows of nz floats, pitch_z floats apart: columns nz .. pitch_z-1 are outside the tensor (zero filled on load)
    const cuuint64_t gdim[3]    = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nx};
    const cuuint64_t gstride[2] = {(cuuint64_t)pitch_z * 4, (cuuint64_t)pitch_z * ny * 4};
    const cuuint32_t box[3]     = {(cuuint32_t)box_z, (cuuint32_t)box_y, (cuuint32_t)box_x};
    const CUresult r = encode_tiled(map, const_cast<float*>(base), gdim, gstride, box);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult " + std::to_string((int)r) + ")");
        return 1;
    }
    return 0;
}

}  // namespace pdx

using namespace pdx;

struct pdx_fd_plan {
    pdx_fd_desc d;
    int kernel;
    FdArgs base;
    LutArgs lut;
    bool has_lut;
    const float* avec;
    int box_y, box_z;
    std::unordered_map<const void*, CUtensorMap> maps;
    std::unordered_map<const void*, CUtensorMap> smaps;   // state arrays (box without halo)
    bool tma_states;
    int64_t launches;
    // SM-resident multi-step path (2-D grids that fit on chip)
    bool resident_ok;
    bool resident_auto;
    bool want_resident;
    ResidentGeom rg;
    unsigned int* res_flags;
    float* res_halo;
    unsigned int res_base;
    // anisotropic operator on the tile kernel: tensor components formed once by pdx_fd_plan_prepare_aniso
    bool aniso_tile_ok;
    float* aniso_C;
    const float* aniso_diff;
    const float* aniso_avec;
    int aniso_chunk;
    CUtensorMap aniso_maps[6];
    ~pdx_fd_plan() {
        if (res_flags) cudaFree(res_flags);
        if (res_halo) cudaFree(res_halo);
        if (aniso_C) cudaFree(aniso_C);
    }
};

static int get_state_map(pdx_fd_plan* p, const float* ptr, const CUtensorMap** out) {
    auto it = p->smaps.find(ptr);
    if (it == p->smaps.end()) {
        CUtensorMap m;
        int by, bz;
        fd_tma_state_box(&by, &bz);
        if (make_tensor_map_3d(&m, ptr, p->d.nx, p->d.ny, p->d.nz, p->base.pz, 1, by, bz)) return 1;
        if (p->smaps.size() > 64) p->smaps.clear();
        it = p->smaps.emplace(ptr, m).first;
    }
    *out = &it->second;
    return 0;
}

static int get_map(pdx_fd_plan* p, const float* ptr, const CUtensorMap** out) {
    auto it = p->maps.find(ptr);
    if (it == p->maps.end()) {
        CUtensorMap m;
        if (make_tensor_map_3d(&m, ptr, p->d.nx, p->d.ny, p->d.nz, p->base.pz, 1, p->box_y, p->box_z)) return 1;
        if (p->maps.size() > 64) p->maps.clear();
        it = p->maps.emplace(ptr, m).first;
    }
    *out = &it->second;
    return 0;
}

extern "C" {

int pdx_version(void) { return PDX_VERSION; }
const char* pdx_last_error(void) { return g_err.c_str(); }
int64_t pdx_launch_count(void) { return g_launches.load(); }

int pdx_csr_hint(const int32_t* rowptr, int32_t n, int64_t nnz) {
    if (!rowptr || n <= 0 || nnz < 0) { set_error("pdx_csr_hint: bad arguments"); return 1; }
    const double mean = (double)nnz / n;
    const int G = mean <= 6 ? 4 : mean <= 12 ? 8 : mean <= 24 ? 16 : 32;
    std::lock_guard<std::mutex> lk(g_hint_mu);
    if (g_hint.size() > 256) g_hint.clear();
    g_hint[rowptr] = G;
    return 0;
}

int pdx_sell_hint(const int32_t* slice_ptr, int32_t max_width) {
    if (!slice_ptr || max_width < 0) { set_error("pdx_sell_hint: bad arguments"); return 1; }
    std::lock_guard<std::mutex> lk(g_hint_mu);
    if (g_sell_hint.size() > 256) g_sell_hint.clear();
    if (max_width == 0) g_sell_hint.erase(slice_ptr); else g_sell_hint[slice_ptr] = max_width;
    return 0;
}

int pdx_model_num_params(int model) {
    switch (model) {
        case PDX_MODEL_NONE: return 0;
        case PDX_MODEL_FENTON4V: return FK_NPARAM;
        case PDX_MODEL_MMS2V:
        case PDX_MODEL_MS2V: return MS_NPARAM;
        case PDX_MODEL_CRN:
        case PDX_MODEL_TTP: return lut_model_num_consts(model);
    }
    return -1;
}
int pdx_model_num_states(int model) {
    switch (model) {
        case PDX_MODEL_NONE: return 0;
        case PDX_MODEL_FENTON4V: return 3;
        case PDX_MODEL_MMS2V:
        case PDX_MODEL_MS2V: return 1;
        case PDX_MODEL_CRN: return 19;
        case PDX_MODEL_TTP: return 22;
    }
    return -1;
}

static bool is_lut_model(int model) { return model == PDX_MODEL_CRN || model == PDX_MODEL_TTP; }

static int fill_lut_args(const pdx_lut_model* m, LutArgs* L) {
    if (!m || !is_lut_model(m->model)) { set_error("pdx_lut_model: model must be PDX_MODEL_CRN or PDX_MODEL_TTP"); return 1; }
    if (m->nconsts != lut_model_num_consts(m->model)) { set_error("pdx_lut_model: wrong number of constants"); return 1; }
    static const int min_cols[2][3] = {{32, 5, 3}, {27, 2, 1}};
    std::memset(L, 0, sizeof(*L));
    for (int t = 0; t < 3; ++t) {
        const pdx_lut_table& T = m->tab[t];
        const int need = min_cols[m->model == PDX_MODEL_TTP][t];
        if (!T.data || T.ncols < need || T.stride < T.ncols || T.nrows < 2 || T.mx_idx < 1 || T.mx_idx > T.nrows - 1 ||
            !(T.res > 0.0f)) {
            set_error("pdx_lut_model: bad table descriptor");
            return 1;
        }
        if (((uintptr_t)T.data % 8) != 0 || (T.stride % 2 != 0 && T.stride != 1)) {
            set_error("pdx_lut_model: tables must be 8-byte aligned with an even row stride");
            return 1;
        }
        L->tab[t] = LutTab{T.data, T.stride, T.mn, T.mx, T.res, T.step, T.mx_idx};
    }
    for (int i = 0; i < m->nconsts; ++i) L->c[i] = m->consts[i];
    L->dt = m->dt;
    return 0;
}
int pdx_model_default_params(int model, float* out) {
    const int np = pdx_model_num_params(model);
    if (np < 0 || !out) { set_error("pdx_model_default_params: bad arguments"); return 1; }
    for (int i = 0; i < PDX_MAX_PARAMS; ++i) out[i] = 0.0f;
    if (model == PDX_MODEL_CRN || model == PDX_MODEL_TTP) {
        set_error("pdx_model_default_params: lookup-table models take pdx_lut_model constants");
        return 1;
    }
    const float* src = model == PDX_MODEL_FENTON4V ? kFkDefaults : model == PDX_MODEL_MMS2V ? kMmsDefaults : kMsDefaults;
    for (int i = 0; i < np; ++i) out[i] = src[i];
    return 0;
}

void pdx_fd_desc_init(pdx_fd_desc* d) {
    std::memset(d, 0, sizeof(*d));
    d->ndim = 3;
    d->dx = d->dy = d->dz = 1.0f;
    d->math_mode = PDX_MATH_FAST;
    d->kernel = PDX_KERNEL_AUTO;
    d->x_begin = -1;   // resolved in plan_create
}

static int fill_args(const pdx_fd_desc& d, FdArgs* a) {
    std::memset(a, 0, sizeof(*a));
    a->nx = d.nx; a->ny = d.ny; a->nz = d.nz;
    a->pz = d.pitch_z > 0 ? d.pitch_z : d.nz;
    a->has_x = d.ndim == 3;
    a->xb = d.x_begin; a->xe = d.x_end;
    if (a->has_x) {
        a->xclo = d.x_clamp_lo; a->xchi = d.x_clamp_hi;
        a->dxlo = a->xclo - 1 > 0 ? a->xclo - 1 : 0;
        a->dxhi = a->xchi + 1 < d.nx - 1 ? a->xchi + 1 : d.nx - 1;
    }
    a->yclo = 1; a->ychi = d.ny - 2;
    a->zclo = 1; a->zchi = d.nz - 2;
    a->dirichlet = d.dirichlet; a->dirichlet_value = d.dirichlet_value;
    a->dt = d.dt; a->coef = d.coef;
    a->hx = d.dx; a->hy = d.dy; a->hz = d.dz;
    a->hxsq = d.dx * d.dx; a->hysq = d.dy * d.dy; a->hzsq = d.dz * d.dz;
    a->rhxsq = (float)(1.0 / (double)a->hxsq);
    a->rhysq = (float)(1.0 / (double)a->hysq);
    a->rhzsq = (float)(1.0 / (double)a->hzsq);
    // conv taps, laplace_convolution_homogeneous_isotropic_diffusion.py:26-40: the kernel as
    // written puts 1/dzsq on the axis-0 neighbours and 1/dxsq on the axis-2 neighbours
    const float ix = 1.0f / a->hxsq, iy = 1.0f / a->hysq, iz = 1.0f / a->hzsq;
    a->kx = iz; a->ky = iy; a->kz = ix;
    a->kc = -2.0f * ((ix + iy) + iz);
    if (!is_lut_model(d.model)) fill_model_params(d.model, d.params, d.dt, &a->mp);
    return 0;
}

int pdx_fd_plan_create(const pdx_fd_desc* din, pdx_fd_plan** out) {
    if (!din || !out) { set_error("pdx_fd_plan_create: null argument"); return 1; }
    pdx_fd_desc d = *din;
    if (d.ndim != 2 && d.ndim != 3) { set_error("pdx_fd_plan_create: ndim must be 2 or 3"); return 1; }
    if (d.ndim == 2 && d.nx != 1) { set_error("pdx_fd_plan_create: 2-D fields use nx = 1"); return 1; }
    if (d.ny < 3 || d.nz < 3 || (d.ndim == 3 && d.nx < 3)) {
        set_error("pdx_fd_plan_create: every active axis needs at least 3 cells");
        return 1;
    }
    if (pdx_model_num_states(d.model) < 0) { set_error("pdx_fd_plan_create: unknown model"); return 1; }
    if (d.pitch_z != 0 && d.pitch_z != d.nz) {
        if (d.pitch_z < d.nz) { set_error("pdx_fd_plan_create: pitch_z < nz"); return 1; }
        if (is_lut_model(d.model) || d.lap_kind == PDX_LAP_ANISO || d.kernel == PDX_KERNEL_RESIDENT) {
            set_error("pdx_fd_plan_create: pitched rows are for the closed-form models with the homogeneous / conv / "
                      "heterogeneous operators on the per-step kernels");
            return 1;
        }
    }
    if (d.lap_kind < PDX_LAP_HOMOG || d.lap_kind > PDX_LAP_ANISO) { set_error("pdx_fd_plan_create: unknown lap_kind"); return 1; }
    if ((d.lap_kind == PDX_LAP_CONV || d.lap_kind == PDX_LAP_ANISO) && d.ndim == 2) {
        set_error("pdx_fd_plan_create: the conv and anisotropic operators are 3-D only");
        return 1;
    }
    if (d.math_mode != PDX_MATH_EXACT && d.math_mode != PDX_MATH_FAST) { set_error("pdx_fd_plan_create: unknown math_mode"); return 1; }
    if (d.ndim == 2) {
        d.x_begin = 0; d.x_end = 1; d.x_clamp_lo = 0; d.x_clamp_hi = 0;
    } else if (d.x_begin < 0) {   // stand-alone grid defaults
        d.x_begin = 0; d.x_end = d.nx; d.x_clamp_lo = 1; d.x_clamp_hi = d.nx - 2;
    }
    if (d.x_begin < 0 || d.x_end > d.nx || d.x_begin > d.x_end || d.x_clamp_lo < 0 || d.x_clamp_hi > d.nx - 1 ||
        d.x_clamp_lo > d.x_clamp_hi) {
        set_error("pdx_fd_plan_create: inconsistent x range / clamp");
        return 1;
    }
    int kernel = d.kernel;
    const bool tma_ok = fd_tma_supported(d) && !is_lut_model(d.model) && d.lap_kind != PDX_LAP_ANISO;
    ResidentGeom rg{};
    const bool res_ok = fd_resident_plan(d, &rg);
    if (kernel == PDX_KERNEL_RESIDENT && !res_ok) {
        set_error("pdx_fd_plan_create: the resident kernel takes 2-D grids (nz <= 1024) that fit in the SMs' shared "
                  "memory, homogeneous operator, closed-form models, no Dirichlet pad");
        return 1;
    }
    const bool want_resident = kernel == PDX_KERNEL_RESIDENT;
    if (kernel == PDX_KERNEL_AUTO || want_resident) kernel = tma_ok ? PDX_KERNEL_TMA : PDX_KERNEL_GENERIC;
    if (kernel == PDX_KERNEL_TMA && !tma_ok && d.lap_kind == PDX_LAP_ANISO && fd_aniso_tile_supported(d)) {
        kernel = PDX_KERNEL_GENERIC;      // the tile kernel takes over once pdx_fd_plan_prepare_aniso has run
    }
    if (kernel == PDX_KERNEL_TMA && !tma_ok) {
        set_error("pdx_fd_plan_create: TMA kernel needs nz % 4 == 0, no EXACT conv");
        return 1;
    }
    pdx_fd_plan* p = new (std::nothrow) pdx_fd_plan();
    if (!p) { set_error("out of memory"); return 1; }
    p->d = d;
    p->kernel = kernel;
    p->launches = 0;
    p->resident_ok = res_ok;
    p->resident_auto = false;
    p->rg = rg;
    p->res_flags = nullptr;
    p->res_halo = nullptr;
    p->res_base = 0;
    if (res_ok && (want_resident || d.kernel == PDX_KERNEL_AUTO)) {
        const char* e = std::getenv("PDX_RESIDENT");            // A/B switch: PDX_RESIDENT=0 keeps AUTO on the per-step kernels
        p->resident_auto = want_resident || !(e && e[0] == '0');
        if (check_cuda(cudaMalloc(&p->res_flags, sizeof(unsigned int) * PDX_RESIDENT_FLAG_STRIDE * (size_t)rg.ncta), "cudaMalloc(resident flags)") ||
            check_cuda(cudaMemset(p->res_flags, 0, sizeof(unsigned int) * PDX_RESIDENT_FLAG_STRIDE * (size_t)rg.ncta), "clear resident flags") ||
            check_cuda(cudaMalloc(&p->res_halo, sizeof(float) * 4 * (size_t)rg.ncta * d.nz), "cudaMalloc(resident halo)")) {
            delete p;
            return 1;
        }
    }
    p->want_resident = want_resident;
    p->aniso_tile_ok = fd_aniso_tile_supported(d) && d.kernel != PDX_KERNEL_GENERIC;
    {   // A/B switch: PDX_ANISO_TILE=0 keeps the anisotropic operator on the generic kernel
        const char* e = std::getenv("PDX_ANISO_TILE");
        if (e && e[0] == '0') p->aniso_tile_ok = false;
    }
    p->aniso_C = nullptr;
    p->aniso_diff = p->aniso_avec = nullptr;
    p->aniso_chunk = d.x_chunk > 0 ? d.x_chunk : fd_aniso_auto_chunk(d);
    p->has_lut = false;
    p->avec = nullptr;
    {   // states staged by TMA unless PDX_TMA_STATES=0 (A/B switch for profiling)
        const char* e = std::getenv("PDX_TMA_STATES");
        p->tma_states = !(e && e[0] == '0');
    }
    fill_args(d, &p->base);
    int ty, tz;
    fd_tma_tile_geometry(d, &ty, &tz, &p->box_y, &p->box_z);
    p->base.nty = (d.ny + ty - 1) / ty;
    p->base.ntz = (d.nz + tz - 1) / tz;
    p->base.x_chunk = d.x_chunk > 0 ? d.x_chunk : fd_tma_auto_chunk(d);
    {   // CTA rasterisation of the TMA kernel (pdx_fd_tma.cu): tile groups + alternating march direction once an
        // x-layer of CTAs is many waves (its halo planes would fall out of L2); PDX_FD_RASTER=<G> overrides (0 = off)
        const char* e = std::getenv("PDX_FD_RASTER");
        const long long layer = (long long)p->base.ntz * p->base.nty;
        p->base.raster = e ? std::atoi(e) : (layer > 1024 ? 16 : 0);
        if (p->base.raster < 0) p->base.raster = 0;
    }
    if (kernel == PDX_KERNEL_TMA) {
        // set the kernel's shared-memory attribute now (an empty range launches nothing), so
        // that the first real step can already be inside a CUDA-graph capture
        FdArgs a = p->base;
        a.xb = a.xe = 0;
        CUtensorMap dummy;
        std::memset(&dummy, 0, sizeof(dummy));
        pdx_fd_desc dd = d;
        if (dd.lap_kind == PDX_LAP_CONV) dd.lap_kind = PDX_LAP_HOMOG;
        const CUtensorMap* none[PDX_MAX_STATES] = {nullptr, nullptr, nullptr, nullptr};
        if (launch_fd_tma(dd, a, dummy, dummy, nullptr, nullptr) || launch_fd_tma(dd, a, dummy, dummy, none, nullptr)) {
            delete p;
            return 1;
        }
    }
    *out = p;
    return 0;
}

void pdx_fd_plan_destroy(pdx_fd_plan* p) { delete p; }

int pdx_fd_plan_set_lut(pdx_fd_plan* p, const pdx_lut_model* m) {
    if (!p || !m) { set_error("pdx_fd_plan_set_lut: null argument"); return 1; }
    if (m->model != p->d.model) { set_error("pdx_fd_plan_set_lut: model differs from the plan's descriptor"); return 1; }
    if (fill_lut_args(m, &p->lut)) return 1;
    p->has_lut = true;
    return 0;
}

int pdx_fd_plan_set_avec(pdx_fd_plan* p, const float* AVEC) {
    if (!p || !AVEC) { set_error("pdx_fd_plan_set_avec: null argument"); return 1; }
    p->avec = AVEC;
    return 0;
}
int pdx_fd_plan_aniso_tile_capable(const pdx_fd_plan* p) { return p && p->aniso_tile_ok ? 1 : 0; }

int pdx_fd_plan_prepare_aniso(pdx_fd_plan* p, const float* DIFF, const float* AVEC, void* stream) {
    if (!p || !DIFF || !AVEC) { set_error("pdx_fd_plan_prepare_aniso: null argument"); return 1; }
    if (!p->aniso_tile_ok) {
        set_error("pdx_fd_plan_prepare_aniso: the tile kernel takes 3-D grids with nz % 4 == 0, the closed-form models, "
                  "no Dirichlet pad (and kernel != GENERIC)");
        return 1;
    }
    const pdx_fd_desc& d = p->d;
    const int64_t n = (int64_t)d.nx * d.ny * d.nz;
    if (!p->aniso_C) {
        if (check_cuda(cudaMalloc(&p->aniso_C, sizeof(float) * 6 * (size_t)n), "cudaMalloc(anisotropic tensor)")) return 1;
        for (int c = 0; c < 6; ++c)
            if (make_tensor_map_3d(&p->aniso_maps[c], p->aniso_C + (size_t)c * n, d.nx, d.ny, d.nz, d.nz, 1, p->box_y, p->box_z)) return 1;
    }
    if (launch_aniso_tensor(DIFF, AVEC, p->aniso_C, n, (cudaStream_t)stream)) return 1;
    p->aniso_diff = DIFF;
    p->aniso_avec = AVEC;
    p->avec = AVEC;
    // opt in to the kernel's shared memory now, so that the first step may already be inside a graph capture
    FdArgs a = p->base;
    a.xb = a.xe = 0;
    CUtensorMap dummy;
    std::memset(&dummy, 0, sizeof(dummy));
    return launch_fd_aniso_tile(d, a, dummy, p->aniso_maps, nullptr, nullptr);
}

int pdx_fd_plan_kernel(const pdx_fd_plan* p) { return p ? (p->want_resident ? (int)PDX_KERNEL_RESIDENT : p->kernel) : -1; }
int pdx_fd_plan_resident_capable(const pdx_fd_plan* p) { return p && p->resident_ok && p->resident_auto ? 1 : 0; }
int64_t pdx_fd_plan_launches(const pdx_fd_plan* p) { return p ? p->launches : -1; }

static int step_once(pdx_fd_plan* p, const float* Uin, float* Uout, float* const* states, const float* DIFF, int xb,
                     int xe, cudaStream_t s, float* mirror_lo = nullptr, float* mirror_hi = nullptr) {
    const pdx_fd_desc& d = p->d;
    if (!Uin || !Uout || Uin == Uout) { set_error("pdx_fd_step: U_in/U_out must be distinct non-null buffers"); return 1; }
    const int ns = pdx_model_num_states(d.model);
    if (ns > 0 && !states) { set_error("pdx_fd_step: states required"); return 1; }
    if ((d.lap_kind == PDX_LAP_HETEROG || d.lap_kind == PDX_LAP_ANISO) && !DIFF) {
        set_error("pdx_fd_step: DIFF required for the heterogeneous / anisotropic operator");
        return 1;
    }
    if (d.lap_kind == PDX_LAP_ANISO && !p->avec) { set_error("pdx_fd_step: call pdx_fd_plan_set_avec first"); return 1; }
    FdArgs a = p->base;
    a.Uin = Uin; a.Uout = Uout; a.DIFF = DIFF; a.AVEC = p->avec;
    a.mirror_lo = mirror_lo; a.mirror_hi = mirror_hi;
    a.mirror_lo_plane = d.x_begin; a.mirror_hi_plane = d.x_end - 1;
    const bool lut = is_lut_model(d.model);
    if (lut && !p->has_lut) { set_error("pdx_fd_step: call pdx_fd_plan_set_lut first"); return 1; }
    for (int i = 0; i < ns; ++i) {
        if (!states[i]) { set_error("pdx_fd_step: null state array"); return 1; }
        if (lut) p->lut.st[i] = states[i]; else a.st[i] = states[i];
    }
    a.xb = xb; a.xe = xe;
    if (xe <= xb) return 0;
    if (lut) {
        const int rc = d.math_mode == PDX_MATH_FAST ? launch_fd_lut_fast(d, a, p->lut, s) : launch_fd_lut_exact(d, a, p->lut, s);
        if (!rc) p->launches += 1;
        return rc;
    }
    if (d.lap_kind == PDX_LAP_ANISO && p->aniso_C && DIFF == p->aniso_diff && p->avec == p->aniso_avec) {
        bool aligned = ((uintptr_t)Uin % 16 == 0) && ((uintptr_t)Uout % 16 == 0);
        for (int i = 0; i < ns; ++i) aligned = aligned && ((uintptr_t)states[i] % 16 == 0);
        if (aligned) {
            const CUtensorMap* mU = nullptr;
            if (get_map(p, Uin, &mU)) return 1;
            a.x_chunk = p->aniso_chunk;
            const CUtensorMap* mS[PDX_MAX_STATES] = {nullptr, nullptr, nullptr, nullptr};
            for (int i = 0; i < ns; ++i)
                if (get_state_map(p, states[i], &mS[i])) return 1;
            const int rc = launch_fd_aniso_tile(d, a, *mU, p->aniso_maps, mS, s);
            if (!rc) p->launches += 1;
            return rc;
        }
    }
    int kernel = p->kernel;
    if (kernel == PDX_KERNEL_TMA) {
        bool aligned = ((uintptr_t)Uin % 16 == 0) && ((uintptr_t)Uout % 16 == 0) &&
                       (!DIFF || (uintptr_t)DIFF % 16 == 0);
        for (int i = 0; i < ns; ++i) aligned = aligned && ((uintptr_t)states[i] % 16 == 0);
        if (!aligned) {
            if (d.kernel == PDX_KERNEL_TMA) { set_error("pdx_fd_step: TMA kernel needs 16-byte aligned arrays"); return 1; }
            kernel = PDX_KERNEL_GENERIC;
        }
    }
    int rc;
    if (kernel == PDX_KERNEL_TMA) {
        const CUtensorMap* mU = nullptr;
        const CUtensorMap* mD = nullptr;
        if (get_map(p, Uin, &mU)) return 1;
        if (d.lap_kind == PDX_LAP_HETEROG) { if (get_map(p, DIFF, &mD)) return 1; } else mD = mU;
        pdx_fd_desc dd = d;
        if (dd.lap_kind == PDX_LAP_CONV) {
            // FAST conv alias of the 7-point kernel.  The reference's conv kernel puts 1/dz^2 on
            // the axis-0 taps and 1/dx^2 on the axis-2 taps (laplace_convolution...py:31-39), so the
            // alias swaps the two spacings to reproduce it for DX != DZ.
            dd.lap_kind = PDX_LAP_HOMOG;
            std::swap(a.hxsq, a.hzsq);
            std::swap(a.rhxsq, a.rhzsq);
        }
        const CUtensorMap* mS[PDX_MAX_STATES] = {nullptr, nullptr, nullptr, nullptr};
        // 3-D: planes of the x-march; 2-D: the y-tiles a CTA marches over (x_chunk > 1) - a single tile per CTA has
        // nothing to pipeline and keeps the register path
        const bool use_tst = p->tma_states && ns > 0 && (a.has_x || a.x_chunk > 1);
        if (use_tst)
            for (int i = 0; i < ns; ++i)
                if (get_state_map(p, states[i], &mS[i])) return 1;
        rc = launch_fd_tma(dd, a, *mU, *mD, use_tst ? mS : nullptr, s);
    } else {
        rc = launch_fd_generic(d, a, s);
    }
    if (!rc) p->launches += 1;
    return rc;
}

int pdx_fd_step(pdx_fd_plan* p, float* U_a, float* U_b, float* const* states, const float* DIFF, int nsteps,
                void* stream) {
    if (!p) { set_error("pdx_fd_step: null plan"); return 1; }
    if (p->resident_ok && p->resident_auto && nsteps >= (p->want_resident ? 1 : 8)) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing((cudaStream_t)stream, &cap);
        if (cap == cudaStreamCaptureStatusNone) {
            // the whole call is ONE launch: U and the gates stay in shared memory for nsteps steps
            const int ns = pdx_model_num_states(p->d.model);
            if (!U_a || !U_b || U_a == U_b) { set_error("pdx_fd_step: U_a/U_b must be distinct non-null buffers"); return 1; }
            if (ns > 0 && !states) { set_error("pdx_fd_step: states required"); return 1; }
            FdArgs a = p->base;
            a.Uin = U_a;
            a.Uout = (nsteps & 1) ? U_b : U_a;
            for (int i = 0; i < ns; ++i) {
                if (!states[i]) { set_error("pdx_fd_step: null state array"); return 1; }
                a.st[i] = states[i];
            }
            const int rc = launch_fd_resident(p->d, a, p->rg, nsteps, p->res_base, p->res_flags, p->res_halo,
                                              (cudaStream_t)stream);
            if (!rc) { p->res_base += (unsigned int)nsteps; p->launches += 1; }
            return rc;
        }
    }
    for (int k = 0; k < nsteps; ++k) {
        const float* in = (k & 1) ? U_b : U_a;
        float* outp = (k & 1) ? U_a : U_b;
        if (step_once(p, in, outp, states, DIFF, p->d.x_begin, p->d.x_end, (cudaStream_t)stream)) return 1;
    }
    return 0;
}

int pdx_fd_step_mirror(pdx_fd_plan* p, const float* U_in, float* U_out, float* const* states, const float* DIFF,
                       float* mirror_lo, float* mirror_hi, void* stream) {
    if (!p) { set_error("pdx_fd_step_mirror: null plan"); return 1; }
    if (p->d.ndim != 3) { set_error("pdx_fd_step_mirror: slab decomposition is 3-D only"); return 1; }
    if (((uintptr_t)mirror_lo | (uintptr_t)mirror_hi) % 16 != 0) { set_error("pdx_fd_step_mirror: mirror planes must be 16-byte aligned"); return 1; }
    return step_once(p, U_in, U_out, states, DIFF, p->d.x_begin, p->d.x_end, (cudaStream_t)stream, mirror_lo, mirror_hi);
}

int pdx_fd_step_range(pdx_fd_plan* p, const float* U_in, float* U_out, float* const* states, const float* DIFF,
                      int xb, int xe, void* stream) {
    if (!p) { set_error("pdx_fd_step_range: null plan"); return 1; }
    if (xb < p->d.x_begin || xe > p->d.x_end) { set_error("pdx_fd_step_range: range outside the plan's planes"); return 1; }
    return step_once(p, U_in, U_out, states, DIFF, xb, xe, (cudaStream_t)stream);
}

int pdx_fd_laplace(int ndim, int nx, int ny, int nz, float dx, float dy, float dz, int lap_kind, int math_mode,
                   const float* X0, const float* DIFF, float* out, void* stream) {
    if (!X0 || !out) { set_error("pdx_fd_laplace: null array"); return 1; }
    if ((ndim != 2 && ndim != 3) || (ndim == 2 && nx != 1) || nx < 1 || ny < 1 || nz < 1) {
        set_error("pdx_fd_laplace: bad shape");
        return 1;
    }
    if (lap_kind == PDX_LAP_HETEROG && !DIFF) { set_error("pdx_fd_laplace: DIFF required"); return 1; }
    if (lap_kind == PDX_LAP_CONV && ndim == 2) { set_error("pdx_fd_laplace: the conv operator is 3-D only"); return 1; }
    if (lap_kind < PDX_LAP_HOMOG || lap_kind > PDX_LAP_HETEROG) { set_error("pdx_fd_laplace: unknown lap_kind (use pdx_fd_laplace_aniso)"); return 1; }
    pdx_fd_desc d;
    pdx_fd_desc_init(&d);
    d.ndim = ndim; d.nx = nx; d.ny = ny; d.nz = nz; d.dx = dx; d.dy = dy; d.dz = dz;
    d.lap_kind = lap_kind; d.math_mode = math_mode; d.model = PDX_MODEL_NONE;
    d.x_begin = 0; d.x_end = nx; d.x_clamp_lo = 0; d.x_clamp_hi = nx - 1;
    FdArgs a;
    fill_args(d, &a);
    a.yclo = 0; a.ychi = ny - 1; a.zclo = 0; a.zchi = nz - 1;   // pure symmetric pad, no enforce_boundary
    a.dxlo = 0; a.dxhi = nx - 1;
    a.Uin = X0; a.Uout = out; a.DIFF = DIFF;
    return launch_fd_laplace_only(lap_kind, math_mode, a, (cudaStream_t)stream);
}

int pdx_fd_laplace_aniso(int nx, int ny, int nz, float dx, float dy, float dz, const float* X0, const float* DIFF0,
                         const float* AVEC0, float* out, void* stream) {
    if (!X0 || !DIFF0 || !AVEC0 || !out || nx < 1 || ny < 1 || nz < 1) { set_error("pdx_fd_laplace_aniso: bad arguments"); return 1; }
    pdx_fd_desc d;
    pdx_fd_desc_init(&d);
    d.ndim = 3; d.nx = nx; d.ny = ny; d.nz = nz; d.dx = dx; d.dy = dy; d.dz = dz;
    d.lap_kind = PDX_LAP_ANISO; d.math_mode = PDX_MATH_EXACT; d.model = PDX_MODEL_NONE;
    d.x_begin = 0; d.x_end = nx; d.x_clamp_lo = 0; d.x_clamp_hi = nx - 1;
    FdArgs a;
    fill_args(d, &a);
    a.yclo = 0; a.ychi = ny - 1; a.zclo = 0; a.zchi = nz - 1;   // pure symmetric pad, no enforce_boundary
    a.dxlo = 0; a.dxhi = nx - 1;
    a.Uin = X0; a.Uout = out; a.DIFF = DIFF0; a.AVEC = AVEC0;
    return launch_fd_laplace_aniso(a, (cudaStream_t)stream);
}

int pdx_ionic_lut_step(const pdx_lut_model* m, int64_t n, const float* U, float* const* states, float* dU_out,
                       const float* I0, float* rhs0_out, void* stream) {
    if (!m || !U || !states || n < 0) { set_error("pdx_ionic_lut_step: bad arguments"); return 1; }
    if (n == 0) return 0;
    LutArgs L;
    if (fill_lut_args(m, &L)) return 1;
    const int ns = pdx_model_num_states(m->model);
    for (int i = 0; i < ns; ++i) {
        if (!states[i]) { set_error("pdx_ionic_lut_step: null state array"); return 1; }
        L.st[i] = states[i];
    }
    return m->math_mode == PDX_MATH_FAST ? launch_lut_ionic_fast(m->model, L, n, U, dU_out, I0, rhs0_out, (cudaStream_t)stream)
                                         : launch_lut_ionic_exact(m->model, L, n, U, dU_out, I0, rhs0_out, (cudaStream_t)stream);
}

int pdx_fd_enforce_boundary(int ndim, int nx, int ny, int nz, int dirichlet, float value, const float* X, float* out,
                            void* stream) {
    if (!X || !out || X == out) { set_error("pdx_fd_enforce_boundary: need distinct non-null arrays"); return 1; }
    if ((ndim != 2 && ndim != 3) || (ndim == 2 && nx != 1) || ny < 3 || nz < 3 || (ndim == 3 && nx < 3)) {
        set_error("pdx_fd_enforce_boundary: bad shape");
        return 1;
    }
    pdx_fd_desc d;
    pdx_fd_desc_init(&d);
    d.ndim = ndim; d.nx = nx; d.ny = ny; d.nz = nz;
    d.x_begin = 0; d.x_end = nx;
    d.x_clamp_lo = ndim == 3 ? 1 : 0; d.x_clamp_hi = ndim == 3 ? nx - 2 : 0;
    d.dirichlet = dirichlet; d.dirichlet_value = value;
    FdArgs a;
    fill_args(d, &a);
    if (ndim == 2) { a.xclo = 0; a.xchi = 0; }
    a.Uin = X; a.Uout = out;
    return launch_enforce_boundary(a, (cudaStream_t)stream);
}

int pdx_ionic_step(int model, int math_mode, const float* params, const float* const* param_arrays, int64_t n,
                   const float* U, float* const* states, float* dU_out, float dt, const float* I0, float* rhs0_out,
                   void* stream) {
    if (!params || !U || !states || n < 0) { set_error("pdx_ionic_step: bad arguments"); return 1; }
    if (n == 0) return 0;
    return launch_ionic(model, math_mode, params, param_arrays, n, U, states, dU_out, dt, I0, rhs0_out,
                        (cudaStream_t)stream);
}

int pdx_stim_apply(float* U, const float* mask, float intensity, int64_t n, void* stream) {
    if (!U || !mask || n < 0) { set_error("pdx_stim_apply: bad arguments"); return 1; }
    if (n == 0) return 0;
    return launch_stim(U, mask, intensity, n, (cudaStream_t)stream);
}

int pdx_stim_apply_max(float* U, const float* mask, float intensity, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!U || !mask))) { set_error("pdx_stim_apply_max: bad arguments"); return 1; }
    if (n == 0) return 0;
    return launch_stim_max(U, mask, intensity, n, (cudaStream_t)stream);
}

}  // extern "C"
