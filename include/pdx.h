/*
 * pdx.h - C ABI of libpdx.so: B200 (sm_100a) kernels for PDEnsorflow's monodomain
 * time-stepping hot path.
 *
 * The reference (cesare-corrado/PDEnsorflow) has no FFI layer: its hot path is reached
 * through three Python contracts (operator functions, the IonicModel protocol, the
 * HeatSolver/MonodomainSolver classes) that bottom out in TensorFlow ops.  The entry
 * points below are what a Python shim for those contracts binds with ctypes; each one
 * cites the reference interface it replaces (paths relative to
 * /root/reference/PDEnsorflow/).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *  - plain C: pointers, sizes, POD structs.  No torch / C++ types.
 *  - all field buffers are DEVICE pointers owned by the caller (fp32, contiguous,
 *    C order, last axis contiguous); the library never frees or retains them beyond
 *    the call (tensor-map caches are keyed by pointer value only).
 *  - every compute call takes a cudaStream_t (passed as void*) and is asynchronous and
 *    stream ordered, so it can be captured into a CUDA graph.
 *  - return value: 0 = ok, non-zero = error; pdx_last_error() gives the message
 *    (thread local).  No CPU fallback exists: without a CUDA device every compute
 *    call fails.
 *  - axis naming: arrays are [nx][ny][nz], nz contiguous.  The reference's
 *    [width, height, depth] maps to nx=width, ny=height, nz=depth (DX,DY,DZ likewise);
 *    2-D fields [W,H] use nx=1, ny=W, nz=H (dy=DX, dz=DY).
 */
#ifndef PDX_H
#define PDX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDX_VERSION 120

/* ionic models: gpuSolve/ionic/{fenton4v,mms2v,ms2v,courtemanche_ramirez_nattel,
 * ten_tusscher_panfilov}.py.  CRN and TTP are the lookup-table models (pdx_lut_model below). */
enum { PDX_MODEL_NONE = 0, PDX_MODEL_FENTON4V = 1, PDX_MODEL_MMS2V = 2, PDX_MODEL_MS2V = 3,
       PDX_MODEL_CRN = 4, PDX_MODEL_TTP = 5 };
/* diffusion operators: gpuSolve/diffop3D/__init__.py:25-28, diffop2D/__init__.py:20-21 */
enum { PDX_LAP_HOMOG = 0, PDX_LAP_CONV = 1, PDX_LAP_HETEROG = 2, PDX_LAP_ANISO = 3 };
/* arithmetic: EXACT = op-for-op fp32 in the reference's evaluation order, IEEE division,
 * no FMA contraction, correctly rounded tanh.  FAST = same formulas with reciprocal
 * multiplies, FMA contraction and ex2/rcp-based tanh (|err| ~1e-7); meets the
 * north-star tolerance (rel-L2 <= 1e-5) and is what bench.py times. */
enum { PDX_MATH_EXACT = 0, PDX_MATH_FAST = 1 };
/* kernel selection for the fused FD step.  RESIDENT: 2-D grids that fit in the shared memory of the
 * SMs (up to ~1.5 M nodes with Fenton-Karma) are advanced by ONE launch per pdx_fd_step call that keeps
 * U and the gates on chip for all nsteps (pdx_fd_resident.cu); homogeneous operator, closed-form
 * models.  AUTO picks it for such grids when a call asks for >= 8 steps outside stream capture. */
enum { PDX_KERNEL_AUTO = 0, PDX_KERNEL_GENERIC = 1, PDX_KERNEL_TMA = 2, PDX_KERNEL_RESIDENT = 3 };

#define PDX_MAX_PARAMS 24
#define PDX_MAX_STATES 4          /* closed-form models (FENTON4V / MMS2V / MS2V) */
#define PDX_LUT_MAX_STATES 24     /* lookup-table models: CRN 19, TTP 22 state arrays */
#define PDX_LUT_MAX_CONSTS 64

/* Parameter vector layouts (fp32, the reference's defaults are returned by
 * pdx_model_default_params):
 *  FENTON4V (fenton4v.py:56-78): tau_vp tau_vn tau_wp tau_wn1 tau_wn2 tau_d tau_si tau_so
 *      tau_a u_c u_w u_0 u_m u_csi u_so r_sp r_sn k_ a_so b_so c_so vmin DV     (23)
 *  MMS2V (mms2v.py:47-55):  tau_in tau_out tau_open tau_close u_gate u_crit vmin DV (8)
 *  MS2V  (ms2v.py:46-53):   tau_in tau_out tau_open tau_close u_gate  -     vmin DV (8)
 *  (DV is the models' _DV attribute = vmax - vmin formed in fp32 at construction.)
 * State arrays: FENTON4V V,W,S (init 1,1,0); MMS2V/MS2V H (init 1). */
/* For CRN / TTP pdx_model_num_params returns the length of pdx_lut_model.consts and
 * pdx_model_num_states 19 / 22; pdx_model_default_params does not apply (the constants are
 * folded from the model's Python attributes by the host, as the reference folds them). */
int pdx_model_num_params(int model);
int pdx_model_num_states(int model);
int pdx_model_default_params(int model, float* out /* PDX_MAX_PARAMS */);

/* Descriptor of one fused explicit FD step:
 *   U1 = (U0 + dt*dU(U)) + coef*lap(U0)       U0 = enforce_boundary(U)
 * replaces the body of the examples' solve(): Tests/FD/Fenton/fenton.py:90-99,
 * Tests/FD/Fenton_sphere/fenton.py:109-115, Tests/FD/HeatEquation/heat.py:86-94
 * = enforce_boundary (fenton.py:46-53) + IonicModel.differentiate
 * (fenton4v.py:168-185 / mms2v.py:95-105 / ms2v.py:90-100) + laplace_* + Euler update.
 *
 * Slab decomposition: the arrays hold nx planes of which [x_begin, x_end) are advanced;
 * plane indices used by the stencil are clamped to [x_clamp_lo, x_clamp_hi].  A
 * stand-alone grid uses x_begin=0, x_end=nx, clamp [1, nx-2]; an interior slab with one
 * ghost plane per side uses x_begin=1, x_end=nx-1, clamp [0, nx-1] (ghost planes are
 * filled by the caller's halo exchange). */
typedef struct pdx_fd_desc {
    int32_t ndim;            /* 2 or 3 */
    int32_t nx, ny, nz;      /* array extents (nx = 1 for 2-D) */
    float   dx, dy, dz;      /* spacings of the nx/ny/nz axes (dx ignored for 2-D) */
    float   dt;              /* f32(dt): multiplies dU and advances the gates */
    float   coef;            /* f32(diff*dt) for HOMOG/CONV, f32(dt) for HETEROG */
    int32_t lap_kind;        /* PDX_LAP_* */
    int32_t model;           /* PDX_MODEL_* (NONE = heat equation) */
    int32_t math_mode;       /* PDX_MATH_* */
    int32_t kernel;          /* PDX_KERNEL_* */
    int32_t x_begin, x_end;  /* planes advanced by a step */
    int32_t x_clamp_lo, x_clamp_hi;
    int32_t dirichlet;       /* 0: symmetric clamp (Neumann); 1: constant boundary
                                (Tests/FD/LaplaceSolver/laplaceSolver.py:44-52) */
    float   dirichlet_value;
    int32_t x_chunk;         /* planes marched per CTA by the TMA kernel; 0 = auto */
    float   params[PDX_MAX_PARAMS];
    int32_t pitch_z;         /* floats between consecutive rows of EVERY array of the plan (U, states, DIFF); 0 = nz
                                (contiguous, the reference's layout).  The tile kernel needs rows that are 16-byte
                                multiples: a grid with nz % 4 != 0 runs on it when the caller keeps its arrays with
                                pitch_z = nz rounded up to a multiple of 4 (pdensorflow_b200.fd.FDStepper does); the
                                pad columns are scratch.  Closed-form models, 2-D / 3-D, not the SM-resident kernel. */
} pdx_fd_desc;

typedef struct pdx_fd_plan pdx_fd_plan;

/* ---- lookup-table ionic models (CourtemancheRamirezNattel, TenTusscherPanfilov) ------------
 * The reference builds three uniform tables per model on the host in NumPy at the model's dt
 * (construct_tables: courtemanche_ramirez_nattel.py:255-391, ten_tusscher_panfilov.py:234-371)
 * and reads them by linear interpolation (_interpolate, :241-252 / :220-231).  The host shim
 * does the same and hands the tables (device, row-major, `stride` floats per row) plus the
 * scalar constants (Python-scalar sub-expressions folded in float64, then rounded to fp32 - TF's
 * rule) to the kernels.  Layouts (columns / constants / state order):
 *  CRN tab[0] V   [-200,200) res 0.1, 32 cols: (A,B) pairs m h j oa oi ua ui xr xs d f w, then GKur
 *                 INaK IbNa GCaL*(V-65) NaCa_a NaCa_b V*GbCa GNa*(V-ENa)
 *      tab[1] Cai [3e-4,30) res 3e-4, 5 cols: Iup_ca buf IpCa_ca conCa fCa_A
 *      tab[2] fn  [-2e-11,1e-10) res 2e-15, 3 cols: u_A v_A v_B
 *      consts: RT/F Ko kACh Gto factorGKur GKr GKs GK1 factorGtr tau_tr factorGrel k_rel
 *              maxIup/maxCaup C_dCa_rel KmCsqn F*Voli C_dCaup C_B1d C_B1e C_Fn1 C_Fn2 B_fCa B_u  (23)
 *      states: Ca_rel Ca_up Cai Ki d f f_Ca h j m oa oi u ua ui v w xr xs             (19)
 *  TTP tab[0] V   [-800,800) res 0.05, 27 cols: (A,B) pairs D F2 F H J M R S Xr1 Xr2 Xs, then a2
 *                 NaCa_A NaCa_B rec_iNaK rec_ipK
 *      tab[1] CaSS [1e-5,10) res 1e-5, 2 cols: FCaSS_A FCaSS_B
 *      tab[2] V-Ek [-800,800) res 0.05, 1 col: rec_iK1
 *      consts: RTONF 0.5*RTONF Cao Ko Ko+pKNa*Nao pKNa Nao GpCa KpCa Fconst F_RT 0.5*F_RT sqrt(Ko/5.4)
 *              GNa pmf_INaK KmNa GbCa GbNa GpK GK1 Vleak Vmaxup Kup^2 Vxfer invVcF*Cm maxsr maxsr-minsr EC
 *              1/(2VcF) Cm Vsr/Vc Bufc*Kbufc Kbufc k4 k2 k1 k3 Vrel Bufsr*Kbufsr Kbufsr Vc/Vss Vsr/Vss
 *              1/(2VssF) Bufss*Kbufss Kbufss                                            (45)
 *      states: GCaL GKr GKs Gto CaSR CaSS Cai F F2 FCaSS H J Ki M Nai R R_bar S Xr1 Xr2 Xs D   (22)
 * Arithmetic is op-for-op fp32 in the reference's order in both modes; math_mode (pdx_lut_model.math_mode
 * for the pointwise step, pdx_fd_desc.math_mode for the fused step) only selects how the 4-5 exp / log /
 * expm1 per node are evaluated: PDX_MATH_EXACT = correctly rounded (fp64 evaluation, bit-comparable with
 * the oracle, compute bound), PDX_MATH_FAST = fp32 expf / logf / expm1f (<= 2 ulp, HBM bound). */
typedef struct pdx_lut_table {
    const float* data;       /* device, nrows x stride */
    int32_t nrows, ncols, stride;
    float   mn, mx, res, step;   /* f32(mn), f32(mx), f32(res), f32(1.0/res) */
    int32_t mx_idx;              /* int((mx-mn)*step) - 1 (float64, as the reference computes it) */
} pdx_lut_table;

typedef struct pdx_lut_model {
    int32_t model;               /* PDX_MODEL_CRN / PDX_MODEL_TTP */
    int32_t nconsts;
    float   dt;                  /* f32(dt) of the forward-Euler concentration updates */
    pdx_lut_table tab[3];
    float   consts[PDX_LUT_MAX_CONSTS];
    int32_t math_mode;           /* PDX_MATH_EXACT (0, default) / PDX_MATH_FAST: pointwise step only */
} pdx_lut_model;

/* IonicModel.differentiate for CRN / TTP (courtemanche_ramirez_nattel.py:423-514,
 * ten_tusscher_panfilov.py:405-523): writes dU (may be NULL), advances the states in place;
 * if rhs0_out != NULL also writes U + dt*(dU + I0) (monodomainSolver.py:100-103). */
int pdx_ionic_lut_step(const pdx_lut_model* m, int64_t n, const float* U, float* const* states,
                       float* dU_out, const float* I0, float* rhs0_out, void* stream);
/* attach the tables/constants to a plan whose descriptor names PDX_MODEL_CRN / PDX_MODEL_TTP
 * (required before pdx_fd_step; states[] then holds 19 / 22 arrays) */
int pdx_fd_plan_set_lut(pdx_fd_plan* p, const pdx_lut_model* m);

void pdx_fd_desc_init(pdx_fd_desc* d);                      /* zero + defaults */
int  pdx_fd_plan_create(const pdx_fd_desc* d, pdx_fd_plan** out);
void pdx_fd_plan_destroy(pdx_fd_plan* p);
/* which kernel a plan resolved to for single steps (PDX_KERNEL_GENERIC / PDX_KERNEL_TMA /
 * PDX_KERNEL_RESIDENT) */
int  pdx_fd_plan_kernel(const pdx_fd_plan* p);
/* 1 when multi-step calls on this plan can use the SM-resident kernel */
int  pdx_fd_plan_resident_capable(const pdx_fd_plan* p);
/* number of kernel launches issued through this plan so far */
int64_t pdx_fd_plan_launches(const pdx_fd_plan* p);

/* nsteps fused steps, ping-ponging between U_a and U_b: step k reads (k even ? U_a : U_b)
 * and writes the other; the result is in U_b when nsteps is odd, in U_a when even.
 * states[]: the model's state arrays, updated in place.  DIFF: diffusivity field for
 * PDX_LAP_HETEROG (else NULL).  All arrays have the descriptor's [nx][ny][nz] shape.
 * PDX_LAP_ANISO: DIFF is [nx][ny][nz][2] (sigma_t, sigma_l - sigma_t) and the fibre dyadic
 * [nx][ny][nz][6] is attached with pdx_fd_plan_set_avec. */
int pdx_fd_step(pdx_fd_plan* p, float* U_a, float* U_b, float* const* states,
                const float* DIFF, int nsteps, void* stream);
int pdx_fd_plan_set_avec(pdx_fd_plan* p, const float* AVEC);
/* PDX_LAP_ANISO on the TMA tile kernel (pdx_fd_aniso.cu).  laplace_heterogeneous_anisotropic_diffusion.py:40-45 forms
 * C = sigma_t I + (sigma_l - sigma_t) a (x) a from DIFF / AVEC at every call; the fields are constants of a run, so
 * pdx_fd_plan_prepare_aniso forms the six components once (same fp32 operations) into plan-owned [nx][ny][nz] arrays
 * (stream ordered) and attaches AVEC.  Following pdx_fd_step* calls that pass the SAME DIFF pointer run the tile
 * kernel; any other DIFF pointer, or a plan for which pdx_fd_plan_aniso_tile_capable() is 0 (nz % 4 != 0, Dirichlet
 * pad, lookup-table model, kernel = GENERIC), keeps the one-thread-per-node kernel.  Call it again after changing
 * the contents of DIFF or AVEC. */
int pdx_fd_plan_aniso_tile_capable(const pdx_fd_plan* p);
int pdx_fd_plan_prepare_aniso(pdx_fd_plan* p, const float* DIFF, const float* AVEC, void* stream);
/* one step restricted to planes [xb, xe) (boundary-first / interior split used to
 * overlap the halo exchange); same clamps as the plan. */
int pdx_fd_step_range(pdx_fd_plan* p, const float* U_in, float* U_out, float* const* states,
                      const float* DIFF, int xb, int xe, void* stream);

/* One full step of a slab with the halo exchange FUSED into it: besides U_out, the new values of
 * the slab's first owned plane (x_begin) are stored to mirror_lo and those of its last owned plane
 * (x_end - 1) to mirror_hi - plane-sized ([ny][nz]) buffers that are the neighbouring ranks' ghost
 * planes, mapped into this process as NVLink peer memory (CUDA IPC / symmetric memory).  Either
 * may be NULL (global boundary).  The caller orders steps across ranks with one barrier per step
 * (pdensorflow_b200.fd.DistributedFDStepper, halo='peer').  New functionality: the reference has
 * no multi-GPU path. */
int pdx_fd_step_mirror(pdx_fd_plan* p, const float* U_in, float* U_out, float* const* states,
                       const float* DIFF, float* mirror_lo, float* mirror_hi, void* stream);

/* Stand-alone operators (pure functions, new output array) - contract 1:
 * gpuSolve/diffop3D/laplace_homogeneous_isotropic_diffusion.py:3-52,
 * laplace_convolution_homogeneous_isotropic_diffusion.py:4-47,
 * laplace_heterogeneous_isotropic_diffusion.py:3-35, and the diffop2D analogues
 * (diffop2D/laplace_homogeneous_isotropic_diffusion.py:3-36,
 *  laplace_heterogeneous_isotropic_diffusion.py:3-32).  Symmetric pad of X0 (and DIFF). */
int pdx_fd_laplace(int ndim, int nx, int ny, int nz, float dx, float dy, float dz,
                   int lap_kind, int math_mode, const float* X0, const float* DIFF,
                   float* out, void* stream);
/* laplace_heterogeneous_anisotropic_diffusion.py:3-95: DIFF0 [nx][ny][nz][2], AVEC0 [nx][ny][nz][6] */
int pdx_fd_laplace_aniso(int nx, int ny, int nz, float dx, float dy, float dz, const float* X0,
                         const float* DIFF0, const float* AVEC0, float* out, void* stream);
/* enforce_boundary: Tests/FD/Fenton/fenton.py:46-53 (dirichlet=0) /
 * LaplaceSolver/laplaceSolver.py:44-52 (dirichlet=1) */
int pdx_fd_enforce_boundary(int ndim, int nx, int ny, int nz, int dirichlet, float value,
                            const float* X, float* out, void* stream);

/* IonicModel.differentiate(U) - contract 2 (fenton4v.py:168-185, mms2v.py:95-105,
 * ms2v.py:90-100): writes dU, advances states in place by dt.  param_arrays (may be
 * NULL): per-parameter device pointers to per-node values installed through
 * set_parameter (ionicmodel.py:56-62); NULL entries use params[i].
 * If U1_out != NULL the kernel also writes rhs0 = U + dt*(dU + I0) (I0 may be NULL):
 * the RHS0 of MonodomainSolver.solve_step (monodomainSolver.py:100-103). */
int pdx_ionic_step(int model, int math_mode, const float* params,
                   const float* const* param_arrays, int64_t n, const float* U,
                   float* const* states, float* dU_out, float dt,
                   const float* I0, float* rhs0_out, void* stream);

/* Stimulus application of the FD examples (Tests/FD/Fenton/fenton.py:154-156):
 * U = mask ? max(U, intensity*mask) : U, mask as produced by Stimulus.set_stimregion
 * (force_terms/stimulus.py:73-80). */
int pdx_stim_apply(float* U, const float* mask, float intensity, int64_t n, void* stream);
/* The heat examples' form (Tests/FD/HeatEquation/heat.py:143-144): U = maximum(U, intensity*mask) on EVERY node. */
int pdx_stim_apply_max(float* U, const float* mask, float intensity, int64_t n, void* stream);

/* ------------------------------------------------------------------ FEM path */
/* y = A x, CSR fp32 / int32 (replaces tf.raw_ops.SparseMatrixMatMul with a [N,1]
 * operand: linearsolvers/conjgrad.py:99-101, monodomainSolver.py:104, heatSolver.py:132) */
int pdx_csr_spmv(int32_t n, const int32_t* rowptr, const int32_t* col, const float* val,
                 const float* x, float* y, void* stream);

/* Jacobi-preconditioned CG - contract 3 (linearsolvers/conjgrad.py:172-203,257-312;
 * jacobi_precond.py:19-41).  One persistent cooperative kernel runs the whole solve:
 * r=b-Ax; z=Minv*r; p=z; loop {q=Ap; alpha=rz/(p.q); x+=alpha p; r-=alpha q; res=r.r;
 * z=Minv*r; rz'=r.z; beta=rz'/rz; p=z+beta p; every check_every its: stop if !finite(res)
 * or res < toll^2}.  x holds X0 on entry and the solution on exit.
 * info (device, 2 floats + 2 ints packed as 4 x 4 bytes): [0] int niters, [1] float res,
 * [2] int flags (bit0 = maxiter reached, bit1 = non-finite), [3] reserved. */
typedef struct pdx_pcg pdx_pcg;
int  pdx_pcg_create(int32_t n, const int32_t* rowptr, const int32_t* col, const float* val,
                    const float* minv /* may be NULL: no preconditioner */, pdx_pcg** out);
void pdx_pcg_destroy(pdx_pcg* s);
int  pdx_pcg_solve(pdx_pcg* s, const float* b, float* x, int maxiter, double toll,
                   int check_every, int32_t* info /* device, 4 words */, void* stream);
/* The same solve with the right-hand side of MonodomainSolver.solve_step / HeatSolver.solve_step formed inside
 * the kernel: b = MASS * rhs0 (monodomainSolver.py:104, heatSolver.py:132), mval = MASS values on the system
 * matrix' pattern.  One launch and one vector round trip less per time step. */
int  pdx_pcg_solve_rhs0(pdx_pcg* s, const float* mval, const float* rhs0, float* x, int maxiter, double toll,
                        int check_every, int32_t* info /* device, 4 words */, void* stream);

/* Optional: tell the library the size of a CSR matrix once, when it is uploaded (the kernels pick their lanes-per-row
 * from nnz / n; without the hint pdx_fem_explicit_step uses 8 lanes - it never reads rowptr back). */
int pdx_csr_hint(const int32_t* rowptr, int32_t n, int64_t nnz);
/* Registers the widest slice (entries per row) of a SELL-32 matrix by its slice_ptr address (0 forgets it).  With it
 * pdx_fem_explicit_step_sell / _sync run blocks of 10^6 rows and more on a persistent kernel that streams the slices
 * through shared memory with the copy engine; without it (or for small blocks) the one-thread-per-row kernel runs.
 * No device read at step time, so steps stay capturable. */
int pdx_sell_hint(const int32_t* slice_ptr, int32_t max_width);

/* Explicit mass-lumped FEM step named by the north star (no reference implementation;
 * SURVEY.md section 8a row E): U1 = U + dt*(dU + I0) - dt * mlinv * (K U), fused with the
 * ionic update (row-per-warp CSR). */
int pdx_fem_explicit_step(int model, int math_mode, const float* params, int32_t n,
                          const int32_t* rowptr, const int32_t* col, const float* kval,
                          const float* mlinv, const float* U_in, float* U_out,
                          float* const* states, const float* I0, float dt, void* stream);

/* Row-block partitioned form of the same step (config 5: node-block partition with ghost-node
 * exchange).  This rank owns global rows [row_lo, row_lo + n_local); rowptr is local, col holds
 * GLOBAL column indices; U_in / U_out are full-length (global) buffers of which the rank's own rows
 * and its ghost columns are valid; states, mlinv and I0 are local (n_local).  The new values of the
 * first send_lo and last send_hi local rows are also stored at their global index into mirror_lo /
 * mirror_hi = the neighbouring ranks' U_out buffers mapped as NVLink peer memory, i.e. the ghost
 * exchange is fused into the step; the caller orders steps with one barrier per step
 * (pdensorflow_b200.fem_dist.PartitionedExplicitFEM). */
int pdx_fem_explicit_step_part(int model, int math_mode, const float* params, int32_t n_local, int32_t row_lo,
                               const int32_t* rowptr, const int32_t* col, const float* kval, const float* mlinv,
                               const float* U_in, float* U_out, float* const* states, const float* I0, float dt,
                               int32_t send_lo, float* mirror_lo, int32_t send_hi, float* mirror_hi, void* stream);

/* The same partitioned step on a SELL-32 matrix (sliced ELLPACK, slice height 32 = one warp, one
 * thread per row): slice_ptr int32 [ceil(n_local/32) + 1] gives each slice's offset into col / kval,
 * entry k of row 32s + l sits at slice_ptr[s] + 32k + l, short rows are padded with (col = any valid
 * global index, val = 0).  Every matrix access is a coalesced 128-byte warp load and every lane runs
 * the ionic update; this is the layout pdensorflow_b200.fem_dist uses for config 5. */
int pdx_fem_explicit_step_sell(int model, int math_mode, const float* params, int32_t n_local, int32_t row_lo,
                               const int32_t* slice_ptr, const int32_t* col, const float* kval, const float* mlinv,
                               const float* U_in, float* U_out, float* const* states, const float* I0, float dt,
                               int32_t send_lo, float* mirror_lo, int32_t send_hi, float* mirror_hi, void* stream);

/* The partitioned step (sell = 0: CSR as _part, 1: SELL-32 as _sell) ordered by NEIGHBOUR FLAGS instead of a barrier
 * between steps.  ctrl: this rank's control block, uint32[8] in memory the neighbours can write (NVLink peer /
 * symmetric memory), zeroed once before the first step with every rank's initial U in place; ctrl_lo / ctrl_hi: the
 * neighbours' blocks (NULL at the ends of the chain).  Only the CTAs that hold the first send_lo / last send_hi rows
 * take part (they read the ghost rows and store into the neighbour): they wait until that neighbour has published as
 * many steps as this rank (its ghost rows are in U_in, and it no longer reads the buffer this step overwrites), and
 * the last of them publishes the step into the neighbour's block.  One GPU per rank (or all ranks on one stream): the
 * waiting CTAs occupy their SMs. */
int pdx_fem_explicit_step_sync(int sell, int model, int math_mode, const float* params, int32_t n_local, int32_t row_lo,
                               const int32_t* rowptr, const int32_t* col, const float* kval, const float* mlinv,
                               const float* U_in, float* U_out, float* const* states, const float* I0, float dt,
                               int32_t send_lo, float* mirror_lo, int32_t send_hi, float* mirror_hi, uint32_t* ctrl,
                               uint32_t* ctrl_lo, uint32_t* ctrl_hi, void* stream);

/* ------------------------------------------------------------------ FEM path across GPUs */
/* Row-block partitioned Jacobi-PCG = the semi-implicit step of MonodomainSolver / HeatSolver
 * (monodomainSolver.py:98-108, heatSolver.py:127-135, conjgrad.py:172-203,257-312) with the mesh nodes
 * split into contiguous row blocks, one per GPU (the reference has no multi-GPU
 * path).  One persistent cooperative kernel per rank per solve; the ranks communicate from inside the
 * kernel through NVLink peer memory only (8-byte (epoch,value) words for the two dot products of an
 * iteration, z = Minv*r of the edge rows for the ghosts) - see pdx_fem_part.cu.
 *
 * Vectors multiplied by the matrix are WINDOWS of ghost_lo + n_local + ghost_hi floats: the ghost_lo
 * rows owned by rank-1 just below this rank's first row, the owned rows, the ghost_hi rows of rank+1
 * just above.  col[] holds window-local indices.  send_lo / send_hi = how many of this rank's first /
 * last rows the lower / upper neighbour holds as ghosts (= that neighbour's ghost_hi / ghost_lo);
 * lower_window_index = the index of this rank's row 0 in the lower neighbour's window
 * (its ghost_lo + n_local), upper_window_index = the index of this rank's row n_local - send_hi in the
 * upper neighbour's window (its ghost_lo - send_hi).
 *
 * Workspace: every rank allocates pdx_pcg_part_workspace_bytes(max_window) bytes (max_window = the
 * largest window over the ranks, so that the layout is the same everywhere) of peer-accessible device
 * memory (symmetric memory / CUDA IPC / plain device memory for ranks that share a GPU) and passes all
 * `world` base pointers, as mapped in THIS process, to create.  It holds what the neighbours write:
 * the exchange words and the p / z windows.  x, b and rhs0 are ordinary caller-owned device arrays.
 * All ranks must have returned from create before any rank calls solve (create clears the block).
 * world = 1 with plain device memory is the single-GPU solver for large systems (SELL-32 storage,
 * gpuSolve.linearsolvers.ConjGrad uses it above 10^5 rows). */
#define PDX_PART_MAX_WORLD 16
typedef struct pdx_pcg_part_desc {
    int32_t n_local, ghost_lo, ghost_hi;
    int32_t rank, world;
    int32_t send_lo, send_hi;
    int32_t lower_window_index, upper_window_index;
    int32_t max_blocks;          /* cap on the cooperative grid (0 = auto); ranks sharing one GPU must fit together */
    int32_t timeout_ms;          /* a rank waiting longer than this for a peer gives up (0 = 10000) */
    int32_t sell;                /* 0: CSR.  1: SELL-32 as in pdx_fem_explicit_step_sell - rowptr is slice_ptr
                                    [ceil(n_local/32) + 1], col / aval / mval are slice-wise column-major, rows
                                    padded with (col = any valid window index, val = 0) */
    const int32_t* rowptr;       /* n_local + 1 */
    const int32_t* col;          /* window-local */
    const float*   aval;         /* A = MASS + dt*STIFFNESS, rows of this rank */
    const float*   mval;         /* MASS on the same pattern, or NULL */
    const float*   minv;         /* 1/diag(A) of the owned rows, or NULL */
} pdx_pcg_part_desc;
typedef struct pdx_pcg_part pdx_pcg_part;
int64_t pdx_pcg_part_workspace_bytes(int32_t max_window);
int     pdx_pcg_part_create(const pdx_pcg_part_desc* d, void* const* peer_workspaces /* world */,
                            int32_t max_window, pdx_pcg_part** out);
void    pdx_pcg_part_destroy(pdx_pcg_part* s);
int32_t pdx_pcg_part_blocks(const pdx_pcg_part* s);
int32_t pdx_pcg_part_threads(const pdx_pcg_part* s);        /* threads per block of the cooperative grid */
int32_t pdx_pcg_part_stage_bytes(const pdx_pcg_part* s);    /* > 0: q = A p streams the SELL slices through a shared-memory
                                                               ring of this slot size (copy engine), 0: per-thread loads */
/* Diagnostics: when device_stamps != NULL, thread 0 of every following solve writes %globaltimer (ns) at each phase
 * boundary: [0] end of the set-up product, [1] after its all-reduce, then per iteration: end of q = A p, after the
 * all-reduce, end of the x / r update, after the all-reduce, end of the p update, after the closing grid barrier. */
int     pdx_pcg_part_set_profile(pdx_pcg_part* s, uint64_t* device_stamps, int32_t nstamps);
/* Solve A x = b.  x: WINDOW (ghost_lo + n_local + ghost_hi floats): guess on entry - ghosts included,
 * they must equal the owners' values - solution on exit, ghosts kept equal to the owners' values by the
 * kernel.  b: n_local values, or NULL to use b = MASS * rhs0 with the rhs0 WINDOW (ghosts included:
 * fem_dist runs the ionic update over the window).  Every rank must call it with the same maxiter /
 * toll / check_every.  info as pdx_pcg_solve, plus flag bit2 = a peer did not answer within
 * timeout_ms (fatal for the handle). */
int     pdx_pcg_part_solve(pdx_pcg_part* s, const float* b, const float* rhs0, float* x, int maxiter,
                           double toll, int check_every, int32_t* info /* device, 4 words */, void* stream);

/* ------------------------------------------------------------------ FEM assembly on the device */
/* P1 mass / stiffness assembly (gpuSolve/matrices/globalMatrices.py:82-287, localMass.py,
 * localStiffness.py): one thread per element evaluates the contravariant basis, the measure and the
 * nv x nv local matrices in fp64 and adds them (fp64 atomics) into CSR-ordered accumulators at the
 * positions of (row, col) in the given pattern (columns sorted inside each row).
 *  nv 2 / 3 / 4 = Edges / Trias / Tetras; elems [ne][elem_stride] int32 holds the nv vertex ids first
 *  (the reference keeps the region id in the last column); pts fp64 [npt][3]; rowid (or NULL) maps a
 *  vertex id to its matrix row (RCM renumbering: iperm); sigma fp64 [ne][6] = xx xy xz yy yz zz of the
 *  diffusion tensor (NULL = identity).  mass_acc / stiff_acc: fp64 [nnz], zeroed by the caller, either may
 *  be NULL.  *status (device int32) is set to 1 if an entry is missing from the pattern.
 * pdx_round_to_f32 turns the accumulators into the fp32 value arrays the solvers use. */
int pdx_fem_assemble(int nv, int64_t ne, const int32_t* elems, int elem_stride, const double* pts,
                     const int32_t* rowid, const double* sigma, const int32_t* rowptr, const int32_t* col,
                     double* mass_acc, double* stiff_acc, int32_t* status, void* stream);
int pdx_round_to_f32(const double* in, float* out, int64_t n, void* stream);

/* ---- matrix-format and ghost plumbing on the device (the reference does all of it on the
 * host: gpuSolve/matrices/globalMatrices.py:54-74,612-675, and has no multi-GPU path) -------------------------------
 * CSR -> SELL-32 (the layout of pdx_fem_explicit_step_sell / pdx_pcg_part with sell = 1): pdx_sell_slice_widths writes
 * the longest row of every 32-row slice; the caller forms slice_ptr = exclusive scan of 32 * width (int32) and
 * allocates slice_ptr[nslices] entries; pdx_sell_fill scatters the rows column-major into their slices and pads with
 * (col = row_lo + own row, val = 0).  rowptr is local (n + 1), col is copied verbatim. */
int pdx_sell_slice_widths(int32_t n, const int32_t* rowptr, int32_t* widths /* ceil(n / 32) */, void* stream);
int pdx_sell_fill(int32_t n, int32_t row_lo, const int32_t* rowptr, const int32_t* col, const float* val,
                  const int32_t* slice_ptr, int32_t* scol, float* sval, void* stream);
/* General ghost exchange of a partitioned step: peer_bufs[peer[i]][rows[i]] = src[rows[i]] for i < count.  peer_bufs
 * is a DEVICE array of the ranks' full-length buffers (NVLink peer mappings); rows are global indices. */
int pdx_push_rows(const float* src, int32_t count, const int32_t* rows, const int32_t* peer, float* const* peer_bufs,
                  void* stream);

/* ------------------------------------------------------------------ misc */
int         pdx_version(void);
const char* pdx_last_error(void);
/* total kernel launches issued by this library in this process */
int64_t     pdx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PDX_H */
