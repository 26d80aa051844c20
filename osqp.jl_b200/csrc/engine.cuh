// engine.cuh -- data layout in HBM and host<->kernel contracts of the B200 OSQP engine.
//
// Replaces libosqp 0.6.2 behind the reference's ccall boundary (src/interface.jl:146-715).
// Everything the ADMM loop touches is resident in HBM for the lifetime of the workspace:
//   * A   (m x n) as CSR  : row-parallel  t = A v
//   * A'  (n x m) as CSR  : = the caller's CSC arrays verbatim; row-parallel  A' w
//   * P   (n x n) as full symmetric CSR (the caller passes the upper triangle only,
//         src/interface.jl:102-104)
//   fp64 values + int32 column indices (12 B / nnz), int32 row pointers.
//   "val0" are the unscaled values as passed (kept so osqp_update_P/A can re-equilibrate).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace osqpb200 {

constexpr int kMaxBlocks = 1184;      // 148 SMs x 8
constexpr int kRedSlots = 32;         // scalars reduced per grid barrier
constexpr int kLogRows = 64;          // verbose table rows buffered per solve

struct CsrDev {
  int rows = 0, cols = 0;
  long long nnz = 0;
  int *rowptr = nullptr;   // rows + 1
  int *col = nullptr;      // nnz
  double *val = nullptr;   // nnz, scaled working copy
  double *val0 = nullptr;  // nnz, as passed by the caller
  int lanes = 32;          // lanes cooperating on one row (power of two <= 32)
};

// Column-blocked copy of a CSR matrix for the hot SpMV phases: the dense input vector is staged
// one block of W columns at a time in shared memory (TMA bulk copy) and gathered from there, so
// column indices are block-local 16-bit values (10 B / nnz instead of 12) and no random 8-byte
// gather ever reaches L2.  Block cb holds, row by row, the entries with cb*W <= col < (cb+1)*W.
struct BlkDev {
  int nb = 0;                    // number of column blocks
  int W = 0;                     // columns per block (multiple of 32, W*8 bytes fit in shared memory)
  int rows = 0, cols = 0;
  int lanes = 8;                 // threads cooperating on one row segment
  int *rowptr = nullptr;         // [nb][rows+1] absolute positions into col/val (block-major storage)
  unsigned short *col = nullptr; // column index local to the block
  double *val = nullptr;         // scaled values
  int *from_csr = nullptr;       // CSR position -> blocked position (value refresh after re-scaling)
  // element span of the rows a thread block owns inside column block cb: [span[cb*grid+b], span[cb*grid+b+1])
  // (owner passes of A and P); used to bulk-prefetch the next pass into L2
  int *span = nullptr;
};

// Persistent solver state that survives between launches (device memory).
struct DevState {
  double rho;              // current scalar rho (settings->rho)
  double c, cinv;          // cost scaling
  long long rho_updates;
  long long adaptive_interval;  // 0 until fixed (settings->adaptive_rho_interval)
  int needs_refresh;       // PCG residual/z_tilde recurrences invalid (matrix, rho or iterate change)
  int pd_check_failed;     // setup-time curvature probe found p'(P+sigma I)p <= 0
  int ctype_changed;       // update_rho_vec saw a constraint type change
  int pad;
};

// What one osqp_solve launch reports back (device -> pinned host).
struct DevInfo {
  long long iter;
  long long status_val;
  double obj_val, pri_res, dua_res;
  double rho_estimate;
  long long rho_updates;
  double rho;                   // rho at exit
  long long adaptive_interval;  // interval at exit
  long long cg_iters;           // total PCG iterations in this solve
  long long cg_solves;          // ADMM iterations that ran the PCG
  long long checks;             // update_info evaluations
  long long refreshes;          // full residual / z_tilde rebuilds
  double elapsed_s;             // device-side wall time of the loop (globaltimer)
  long long log_rows;
  double log[kLogRows][6];      // iter, obj, pri_res, dua_res, rho, time
};

struct DevPtrs {
  int n = 0, m = 0;
  CsrDev A, At, P;
  // problem vectors (scaled working copies + originals as passed)
  double *q = nullptr, *l = nullptr, *u = nullptr;
  double *q0 = nullptr, *l0 = nullptr, *u0 = nullptr;
  double *Pdiag = nullptr;              // diag of scaled P
  double *rho_vec = nullptr, *rho_inv = nullptr;
  int *ctype = nullptr;
  double *D = nullptr, *Dinv = nullptr, *E = nullptr, *Einv = nullptr;
  double *dtmp = nullptr, *etmp = nullptr;
  // ADMM iterates
  double *x = nullptr, *z = nullptr, *y = nullptr;
  double *xt = nullptr, *zt = nullptr;  // x_tilde (PCG iterate, kept as warm start), z_tilde = A x_tilde
  double *dx = nullptr, *dy = nullptr;  // last step; after an infeasible exit: the certificate
  double *wv = nullptr;                 // m: rho .* z - y
  // PCG
  double *r = nullptr, *b = nullptr, *uu = nullptr, *p = nullptr, *s = nullptr, *w = nullptr, *Minv = nullptr;  // n
  double *t = nullptr, *tr = nullptr, *Ap = nullptr;                                                         // m
  // polish scratch
  double *pol_x = nullptr, *pol_rhs = nullptr;   // n
  double *pol_y = nullptr, *pol_z = nullptr, *pol_rho = nullptr, *pol_b = nullptr;  // m
  // results
  double *sol_x = nullptr, *sol_y = nullptr;
  // column-blocked copies for the hot phases (blocked == 0: fall back to the CSR + L1 gather path)
  int blocked = 0;
  BlkDev Ab, Pb, Atb;
  double *Pu = nullptr;          // n: P u of the current PCG iteration
  double *partAt = nullptr;      // [Atb.nb][n] partial sums of A' w per column block
  int at_ntiles = 0;             // A' tiles (column block x row range), tile t is processed by block t % grid
  int *at_tile_cb = nullptr, *at_tile_r0 = nullptr, *at_tile_r1 = nullptr;
  int *at_tile_lo = nullptr, *at_tile_hi = nullptr;  // element span of each tile in Atb.val / Atb.col
  int smem_x_elems = 0;          // doubles of dynamic shared memory for the staged vector tile
  int smem_rows = 0;             // doubles of dynamic shared memory for per-row running sums
  // work partition: block b owns rows [m_start[b], m_start[b+1]) of A and [n_start[b], n_start[b+1]) of P/A'
  int *m_start = nullptr, *n_start = nullptr;
  // grid barrier + reductions
  unsigned *bar = nullptr;  // [0] arrival count, [1] generation
  double *red = nullptr;    // [2][kRedSlots][grid]
  DevState *state = nullptr;
  DevInfo *info = nullptr;
  unsigned long long *dbg = nullptr;  // [grid][16] globaltimer probes (spmv_blk_kernel only)
};

struct SolveCfg {
  double sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  long long max_iter;
  long long check_termination;
  int scaling;             // settings->scaling != 0
  int scaled_termination;
  int adaptive_rho;
  double adaptive_rho_tolerance;
  double adaptive_time_s;  // adaptive_rho_fraction * setup_time (automatic interval)
  int warm_start;
  int verbose;
  double time_limit_s;     // <= 0: none; budget left for this solve (base time already subtracted)
  // PCG controls (engine-specific; include/osqp_b200.h)
  double pcg_eta;          // PCG stops at |r|inf <= max(pcg_eta * |r0|inf, pcg_floor * |b|inf)
  double pcg_floor;
  int pcg_max_iter;
  int refresh_every;       // recompute z_tilde = A x_tilde and r = b - K x_tilde every k ADMM iterations (1 = always)
};

struct PolishCfg {
  double delta, penalty;
  int refine_iter;         // outer multiplier steps
  double pcg_rel_tol;
  int pcg_max_iter;
  int scaling, scaled_termination;
};

struct PolishOut {
  long long n_active;
  double obj_val, pri_res, dua_res;
  long long cg_iters;
  int success;
  int pad;
};

struct LaunchGeom {
  int grid = 1, block = 1024;
  size_t dyn_smem = 0;
};

// ---- host wrappers implemented in kernels.cu (all asynchronous on `st`)
cudaError_t launch_scale_data(const DevPtrs &d, int scaling_iters, double sigma, cudaStream_t st);
cudaError_t launch_scale_vectors(const DevPtrs &d, int do_q, int do_bounds, cudaStream_t st);
cudaError_t launch_set_rho_vec(const DevPtrs &d, double rho, int detect_change_only, cudaStream_t st);
cudaError_t launch_apply_rho(const DevPtrs &d, double rho, cudaStream_t st);
cudaError_t launch_precond(const DevPtrs &d, double sigma, cudaStream_t st);
cudaError_t launch_pd_probe(const DevPtrs &d, LaunchGeom g, double sigma, int max_it, cudaStream_t st);
cudaError_t launch_warm_start(const DevPtrs &d, const double *x_in, const double *y_in, int scaling, cudaStream_t st);
cudaError_t launch_cold_start(const DevPtrs &d, cudaStream_t st);
cudaError_t launch_scatter_values(double *dst, const double *vals, const long long *idx, const int *map,
                                  long long k, cudaStream_t st);
cudaError_t launch_solve(const DevPtrs &d, const SolveCfg &cfg, LaunchGeom g, cudaStream_t st);
cudaError_t launch_polish(const DevPtrs &d, const PolishCfg &cfg, const SolveCfg &sc, PolishOut *out, LaunchGeom g,
                          cudaStream_t st);
cudaError_t launch_spmv(const DevPtrs &d, int which, const double *in, double *out, double sigma, LaunchGeom g,
                        cudaStream_t st);
int max_coop_blocks_per_sm(int block, size_t dyn_smem);
cudaError_t configure_dyn_smem(size_t dyn_smem);
cudaError_t launch_fill_blocked(const DevPtrs &d, cudaStream_t st);

}  // namespace osqpb200
