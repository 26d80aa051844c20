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
constexpr int kFxSlots = 8;           // fixed-point accumulators per bank (the last one is the overflow flag)
constexpr int kBarBytes = 16 + 3 * kFxSlots * 8;  // DevPtrs::bar: arrival counter, then 3 banks of accumulators
constexpr int kLogRows = 64;          // verbose table rows buffered per solve
constexpr int kPhases = 16;           // phase classes timed by block 0 (include/osqp_b200.h OSQPB200Profile.phase_us)

struct CsrDev {
  int rows = 0, cols = 0;
  long long nnz = 0;
  int *rowptr = nullptr;   // rows + 1
  int *col = nullptr;      // nnz
  double *val = nullptr;   // nnz, scaled working copy
  double *val0 = nullptr;  // nnz, as passed by the caller
  int lanes = 32;          // lanes cooperating on one row (power of two <= 32)
};

// Tile stream: the hot-path storage of a (row-stacked) sparse matrix M (rows x cols) that is multiplied with a
// dense vector of length `cols` once per phase (S_A = [A; P] against an n-vector, S_T = A' against an m-vector).
//
//   * 2-D split.  The columns are cut into `ngroups` groups; every thread block belongs to ONE group and stages
//     only that group's slice of the dense vector (<= kSliceMax doubles) in shared memory with TMA bulk copies, so
//     the vector is pulled from L2 once per phase per SM.
//   * Inside a group the rows are cut into contiguous, nnz-balanced ranges, one per warp (kWarps per block).  A
//     warp's entries are stored contiguously, row by row.  There is no row pointer: an entry is (fp64 value, u16
//     word = column local to the group slice); every row segment is padded with zero entries to a multiple of 4 (a
//     "quad"), and bit 15 of the LAST word of a quad marks the end of a row.  Each lane streams whole quads with two
//     16 B value loads and one 8 B column load: 10 B per stored entry.  Inside a chunk (32 quads) the values are
//     interleaved so that each of the two value loads of a warp covers one contiguous 512 B row (whole sectors).
//   * A lane sums its quad; row sums are formed with a warp-segmented scan over the per-lane flags (carry across
//     chunks) and written straight to part[group][row].  The owner blocks add the `ngroups` partial vectors in the
//     element-wise phase that follows the next grid barrier (fixed order -> run-to-run deterministic).
#ifndef OSQP_B200_WARPS
#define OSQP_B200_WARPS 16
#endif
constexpr int kWarps = OSQP_B200_WARPS;  // warps per thread block of the cooperative kernels (kThreads / 32)
constexpr int kSliceMax = 27648;    // doubles per staged slice (216 KB); local columns are 15-bit
constexpr int kSplitQuads = 64;     // smallest piece a long row is cut into (build_tile_stream: stream rows)

struct TileStreamDev {
  int rows = 0, cols = 0;        // stacked rows, length of the gathered vector
  int ngroups = 0;
  int pf_chunks = 0;             // chunks (32 quads = 1280 B) a warp prefetches into L2 ahead of its register loads
  long long nelem = 0;           // padded stream length in entries (multiple of 4)
  double *val = nullptr;         // [nelem] scaled values (zero on padding)
  float *val32 = nullptr;        // [nelem] the same values rounded to fp32, in entry order (DevPtrs::mat32), or nullptr
  unsigned short *cf = nullptr;  // [nelem] local column; bit 15 of every 4th word: row ends with this quad
  int *from_csr = nullptr;       // stacked CSR position -> stream position (value refresh after re-scaling)
  int *blk_group = nullptr;      // [grid]
  int paired = 0;                // 1: blocks 2p / 2p+1 form a cluster, stream the same rows of groups 0 / 1 and combine
                                 //    their partial row sums through distributed shared memory (no `part` traffic)
  int *blk_row0 = nullptr;       // [grid] stacked rows [blk_row0, blk_row1) streamed by the block
  int *blk_row1 = nullptr;
  int *grp_col0 = nullptr;       // [ngroups + 1] column range of each group (multiples of 32)
  int *w_row0 = nullptr;         // [grid * kWarps] first stacked row of warp i
  int *w_q0 = nullptr;           // [grid * kWarps + 1] first quad of warp i's stream (a multiple of 32 = one chunk)
  int *w_qn = nullptr;           // [grid * kWarps] quads in warp i's stream
  // Lane-row layout (lane_rows == 1, the default): the stream rows of a block are sorted by length and dealt 32 at a
  // time into SLICES; lane l of a slice owns stream row sl_row[slice * 32 + l] (-1: none) and finds quad t of it in
  // chunk t of the slice, so a row sum is a private accumulator of the lane (no segmented scan, no row-end flags).
  // A slice is sl_len[slice] chunks long (its longest row; shorter rows are zero-padded); warp i streams slices
  // [w_s0[i], w_s0[i + 1]) back to back from quad w_q0[i].  lane_rows == 0: the scan layout described above.
  int lane_rows = 0;
  int *w_s0 = nullptr;           // [grid * kWarps + 1]
  int *sl_len = nullptr;         // [slices]
  int *sl_row = nullptr;         // [slices * 32] stream-row index inside the group
  int split = 0;                 // 1: some rows are cut into several stream rows (engine.cuh kSplitQuads)
  int srows = 0;                 // stream rows per group (stride of `part`); == rows when nothing is split
  int *sr_ptr = nullptr;         // [ngroups][rows + 1] first stream row of each row (identity when split == 0)
  double *part = nullptr;        // [ngroups][srows] partial sums of the stream rows of the last phase
};

// Low-rank part of the PCG preconditioner (Woodbury).  K = P + sigma I + A' diag(rho) A is Jacobi-friendly except for
// a FEW coupling rows of A that carry the equality weight 1e3 rho and touch many columns (budget / factor rows of a
// portfolio, linking constraints): each puts an outlier eigenvalue into the Jacobi-scaled K and costs about one PCG
// iteration per ADMM step.  With W = those rows (w <= kWoodMax), D = diag(P) + sigma + diag of the remaining rows and
// S = diag(sqrt(rho_W)):
//     M = D + A_W' S^2 A_W ,   M^{-1} = D^{-1} - D^{-1} A_W' S C^{-1} S A_W D^{-1} ,   C = I + S A_W D^{-1} A_W' S  (w x w)
// C^{-1} is formed explicitly on the device whenever rho changes (admm_kernel / polish_kernel, kernels.cu
// wood_refresh); applying M^{-1} costs two small products with compact copies of A_W and A_W' and one w x w product.
constexpr int kWoodMax = 512;
constexpr int kWoodCols = 8;   // columns of C assembled per round of wood_refresh

struct WoodDev {
  int w = 0;                 // rows in the set (0: plain Jacobi)
  int ld = 0;                // leading dimension of C / Cinv
  int *rows = nullptr;       // [w] row of A of member a
  int *idx = nullptr;        // [m] member index of a row of A, -1 otherwise
  int *rp = nullptr, *ci = nullptr, *src = nullptr;     // A_W  as CSR (w rows over the n columns); src: position in A.val
  double *val = nullptr;
  int *trp = nullptr, *tci = nullptr, *tsrc = nullptr;  // A_W' as CSR (n rows over the w members)
  double *tval = nullptr;
  double *C = nullptr, *Cinv = nullptr;  // [ld * ld] C (then its Cholesky factor), explicit inverse
  double *s = nullptr, *t = nullptr, *g = nullptr;  // [ld] sqrt(penalty), S A_W v, S C^{-1} t
  double *v = nullptr;       // [n] D^{-1} r
  double *vk = nullptr;      // [kWoodCols][n + 8] scratch of wood_refresh
};

// Slack elimination in the preconditioner.  An equality row i whose column j(i) appears in no other row of A and only
// on the diagonal of P (y in "A_d x - y = b": Lasso, regression, soft constraints) puts K_yy = P_jj + sigma + rho_i a_ij^2
// on the diagonal of K and couples y_j to the rest only through row i.  The block factorisation
//   K = [I  K_xy K_yy^-1; 0  I] diag(S, K_yy) [I 0; K_yy^-1 K_yx  I],   S = K_xx - K_xy K_yy^-1 K_yx
// has a Schur complement S that is K_xx with the equality weight of those rows REDUCED to
// rho_i (P_jj + sigma) / K_yy (about 2 instead of 100 for the Lasso): Jacobi on S works where Jacobi on K needs one
// iteration per outlier eigenvalue.  M^-1 r = block back-substitution with diag(S)^-1 in place of S^-1; it costs one
// extra A' stream phase per PCG iteration (the A phase it needs doubles as the A phase of the next K-apply).
// Used when there are too many such rows for the dense Woodbury correction (kernels.cu pcg_run_stream_slack).
struct SlackDev {
  int rows = 0;            // equality rows with a private slack column (0: off)
  int *col = nullptr;      // [m] slack column of row i, -1 if none
  int *pos = nullptr;      // [m] position of a_ij in A.val (CSR)
  double *rho_eff = nullptr;  // [m] rho with the reduced weight on the slack rows
  float *g32 = nullptr;    // [m] gather vector of the extra A' phase (zero outside the slack rows)
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
  int pd_certified;        // k_gershgorin: every row of the scaled P + sigma I is strictly diagonally dominant (=> PD)
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
  double phase_us[kPhases];     // block 0's wall time per phase class (PhaseClock in kernels.cu)
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
  // tile streams for the hot phases (blocked == 0: fall back to the CSR + L1 gather path everywhere)
  int blocked = 0;
  int info_streams = 0;          // 1: update_info and the residual refresh also run on the tile streams
  TileStreamDev SA, ST;          // [A; P] against an n-vector, A' against an m-vector
  // mat32 == 1: the PCG phases of the fixed-mode kernels (kernels_fast.cu) stream SA.val32 / ST.val32 instead of the
  // fp64 values: 6 instead of 10 bytes per entry, so that both streams of a PCG iteration stay resident in L2.  The
  // PCG then iterates on K~ = K(1 + 6e-8): every right-hand side, the z~ = A x~ and residual rebuilds of a refresh and
  // update_info stay on the fp64 values, so the recurrences r += b - b_old, z~ += alpha A p only ever carry
  // (K - K~)(x~ - x~ at the last refresh), which vanishes as the iterates converge.
  int mat32 = 0;
  // fp32 shadows of the two vectors the PCG phases gather (u = M^{-1} r and rho .* (A u)): the staged slices of
  // these phases are fp32 (half the L2 -> SM traffic of the staging); u is rounded BEFORE it enters the recurrences, so
  // CG stays exact for a preconditioner perturbed at the 6e-8 level; the rounding of rho .* (A u) perturbs one K-apply
  // by 6e-8 of that step's own residual change (kernels.cu store_u / store_tr).  Values stay fp64 and sums accumulate
  // in fp64.  f32_slices == 0: everything fp64 (OSQP_B200_F32_SLICES=0).
  int f32_slices = 0;
  float *uu32 = nullptr, *tr32 = nullptr;  // n, m
  WoodDev W;                     // low-rank part of the preconditioner (w == 0: none)
  SlackDev SL;                   // slack elimination in the preconditioner (rows == 0: none); exclusive with W
  double *Pu = nullptr;          // n: P u of the current PCG iteration
  int smem_x_elems = 0;          // doubles of dynamic shared memory for the staged slice
  int smem_rows = 0;             // doubles of dynamic shared memory for the row sums of a cluster pair (0: unpaired)
  // work partition: block b owns rows [m_start[b], m_start[b+1]) of A and [n_start[b], n_start[b+1]) of P/A'
  int *m_start = nullptr, *n_start = nullptr;
  // grid barrier + reductions
  unsigned *bar = nullptr;  // kBarBytes: [0] arrival count, from byte 16: fixed-point accumulator banks
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
  double pcg_eta_e;        // > 0: also continue while the last energy-norm decrement exceeds pcg_eta_e^2 of the total; < 0: automatic
  int pcg_max_iter;
  int refresh_every;       // recompute z_tilde = A x_tilde and r = b - K x_tilde every k ADMM iterations (1 = always)
  int wood_refresh;        // the Woodbury data (WoodDev C^{-1}, s, D) are stale: rebuild before the first PCG
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
  int cluster = 1;  // thread blocks per cluster of the cooperative launches (2: TileStreamDev::paired)
  int fast = 1;     // 0: never use the fixed-mode compilations of the ADMM / polish kernels (OSQP_B200_FAST_KERNELS=0)
};

// ---- host wrappers implemented in kernels.cu (all asynchronous on `st`)
cudaError_t launch_scale_data(const DevPtrs &d, int scaling_iters, double sigma, cudaStream_t st);
cudaError_t launch_scale_vectors(const DevPtrs &d, int do_q, int do_bounds, cudaStream_t st);
cudaError_t launch_set_rho_vec(const DevPtrs &d, double rho, int detect_change_only, cudaStream_t st);
cudaError_t launch_apply_rho(const DevPtrs &d, double rho, cudaStream_t st);
cudaError_t launch_precond(const DevPtrs &d, double sigma, cudaStream_t st);
cudaError_t launch_pd_probe(const DevPtrs &d, LaunchGeom g, double sigma, int max_it, int trials, cudaStream_t st);
cudaError_t launch_gershgorin(const DevPtrs &d, double sigma, cudaStream_t st);
cudaError_t launch_warm_start(const DevPtrs &d, const double *x_in, const double *y_in, int scaling, cudaStream_t st);
cudaError_t launch_cold_start(const DevPtrs &d, cudaStream_t st);
cudaError_t launch_scatter_values(double *dst, const double *vals, const long long *idx, const int *map,
                                  long long k, cudaStream_t st);
cudaError_t launch_solve(const DevPtrs &d, const SolveCfg &cfg, LaunchGeom g, cudaStream_t st);
// kernels_fast.cu / kernels_fast2.cu / kernels_fast3.cu: admm_kernel / polish_kernel compiled with the storage mode fixed (kernels.cu fast_mode)
int fast_mode(const DevPtrs &d, const LaunchGeom &g);
cudaError_t launch_solve_fast(const DevPtrs &d, const SolveCfg &cfg, LaunchGeom g, cudaStream_t st);
cudaError_t launch_polish_fast(const DevPtrs &d, const PolishCfg &cfg, const SolveCfg &sc, PolishOut *out, LaunchGeom g,
                               cudaStream_t st);
void kernels_fast(const void **admm, const void **polish);
cudaError_t launch_solve_fast2(const DevPtrs &d, const SolveCfg &cfg, LaunchGeom g, cudaStream_t st);
cudaError_t launch_polish_fast2(const DevPtrs &d, const PolishCfg &cfg, const SolveCfg &sc, PolishOut *out, LaunchGeom g,
                                cudaStream_t st);
void kernels_fast2(const void **admm, const void **polish);
cudaError_t launch_solve_fast3(const DevPtrs &d, const SolveCfg &cfg, LaunchGeom g, cudaStream_t st);
cudaError_t launch_polish_fast3(const DevPtrs &d, const PolishCfg &cfg, const SolveCfg &sc, PolishOut *out, LaunchGeom g,
                                cudaStream_t st);
void kernels_fast3(const void **admm, const void **polish);
cudaError_t launch_polish(const DevPtrs &d, const PolishCfg &cfg, const SolveCfg &sc, PolishOut *out, LaunchGeom g,
                          cudaStream_t st);
cudaError_t launch_spmv(const DevPtrs &d, int which, const double *in, double *out, double sigma, LaunchGeom g,
                        cudaStream_t st);
int max_coop_blocks_per_sm(int block, size_t dyn_smem);
int coop_threads();          // threads per block the cooperative kernels are compiled for (kernels.cu kThreads)
size_t coop_static_smem();   // static shared memory of admm_kernel / polish_kernel (the larger)
cudaError_t configure_dyn_smem(size_t dyn_smem);
cudaError_t raise_dyn_smem(const void *func, size_t bytes);  // never lowers the per-device attribute
cudaError_t launch_fill_blocked(const DevPtrs &d, cudaStream_t st);
cudaError_t launch_fill_wood(const DevPtrs &d, cudaStream_t st);
cudaError_t launch_reduce_selftest(const DevPtrs &d, LaunchGeom g, double ref, double *out, cudaStream_t st);
cudaError_t launch_barrier_bench(const DevPtrs &d, LaunchGeom g, int iters, int mode, double *sink,
                                 unsigned long long *ns_out, cudaStream_t st);
int max_active_clusters(int csize, int block, size_t dyn_smem);
cudaError_t launch_membench(const void *buf, long long bytes, int pattern, int depth, int grid, double *sink,
                            cudaStream_t st);

}  // namespace osqpb200
