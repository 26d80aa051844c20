/*
 * osqp_b200.h -- engine-specific extensions exported by osqp.jl_b200/lib/libosqp.so next to
 * the 30 reference symbols of osqp.h.  Nothing in the reference calls these; they exist for
 * measurement (bench.py, profiles/), for tuning the inner solver, and for the batched path
 * that the reference has no API for (BASELINE.json config 5, SURVEY.md 8b "Batch extension").
 */
#ifndef OSQP_B200_EXT_H
#define OSQP_B200_EXT_H
#include "osqp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  c_int   device;        /* CUDA device ordinal the workspace lives on */
  c_int   grid;          /* persistent-kernel grid (blocks) */
  c_int   block;         /* threads per block */
  c_int   lanes_A;       /* threads cooperating on one row of A */
  c_int   lanes_N;       /* threads cooperating on one row of P / A' */
  c_int   nnz_A;         /* stored non-zeros of A */
  c_int   nnz_P_full;    /* stored non-zeros of the full symmetric P */
  c_int   launches;      /* kernels launched for this workspace since setup */
  c_int   admm_iters;    /* last solve: ADMM iterations executed */
  c_int   pcg_iters;     /* last solve: total PCG iterations */
  c_int   info_evals;    /* last solve: residual/termination evaluations */
  c_int   refreshes;     /* last solve: full residual rebuilds */
  c_float kernel_ms;     /* last solve: CUDA-event duration of the admm_kernel launch */
  c_float polish_ms;     /* last solve: CUDA-event duration of the polish_kernel launch (0 if none) */
  c_float alg_bytes;     /* last solve: algorithmic bytes of the admm_kernel launch (DESIGN.md model) */
  c_float spmv_bytes_A;  /* algorithmic bytes of one A v, A'w, (P+sigma I)v  (12 B/nnz CSR model) */
  c_float spmv_bytes_At;
  c_float spmv_bytes_P;
  /* last solve: wall time (us) thread block 0 spent per phase class of the admm_kernel launch.
   * PCG iteration: 0 stream [A;P]u | 1 grid barrier | 2 combine partials -> t, rho.*t, Pu | 3 reduce + barrier |
   * 4 stream A'(rho.*t) | 5 grid barrier | 6 vector recurrences | 7 reduce + barrier.  ADMM step: 8 x,z,y update,
   * rhs vector, barriers | 9 stream A' rhs + barrier | 10 rhs/residual + reduce | 11 residual refresh |
   * 12 update_info | 13 rho update | 14 epilogue | 15 unused */
  c_float phase_us[16];
  c_int   streams;       /* 1: the hot phases run on tile streams; 0: CSR path (small / declined problems) */
  c_int   groups_A;      /* column groups of the [A; P] stream and of the A' stream */
  c_int   groups_At;
  c_int   paired;        /* 1: [A; P] stream runs as cluster pairs with the DSMEM combine */
  c_int   fast_kernels;  /* last solve: 1 / 2 = ran on a fixed-mode compilation of the kernels (csrc/kernels_fast.cu, _fast2.cu) */
} OSQPB200Profile;

/* Measurement of the last osqp_solve on this workspace. */
c_int osqp_b200_get_profile(const OSQPWorkspace *work, OSQPB200Profile *out);

/* Inner-solver controls.  Each ADMM step's PCG stops at ||r||inf <= max(eta*||r0||inf, floor_rel*||b||inf)
 * where r0 is the residual the step starts from (defaults 1e-3, 1e-13; DESIGN.md "inner accuracy").
 * Values <= 0 keep the current one.  refresh_every: rebuild z~ = A x~ and r = b - K x~ from scratch
 * every k ADMM iterations (0: only when K or the iterates were changed from outside). */
c_int osqp_b200_set_pcg(OSQPWorkspace *work, c_float eta, c_float floor_rel, c_int max_iter, c_int refresh_every);

/* Standalone SpMV with the resident (scaled) matrices, same device code and work split as the ADMM
 * kernel.  which: 0 out=A in | 1 out=A' in | 2 out=(P+sigma I) in on the hot-path format (column-blocked,
 * vector tile staged in shared memory by TMA); 10/11/12: same products on the CSR + L1-gather path used by
 * the rare phases.  Host buffers; runs `reps` launches, returns the mean CUDA-event time per launch. */
c_int osqp_b200_spmv(OSQPWorkspace *work, c_int which, const c_float *in_host, c_float *out_host, c_int reps,
                     c_float *ms_per_rep);

/* Scaling computed at setup: D (n), E (m), c -- for parity checks against the oracle. */
c_int osqp_b200_get_scaling(OSQPWorkspace *work, c_float *D, c_float *E, c_float *c);

/* Debug: per-block globaltimer probes written by the last blocked osqp_b200_spmv launch ([grid][16]). */
c_int osqp_b200_debug_read(OSQPWorkspace *work, unsigned long long *out, c_int count);

/* Host-only self-test of the tile-stream storage (DESIGN.md "tile streams"): builds the stream of a CSR matrix
 * with the setup code and replays the device reduction on the CPU; y_out = M x.  Returns 0, 2 if the builder
 * declines the matrix (padding / capacity), > 2 on a format violation.  Needs no GPU. */
c_int osqp_b200_stream_selftest(c_int rows, c_int cols, const c_int *rowptr, const c_int *col, const c_float *val,
                                const c_float *x, c_int grid, c_int ngroups, c_int paired, c_float *y_out,
                                c_float *padding_ratio);

/* ---- batched engine (BASELINE.json config 5; the reference has no batch API -- SURVEY.md 8b "Batch extension") ----
 * `count` independent QPs that share ONE sparsity pattern (pattern->P upper triangular CSC, pattern->A CSC; the
 * value arrays of `pattern` are ignored) and differ in their values: Px [count][nnz(P)], Ax [count][nnz(A)],
 * q [count][n], l/u [count][m], row-major, in the order of the pattern's CSC arrays.  One thread block per QP, all
 * data in shared memory, exact (dense Cholesky) KKT solve; libosqp 0.6.2 semantics per QP except: no polish, no
 * time limit, adaptive_rho_interval = 0 means 50 (there is no per-QP wall clock).  n, m <= 256.
 * A batch lives on one GPU; shard a large batch over GPUs by calling osqp_batch_setup once per device with a
 * contiguous block of the QPs (osqp.jl_b200/batch.py: shard_range) -- no data is exchanged between shards. */
typedef struct OSQPB200Batch OSQPB200Batch;
typedef struct {
  c_int   iter;
  c_int   status_val;    /* same codes as OSQPInfo.status_val (src/constants.jl:9-21) */
  c_float obj_val, pri_res, dua_res, rho_estimate;
  c_int   rho_updates;
} OSQPB200BatchInfo;

c_int osqp_batch_setup(OSQPB200Batch **out, c_int count, const OSQPData *pattern, const c_float *Px,
                       const c_float *Ax, const c_float *q, const c_float *l, const c_float *u,
                       const OSQPSettings *settings);
/* x_out [count][n], y_out [count][m] (NaN without a solution; for a dual / primal infeasible QP the certificate
 * delta_x / delta_y is returned in x_out / y_out), info_out [count].  Iterates stay resident for warm starts. */
c_int osqp_batch_solve(OSQPB200Batch *b, c_float *x_out, c_float *y_out, OSQPB200BatchInfo *info_out);
/* The same solve without the host-side copy: *x, *y, *info point at the engine's pinned host mirrors, valid until the
 * next call on the batch (the ownership rule of workspace->solution in the single-QP ABI, src/interface.jl:179-191). */
c_int osqp_batch_solve_view(OSQPB200Batch *b, const c_float **x, const c_float **y, const OSQPB200BatchInfo **info);
/* Device pointers of the last solve's x* [count][n] and y* [count][m] (valid until the next call on the batch): what a
 * multi-GPU caller hands to its all-gather without a round trip through host memory. */
c_int osqp_batch_device_solution(OSQPB200Batch *b, c_float **x_dev, c_float **y_dev);
/* new q / l / u for every QP (NULL: keep); the MPC re-solve pattern of src/modcaches.jl:166-179 */
c_int osqp_batch_update(OSQPB200Batch *b, const c_float *q, const c_float *l, const c_float *u);
/* Pinned host staging for the inputs of osqp_batch_update: q [count][n], l, u [count][m].  A caller that writes its
 * new data there and passes the same pointers to osqp_batch_update gets a PCIe-speed H2D without a pageable bounce. */
c_int osqp_batch_input_view(OSQPB200Batch *b, c_float **q, c_float **l, c_float **u);
c_int osqp_batch_warm_start(OSQPB200Batch *b, const c_float *x, const c_float *y);
/* max_iter, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, alpha, check_termination, warm_start, scaled_termination */
c_int osqp_batch_update_setting(OSQPB200Batch *b, const char *name, c_float value);
c_float osqp_batch_last_kernel_ms(const OSQPB200Batch *b);
c_int osqp_batch_cleanup(OSQPB200Batch *b);

c_int osqp_b200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
