/*
 * osqp.h -- C ABI of the B200-native OSQP ADMM engine (libosqp.so drop-in).
 *
 * Every struct below is layout-exact with the Julia mirrors in the reference
 * (osqp/OSQP.jl v0.8.1) and every function is one of the 30 symbols that
 * reference ccall's.  Citations are <file>:<line> relative to the reference
 * tree.  Two shared libraries export this ABI:
 *
 *   osqp.jl_b200/lib/libosqp.so   -- the product: sm_100a CUDA engine
 *   oracle/liboracle_osqp.so      -- the CPU oracle (test infrastructure only)
 *
 * Scalar types: src/types.jl:5-9  (Cc_int = Clonglong on 64-bit; c_float = Cdouble).
 */
#ifndef OSQP_B200_OSQP_H
#define OSQP_B200_OSQP_H

#ifdef __cplusplus
extern "C" {
#endif

typedef long long c_int;   /* src/types.jl:5-9 */
typedef double    c_float;

/* ---- constants (src/constants.jl:1-21) ---------------------------------- */
#define OSQP_INFTY 1e30                 /* src/constants.jl:5 */
#define QDLDL_SOLVER 0                  /* src/constants.jl:1 */
#define MKL_PARDISO_SOLVER 1            /* src/constants.jl:2 */

#define OSQP_DUAL_INFEASIBLE_INACCURATE   (4)   /* src/constants.jl:10 */
#define OSQP_PRIMAL_INFEASIBLE_INACCURATE (3)
#define OSQP_SOLVED_INACCURATE            (2)
#define OSQP_SOLVED                       (1)
#define OSQP_MAX_ITER_REACHED             (-2)
#define OSQP_PRIMAL_INFEASIBLE            (-3)
#define OSQP_DUAL_INFEASIBLE              (-4)
#define OSQP_SIGINT                       (-5)
#define OSQP_TIME_LIMIT_REACHED           (-6)
#define OSQP_NON_CVX                      (-7)
#define OSQP_UNSOLVED                     (-10) /* src/constants.jl:20 */

/* ---- structs ------------------------------------------------------------ */

/* src/types.jl:11-19 (Ccsc), sizeof 56.  Julia passes nz = -1 (CSC), :46.
 * Indices are 0-based Int64 (src/types.jl:39-44).                          */
typedef struct {
  c_int    nzmax;
  c_int    m;
  c_int    n;
  c_int   *p;
  c_int   *i;
  c_float *x;
  c_int    nz;
} csc;

/* src/types.jl:101-109 (Data), sizeof 56. P is upper-triangular
 * (src/interface.jl:102-104); l/u clamped to +-1e30 (src/interface.jl:107-108). */
typedef struct {
  c_int    n;
  c_int    m;
  csc     *P;
  csc     *A;
  c_float *q;
  c_float *l;
  c_float *u;
} OSQPData;

/* src/types.jl:111-134 (Settings), sizeof 176; linsys_solver is a 4-byte enum. */
typedef struct {
  c_float rho;
  c_float sigma;
  c_int   scaling;
  c_int   adaptive_rho;
  c_int   adaptive_rho_interval;
  c_float adaptive_rho_tolerance;
  c_float adaptive_rho_fraction;
  c_int   max_iter;
  c_float eps_abs;
  c_float eps_rel;
  c_float eps_prim_inf;
  c_float eps_dual_inf;
  c_float alpha;
  int     linsys_solver;
  c_float delta;
  c_int   polish;
  c_int   polish_refine_iter;
  c_int   verbose;
  c_int   scaled_termination;
  c_int   check_termination;
  c_int   warm_start;
  c_float time_limit;
} OSQPSettings;

/* src/types.jl:81-99 (CInfo), sizeof 136. */
typedef struct {
  c_int   iter;
  char    status[32];
  c_int   status_val;
  c_int   status_polish;
  c_float obj_val;
  c_float pri_res;
  c_float dua_res;
  c_float setup_time;
  c_float solve_time;
  c_float update_time;
  c_float polish_time;
  c_float run_time;
  c_int   rho_updates;
  c_float rho_estimate;
} OSQPInfo;

/* src/types.jl:74-77 (Solution), sizeof 16. */
typedef struct {
  c_float *x;
  c_float *y;
} OSQPSolution;

/* src/types.jl:173-217 (Workspace): Julia unsafe_load's the first 240 bytes
 * and dereferences data, delta_y, delta_x, solution, info
 * (src/interface.jl:176-205, 744-746).  Those five are valid HOST pointers
 * until the next call on the workspace; all other pointer fields are opaque
 * (NULL in the CUDA engine: the iterates live in HBM).                      */
typedef struct {
  OSQPData     *data;          /* @0   host; n, m valid; P/A NULL in the engine */
  void         *linsys_solver; /* @8   opaque */
  void         *pol;           /* @16  opaque */
  c_float      *rho_vec;       /* @24 */
  c_float      *rho_inv_vec;   /* @32 */
  c_int        *constr_type;   /* @40 */
  c_float      *x;             /* @48 */
  c_float      *y;             /* @56 */
  c_float      *z;             /* @64 */
  c_float      *xz_tilde;      /* @72 */
  c_float      *x_prev;        /* @80 */
  c_float      *z_prev;        /* @88 */
  c_float      *Ax;            /* @96 */
  c_float      *Px;            /* @104 */
  c_float      *Aty;           /* @112 */
  c_float      *delta_y;       /* @120 host, m doubles (src/interface.jl:200) */
  c_float      *Atdelta_y;     /* @128 */
  c_float      *delta_x;       /* @136 host, n doubles (src/interface.jl:205) */
  c_float      *Pdelta_x;      /* @144 */
  c_float      *Adelta_x;      /* @152 */
  c_float      *D_temp;        /* @160 */
  c_float      *D_temp_A;      /* @168 */
  c_float      *E_temp;        /* @176 */
  OSQPSettings *settings;      /* @184 host */
  void         *scaling;       /* @192 opaque */
  OSQPSolution *solution;      /* @200 host; x (n), y (m) host doubles */
  OSQPInfo     *info;          /* @208 host */
  void         *timer;         /* @216 opaque */
  c_int         first_run;     /* @224 */
  c_int         summary_printed; /* @232 */
} OSQPWorkspace;

/* ---- layout checks (compile time): sizes and offsets of src/types.jl:11-217 -------------
 * OSQP.jl unsafe_load's these structs field by field; a compiler or edit that moves a field
 * must not build.  (C11 / C++11; tests/test_abi_layout.py checks the Python mirror.)        */
#include <stddef.h>
#if defined(__cplusplus)
#define OSQP_LAYOUT_ASSERT(cond, msg) static_assert(cond, msg)
#else
#define OSQP_LAYOUT_ASSERT(cond, msg) _Static_assert(cond, msg)
#endif
#define OSQP_LAYOUT_AT(T, f, off) OSQP_LAYOUT_ASSERT(offsetof(T, f) == (off), #T "." #f " must sit at byte " #off)
OSQP_LAYOUT_ASSERT(sizeof(c_int) == 8 && sizeof(c_float) == 8, "c_int = Int64, c_float = Float64 (src/types.jl:5-9)");
OSQP_LAYOUT_ASSERT(sizeof(void *) == 8, "64-bit pointers");
OSQP_LAYOUT_ASSERT(sizeof(csc) == 56, "Ccsc is 56 bytes");
OSQP_LAYOUT_AT(csc, nzmax, 0); OSQP_LAYOUT_AT(csc, m, 8); OSQP_LAYOUT_AT(csc, n, 16); OSQP_LAYOUT_AT(csc, p, 24);
OSQP_LAYOUT_AT(csc, i, 32); OSQP_LAYOUT_AT(csc, x, 40); OSQP_LAYOUT_AT(csc, nz, 48);
OSQP_LAYOUT_ASSERT(sizeof(OSQPData) == 56, "Data is 56 bytes");
OSQP_LAYOUT_AT(OSQPData, n, 0); OSQP_LAYOUT_AT(OSQPData, m, 8); OSQP_LAYOUT_AT(OSQPData, P, 16);
OSQP_LAYOUT_AT(OSQPData, A, 24); OSQP_LAYOUT_AT(OSQPData, q, 32); OSQP_LAYOUT_AT(OSQPData, l, 40);
OSQP_LAYOUT_AT(OSQPData, u, 48);
OSQP_LAYOUT_ASSERT(sizeof(OSQPSettings) == 176, "Settings is 176 bytes");
OSQP_LAYOUT_AT(OSQPSettings, rho, 0); OSQP_LAYOUT_AT(OSQPSettings, sigma, 8); OSQP_LAYOUT_AT(OSQPSettings, scaling, 16);
OSQP_LAYOUT_AT(OSQPSettings, adaptive_rho, 24); OSQP_LAYOUT_AT(OSQPSettings, adaptive_rho_interval, 32);
OSQP_LAYOUT_AT(OSQPSettings, adaptive_rho_tolerance, 40); OSQP_LAYOUT_AT(OSQPSettings, adaptive_rho_fraction, 48);
OSQP_LAYOUT_AT(OSQPSettings, max_iter, 56); OSQP_LAYOUT_AT(OSQPSettings, eps_abs, 64);
OSQP_LAYOUT_AT(OSQPSettings, eps_rel, 72); OSQP_LAYOUT_AT(OSQPSettings, eps_prim_inf, 80);
OSQP_LAYOUT_AT(OSQPSettings, eps_dual_inf, 88); OSQP_LAYOUT_AT(OSQPSettings, alpha, 96);
OSQP_LAYOUT_AT(OSQPSettings, linsys_solver, 104); OSQP_LAYOUT_AT(OSQPSettings, delta, 112);
OSQP_LAYOUT_AT(OSQPSettings, polish, 120); OSQP_LAYOUT_AT(OSQPSettings, polish_refine_iter, 128);
OSQP_LAYOUT_AT(OSQPSettings, verbose, 136); OSQP_LAYOUT_AT(OSQPSettings, scaled_termination, 144);
OSQP_LAYOUT_AT(OSQPSettings, check_termination, 152); OSQP_LAYOUT_AT(OSQPSettings, warm_start, 160);
OSQP_LAYOUT_AT(OSQPSettings, time_limit, 168);
OSQP_LAYOUT_ASSERT(sizeof(OSQPInfo) == 136, "CInfo is 136 bytes");
OSQP_LAYOUT_AT(OSQPInfo, iter, 0); OSQP_LAYOUT_AT(OSQPInfo, status, 8); OSQP_LAYOUT_AT(OSQPInfo, status_val, 40);
OSQP_LAYOUT_AT(OSQPInfo, status_polish, 48); OSQP_LAYOUT_AT(OSQPInfo, obj_val, 56); OSQP_LAYOUT_AT(OSQPInfo, pri_res, 64);
OSQP_LAYOUT_AT(OSQPInfo, dua_res, 72); OSQP_LAYOUT_AT(OSQPInfo, setup_time, 80); OSQP_LAYOUT_AT(OSQPInfo, solve_time, 88);
OSQP_LAYOUT_AT(OSQPInfo, update_time, 96); OSQP_LAYOUT_AT(OSQPInfo, polish_time, 104); OSQP_LAYOUT_AT(OSQPInfo, run_time, 112);
OSQP_LAYOUT_AT(OSQPInfo, rho_updates, 120); OSQP_LAYOUT_AT(OSQPInfo, rho_estimate, 128);
OSQP_LAYOUT_ASSERT(sizeof(OSQPSolution) == 16, "Solution is 16 bytes");
OSQP_LAYOUT_AT(OSQPSolution, x, 0); OSQP_LAYOUT_AT(OSQPSolution, y, 8);
OSQP_LAYOUT_ASSERT(sizeof(OSQPWorkspace) == 240, "Julia reads 240 bytes of Workspace");
OSQP_LAYOUT_AT(OSQPWorkspace, data, 0); OSQP_LAYOUT_AT(OSQPWorkspace, rho_vec, 24); OSQP_LAYOUT_AT(OSQPWorkspace, x, 48);
OSQP_LAYOUT_AT(OSQPWorkspace, delta_y, 120); OSQP_LAYOUT_AT(OSQPWorkspace, delta_x, 136);
OSQP_LAYOUT_AT(OSQPWorkspace, settings, 184); OSQP_LAYOUT_AT(OSQPWorkspace, scaling, 192);
OSQP_LAYOUT_AT(OSQPWorkspace, solution, 200); OSQP_LAYOUT_AT(OSQPWorkspace, info, 208);
OSQP_LAYOUT_AT(OSQPWorkspace, timer, 216); OSQP_LAYOUT_AT(OSQPWorkspace, first_run, 224);
OSQP_LAYOUT_AT(OSQPWorkspace, summary_printed, 232);

/* ---- the 30 symbols ------------------------------------------------------
 * All return 0 on success; non-zero => Julia raises ErrorException.          */

void        osqp_set_default_settings(OSQPSettings *settings);      /* src/types.jl:138-143 */
c_int       osqp_setup(OSQPWorkspace **workp, const OSQPData *data,
                       const OSQPSettings *settings);               /* src/interface.jl:146-153 */
c_int       osqp_solve(OSQPWorkspace *work);                        /* src/interface.jl:170-175 */
const char *osqp_version(void);                                     /* src/interface.jl:220 */
c_int       osqp_cleanup(OSQPWorkspace *work);                      /* src/interface.jl:224-229; NULL ok */

c_int osqp_update_lin_cost(OSQPWorkspace *work, const c_float *q_new);      /* src/interface.jl:240-246 */
c_int osqp_update_lower_bound(OSQPWorkspace *work, const c_float *l_new);   /* src/interface.jl:258-264 */
c_int osqp_update_upper_bound(OSQPWorkspace *work, const c_float *u_new);   /* src/interface.jl:276-282 */
c_int osqp_update_bounds(OSQPWorkspace *work, const c_float *l_new,
                         const c_float *u_new);                             /* src/interface.jl:302-309 */
c_int osqp_update_P(OSQPWorkspace *work, const c_float *Px_new,
                    const c_int *Px_new_idx, c_int P_new_n);                /* src/interface.jl:336-344 */
c_int osqp_update_A(OSQPWorkspace *work, const c_float *Ax_new,
                    const c_int *Ax_new_idx, c_int A_new_n);                /* src/interface.jl:357-365 */
c_int osqp_update_P_A(OSQPWorkspace *work, const c_float *Px_new,
                      const c_int *Px_new_idx, c_int P_new_n,
                      const c_float *Ax_new, const c_int *Ax_new_idx,
                      c_int A_new_n);                                       /* src/interface.jl:381-400 */

c_int osqp_warm_start(OSQPWorkspace *work, const c_float *x, const c_float *y); /* src/interface.jl:708-715 */
c_int osqp_warm_start_x(OSQPWorkspace *work, const c_float *x);                 /* src/interface.jl:675-681 */
c_int osqp_warm_start_y(OSQPWorkspace *work, const c_float *y);                 /* src/interface.jl:689-695 */

c_int osqp_update_max_iter(OSQPWorkspace *work, c_int max_iter_new);              /* src/interface.jl:475-481 */
c_int osqp_update_eps_abs(OSQPWorkspace *work, c_float eps_abs_new);              /* src/interface.jl:488-494 */
c_int osqp_update_eps_rel(OSQPWorkspace *work, c_float eps_rel_new);              /* src/interface.jl:501-507 */
c_int osqp_update_eps_prim_inf(OSQPWorkspace *work, c_float eps_prim_inf_new);    /* src/interface.jl:514-520 */
c_int osqp_update_eps_dual_inf(OSQPWorkspace *work, c_float eps_dual_inf_new);    /* src/interface.jl:527-533 */
c_int osqp_update_rho(OSQPWorkspace *work, c_float rho_new);                      /* src/interface.jl:540-546 */
c_int osqp_update_alpha(OSQPWorkspace *work, c_float alpha_new);                  /* src/interface.jl:553-559 */
c_int osqp_update_delta(OSQPWorkspace *work, c_float delta_new);                  /* src/interface.jl:566-572 */
c_int osqp_update_polish(OSQPWorkspace *work, c_int polish_new);                  /* src/interface.jl:579-585 */
c_int osqp_update_polish_refine_iter(OSQPWorkspace *work, c_int polish_refine_iter_new); /* src/interface.jl:592-598 */
c_int osqp_update_verbose(OSQPWorkspace *work, c_int verbose_new);                /* src/interface.jl:605-611 */
c_int osqp_update_scaled_termination(OSQPWorkspace *work, c_int scaled_termination_new); /* src/interface.jl:618-624 */
c_int osqp_update_check_termination(OSQPWorkspace *work, c_int check_termination_new);   /* src/interface.jl:631-637 */
c_int osqp_update_warm_start(OSQPWorkspace *work, c_int warm_start_new);          /* src/interface.jl:644-650 */
c_int osqp_update_time_limit(OSQPWorkspace *work, c_float time_limit_new);        /* src/interface.jl:657-663 */

#ifdef __cplusplus
}
#endif
#endif /* OSQP_B200_OSQP_H */
