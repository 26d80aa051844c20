// oracle/osqp_oracle.cpp -- TEST INFRASTRUCTURE, NOT THE PRODUCT.
//
// CPU restatement of the solver the reference binds: libosqp v0.6.2
// (OSQP_jll "=0.6.2", reference Project.toml:13,18, loaded at src/OSQP.jl:7).
// Its C source is NOT under /root/reference (un-vendored binary dependency),
// so the arithmetic below restates the published OSQP algorithm (Stellato et
// al., "OSQP: an operator splitting solver for quadratic programs", Math.
// Prog. Comp. 2020) with the v0.6.2 behavioural details listed in SURVEY.md
// Appendix A, and the ABI follows the reference's own call sites:
//   structs  -> src/types.jl:11-217          symbols -> src/interface.jl:146-715
//
// Parity status: PINNED on the reference's own known-answer tests (test/basic.jl,
// polishing.jl incl. the Mosek JLD2 fixture, dual_/primal_infeasibility.jl,
// non_convex.jl, unconstrained.jl, warm_start.jl invariants) -- see
// tests/test_reference_suite.py.  UNPINNED at the BASELINE.json sizes (the
// reference holds no test larger than 100x500) and against a live libosqp
// (none exists in this container).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// Two linear-system backends:
//   direct (default): sparse LDL^T of the (n+m) quasi-definite KKT matrix,
//                     single-threaded like libosqp+QDLDL  (ldl.hpp)
//   pcg             : Jacobi-PCG on the reduced system P+sigma*I+A'diag(rho)A
//                     with OpenMP SpMV -- the "best effort CPU" comparator and
//                     the only one that fits BASELINE config 2 in memory.

#include "../include/osqp.h"
#include "ldl.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using oracle::idx_t;
typedef std::vector<double> vec;

constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kDivisionTol = 1e-30;
constexpr int kPrintInterval = 200;
const double kNaN = std::numeric_limits<double>::quiet_NaN();

// process-global backend selection (see osqp_oracle_configure)
int g_linsys_mode = 0;  // 0 direct, 1 pcg (residual criterion), 2 pcg (preconditioned-residual criterion), 3 pcg (relative-to-initial-residual criterion)
double g_pcg_tol = 1e-9;
idx_t g_pcg_max_iter = 0;  // 0 => 10 * n, capped

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct Csc {
  idx_t m = 0, n = 0;
  std::vector<idx_t> p, i;
  vec x;
  idx_t nnz() const { return p.empty() ? 0 : p[n]; }
};

// ---------------------------------------------------------------- lin_alg
double norm_inf(const vec &v) {
  double r = 0;
  for (double a : v) r = std::max(r, std::fabs(a));
  return r;
}
double scaled_norm_inf(const vec &s, const vec &v) {
  double r = 0;
  for (size_t k = 0; k < v.size(); k++) r = std::max(r, std::fabs(s[k] * v[k]));
  return r;
}
double dot(const vec &a, const vec &b) {
  double r = 0;
  for (size_t k = 0; k < a.size(); k++) r += a[k] * b[k];
  return r;
}
// y (+)= A x   [libosqp mat_vec]
void mat_vec(const Csc &A, const double *x, double *y, int plus_eq) {
  if (!plus_eq) for (idx_t i = 0; i < A.m; i++) y[i] = 0;
  else if (plus_eq == -1) {
    for (idx_t j = 0; j < A.n; j++)
      for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) y[A.i[k]] -= A.x[k] * x[j];
    return;
  }
  for (idx_t j = 0; j < A.n; j++)
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) y[A.i[k]] += A.x[k] * x[j];
}
// y (+)= A' x, optionally skipping the diagonal [libosqp mat_tpose_vec]
void mat_tpose_vec(const Csc &A, const double *x, double *y, int plus_eq, int skip_diag) {
  if (!plus_eq) for (idx_t j = 0; j < A.n; j++) y[j] = 0;
  double sgn = (plus_eq == -1) ? -1.0 : 1.0;
  for (idx_t j = 0; j < A.n; j++) {
    double s = 0;
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) {
      if (skip_diag && A.i[k] == j) continue;
      s += A.x[k] * x[A.i[k]];
    }
    y[j] += sgn * s;
  }
}
// y = P x for P stored as upper triangle
void sym_mat_vec(const Csc &P, const double *x, double *y) {
  mat_vec(P, x, y, 0);
  mat_tpose_vec(P, x, y, 1, 1);
}
double quad_form(const Csc &P, const double *x) {
  double q = 0;
  for (idx_t j = 0; j < P.n; j++)
    for (idx_t k = P.p[j]; k < P.p[j + 1]; k++) {
      idx_t i = P.i[k];
      if (i == j) q += 0.5 * P.x[k] * x[i] * x[i];
      else if (i < j) q += P.x[k] * x[i] * x[j];
    }
  return q;
}
void inf_norm_cols(const Csc &A, vec &E) {
  for (idx_t j = 0; j < A.n; j++) {
    double r = 0;
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) r = std::max(r, std::fabs(A.x[k]));
    E[j] = r;
  }
}
void inf_norm_rows(const Csc &A, vec &E) {
  std::fill(E.begin(), E.end(), 0.0);
  for (idx_t j = 0; j < A.n; j++)
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) E[A.i[k]] = std::max(E[A.i[k]], std::fabs(A.x[k]));
}
void inf_norm_cols_sym_triu(const Csc &P, vec &E) {
  std::fill(E.begin(), E.end(), 0.0);
  for (idx_t j = 0; j < P.n; j++)
    for (idx_t k = P.p[j]; k < P.p[j + 1]; k++) {
      idx_t i = P.i[k];
      double a = std::fabs(P.x[k]);
      E[j] = std::max(E[j], a);
      if (i != j) E[i] = std::max(E[i], a);
    }
}
void premult_diag(Csc &A, const vec &d) {
  for (idx_t j = 0; j < A.n; j++)
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) A.x[k] *= d[A.i[k]];
}
void postmult_diag(Csc &A, const vec &d) {
  for (idx_t j = 0; j < A.n; j++)
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) A.x[k] *= d[j];
}

// ---------------------------------------------------------------- linear systems
struct LinSys {
  virtual ~LinSys() {}
  // ADMM mode: in b=[rhs_x; rhs_z], out b=[x_tilde; z_tilde].  Polish mode: b <- K^{-1} b.
  virtual int solve(double *b) = 0;
  virtual int update_matrices(const Csc &P, const Csc &A) = 0;
  virtual int update_rho_vec(const vec &rho_vec) = 0;
  virtual void warm_hint(const double *) {}
  double stat_a = 0, stat_b = 0;  // direct: nnz(L), factor count; pcg: total cg its, solves
};

// KKT = [[P + sigma I, A'],[A, -diag(1/rho)]], upper triangle, min-degree permuted.
struct DirectLdl : LinSys {
  idx_t n, m;
  bool polish;
  double sigma;
  vec rho_inv;  // per row (polish: all = sigma)
  oracle::SymCsc K, Kp;
  std::vector<idx_t> PtoK, AtoK, diagPK, rhoK, KtoKp, perm, pinv;
  oracle::Ldl ldl;
  vec sol, bp;

  int init(const Csc &P, const Csc &A, double sigma_, const vec *rho_vec, bool polish_) {
    n = P.n; m = A.m; sigma = sigma_; polish = polish_;
    rho_inv.assign(m, sigma);
    if (rho_vec) for (idx_t i = 0; i < m; i++) rho_inv[i] = 1.0 / (*rho_vec)[i];
    // row-major view of A for the (1,2) block columns
    std::vector<idx_t> rp(m + 1, 0), rj(A.nnz()), rk(A.nnz());
    for (idx_t k = 0; k < A.nnz(); k++) rp[A.i[k] + 1]++;
    for (idx_t i = 0; i < m; i++) rp[i + 1] += rp[i];
    {
      std::vector<idx_t> w(rp.begin(), rp.end() - 1);
      for (idx_t j = 0; j < n; j++)
        for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) { idx_t q = w[A.i[k]]++; rj[q] = j; rk[q] = k; }
    }
    K.n = n + m;
    K.p.assign(n + m + 1, 0);
    K.i.clear(); K.x.clear();
    PtoK.assign(P.nnz(), -1); AtoK.assign(A.nnz(), -1);
    diagPK.assign(n, -1); rhoK.assign(m, -1);
    for (idx_t j = 0; j < n; j++) {
      bool has_diag = false;
      for (idx_t k = P.p[j]; k < P.p[j + 1]; k++) {
        idx_t i = P.i[k];
        PtoK[k] = (idx_t)K.i.size();
        K.i.push_back(i);
        if (i == j) { K.x.push_back(P.x[k] + sigma); diagPK[j] = PtoK[k]; has_diag = true; }
        else K.x.push_back(P.x[k]);
      }
      if (!has_diag) { diagPK[j] = (idx_t)K.i.size(); K.i.push_back(j); K.x.push_back(sigma); }
      K.p[j + 1] = (idx_t)K.i.size();
    }
    for (idx_t r = 0; r < m; r++) {
      for (idx_t q = rp[r]; q < rp[r + 1]; q++) {
        AtoK[rk[q]] = (idx_t)K.i.size();
        K.i.push_back(rj[q]);
        K.x.push_back(A.x[rk[q]]);
      }
      rhoK[r] = (idx_t)K.i.size();
      K.i.push_back(n + r);
      K.x.push_back(-rho_inv[r]);
      K.p[n + r + 1] = (idx_t)K.i.size();
    }
    perm = oracle::min_degree_order(n + m, K.p, K.i);
    pinv.resize(n + m);
    for (idx_t k = 0; k < n + m; k++) pinv[perm[k]] = k;
    Kp = oracle::sym_permute(K, pinv, KtoKp);
    // sym_permute does not sort rows within a column; the up-looking factor does not need it.
    ldl.symbolic(Kp);
    sol.resize(n + m); bp.resize(n + m);
    stat_a = (double)ldl.Lp[n + m];
    return factor();
  }
  int factor() {
    stat_b += 1;
    idx_t npos = ldl.numeric(Kp);
    if (npos < 0) {
      fprintf(stderr, "ERROR in LDL factorisation: zero pivot in the KKT matrix\n");
      return -1;
    }
    if (npos != n) {
      fprintf(stderr, "ERROR in LDL factorisation: KKT matrix is not quasi-definite "
                      "(the problem seems to be non-convex)\n");
      return -2;
    }
    return 0;
  }
  int solve(double *b) override {
    idx_t N = n + m;
    for (idx_t k = 0; k < N; k++) bp[k] = b[perm[k]];
    ldl.solve(bp.data());
    for (idx_t k = 0; k < N; k++) sol[perm[k]] = bp[k];
    if (polish) { for (idx_t k = 0; k < N; k++) b[k] = sol[k]; return 0; }
    for (idx_t j = 0; j < n; j++) b[j] = sol[j];
    for (idx_t i = 0; i < m; i++) b[n + i] += rho_inv[i] * sol[n + i];
    return 0;
  }
  int update_matrices(const Csc &P, const Csc &A) override {
    for (idx_t k = 0; k < P.nnz(); k++) Kp.x[KtoKp[PtoK[k]]] = P.x[k];
    for (idx_t j = 0; j < n; j++) {
      // diagonal: P_jj (if stored) + sigma
      idx_t q = KtoKp[diagPK[j]];
      bool stored = false;
      for (idx_t k = P.p[j]; k < P.p[j + 1]; k++) if (P.i[k] == j) { Kp.x[q] = P.x[k] + sigma; stored = true; }
      if (!stored) Kp.x[q] = sigma;
    }
    for (idx_t k = 0; k < A.nnz(); k++) Kp.x[KtoKp[AtoK[k]]] = A.x[k];
    return factor();
  }
  int update_rho_vec(const vec &rho_vec) override {
    for (idx_t i = 0; i < m; i++) {
      rho_inv[i] = 1.0 / rho_vec[i];
      Kp.x[KtoKp[rhoK[i]]] = -rho_inv[i];
    }
    return factor();
  }
};

// Reduced-KKT Jacobi-PCG:  (P + sigma I + A' diag(rho) A) x = rhs_x + A' (rho .* rhs_z).
struct ReducedPcg : LinSys {
  idx_t n, m;
  double sigma, tol;
  bool precond_norm = false, rel_init = false;
  idx_t max_it;
  vec rho;
  // CSR of A, CSR of A' (= CSC of A), full symmetric P in CSR
  std::vector<idx_t> Ap, Aj, Tp, Tj, Pp, Pj;
  vec Ax, Tx, Px, Pdiag, Minv;
  vec xk, r, d, p, Kp_, t, rhs;

  int init(const Csc &P, const Csc &A, double sigma_, const vec &rho_vec) {
    n = P.n; m = A.m; sigma = sigma_; rho = rho_vec;
    tol = g_pcg_tol;
    precond_norm = (g_linsys_mode == 2);
    rel_init = (g_linsys_mode == 3);
    max_it = g_pcg_max_iter > 0 ? g_pcg_max_iter : std::max<idx_t>(20, std::min<idx_t>(10 * n, 5000));
    xk.assign(n, 0.0); r.resize(n); d.resize(n); p.resize(n); Kp_.resize(n); t.resize(m); rhs.resize(n);
    return update_matrices(P, A);
  }
  int update_matrices(const Csc &P, const Csc &A) override {
    Tp = A.p; Tj = A.i; Tx = A.x;
    Ap.assign(m + 1, 0);
    for (idx_t k = 0; k < A.nnz(); k++) Ap[A.i[k] + 1]++;
    for (idx_t i = 0; i < m; i++) Ap[i + 1] += Ap[i];
    Aj.resize(A.nnz()); Ax.resize(A.nnz());
    {
      std::vector<idx_t> w(Ap.begin(), Ap.end() - 1);
      for (idx_t j = 0; j < n; j++)
        for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) { idx_t q = w[A.i[k]]++; Aj[q] = j; Ax[q] = A.x[k]; }
    }
    Pp.assign(n + 1, 0);
    for (idx_t j = 0; j < n; j++)
      for (idx_t k = P.p[j]; k < P.p[j + 1]; k++) {
        Pp[j + 1]++;
        if (P.i[k] != j) Pp[P.i[k] + 1]++;
      }
    for (idx_t j = 0; j < n; j++) Pp[j + 1] += Pp[j];
    Pj.resize(Pp[n]); Px.resize(Pp[n]);
    Pdiag.assign(n, 0.0);
    {
      std::vector<idx_t> w(Pp.begin(), Pp.end() - 1);
      for (idx_t j = 0; j < n; j++)
        for (idx_t k = P.p[j]; k < P.p[j + 1]; k++) {
          idx_t i = P.i[k];
          idx_t q = w[j]++; Pj[q] = i; Px[q] = P.x[k];
          if (i != j) { q = w[i]++; Pj[q] = j; Px[q] = P.x[k]; }
          else Pdiag[j] += P.x[k];
        }
    }
    build_precond();
    return 0;
  }
  void build_precond() {
    Minv.resize(n);
#pragma omp parallel for schedule(static)
    for (idx_t j = 0; j < n; j++) {
      double s = Pdiag[j] + sigma;
      for (idx_t k = Tp[j]; k < Tp[j + 1]; k++) s += rho[Tj[k]] * Tx[k] * Tx[k];
      Minv[j] = 1.0 / s;
    }
  }
  int update_rho_vec(const vec &rho_vec) override { rho = rho_vec; build_precond(); return 0; }
  void apply_A(const double *v, double *out) const {
#pragma omp parallel for schedule(static)
    for (idx_t i = 0; i < m; i++) {
      double s = 0;
      for (idx_t k = Ap[i]; k < Ap[i + 1]; k++) s += Ax[k] * v[Aj[k]];
      out[i] = s;
    }
  }
  // out = P v + sigma v + A' w
  void apply_PAt(const double *v, const double *w, double *out) const {
#pragma omp parallel for schedule(static)
    for (idx_t j = 0; j < n; j++) {
      double s = sigma * v[j];
      for (idx_t k = Pp[j]; k < Pp[j + 1]; k++) s += Px[k] * v[Pj[k]];
      for (idx_t k = Tp[j]; k < Tp[j + 1]; k++) s += Tx[k] * w[Tj[k]];
      out[j] = s;
    }
  }
  void apply_K(const double *v, double *out) {
    apply_A(v, t.data());
#pragma omp parallel for schedule(static)
    for (idx_t i = 0; i < m; i++) t[i] *= rho[i];
    apply_PAt(v, t.data(), out);
  }
  int solve(double *b) override {
    stat_b += 1;
    // rhs = b_x + A'(rho .* b_z)
#pragma omp parallel for schedule(static)
    for (idx_t i = 0; i < m; i++) t[i] = rho[i] * b[n + i];
#pragma omp parallel for schedule(static)
    for (idx_t j = 0; j < n; j++) {
      double s = b[j];
      for (idx_t k = Tp[j]; k < Tp[j + 1]; k++) s += Tx[k] * t[Tj[k]];
      rhs[j] = s;
    }
    double bnorm = 0;
    if (precond_norm) for (idx_t j = 0; j < n; j++) bnorm = std::max(bnorm, std::fabs(Minv[j] * rhs[j]));
    else bnorm = norm_inf(rhs);
    double thresh = std::max(tol * bnorm, 1e-300);
    apply_K(xk.data(), Kp_.data());
    double rz = 0, rn = 0;
#pragma omp parallel for reduction(+ : rz) reduction(max : rn) schedule(static)
    for (idx_t j = 0; j < n; j++) {
      r[j] = rhs[j] - Kp_[j];
      d[j] = Minv[j] * r[j];
      p[j] = d[j];
      rz += r[j] * d[j];
      rn = std::max(rn, std::fabs(precond_norm ? d[j] : r[j]));
    }
    idx_t it = 0;
    if (rel_init) thresh = std::max(tol * rn, 1e-13 * bnorm);
    while (rn > thresh && it < max_it) {
      apply_K(p.data(), Kp_.data());
      double pKp = 0;
#pragma omp parallel for reduction(+ : pKp) schedule(static)
      for (idx_t j = 0; j < n; j++) pKp += p[j] * Kp_[j];
      if (!(pKp > 0)) break;  // negative curvature / breakdown
      double a = rz / pKp, rz_new = 0;
      rn = 0;
#pragma omp parallel for reduction(+ : rz_new) reduction(max : rn) schedule(static)
      for (idx_t j = 0; j < n; j++) {
        xk[j] += a * p[j];
        r[j] -= a * Kp_[j];
        d[j] = Minv[j] * r[j];
        rz_new += r[j] * d[j];
        rn = std::max(rn, std::fabs(precond_norm ? d[j] : r[j]));
      }
      double beta = rz_new / rz;
      rz = rz_new;
#pragma omp parallel for schedule(static)
      for (idx_t j = 0; j < n; j++) p[j] = d[j] + beta * p[j];
      it++;
    }
    stat_a += (double)it;
    for (idx_t j = 0; j < n; j++) b[j] = xk[j];
    apply_A(xk.data(), b + n);
    return 0;
  }
};

// ---------------------------------------------------------------- workspace
struct Work {
  OSQPWorkspace pub;  // MUST be first: the ABI pointer is &pub
  OSQPData data_pub;
  OSQPSettings st;
  OSQPInfo info;
  OSQPSolution sol_pub;
  idx_t n = 0, m = 0;
  Csc P, A;  // scaled working copies
  csc P_pub, A_pub;
  vec q, l, u;
  vec rho_vec, rho_inv_vec;
  std::vector<c_int> constr_type;
  vec x, y, z, xz_tilde, x_prev, z_prev, Ax, Px, Aty, delta_y, Atdelta_y, delta_x, Pdelta_x, Adelta_x;
  vec D, Dinv, E, Einv, D_temp, D_temp_A, E_temp;
  double c = 1, cinv = 1;
  vec sol_x, sol_y;
  std::unique_ptr<LinSys> lin;
  int linsys_mode = 0;
  // polish
  vec pol_x, pol_z, pol_y;
  double pol_obj = 0, pol_pri = 0, pol_dua = 0;
  // bookkeeping
  double timer0 = 0;
  bool clear_update_time = false, rho_update_from_solve = false;
  bool first_run = true, summary_printed = false;
};

Work *W(OSQPWorkspace *w) { return reinterpret_cast<Work *>(w); }

void update_status(OSQPInfo &info, c_int v) {
  info.status_val = v;
  const char *s = "unsolved";
  switch (v) {
    case OSQP_SOLVED: s = "solved"; break;
    case OSQP_SOLVED_INACCURATE: s = "solved inaccurate"; break;
    case OSQP_PRIMAL_INFEASIBLE: s = "primal infeasible"; break;
    case OSQP_PRIMAL_INFEASIBLE_INACCURATE: s = "primal infeasible inaccurate"; break;
    case OSQP_DUAL_INFEASIBLE: s = "dual infeasible"; break;
    case OSQP_DUAL_INFEASIBLE_INACCURATE: s = "dual infeasible inaccurate"; break;
    case OSQP_MAX_ITER_REACHED: s = "maximum iterations reached"; break;
    case OSQP_TIME_LIMIT_REACHED: s = "run time limit reached"; break;
    case OSQP_SIGINT: s = "interrupted"; break;
    case OSQP_NON_CVX: s = "problem non convex"; break;
    default: break;
  }
  memset(info.status, 0, sizeof(info.status));
  strncpy(info.status, s, sizeof(info.status) - 1);
}

void reset_info(OSQPInfo &info) {
  info.solve_time = 0;
  info.polish_time = 0;
  update_status(info, OSQP_UNSOLVED);
  info.rho_updates = 0;
}

void limit_scaling(vec &v) {
  for (double &a : v) {
    a = a < kMinScaling ? 1.0 : a;
    a = a > kMaxScaling ? kMaxScaling : a;
  }
}
double limit_scaling1(double a) {
  a = a < kMinScaling ? 1.0 : a;
  return a > kMaxScaling ? kMaxScaling : a;
}

// modified Ruiz equilibration + cost normalisation.
void scale_data(Work &w) {
  idx_t n = w.n, m = w.m;
  w.c = 1.0;
  w.D.assign(n, 1.0); w.Dinv.assign(n, 1.0); w.E.assign(m, 1.0); w.Einv.assign(m, 1.0);
  for (c_int it = 0; it < w.st.scaling; it++) {
    inf_norm_cols_sym_triu(w.P, w.D_temp);
    inf_norm_cols(w.A, w.D_temp_A);
    for (idx_t j = 0; j < n; j++) w.D_temp[j] = std::max(w.D_temp[j], w.D_temp_A[j]);
    inf_norm_rows(w.A, w.E_temp);
    limit_scaling(w.D_temp);
    limit_scaling(w.E_temp);
    for (double &a : w.D_temp) a = 1.0 / std::sqrt(a);
    for (double &a : w.E_temp) a = 1.0 / std::sqrt(a);
    premult_diag(w.P, w.D_temp); postmult_diag(w.P, w.D_temp);
    premult_diag(w.A, w.E_temp); postmult_diag(w.A, w.D_temp);
    for (idx_t j = 0; j < n; j++) w.q[j] *= w.D_temp[j];
    for (idx_t j = 0; j < n; j++) w.D[j] *= w.D_temp[j];
    for (idx_t i = 0; i < m; i++) w.E[i] *= w.E_temp[i];
    // cost normalisation
    inf_norm_cols_sym_triu(w.P, w.D_temp);
    double c_temp = 0;
    for (idx_t j = 0; j < n; j++) c_temp += w.D_temp[j];
    c_temp /= (double)n;
    double nq = limit_scaling1(norm_inf(w.q));
    c_temp = limit_scaling1(std::max(c_temp, nq));
    c_temp = 1.0 / c_temp;
    for (double &a : w.P.x) a *= c_temp;
    for (double &a : w.q) a *= c_temp;
    w.c *= c_temp;
  }
  w.cinv = 1.0 / w.c;
  for (idx_t j = 0; j < n; j++) w.Dinv[j] = 1.0 / w.D[j];
  for (idx_t i = 0; i < m; i++) w.Einv[i] = 1.0 / w.E[i];
  for (idx_t i = 0; i < m; i++) { w.l[i] *= w.E[i]; w.u[i] *= w.E[i]; }
}

void unscale_data(Work &w) {
  for (double &a : w.P.x) a *= w.cinv;
  premult_diag(w.P, w.Dinv); postmult_diag(w.P, w.Dinv);
  for (idx_t j = 0; j < w.n; j++) w.q[j] *= w.cinv * w.Dinv[j];
  premult_diag(w.A, w.Einv); postmult_diag(w.A, w.Dinv);
  for (idx_t i = 0; i < w.m; i++) { w.l[i] *= w.Einv[i]; w.u[i] *= w.Einv[i]; }
}

// SURVEY 8a row a3
void set_rho_vec(Work &w) {
  w.st.rho = std::min(std::max(w.st.rho, kRhoMin), kRhoMax);
  for (idx_t i = 0; i < w.m; i++) {
    if (w.l[i] < -OSQP_INFTY * kMinScaling && w.u[i] > OSQP_INFTY * kMinScaling) {
      w.constr_type[i] = -1; w.rho_vec[i] = kRhoMin;
    } else if (w.u[i] - w.l[i] < kRhoTol) {
      w.constr_type[i] = 1; w.rho_vec[i] = kRhoEqOverIneq * w.st.rho;
    } else {
      w.constr_type[i] = 0; w.rho_vec[i] = w.st.rho;
    }
    w.rho_inv_vec[i] = 1.0 / w.rho_vec[i];
  }
}
int update_rho_vec(Work &w) {
  bool changed = false;
  for (idx_t i = 0; i < w.m; i++) {
    c_int t; double r;
    if (w.l[i] < -OSQP_INFTY * kMinScaling && w.u[i] > OSQP_INFTY * kMinScaling) { t = -1; r = kRhoMin; }
    else if (w.u[i] - w.l[i] < kRhoTol) { t = 1; r = kRhoEqOverIneq * w.st.rho; }
    else { t = 0; r = w.st.rho; }
    if (w.constr_type[i] != t) {
      w.constr_type[i] = t; w.rho_vec[i] = r; w.rho_inv_vec[i] = 1.0 / r; changed = true;
    }
  }
  if (changed) return w.lin->update_rho_vec(w.rho_vec);
  return 0;
}

void cold_start(Work &w) {
  std::fill(w.x.begin(), w.x.end(), 0.0);
  std::fill(w.z.begin(), w.z.end(), 0.0);
  std::fill(w.y.begin(), w.y.end(), 0.0);
}

// ---- ADMM steps
void update_xz_tilde(Work &w) {
  idx_t n = w.n, m = w.m;
  for (idx_t j = 0; j < n; j++) w.xz_tilde[j] = w.st.sigma * w.x_prev[j] - w.q[j];
  for (idx_t i = 0; i < m; i++) w.xz_tilde[n + i] = w.z_prev[i] - w.rho_inv_vec[i] * w.y[i];
  w.lin->solve(w.xz_tilde.data());
}
void update_x(Work &w) {
  double a = w.st.alpha;
  for (idx_t j = 0; j < w.n; j++) {
    w.x[j] = a * w.xz_tilde[j] + (1.0 - a) * w.x_prev[j];
    w.delta_x[j] = w.x[j] - w.x_prev[j];
  }
}
void update_z(Work &w) {
  double a = w.st.alpha;
  for (idx_t i = 0; i < w.m; i++) {
    double v = a * w.xz_tilde[w.n + i] + (1.0 - a) * w.z_prev[i] + w.rho_inv_vec[i] * w.y[i];
    w.z[i] = std::min(std::max(v, w.l[i]), w.u[i]);
  }
}
void update_y(Work &w) {
  double a = w.st.alpha;
  for (idx_t i = 0; i < w.m; i++) {
    w.delta_y[i] = w.rho_vec[i] * (a * w.xz_tilde[w.n + i] + (1.0 - a) * w.z_prev[i] - w.z[i]);
    w.y[i] += w.delta_y[i];
  }
}

// ---- residuals (row a9)
double compute_obj_val(Work &w, const vec &x) {
  double o = quad_form(w.P, x.data()) + dot(w.q, x);
  if (w.st.scaling) o *= w.cinv;
  return o;
}
double compute_pri_res(Work &w, const vec &x, const vec &z) {
  mat_vec(w.A, x.data(), w.Ax.data(), 0);
  for (idx_t i = 0; i < w.m; i++) w.z_prev[i] = w.Ax[i] - z[i];
  if (w.st.scaling && !w.st.scaled_termination) return scaled_norm_inf(w.Einv, w.z_prev);
  return norm_inf(w.z_prev);
}
double compute_pri_tol(Work &w, double eps_abs, double eps_rel) {
  double mx;
  if (w.st.scaling && !w.st.scaled_termination)
    mx = std::max(scaled_norm_inf(w.Einv, w.z), scaled_norm_inf(w.Einv, w.Ax));
  else
    mx = std::max(norm_inf(w.z), norm_inf(w.Ax));
  return eps_abs + eps_rel * mx;
}
double compute_dua_res(Work &w, const vec &x, const vec &y) {
  w.x_prev = w.q;
  sym_mat_vec(w.P, x.data(), w.Px.data());
  for (idx_t j = 0; j < w.n; j++) w.x_prev[j] += w.Px[j];
  if (w.m > 0) {
    mat_tpose_vec(w.A, y.data(), w.Aty.data(), 0, 0);
    for (idx_t j = 0; j < w.n; j++) w.x_prev[j] += w.Aty[j];
  }
  if (w.st.scaling && !w.st.scaled_termination) return w.cinv * scaled_norm_inf(w.Dinv, w.x_prev);
  return norm_inf(w.x_prev);
}
double compute_dua_tol(Work &w, double eps_abs, double eps_rel) {
  double mx;
  if (w.st.scaling && !w.st.scaled_termination) {
    mx = std::max(scaled_norm_inf(w.Dinv, w.q),
                  std::max(scaled_norm_inf(w.Dinv, w.Aty), scaled_norm_inf(w.Dinv, w.Px)));
    mx *= w.cinv;
  } else {
    mx = std::max(norm_inf(w.q), std::max(norm_inf(w.Aty), norm_inf(w.Px)));
  }
  return eps_abs + eps_rel * mx;
}

void update_info(Work &w, c_int iter, bool compute_objective, bool polish) {
  if (polish) {
    if (compute_objective) w.pol_obj = compute_obj_val(w, w.pol_x);
    w.pol_pri = (w.m == 0) ? 0.0 : compute_pri_res(w, w.pol_x, w.pol_z);
    w.pol_dua = compute_dua_res(w, w.pol_x, w.pol_y);
    w.info.polish_time = now_s() - w.timer0;
  } else {
    w.info.iter = iter;
    if (compute_objective) w.info.obj_val = compute_obj_val(w, w.x);
    w.info.pri_res = (w.m == 0) ? 0.0 : compute_pri_res(w, w.x, w.z);
    w.info.dua_res = compute_dua_res(w, w.x, w.y);
    w.info.solve_time = now_s() - w.timer0;
  }
  w.summary_printed = false;
}

// ---- infeasibility (row a10)
bool is_primal_infeasible(Work &w, double eps) {
  idx_t m = w.m;
  for (idx_t i = 0; i < m; i++) {
    if (w.u[i] > OSQP_INFTY * kMinScaling) {
      if (w.l[i] < -OSQP_INFTY * kMinScaling) w.delta_y[i] = 0.0;
      else w.delta_y[i] = std::min(w.delta_y[i], 0.0);
    } else if (w.l[i] < -OSQP_INFTY * kMinScaling) {
      w.delta_y[i] = std::max(w.delta_y[i], 0.0);
    }
  }
  bool unscale = w.st.scaling && !w.st.scaled_termination;
  double nrm = unscale ? scaled_norm_inf(w.E, w.delta_y) : norm_inf(w.delta_y);
  if (nrm > kDivisionTol) {
    double lhs = 0;
    for (idx_t i = 0; i < m; i++)
      lhs += w.u[i] * std::max(w.delta_y[i], 0.0) + w.l[i] * std::min(w.delta_y[i], 0.0);
    if (lhs < -eps * nrm) {
      mat_tpose_vec(w.A, w.delta_y.data(), w.Atdelta_y.data(), 0, 0);
      if (unscale) for (idx_t j = 0; j < w.n; j++) w.Atdelta_y[j] *= w.Dinv[j];
      return norm_inf(w.Atdelta_y) < eps * nrm;
    }
  }
  return false;
}
bool is_dual_infeasible(Work &w, double eps) {
  bool unscale = w.st.scaling && !w.st.scaled_termination;
  double nrm = unscale ? scaled_norm_inf(w.D, w.delta_x) : norm_inf(w.delta_x);
  double cs = unscale ? w.c : 1.0;
  if (nrm > kDivisionTol) {
    if (dot(w.q, w.delta_x) < -cs * eps * nrm) {
      sym_mat_vec(w.P, w.delta_x.data(), w.Pdelta_x.data());
      if (unscale) for (idx_t j = 0; j < w.n; j++) w.Pdelta_x[j] *= w.Dinv[j];
      if (norm_inf(w.Pdelta_x) < cs * eps * nrm) {
        mat_vec(w.A, w.delta_x.data(), w.Adelta_x.data(), 0);
        if (unscale) for (idx_t i = 0; i < w.m; i++) w.Adelta_x[i] *= w.Einv[i];
        for (idx_t i = 0; i < w.m; i++) {
          if ((w.u[i] < OSQP_INFTY * kMinScaling && w.Adelta_x[i] > eps * nrm) ||
              (w.l[i] > -OSQP_INFTY * kMinScaling && w.Adelta_x[i] < -eps * nrm))
            return false;
        }
        return true;
      }
    }
  }
  return false;
}

bool check_termination(Work &w, bool approximate) {
  double eps_abs = w.st.eps_abs, eps_rel = w.st.eps_rel;
  double eps_pinf = w.st.eps_prim_inf, eps_dinf = w.st.eps_dual_inf;
  if (w.info.pri_res > OSQP_INFTY || w.info.dua_res > OSQP_INFTY ||
      std::isnan(w.info.pri_res) || std::isnan(w.info.dua_res)) {
    update_status(w.info, OSQP_NON_CVX);
    w.info.obj_val = kNaN;
    return true;
  }
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10; }
  bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
  if (w.m == 0) prim_ok = true;
  else {
    double eps_prim = compute_pri_tol(w, eps_abs, eps_rel);
    if (w.info.pri_res < eps_prim) prim_ok = true;
    else prim_inf = is_primal_infeasible(w, eps_pinf);
  }
  double eps_dual = compute_dua_tol(w, eps_abs, eps_rel);
  if (w.info.dua_res < eps_dual) dual_ok = true;
  else dual_inf = is_dual_infeasible(w, eps_dinf);

  bool unscale = w.st.scaling && !w.st.scaled_termination;
  if (prim_ok && dual_ok) {
    update_status(w.info, approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED);
    return true;
  } else if (prim_inf) {
    update_status(w.info, approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE);
    if (unscale) for (idx_t i = 0; i < w.m; i++) w.delta_y[i] *= w.E[i];
    w.info.obj_val = OSQP_INFTY;
    return true;
  } else if (dual_inf) {
    update_status(w.info, approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE);
    if (unscale) for (idx_t j = 0; j < w.n; j++) w.delta_x[j] *= w.D[j];
    w.info.obj_val = -OSQP_INFTY;
    return true;
  }
  return false;
}

// ---- adaptive rho (row a11)
double compute_rho_estimate(Work &w) {
  double pri = norm_inf(w.z_prev), dua = norm_inf(w.x_prev);
  double pn = std::max(norm_inf(w.z), norm_inf(w.Ax));
  pri /= (pn + 1e-10);
  double dn = std::max(norm_inf(w.q), std::max(norm_inf(w.Aty), norm_inf(w.Px)));
  dua /= (dn + 1e-10);
  double est = w.st.rho * std::sqrt(pri / (dua + 1e-10));
  return std::min(std::max(est, kRhoMin), kRhoMax);
}
int set_rho(Work &w, double rho_new) {
  w.st.rho = std::min(std::max(rho_new, kRhoMin), kRhoMax);
  for (idx_t i = 0; i < w.m; i++) {
    if (w.constr_type[i] == 0) { w.rho_vec[i] = w.st.rho; w.rho_inv_vec[i] = 1.0 / w.st.rho; }
    else if (w.constr_type[i] == 1) {
      w.rho_vec[i] = kRhoEqOverIneq * w.st.rho; w.rho_inv_vec[i] = 1.0 / w.rho_vec[i];
    }
  }
  return w.lin->update_rho_vec(w.rho_vec);
}
int adapt_rho(Work &w) {
  double rho_new = compute_rho_estimate(w);
  w.info.rho_estimate = rho_new;
  if (rho_new > w.st.rho * w.st.adaptive_rho_tolerance || rho_new < w.st.rho / w.st.adaptive_rho_tolerance) {
    int e = set_rho(w, rho_new);
    w.info.rho_updates += 1;
    w.rho_update_from_solve = true;
    return e;
  }
  return 0;
}

bool has_solution(const OSQPInfo &info) {
  c_int s = info.status_val;
  return s != OSQP_PRIMAL_INFEASIBLE && s != OSQP_PRIMAL_INFEASIBLE_INACCURATE &&
         s != OSQP_DUAL_INFEASIBLE && s != OSQP_DUAL_INFEASIBLE_INACCURATE && s != OSQP_NON_CVX;
}

void store_solution(Work &w) {
  if (has_solution(w.info)) {
    for (idx_t j = 0; j < w.n; j++) w.sol_x[j] = w.st.scaling ? w.D[j] * w.x[j] : w.x[j];
    for (idx_t i = 0; i < w.m; i++) w.sol_y[i] = w.st.scaling ? w.cinv * w.E[i] * w.y[i] : w.y[i];
  } else {
    std::fill(w.sol_x.begin(), w.sol_x.end(), kNaN);
    std::fill(w.sol_y.begin(), w.sol_y.end(), kNaN);
    c_int s = w.info.status_val;
    if (s == OSQP_PRIMAL_INFEASIBLE || s == OSQP_PRIMAL_INFEASIBLE_INACCURATE) {
      double nv = norm_inf(w.delta_y);
      for (double &a : w.delta_y) a /= nv;
    }
    if (s == OSQP_DUAL_INFEASIBLE || s == OSQP_DUAL_INFEASIBLE_INACCURATE) {
      double nv = norm_inf(w.delta_x);
      for (double &a : w.delta_x) a /= nv;
    }
    cold_start(w);
  }
}

// ---- printing (cosmetic; format is ours)
void print_setup_header(Work &w) {
  printf("-----------------------------------------------------------------\n");
  printf("   OSQP oracle (CPU restatement of libosqp 0.6.2 behaviour)\n");
  printf("-----------------------------------------------------------------\n");
  printf("problem:  variables n = %lld, constraints m = %lld\n", (long long)w.n, (long long)w.m);
  printf("          nnz(P) + nnz(A) = %lld\n", (long long)(w.P.nnz() + w.A.nnz()));
  printf("settings: linear system solver = %s,\n", w.linsys_mode ? "reduced-KKT Jacobi-PCG" : "direct LDL'");
  printf("          eps_abs = %.1e, eps_rel = %.1e,\n", w.st.eps_abs, w.st.eps_rel);
  printf("          eps_prim_inf = %.1e, eps_dual_inf = %.1e,\n", w.st.eps_prim_inf, w.st.eps_dual_inf);
  printf("          rho = %.2e %s, sigma = %.2e, alpha = %.2f, max_iter = %lld\n", w.st.rho,
         w.st.adaptive_rho ? "(adaptive)" : "", w.st.sigma, w.st.alpha, (long long)w.st.max_iter);
  printf("          scaling: %s, polish: %s, warm start: %s\n\n", w.st.scaling ? "on" : "off",
         w.st.polish ? "on" : "off", w.st.warm_start ? "on" : "off");
}
void print_header() { printf("iter   objective    pri res    dua res    rho        time\n"); }
void print_summary(Work &w) {
  printf("%4lld  %11.4e  %9.2e  %9.2e  %9.2e  %9.2es\n", (long long)w.info.iter, w.info.obj_val,
         w.info.pri_res, w.info.dua_res, w.st.rho, w.info.solve_time);
  w.summary_printed = true;
}
void print_footer(Work &w) {
  printf("\nstatus:               %s\n", w.info.status);
  if (w.st.polish && w.info.status_val == OSQP_SOLVED)
    printf("solution polish:      %s\n", w.info.status_polish == 1 ? "successful" : "unsuccessful");
  printf("number of iterations: %lld\n", (long long)w.info.iter);
  if (w.info.status_val == OSQP_SOLVED || w.info.status_val == OSQP_SOLVED_INACCURATE)
    printf("optimal objective:    %.4f\n", w.info.obj_val);
  printf("run time:             %.2es\n", w.info.run_time);
  printf("optimal rho estimate: %.2e\n\n", w.info.rho_estimate);
}

// ---- polish (row a12)
int polish(Work &w) {
  idx_t n = w.n, m = w.m;
  w.timer0 = now_s();
  std::vector<idx_t> A_to_low(m, -1), A_to_upp(m, -1), low_to_A, upp_to_A;
  for (idx_t i = 0; i < m; i++)
    if (w.z[i] - w.l[i] < -w.y[i]) { A_to_low[i] = (idx_t)low_to_A.size(); low_to_A.push_back(i); }
  for (idx_t i = 0; i < m; i++)
    if (w.u[i] - w.z[i] < w.y[i]) { A_to_upp[i] = (idx_t)upp_to_A.size(); upp_to_A.push_back(i); }
  idx_t n_low = (idx_t)low_to_A.size(), n_upp = (idx_t)upp_to_A.size(), mred = n_low + n_upp;
  Csc Ared;
  Ared.m = mred; Ared.n = n; Ared.p.assign(n + 1, 0);
  for (idx_t j = 0; j < n; j++) {
    for (idx_t k = w.A.p[j]; k < w.A.p[j + 1]; k++) {
      idx_t i = w.A.i[k];
      if (A_to_low[i] != -1) { Ared.i.push_back(A_to_low[i]); Ared.x.push_back(w.A.x[k]); }
      else if (A_to_upp[i] != -1) { Ared.i.push_back(A_to_upp[i] + n_low); Ared.x.push_back(w.A.x[k]); }
    }
    Ared.p[j + 1] = (idx_t)Ared.i.size();
  }
  DirectLdl plsh;
  int ef = plsh.init(w.P, Ared, w.st.delta, nullptr, true);
  if (ef) {
    w.info.status_polish = -1;
    return 1;
  }
  vec rhs(n + mred), sol;
  for (idx_t j = 0; j < n; j++) rhs[j] = -w.q[j];
  for (idx_t j = 0; j < n_low; j++) rhs[n + j] = w.l[low_to_A[j]];
  for (idx_t j = 0; j < n_upp; j++) rhs[n + n_low + j] = w.u[upp_to_A[j]];
  sol = rhs;
  plsh.solve(sol.data());
  // iterative refinement against the unregularised KKT
  vec res(n + mred);
  for (c_int it = 0; it < w.st.polish_refine_iter; it++) {
    res = rhs;
    mat_vec(w.P, sol.data(), res.data(), -1);
    mat_tpose_vec(w.P, sol.data(), res.data(), -1, 1);
    mat_tpose_vec(Ared, sol.data() + n, res.data(), -1, 0);
    mat_vec(Ared, sol.data(), res.data() + n, -1);
    plsh.solve(res.data());
    for (idx_t k = 0; k < n + mred; k++) sol[k] += res[k];
  }
  for (idx_t j = 0; j < n; j++) w.pol_x[j] = sol[j];
  mat_vec(w.A, w.pol_x.data(), w.pol_z.data(), 0);
  for (idx_t i = 0; i < m; i++) {
    if (mred == 0) w.pol_y[i] = 0;
    else if (A_to_low[i] != -1) w.pol_y[i] = sol[n + A_to_low[i]];
    else if (A_to_upp[i] != -1) w.pol_y[i] = sol[n + n_low + A_to_upp[i]];
    else w.pol_y[i] = 0;
  }
  // project (z, y) onto the normal cone of [l, u]
  for (idx_t i = 0; i < m; i++) {
    double t = w.pol_z[i] + w.pol_y[i];
    w.pol_z[i] = std::min(std::max(t, w.l[i]), w.u[i]);
    w.pol_y[i] = t - w.pol_z[i];
  }
  update_info(w, 0, true, true);
  bool ok = (w.pol_pri < w.info.pri_res && w.pol_dua < w.info.dua_res) ||
            (w.pol_pri < w.info.pri_res && w.info.dua_res < 1e-10) ||
            (w.pol_dua < w.info.dua_res && w.info.pri_res < 1e-10);
  if (ok) {
    w.info.obj_val = w.pol_obj; w.info.pri_res = w.pol_pri; w.info.dua_res = w.pol_dua;
    w.info.status_polish = 1;
    w.x = w.pol_x; w.z = w.pol_z; w.y = w.pol_y;
    if (w.st.verbose) printf("plsh  %11.4e  %9.2e  %9.2e   --------  %9.2es\n", w.info.obj_val,
                             w.info.pri_res, w.info.dua_res, w.info.polish_time);
  } else {
    w.info.status_polish = -1;
  }
  return 0;
}

// ---- validation
int validate_data(const OSQPData *d) {
  if (!d) { fprintf(stderr, "ERROR in osqp_setup: missing data\n"); return 1; }
  if (!d->P || !d->A || (!d->q && d->n > 0)) { fprintf(stderr, "ERROR in osqp_setup: missing matrix/vector\n"); return 1; }
  if (d->n <= 0 || d->m < 0) { fprintf(stderr, "ERROR in osqp_setup: n must be positive and m nonnegative\n"); return 1; }
  if (d->P->m != d->n || d->P->n != d->n) { fprintf(stderr, "ERROR in osqp_setup: P does not have dimension n x n\n"); return 1; }
  for (c_int j = 0; j < d->n; j++)
    for (c_int k = d->P->p[j]; k < d->P->p[j + 1]; k++)
      if (d->P->i[k] > j) { fprintf(stderr, "ERROR in osqp_setup: P is not upper triangular\n"); return 1; }
  if (d->A->m != d->m || d->A->n != d->n) { fprintf(stderr, "ERROR in osqp_setup: A does not have dimension m x n\n"); return 1; }
  for (c_int i = 0; i < d->m; i++)
    if (d->l[i] > d->u[i]) {
      fprintf(stderr, "ERROR in osqp_setup: lower bound at index %lld is greater than upper bound\n", (long long)i);
      return 1;
    }
  return 0;
}
int validate_settings(const OSQPSettings *s) {
  if (!s) return 1;
  bool bad = s->scaling < 0 || (s->adaptive_rho != 0 && s->adaptive_rho != 1) || s->adaptive_rho_interval < 0 ||
             s->adaptive_rho_fraction <= 0 || s->adaptive_rho_tolerance < 1.0 || s->polish_refine_iter < 0 ||
             s->rho <= 0 || s->sigma <= 0 || s->delta <= 0 || s->max_iter <= 0 || s->eps_abs < 0 ||
             s->eps_rel < 0 || (s->eps_rel == 0 && s->eps_abs == 0) || s->eps_prim_inf <= 0 ||
             s->eps_dual_inf <= 0 || s->alpha <= 0 || s->alpha >= 2 ||
             (s->linsys_solver != QDLDL_SOLVER && s->linsys_solver != MKL_PARDISO_SOLVER) ||
             (s->verbose != 0 && s->verbose != 1) || (s->scaled_termination != 0 && s->scaled_termination != 1) ||
             s->check_termination < 0 || (s->warm_start != 0 && s->warm_start != 1) || s->time_limit < 0;
  if (bad) fprintf(stderr, "ERROR in osqp_setup: invalid settings\n");
  return bad ? 1 : 0;
}

void copy_csc(Csc &dst, const csc *src) {
  dst.m = src->m; dst.n = src->n;
  dst.p.assign(src->p, src->p + src->n + 1);
  idx_t nz = dst.p[dst.n];
  dst.i.assign(src->i, src->i + nz);
  dst.x.assign(src->x, src->x + nz);
}
void publish(Work &w) {
  OSQPWorkspace &p = w.pub;
  w.P_pub = {w.P.nnz(), w.n, w.n, w.P.p.data(), w.P.i.data(), w.P.x.data(), -1};
  w.A_pub = {w.A.nnz(), w.m, w.n, w.A.p.data(), w.A.i.data(), w.A.x.data(), -1};
  w.data_pub = {w.n, w.m, &w.P_pub, &w.A_pub, w.q.data(), w.l.data(), w.u.data()};
  w.sol_pub = {w.sol_x.data(), w.sol_y.data()};
  p.data = &w.data_pub; p.linsys_solver = w.lin.get(); p.pol = nullptr;
  p.rho_vec = w.rho_vec.data(); p.rho_inv_vec = w.rho_inv_vec.data(); p.constr_type = w.constr_type.data();
  p.x = w.x.data(); p.y = w.y.data(); p.z = w.z.data(); p.xz_tilde = w.xz_tilde.data();
  p.x_prev = w.x_prev.data(); p.z_prev = w.z_prev.data();
  p.Ax = w.Ax.data(); p.Px = w.Px.data(); p.Aty = w.Aty.data();
  p.delta_y = w.delta_y.data(); p.Atdelta_y = w.Atdelta_y.data();
  p.delta_x = w.delta_x.data(); p.Pdelta_x = w.Pdelta_x.data(); p.Adelta_x = w.Adelta_x.data();
  p.D_temp = w.D_temp.data(); p.D_temp_A = w.D_temp_A.data(); p.E_temp = w.E_temp.data();
  p.settings = &w.st; p.scaling = nullptr; p.solution = &w.sol_pub; p.info = &w.info; p.timer = nullptr;
  p.first_run = w.first_run; p.summary_printed = w.summary_printed;
}

void begin_update(Work &w) {
  if (w.clear_update_time) { w.clear_update_time = false; w.info.update_time = 0.0; }
}

int make_linsys(Work &w) {
  if (w.linsys_mode >= 1) {
    auto s = std::make_unique<ReducedPcg>();
    int e = s->init(w.P, w.A, w.st.sigma, w.rho_vec);
    w.lin = std::move(s);
    return e;
  }
  auto s = std::make_unique<DirectLdl>();
  int e = s->init(w.P, w.A, w.st.sigma, &w.rho_vec, false);
  w.lin = std::move(s);
  return e;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

// ---- oracle-only extensions (not part of the reference ABI)
void osqp_oracle_configure(c_int linsys_mode, c_float pcg_tol, c_int pcg_max_iter) {
  g_linsys_mode = (int)linsys_mode;
  if (pcg_tol > 0) g_pcg_tol = pcg_tol;
  g_pcg_max_iter = pcg_max_iter;
}
// out[0]=direct: nnz(L) | pcg: total CG iterations; out[1]=factorisations | solves
void osqp_oracle_stats(OSQPWorkspace *work, c_float *out) {
  Work &w = *W(work);
  out[0] = w.lin ? w.lin->stat_a : 0;
  out[1] = w.lin ? w.lin->stat_b : 0;
}
// torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs of bench.py ask for all host threads explicitly
void osqp_oracle_set_num_threads(c_int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads((int)n);
#else
  (void)n;
#endif
}
c_int osqp_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// `count` independent workspaces solved concurrently, one workspace per task (each solve itself is single-threaded,
// like libosqp): the CPU arm of bench.py for the batched configuration -- a plain C loop, so that the comparison is
// not against the interpreter that would otherwise drive the solves one by one
c_int osqp_solve(OSQPWorkspace *work);
c_int osqp_oracle_solve_many(OSQPWorkspace **works, c_int count) {
  c_int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
#endif
  for (c_int k = 0; k < count; k++) bad += osqp_solve(works[k]) != 0;
  return bad;
}
// y = A x / y = A' x on a caller-provided CSC matrix (SpMV parity checks)
void osqp_oracle_mat_vec(const csc *A, const c_float *x, c_float *y, c_int transpose) {
  Csc M;
  copy_csc(M, A);
  if (transpose) mat_tpose_vec(M, x, y, 0, 0);
  else mat_vec(M, x, y, 0);
}
// scaling vectors of a set-up workspace: D (n), E (m), c
void osqp_oracle_get_scaling(OSQPWorkspace *work, c_float *D, c_float *E, c_float *c) {
  Work &w = *W(work);
  for (idx_t j = 0; j < w.n; j++) D[j] = w.st.scaling ? w.D[j] : 1.0;
  for (idx_t i = 0; i < w.m; i++) E[i] = w.st.scaling ? w.E[i] : 1.0;
  *c = w.st.scaling ? w.c : 1.0;
}

// ---- src/types.jl:138-143
void osqp_set_default_settings(OSQPSettings *s) {
  memset(s, 0, sizeof(*s));
  s->rho = 0.1; s->sigma = 1e-6; s->scaling = 10;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 0; s->adaptive_rho_tolerance = 5; s->adaptive_rho_fraction = 0.4;
  s->max_iter = 4000; s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->alpha = 1.6; s->linsys_solver = QDLDL_SOLVER; s->delta = 1e-6; s->polish = 0; s->polish_refine_iter = 3;
  s->verbose = 1; s->scaled_termination = 0; s->check_termination = 25; s->warm_start = 1; s->time_limit = 0;
}

const char *osqp_version(void) { return "0.6.2-oracle"; }

// ---- src/interface.jl:146-153
c_int osqp_setup(OSQPWorkspace **workp, const OSQPData *data, const OSQPSettings *settings) {
  if (workp) *workp = nullptr;
  if (validate_data(data)) return 1;
  if (validate_settings(settings)) return 1;
  double t0 = now_s();
  Work *wp = new Work();
  Work &w = *wp;
  w.n = data->n; w.m = data->m;
  idx_t n = w.n, m = w.m;
  copy_csc(w.P, data->P);
  copy_csc(w.A, data->A);
  w.q.assign(data->q, data->q + n);
  w.l.assign(data->l, data->l + m);
  w.u.assign(data->u, data->u + m);
  w.st = *settings;
  w.linsys_mode = g_linsys_mode;
  w.rho_vec.assign(m, 0); w.rho_inv_vec.assign(m, 0); w.constr_type.assign(m, 0);
  w.x.assign(n, 0); w.z.assign(m, 0); w.y.assign(m, 0); w.xz_tilde.assign(n + m, 0);
  w.x_prev.assign(n, 0); w.z_prev.assign(m, 0);
  w.Ax.assign(m, 0); w.Px.assign(n, 0); w.Aty.assign(n, 0);
  w.delta_y.assign(m, 0); w.Atdelta_y.assign(n, 0); w.delta_x.assign(n, 0); w.Pdelta_x.assign(n, 0);
  w.Adelta_x.assign(m, 0);
  w.D_temp.assign(n, 0); w.D_temp_A.assign(n, 0); w.E_temp.assign(m, 0);
  w.sol_x.assign(n, 0); w.sol_y.assign(m, 0);
  w.pol_x.assign(n, 0); w.pol_z.assign(m, 0); w.pol_y.assign(m, 0);
  if (w.st.scaling) scale_data(w);
  set_rho_vec(w);
  int e = make_linsys(w);
  if (e) {
    fprintf(stderr, "ERROR in osqp_setup: KKT matrix factorization failed (the problem seems to be non-convex)\n");
    delete wp;
    return e == -2 ? 7 : 4;
  }
  memset(&w.info, 0, sizeof(w.info));
  w.info.status_polish = 0;
  update_status(w.info, OSQP_UNSOLVED);
  w.info.iter = 0; w.info.rho_updates = 0; w.info.rho_estimate = w.st.rho;
  w.first_run = true; w.summary_printed = false;
  w.info.setup_time = now_s() - t0;
  publish(w);
  if (w.st.verbose) print_setup_header(w);
  *workp = &w.pub;
  return 0;
}

// ---- src/interface.jl:170-175 (return value ignored by the caller)
c_int osqp_solve(OSQPWorkspace *work) {
  if (!work) { fprintf(stderr, "ERROR in osqp_solve: workspace not initialized\n"); return 1; }
  Work &w = *W(work);
  if (w.clear_update_time) w.info.update_time = 0.0;
  w.rho_update_from_solve = false;
  bool can_check = false, can_print = false;
  w.timer0 = now_s();
  if (w.st.verbose) print_header();
  if (!w.st.warm_start) cold_start(w);
  c_int iter;
  for (iter = 1; iter <= w.st.max_iter; iter++) {
    std::swap(w.x, w.x_prev);
    std::swap(w.z, w.z_prev);
    update_xz_tilde(w);
    update_x(w);
    update_z(w);
    update_y(w);
    {
      double base = w.first_run ? w.info.setup_time : w.info.update_time;
      double run = base + (now_s() - w.timer0);
      if (w.st.time_limit > 0 && run >= w.st.time_limit) {
        update_status(w.info, OSQP_TIME_LIMIT_REACHED);
        if (w.st.verbose) printf("run time limit reached\n");
        can_check = false;
        break;
      }
    }
    can_check = w.st.check_termination && (iter % w.st.check_termination == 0);
    can_print = w.st.verbose && ((iter % kPrintInterval == 0) || iter == 1);
    if (can_check || can_print) {
      update_info(w, iter, true, false);
      if (can_print) print_summary(w);
      if (can_check && check_termination(w, false)) break;
    }
    if (w.st.adaptive_rho && !w.st.adaptive_rho_interval) {
      if (now_s() - w.timer0 > w.st.adaptive_rho_fraction * w.info.setup_time) {
        c_int N = w.st.check_termination ? w.st.check_termination : 25;
        double xx = (double)iter + 0.5 * (double)N;
        w.st.adaptive_rho_interval = (c_int)(xx - std::fmod(xx, (double)N));
        w.st.adaptive_rho_interval = std::max(w.st.adaptive_rho_interval, w.st.check_termination);
      }
    }
    if (w.st.adaptive_rho && w.st.adaptive_rho_interval && (iter % w.st.adaptive_rho_interval == 0)) {
      if (!can_check && !can_print) update_info(w, iter, true, false);
      if (adapt_rho(w)) { fprintf(stderr, "ERROR in osqp_solve: failed rho update\n"); return 1; }
    }
  }
  if (!can_check) {
    if (!can_print) update_info(w, iter - 1, true, false);
    if (w.st.verbose && !w.summary_printed) print_summary(w);
    check_termination(w, false);
  }
  w.info.rho_estimate = compute_rho_estimate(w);
  if (w.info.status_val == OSQP_UNSOLVED) {
    if (!check_termination(w, true)) update_status(w.info, OSQP_MAX_ITER_REACHED);
  }
  if (w.info.status_val == OSQP_TIME_LIMIT_REACHED) {
    if (!check_termination(w, true)) update_status(w.info, OSQP_TIME_LIMIT_REACHED);
  }
  w.info.solve_time = now_s() - w.timer0;
  if (w.st.verbose && !w.summary_printed) print_summary(w);
  if (w.st.polish && w.info.status_val == OSQP_SOLVED) polish(w);
  w.info.run_time = (w.first_run ? w.info.setup_time : w.info.update_time) + w.info.solve_time + w.info.polish_time;
  w.first_run = false;
  w.clear_update_time = true;
  w.rho_update_from_solve = false;
  if (w.st.verbose) print_footer(w);
  store_solution(w);
  publish(w);
  return 0;
}

// ---- src/interface.jl:224-229
c_int osqp_cleanup(OSQPWorkspace *work) {
  if (work) delete W(work);
  return 0;
}

// ---- updates (row a13)
c_int osqp_update_lin_cost(OSQPWorkspace *work, const c_float *q_new) {
  if (!work) return 1;
  Work &w = *W(work);
  begin_update(w);
  double t0 = now_s();
  w.q.assign(q_new, q_new + w.n);
  if (w.st.scaling) for (idx_t j = 0; j < w.n; j++) w.q[j] *= w.D[j] * w.c;
  reset_info(w.info);
  w.info.update_time += now_s() - t0;
  return 0;
}
c_int osqp_update_bounds(OSQPWorkspace *work, const c_float *l_new, const c_float *u_new) {
  if (!work) return 1;
  Work &w = *W(work);
  begin_update(w);
  double t0 = now_s();
  for (idx_t i = 0; i < w.m; i++)
    if (l_new[i] > u_new[i]) {
      fprintf(stderr, "ERROR in osqp_update_bounds: lower bound must be lower than or equal to upper bound\n");
      return 1;
    }
  w.l.assign(l_new, l_new + w.m);
  w.u.assign(u_new, u_new + w.m);
  if (w.st.scaling) for (idx_t i = 0; i < w.m; i++) { w.l[i] *= w.E[i]; w.u[i] *= w.E[i]; }
  reset_info(w.info);
  int e = update_rho_vec(w);
  w.info.update_time += now_s() - t0;
  return e;
}
c_int osqp_update_lower_bound(OSQPWorkspace *work, const c_float *l_new) {
  if (!work) return 1;
  Work &w = *W(work);
  begin_update(w);
  double t0 = now_s();
  w.l.assign(l_new, l_new + w.m);
  if (w.st.scaling) for (idx_t i = 0; i < w.m; i++) w.l[i] *= w.E[i];
  for (idx_t i = 0; i < w.m; i++)
    if (w.l[i] > w.u[i]) {
      fprintf(stderr, "ERROR in osqp_update_lower_bound: upper bound must be greater than or equal to lower bound\n");
      return 1;
    }
  reset_info(w.info);
  int e = update_rho_vec(w);
  w.info.update_time += now_s() - t0;
  return e;
}
c_int osqp_update_upper_bound(OSQPWorkspace *work, const c_float *u_new) {
  if (!work) return 1;
  Work &w = *W(work);
  begin_update(w);
  double t0 = now_s();
  w.u.assign(u_new, u_new + w.m);
  if (w.st.scaling) for (idx_t i = 0; i < w.m; i++) w.u[i] *= w.E[i];
  for (idx_t i = 0; i < w.m; i++)
    if (w.l[i] > w.u[i]) {
      fprintf(stderr, "ERROR in osqp_update_upper_bound: upper bound must be greater than or equal to lower bound\n");
      return 1;
    }
  reset_info(w.info);
  int e = update_rho_vec(w);
  w.info.update_time += now_s() - t0;
  return e;
}

static c_int update_PA(Work &w, const c_float *Px_new, const c_int *Px_idx, c_int P_n, bool doP,
                       const c_float *Ax_new, const c_int *Ax_idx, c_int A_n, bool doA) {
  begin_update(w);
  double t0 = now_s();
  idx_t nnzP = w.P.nnz(), nnzA = w.A.nnz();
  if (doP && Px_idx && P_n > nnzP) {
    fprintf(stderr, "ERROR in osqp_update_P: new number of elements (%lld) greater than elements in P (%lld)\n",
            (long long)P_n, (long long)nnzP);
    return 1;
  }
  if (doA && Ax_idx && A_n > nnzA) {
    fprintf(stderr, "ERROR in osqp_update_A: new number of elements (%lld) greater than elements in A (%lld)\n",
            (long long)A_n, (long long)nnzA);
    return doP ? 2 : 1;
  }
  if (w.st.scaling) unscale_data(w);
  if (doP) {
    if (Px_idx) for (c_int k = 0; k < P_n; k++) w.P.x[Px_idx[k]] = Px_new[k];
    else for (idx_t k = 0; k < nnzP; k++) w.P.x[k] = Px_new[k];
  }
  if (doA) {
    if (Ax_idx) for (c_int k = 0; k < A_n; k++) w.A.x[Ax_idx[k]] = Ax_new[k];
    else for (idx_t k = 0; k < nnzA; k++) w.A.x[k] = Ax_new[k];
  }
  if (w.st.scaling) scale_data(w);
  // NB libosqp keeps rho_vec as is here (constraint types are re-derived only on bound updates)
  int e = w.lin->update_matrices(w.P, w.A);
  if (w.linsys_mode >= 1) w.lin->update_rho_vec(w.rho_vec);
  reset_info(w.info);
  if (e < 0) fprintf(stderr, "ERROR in osqp_update_P/A: new KKT matrix is not quasidefinite\n");
  w.info.update_time += now_s() - t0;
  publish(w);
  return e;
}
c_int osqp_update_P(OSQPWorkspace *work, const c_float *Px_new, const c_int *Px_new_idx, c_int P_new_n) {
  if (!work) return 1;
  return update_PA(*W(work), Px_new, Px_new_idx, P_new_n, true, nullptr, nullptr, 0, false);
}
c_int osqp_update_A(OSQPWorkspace *work, const c_float *Ax_new, const c_int *Ax_new_idx, c_int A_new_n) {
  if (!work) return 1;
  return update_PA(*W(work), nullptr, nullptr, 0, false, Ax_new, Ax_new_idx, A_new_n, true);
}
c_int osqp_update_P_A(OSQPWorkspace *work, const c_float *Px_new, const c_int *Px_new_idx, c_int P_new_n,
                      const c_float *Ax_new, const c_int *Ax_new_idx, c_int A_new_n) {
  if (!work) return 1;
  return update_PA(*W(work), Px_new, Px_new_idx, P_new_n, true, Ax_new, Ax_new_idx, A_new_n, true);
}

// ---- warm start (row a14)
c_int osqp_warm_start(OSQPWorkspace *work, const c_float *x, const c_float *y) {
  if (!work) return 1;
  Work &w = *W(work);
  if (!w.st.warm_start) w.st.warm_start = 1;
  w.x.assign(x, x + w.n);
  w.y.assign(y, y + w.m);
  if (w.st.scaling) {
    for (idx_t j = 0; j < w.n; j++) w.x[j] *= w.Dinv[j];
    for (idx_t i = 0; i < w.m; i++) w.y[i] *= w.Einv[i] * w.c;
  }
  mat_vec(w.A, w.x.data(), w.z.data(), 0);
  return 0;
}
c_int osqp_warm_start_x(OSQPWorkspace *work, const c_float *x) {
  if (!work) return 1;
  Work &w = *W(work);
  if (!w.st.warm_start) w.st.warm_start = 1;
  w.x.assign(x, x + w.n);
  if (w.st.scaling) for (idx_t j = 0; j < w.n; j++) w.x[j] *= w.Dinv[j];
  mat_vec(w.A, w.x.data(), w.z.data(), 0);
  return 0;
}
c_int osqp_warm_start_y(OSQPWorkspace *work, const c_float *y) {
  if (!work) return 1;
  Work &w = *W(work);
  if (!w.st.warm_start) w.st.warm_start = 1;
  w.y.assign(y, y + w.m);
  if (w.st.scaling) for (idx_t i = 0; i < w.m; i++) w.y[i] *= w.Einv[i] * w.c;
  return 0;
}

// ---- settings (row a15)
#define ORACLE_SETTER(NAME, TYPE, FIELD, BADCOND, MSG)                 \
  c_int NAME(OSQPWorkspace *work, TYPE v) {                            \
    if (!work) return 1;                                               \
    if (BADCOND) { fprintf(stderr, "ERROR in " #NAME ": " MSG "\n"); return 1; } \
    W(work)->st.FIELD = v;                                             \
    return 0;                                                          \
  }
ORACLE_SETTER(osqp_update_max_iter, c_int, max_iter, v <= 0, "max_iter must be positive")
ORACLE_SETTER(osqp_update_eps_abs, c_float, eps_abs, v < 0, "eps_abs must be nonnegative")
ORACLE_SETTER(osqp_update_eps_rel, c_float, eps_rel, v < 0, "eps_rel must be nonnegative")
ORACLE_SETTER(osqp_update_eps_prim_inf, c_float, eps_prim_inf, v < 0, "eps_prim_inf must be nonnegative")
ORACLE_SETTER(osqp_update_eps_dual_inf, c_float, eps_dual_inf, v < 0, "eps_dual_inf must be nonnegative")
ORACLE_SETTER(osqp_update_alpha, c_float, alpha, (v <= 0 || v >= 2), "alpha must be between 0 and 2")
ORACLE_SETTER(osqp_update_delta, c_float, delta, v <= 0, "delta must be positive")
ORACLE_SETTER(osqp_update_polish_refine_iter, c_int, polish_refine_iter, v < 0, "polish_refine_iter must be nonnegative")
ORACLE_SETTER(osqp_update_verbose, c_int, verbose, (v != 0 && v != 1), "verbose should be either 0 or 1")
ORACLE_SETTER(osqp_update_scaled_termination, c_int, scaled_termination, (v != 0 && v != 1), "scaled_termination should be either 0 or 1")
ORACLE_SETTER(osqp_update_check_termination, c_int, check_termination, v < 0, "check_termination should be nonnegative")
ORACLE_SETTER(osqp_update_warm_start, c_int, warm_start, (v != 0 && v != 1), "warm_start should be either 0 or 1")
ORACLE_SETTER(osqp_update_time_limit, c_float, time_limit, v < 0, "time_limit must be nonnegative")

c_int osqp_update_polish(OSQPWorkspace *work, c_int v) {
  if (!work) return 1;
  if (v != 0 && v != 1) { fprintf(stderr, "ERROR in osqp_update_polish: polish should be either 0 or 1\n"); return 1; }
  W(work)->st.polish = v;
  W(work)->info.polish_time = 0.0;
  return 0;
}
c_int osqp_update_rho(OSQPWorkspace *work, c_float rho_new) {
  if (!work) return 1;
  Work &w = *W(work);
  if (rho_new <= 0) { fprintf(stderr, "ERROR in osqp_update_rho: rho must be positive\n"); return 1; }
  begin_update(w);
  double t0 = now_s();
  int e = set_rho(w, rho_new);
  w.info.update_time += now_s() - t0;
  return e;
}

}  // extern "C"
