// batch.cu -- batched engine for many small, independent QPs with one shared sparsity pattern
// (BASELINE.json config 5: 8192 MPC QPs, n = 30, m = 60; SURVEY.md 8b "Batch extension", 8e).
//
// One thread block per QP, everything the ADMM loop touches lives in shared memory as dense column-major
// matrices (P, A, the Cholesky factor of K = P + sigma I + A' diag(rho) A) plus the iterates; HBM is read once at
// the start of a solve and written once at the end, so the path is latency/FP64 bound, not HBM bound, and a batch
// shards across GPUs with no exchange at all.  The per-QP algorithm is libosqp 0.6.2's (Ruiz equilibration, rho
// vector, relaxed ADMM step with an exact KKT solve, update_info / check_termination incl. the infeasibility
// certificates, adaptive rho with refactorisation) and shares its decision rules with the single-QP engine
// (admm_rules.cuh).  The reference has no batch API: the entry points are declared in
// include/osqp_b200.h.
#include "admm_rules.cuh"
#include "osqp.h"
#include "osqp_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace osqpb200 {
namespace {

struct BatchDims {
  int n, m;
  int ldp, lda;      // leading dimensions (== 1 mod 16: conflict-free strided fp64 reads across threads)
  int nnzP, nnzA;    // stored entries of the shared patterns (P upper triangle, A)
  long long stride;  // doubles of per-QP state in HBM
  // offsets (doubles) inside the per-QP state
  int oP, oA, oL, oInvd, oQ, oL_, oU, oD, oE, oRho, oX, oZ, oY, oScal;  // oScal: c, rho, rho_updates
};

// shared-memory carve-up (doubles)
struct BatchSmem {
  double *P, *A, *L, *invd, *q, *l, *u, *D, *E, *Dinv, *Einv, *rho, *rhoinv;
  double *x, *z, *y, *xt, *zt, *xp, *zp, *dx, *dy, *tn, *tm, *red;
  int *ctype;
};

__device__ __forceinline__ void carve(BatchSmem &S, const BatchDims &d, double *base) {
  double *p = base;
  auto take = [&](int k) { double *r = p; p += (k + 1) & ~1; return r; };
  S.P = take(d.ldp * d.n); S.A = take(d.lda * d.n); S.L = take(d.ldp * d.n); S.invd = take(d.n);
  S.q = take(d.n); S.l = take(d.m); S.u = take(d.m); S.D = take(d.n); S.E = take(d.m); S.Dinv = take(d.n);
  S.Einv = take(d.m); S.rho = take(d.m); S.rhoinv = take(d.m);
  S.x = take(d.n); S.z = take(d.m); S.y = take(d.m); S.xt = take(d.n); S.zt = take(d.m); S.xp = take(d.n);
  S.zp = take(d.m); S.dx = take(d.n); S.dy = take(d.m); S.tn = take(d.n); S.tm = take(d.m);
  S.red = take(32 * 24 + 24);
  S.ctype = reinterpret_cast<int *>(take((d.m + 1) / 2 + 1));
}
size_t batch_smem_bytes(const BatchDims &d) {
  auto ev = [](int k) { return (size_t)((k + 1) & ~1); };
  size_t k = 2 * ev(d.ldp * d.n) + ev(d.lda * d.n) + 9 * ev(d.n) + 13 * ev(d.m) + ev(32 * 24 + 24) + ev((d.m + 1) / 2 + 1);
  return 8 * k + 64;
}

__device__ __forceinline__ double limit_scaling_b(double a) {
  a = a < kMinScaling ? 1.0 : a;
  return a > kMaxScaling ? kMaxScaling : a;
}

// block-wide reduction of NV scalars (bit k of maxmask: max, else sum); result broadcast to every thread
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], unsigned maxmask, double *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double a = v[k];
    const bool ismax = maxmask & (1u << k);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double t = __shfl_xor_sync(0xffffffffu, a, o);
      a = ismax ? fmax(a, t) : a + t;
    }
    if (lane == 0) scratch[warp * 24 + k] = a;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    const int k = threadIdx.x;
    const bool ismax = maxmask & (1u << k);
    double a = scratch[k];
    for (int w = 1; w < nwarps; w++) a = ismax ? fmax(a, scratch[w * 24 + k]) : a + scratch[w * 24 + k];
    scratch[32 * 24 + k] = a;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = scratch[32 * 24 + k];
  __syncthreads();
}

// K = P + sigma I + A' diag(rho) A (lower triangle incl. diagonal, into L), then in-place Cholesky L L' = K.
// Returns false (uniformly) if a pivot is not positive.
__device__ bool factor_K(const BatchSmem &S, const BatchDims &d, double sigma) {
  const int n = d.n, m = d.m, tid = threadIdx.x, nth = blockDim.x;
  for (int e = tid; e < n * n; e += nth) {
    const int j = e / n, i = e - j * n;  // column j, row i
    if (i < j) continue;
    double a = S.P[j * d.ldp + i] + (i == j ? sigma : 0.0);
    const double *ai = S.A + i * d.lda, *aj = S.A + j * d.lda;
    for (int k = 0; k < m; k++) a = fma(ai[k] * S.rho[k], aj[k], a);
    S.L[j * d.ldp + i] = a;
  }
  __syncthreads();
  bool ok = true;
  for (int k = 0; k < n; k++) {
    const double dkk = S.L[k * d.ldp + k];
    if (!(dkk > 0.0)) { ok = false; break; }  // uniform: every thread reads the same value
    const double r = sqrt(dkk), rinv = 1.0 / r;
    __syncthreads();
    if (tid == 0) { S.L[k * d.ldp + k] = r; S.invd[k] = rinv; }
    for (int i = k + 1 + tid; i < n; i += nth) S.L[k * d.ldp + i] *= rinv;
    __syncthreads();
    // trailing update: column j > k, rows i >= j
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += nth) {
      const int jj = e / rem, ii = e - jj * rem;
      if (ii < jj) continue;
      const int j = k + 1 + jj, i = k + 1 + ii;
      S.L[j * d.ldp + i] -= S.L[k * d.ldp + i] * S.L[k * d.ldp + j];
    }
    __syncthreads();
  }
  return ok;
}

// x <- K^{-1} b with b in S.tn (overwritten by the solution): forward then backward substitution, one barrier per
// column; thread i keeps entry i in a register.
__device__ __forceinline__ void solve_K(const BatchSmem &S, const BatchDims &d) {
  const int n = d.n, tid = threadIdx.x;
  double bi = tid < n ? S.tn[tid] : 0.0;
  for (int k = 0; k < n; k++) {
    if (tid == k) { bi *= S.invd[k]; S.tn[k] = bi; }
    __syncthreads();
    if (tid > k && tid < n) bi -= S.L[k * d.ldp + tid] * S.tn[k];
  }
  __syncthreads();
  for (int k = n - 1; k >= 0; k--) {
    if (tid == k) { bi *= S.invd[k]; S.tn[k] = bi; }
    __syncthreads();
    if (tid < k) bi -= S.L[tid * d.ldp + k] * S.tn[k];
  }
  __syncthreads();
}

__device__ __forceinline__ void set_rho_vec_b(const BatchSmem &S, const BatchDims &d, double rho, bool set_types) {
  for (int i = threadIdx.x; i < d.m; i += blockDim.x) {
    int t = S.ctype[i];
    if (set_types) {
      const double li = S.l[i], ui = S.u[i];
      t = (li < -kInfty * kMinScaling && ui > kInfty * kMinScaling) ? -1 : (ui - li < kRhoTol ? 1 : 0);
      S.ctype[i] = t;
    }
    const double r = t < 0 ? kRhoMin : (t == 1 ? kRhoEqOverIneq * rho : rho);
    S.rho[i] = r;
    S.rhoinv[i] = 1.0 / r;
  }
  __syncthreads();
}

// update_info (row a9) + the infeasibility products (row a10) on the dense matrices
__device__ void batch_info(const BatchSmem &S, const BatchDims &d, const SolveCfg &c, double cost_c, double cost_cinv,
                           InfoScalars &I) {
  const int n = d.n, m = d.m, tid = threadIdx.x;
  const bool unscale = c.scaling && !c.scaled_termination;
  double v[23];
#pragma unroll
  for (int k = 0; k < 23; k++) v[k] = 0.0;
  v[8] = -INFINITY;
  v[9] = -INFINITY;
  if (tid < m) {
    double Ax = 0.0, Adx = 0.0;
    for (int j = 0; j < n; j++) {
      const double a = S.A[j * d.lda + tid];
      Ax = fma(a, S.x[j], Ax);
      Adx = fma(a, S.dx[j], Adx);
    }
    const double zi = S.z[tid], ei = unscale ? S.Einv[tid] : 1.0, Ei = unscale ? S.E[tid] : 1.0;
    const double li = S.l[tid], ui = S.u[tid], dyi = S.dy[tid];
    const double pr = fabs(Ax - zi);
    v[0] = ei * pr; v[1] = pr; v[2] = ei * fabs(zi); v[3] = fabs(zi); v[4] = ei * fabs(Ax); v[5] = fabs(Ax);
    v[6] = Ei * fabs(dyi);
    v[7] = ui * fmax(dyi, 0.0) + li * fmin(dyi, 0.0);
    const double adx = ei * Adx;
    if (ui < kInfty * kMinScaling) v[8] = adx;
    if (li > -kInfty * kMinScaling) v[9] = -adx;
  }
  if (tid < n) {
    double Px = 0.0, Pdx = 0.0, Aty = 0.0, Atdy = 0.0;
    for (int j = 0; j < n; j++) {
      const double a = S.P[j * d.ldp + tid];
      Px = fma(a, S.x[j], Px);
      Pdx = fma(a, S.dx[j], Pdx);
    }
    const double *col = S.A + tid * d.lda;
    for (int i = 0; i < m; i++) {
      Aty = fma(col[i], S.y[i], Aty);
      Atdy = fma(col[i], S.dy[i], Atdy);
    }
    const double qj = S.q[tid], xj = S.x[tid], dxj = S.dx[tid];
    const double di = unscale ? S.Dinv[tid] : 1.0, Dj = unscale ? S.D[tid] : 1.0;
    const double dr = fabs(qj + Px + Aty);
    v[10] = di * dr; v[11] = dr; v[12] = di * fabs(qj); v[13] = fabs(qj); v[14] = di * fabs(Aty); v[15] = fabs(Aty);
    v[16] = di * fabs(Px); v[17] = fabs(Px); v[18] = xj * (0.5 * Px + qj); v[19] = Dj * fabs(dxj); v[20] = qj * dxj;
    v[21] = di * fabs(Pdx); v[22] = di * fabs(Atdy);
  }
  block_reduce<23>(v, 0x7FFFFFu & ~((1u << 7) | (1u << 18) | (1u << 20)), S.red);
  I.pri_t = v[0]; I.pri_r = v[1]; I.nz_t = v[2]; I.nz_r = v[3]; I.nAx_t = v[4]; I.nAx_r = v[5];
  I.ndy_t = v[6]; I.lhs = v[7]; I.maxU_t = v[8]; I.maxNegL_t = v[9];
  I.dua_t = v[10]; I.dua_r = v[11]; I.nq_t = v[12]; I.nq_r = v[13]; I.nAty_t = v[14]; I.nAty_r = v[15];
  I.nPx_t = v[16]; I.nPx_r = v[17]; I.obj = v[18]; I.ndx_t = v[19]; I.qdx = v[20]; I.nPdx_t = v[21];
  I.nAtdy_t = v[22];
  I.obj_val = c.scaling ? I.obj * cost_cinv : I.obj;
  I.pri_res = (m == 0) ? 0.0 : I.pri_t;
  I.dua_res = unscale ? cost_cinv * I.dua_t : I.dua_t;
}

// ------------------------------------------------------------------ setup: dense load, Ruiz, rho vector, factor
__global__ void batch_setup_kernel(BatchDims d, double *state, const long long *Pp, const long long *Pi,
                                   const long long *Ap, const long long *Ai, const double *Px, const double *Ax,
                                   const double *q, const double *l, const double *u, int scaling_iters, double rho0,
                                   double sigma, int *fail) {
  extern __shared__ __align__(16) double smem_b[];
  BatchSmem S;
  carve(S, d, smem_b);
  const int n = d.n, m = d.m, tid = threadIdx.x, nth = blockDim.x;
  const long long b = blockIdx.x;
  for (int e = tid; e < d.ldp * n; e += nth) S.P[e] = 0.0;
  for (int e = tid; e < d.lda * n; e += nth) S.A[e] = 0.0;
  __syncthreads();
  // shared patterns (0-based CSC), per-QP values
  for (int j = tid; j < n; j += nth) {
    for (long long k = Pp[j]; k < Pp[j + 1]; k++) {
      const int i = (int)Pi[k];
      const double a = Px[b * d.nnzP + k];
      S.P[j * d.ldp + i] = a;  // upper triangle entry (i <= j) ...
      S.P[i * d.ldp + j] = a;  // ... and its mirror (columns are private to thread j only for the first store)
    }
  }
  __syncthreads();
  for (int j = tid; j < n; j += nth)
    for (long long k = Ap[j]; k < Ap[j + 1]; k++) S.A[j * d.lda + (int)Ai[k]] = Ax[b * d.nnzA + k];
  for (int j = tid; j < n; j += nth) { S.q[j] = q[b * n + j]; S.D[j] = 1.0; }
  for (int i = tid; i < m; i += nth) { S.l[i] = l[b * m + i]; S.u[i] = u[b * m + i]; S.E[i] = 1.0; }
  __syncthreads();
  double cost = 1.0;
  for (int it = 0; it < scaling_iters; it++) {
    // column norms of [P A'; A 0]
    if (tid < n) {
      double mx = 0.0;
      for (int i = 0; i < n; i++) mx = fmax(mx, fabs(S.P[tid * d.ldp + i]));
      for (int i = 0; i < m; i++) mx = fmax(mx, fabs(S.A[tid * d.lda + i]));
      S.tn[tid] = 1.0 / sqrt(limit_scaling_b(mx));
    }
    if (tid < m) {
      double mx = 0.0;
      for (int j = 0; j < n; j++) mx = fmax(mx, fabs(S.A[j * d.lda + tid]));
      S.tm[tid] = 1.0 / sqrt(limit_scaling_b(mx));
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += nth) {
      const int j = e / n, i = e - j * n;
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      S.P[j * d.ldp + i] = (S.P[j * d.ldp + i] * S.tn[lo]) * S.tn[hi];
    }
    for (int e = tid; e < n * m; e += nth) {
      const int j = e / m, i = e - j * m;
      S.A[j * d.lda + i] = (S.A[j * d.lda + i] * S.tm[i]) * S.tn[j];
    }
    if (tid < n) { S.q[tid] *= S.tn[tid]; S.D[tid] *= S.tn[tid]; }
    if (tid < m) S.E[tid] *= S.tm[tid];
    __syncthreads();
    // cost normalisation
    double r2[2] = {0.0, 0.0};
    if (tid < n) {
      double mx = 0.0;
      for (int i = 0; i < n; i++) mx = fmax(mx, fabs(S.P[tid * d.ldp + i]));
      r2[0] = mx;
      r2[1] = fabs(S.q[tid]);
    }
    block_reduce<2>(r2, 0x2u, S.red);
    double ct = limit_scaling_b(fmax(r2[0] / (double)n, limit_scaling_b(r2[1])));
    ct = 1.0 / ct;
    for (int e = tid; e < n * n; e += nth) { const int j = e / n, i = e - j * n; S.P[j * d.ldp + i] *= ct; }
    if (tid < n) S.q[tid] *= ct;
    cost *= ct;
    __syncthreads();
  }
  if (tid < m) { S.l[tid] *= S.E[tid]; S.u[tid] *= S.E[tid]; }
  __syncthreads();
  set_rho_vec_b(S, d, rho0, true);
  const bool ok = factor_K(S, d, sigma);
  if (!ok && tid == 0) atomicExch(fail, 1);
  // state -> HBM
  double *st = state + b * d.stride;
  for (int e = tid; e < d.ldp * n; e += nth) { st[d.oP + e] = S.P[e]; st[d.oL + e] = S.L[e]; }
  for (int e = tid; e < d.lda * n; e += nth) st[d.oA + e] = S.A[e];
  for (int j = tid; j < n; j += nth) {
    st[d.oInvd + j] = S.invd[j]; st[d.oQ + j] = S.q[j]; st[d.oD + j] = S.D[j]; st[d.oX + j] = 0.0;
  }
  for (int i = tid; i < m; i += nth) {
    st[d.oL_ + i] = S.l[i]; st[d.oU + i] = S.u[i]; st[d.oE + i] = S.E[i]; st[d.oRho + i] = (double)S.ctype[i];
    st[d.oZ + i] = 0.0; st[d.oY + i] = 0.0;
  }
  if (tid == 0) { st[d.oScal] = cost; st[d.oScal + 1] = rho0; st[d.oScal + 2] = 0.0; }
}

// q / bounds / warm-start updates on the resident state
__global__ void batch_update_kernel(BatchDims d, double *state, const double *q, const double *l, const double *u,
                                    const double *x, const double *y, int scaling) {
  extern __shared__ __align__(16) double smem_b[];
  const int n = d.n, m = d.m, tid = threadIdx.x, nth = blockDim.x;
  const long long b = blockIdx.x;
  double *st = state + b * d.stride;
  const double c = st[d.oScal];
  if (q) for (int j = tid; j < n; j += nth) st[d.oQ + j] = (q[b * n + j] * st[d.oD + j]) * c;
  if (l) for (int i = tid; i < m; i += nth) st[d.oL_ + i] = st[d.oE + i] * l[b * m + i];
  if (u) for (int i = tid; i < m; i += nth) st[d.oU + i] = st[d.oE + i] * u[b * m + i];
  if (x) {
    double *xs = smem_b;
    for (int j = tid; j < n; j += nth) { xs[j] = scaling ? x[b * n + j] / st[d.oD + j] : x[b * n + j]; st[d.oX + j] = xs[j]; }
    __syncthreads();
    for (int i = tid; i < m; i += nth) {
      double a = 0.0;
      for (int j = 0; j < n; j++) a = fma(st[d.oA + j * d.lda + i], xs[j], a);
      st[d.oZ + i] = a;
    }
  }
  if (y) for (int i = tid; i < m; i += nth) st[d.oY + i] = scaling ? (y[b * m + i] / st[d.oE + i]) * c : y[b * m + i];
}

// ------------------------------------------------------------------ solve: the whole ADMM loop of one QP per block
__global__ void batch_solve_kernel(BatchDims d, double *state, SolveCfg c, long long adaptive_interval,
                                   int bounds_changed, double *x_out, double *y_out, OSQPB200BatchInfo *info_out) {
  extern __shared__ __align__(16) double smem_b[];
  BatchSmem S;
  carve(S, d, smem_b);
  const int n = d.n, m = d.m, tid = threadIdx.x, nth = blockDim.x;
  const long long b = blockIdx.x;
  double *st = state + b * d.stride;
  for (int e = tid; e < d.ldp * n; e += nth) { S.P[e] = st[d.oP + e]; S.L[e] = st[d.oL + e]; }
  for (int e = tid; e < d.lda * n; e += nth) S.A[e] = st[d.oA + e];
  for (int j = tid; j < n; j += nth) {
    S.invd[j] = st[d.oInvd + j]; S.q[j] = st[d.oQ + j]; S.D[j] = st[d.oD + j]; S.Dinv[j] = 1.0 / S.D[j];
    S.x[j] = c.warm_start ? st[d.oX + j] : 0.0; S.dx[j] = 0.0;
  }
  for (int i = tid; i < m; i += nth) {
    S.l[i] = st[d.oL_ + i]; S.u[i] = st[d.oU + i]; S.E[i] = st[d.oE + i]; S.Einv[i] = 1.0 / S.E[i];
    S.ctype[i] = (int)st[d.oRho + i];
    S.z[i] = c.warm_start ? st[d.oZ + i] : 0.0; S.y[i] = c.warm_start ? st[d.oY + i] : 0.0; S.dy[i] = 0.0;
  }
  const double cost_c = st[d.oScal], cost_cinv = 1.0 / cost_c;
  double rho = st[d.oScal + 1];
  long long rho_updates = (long long)st[d.oScal + 2];
  __syncthreads();
  bool refactor = false;
  if (bounds_changed) {  // update_rho_vec: constraint types may have changed with the bounds
    int changed = 0;
    for (int i = tid; i < m; i += nth) {
      const double li = S.l[i], ui = S.u[i];
      const int t = (li < -kInfty * kMinScaling && ui > kInfty * kMinScaling) ? -1 : (ui - li < kRhoTol ? 1 : 0);
      if (t != S.ctype[i]) { S.ctype[i] = t; changed = 1; }
    }
    refactor = __syncthreads_or(changed);
  }
  set_rho_vec_b(S, d, rho, false);
  long long status = ST_UNSOLVED;
  if (refactor && !factor_K(S, d, c.sigma)) status = ST_NON_CVX;

  InfoScalars I;
  I.pri_res = I.dua_res = I.obj_val = 0.0;
  long long it = 0, info_iter = 0;
  double rho_est = rho;
  bool checked = false;
  if (status == ST_UNSOLVED)
    for (it = 1; it <= c.max_iter; it++) {
      // rhs: tn = sigma x - q + A'(rho z - y)
      if (tid < m) S.tm[tid] = S.rho[tid] * S.z[tid] - S.y[tid];
      if (tid < n) S.xp[tid] = S.x[tid];
      if (tid < m) S.zp[tid] = S.z[tid];
      __syncthreads();
      if (tid < n) {
        double a = c.sigma * S.x[tid] - S.q[tid];
        const double *col = S.A + tid * d.lda;
        for (int i = 0; i < m; i++) a = fma(col[i], S.tm[i], a);
        S.tn[tid] = a;
      }
      __syncthreads();
      solve_K(S, d);  // tn = x_tilde
      if (tid < n) S.xt[tid] = S.tn[tid];
      if (tid < m) {
        double a = 0.0;
        for (int j = 0; j < n; j++) a = fma(S.A[j * d.lda + tid], S.tn[j], a);
        S.zt[tid] = a;
      }
      __syncthreads();
      if (tid < n) {
        const double xn = c.alpha * S.xt[tid] + (1.0 - c.alpha) * S.xp[tid];
        S.dx[tid] = xn - S.xp[tid];
        S.x[tid] = xn;
      }
      if (tid < m) {
        const double zh = c.alpha * S.zt[tid] + (1.0 - c.alpha) * S.zp[tid];
        const double yi = S.y[tid], li = S.l[tid], ui = S.u[tid];
        const double zn = fmin(fmax(zh + S.rhoinv[tid] * yi, li), ui);
        double dyi = S.rho[tid] * (zh - zn);
        S.y[tid] = yi + dyi;
        S.z[tid] = zn;
        if (ui > kInfty * kMinScaling) {
          if (li < -kInfty * kMinScaling) dyi = 0.0;
          else dyi = fmin(dyi, 0.0);
        } else if (li < -kInfty * kMinScaling) {
          dyi = fmax(dyi, 0.0);
        }
        S.dy[tid] = dyi;
      }
      __syncthreads();
      checked = c.check_termination && (it % c.check_termination == 0);
      const bool adapt = c.adaptive_rho && adaptive_interval && (it % adaptive_interval == 0);
      if (checked || adapt) {
        batch_info(S, d, c, cost_c, cost_cinv, I);
        info_iter = it;
        if (checked) {
          status = check_termination(I, c, m, cost_c, cost_cinv, false);
          if (status != ST_UNSOLVED) break;
        }
      }
      if (adapt) {
        const double rho_new = rho_estimate(I, rho);
        rho_est = rho_new;
        if (rho_new > rho * c.adaptive_rho_tolerance || rho_new < rho / c.adaptive_rho_tolerance) {
          rho = fmin(fmax(rho_new, kRhoMin), kRhoMax);
          rho_updates++;
          set_rho_vec_b(S, d, rho, false);
          if (!factor_K(S, d, c.sigma)) { status = ST_NON_CVX; break; }
          refactor = true;
        }
      }
    }
  if (status == ST_UNSOLVED) {  // the loop ran out of iterations (libosqp: last check, then the approximate one)
    if (!checked) {
      batch_info(S, d, c, cost_c, cost_cinv, I);
      info_iter = it - 1;
      status = check_termination(I, c, m, cost_c, cost_cinv, false);
    }
    if (status == ST_UNSOLVED) {
      const long long s2 = check_termination(I, c, m, cost_c, cost_cinv, true);
      status = (s2 != ST_UNSOLVED) ? s2 : ST_MAX_ITER;
    }
  }
  if (status != ST_NON_CVX) rho_est = rho_estimate(I, rho);
  double obj_val = I.obj_val;
  if (status == ST_NON_CVX) obj_val = nan("");
  const bool pinf = (status == ST_PINF || status == ST_PINF_INACC), dinf = (status == ST_DINF || status == ST_DINF_INACC);
  if (pinf) obj_val = kInfty;
  if (dinf) obj_val = -kInfty;
  const bool has_sol = !(pinf || dinf || status == ST_NON_CVX);
  const bool unscale = c.scaling && !c.scaled_termination;
  // store_solution (row a16): solution or NaN + certificate, iterates kept for the next warm start
  for (int j = tid; j < n; j += nth) {
    double out;
    if (has_sol) out = c.scaling ? S.D[j] * S.x[j] : S.x[j];
    else if (dinf) out = (unscale ? S.D[j] : 1.0) * S.dx[j] / I.ndx_t;  // certificate in place of x
    else out = nan("");
    x_out[b * n + j] = out;
    st[d.oX + j] = has_sol ? S.x[j] : 0.0;
  }
  for (int i = tid; i < m; i += nth) {
    double out;
    if (has_sol) out = c.scaling ? cost_cinv * S.E[i] * S.y[i] : S.y[i];
    else if (pinf) out = (unscale ? S.E[i] : 1.0) * S.dy[i] / I.ndy_t;  // certificate in place of y
    else out = nan("");
    y_out[b * m + i] = out;
    st[d.oZ + i] = has_sol ? S.z[i] : 0.0;
    st[d.oY + i] = has_sol ? S.y[i] : 0.0;
    st[d.oRho + i] = (double)S.ctype[i];
  }
  if (refactor) {
    for (int e = tid; e < d.ldp * n; e += nth) st[d.oL + e] = S.L[e];
    for (int j = tid; j < n; j += nth) st[d.oInvd + j] = S.invd[j];
  }
  if (tid == 0) {
    st[d.oScal + 1] = rho;
    st[d.oScal + 2] = (double)rho_updates;
    OSQPB200BatchInfo &o = info_out[b];
    o.iter = info_iter;
    o.status_val = status;
    o.obj_val = obj_val;
    o.pri_res = I.pri_res;
    o.dua_res = I.dua_res;
    o.rho_estimate = rho_est;
    o.rho_updates = rho_updates;
  }
}

// =================================================================== fast path: one WARP per QP (n <= 32, m <= 64)
// Lane j owns entry j of every n-vector and entries j, j + 32 of every m-vector in registers; the matrices stay sparse
// (the shared pattern lives once per block in shared memory, each warp keeps its own values), only the INVERSE of the
// Cholesky factor of K is dense (row-major, ld 33: conflict-free by rows and by columns).  The per-iteration solve is
// two triangular matrix-vector products without any block barrier, so many QPs are in flight per SM.
constexpr int kFastWarps = 4;  // QPs per thread block (1 was measured: fewer resident warps, 25 % slower)
constexpr int kLdl = 33;

struct FastDims {
  int n, m, nnzP, nnzA, nnzPf;  // nnzPf: entries of the full symmetric pattern of P
  long long stride;             // doubles of per-QP state in HBM
  int oPv, oAv, oL, oInvd, oQ, oLo, oUp, oD, oE, oCt, oX, oZ, oY, oScal;
  // shared pattern (shorts), offsets into the pattern array
  int pAcp, pAri, pArp, pAcc, pAcq, pPrp, pPcc, pPcq, plen;
  // ELL view of the same pattern for the products of the ADMM loop (branch-free, fixed trip count): slot t of row i of
  // A is entry pEAp[t * 64 + i] of the value array times entry pEAc[t * 64 + i] of the vector; slot t of column j:
  // pECp / pECr [t * 32 + j]; slot t of row j of the full P: pEPp / pEPc [t * 32 + j].  Empty slots point at the zero
  // stored behind the values (Av[nnzA], Pv[nnzP]).  ell == 0: rows too long for this (RA, CA, RP > kEllMax).
  int ell, RA, CA, RP, pEAp, pEAc, pECp, pECr, pEPp, pEPc;
};
constexpr int kEllMax = 8;

struct FastWarp {  // per-warp shared memory
  double *Av, *Pv, *L, *invd, *vn0, *vn1, *vm0, *vm1;
};

__device__ __forceinline__ int fast_warp_doubles(const FastDims &d) {
  return ((d.nnzA + 2) & ~1) + ((d.nnzP + 2) & ~1) + 32 * kLdl + 32 + 32 + 32 + 64 + 64 + 2;
}
size_t fast_smem_bytes(const FastDims &d) {
  const int per = ((d.nnzA + 2) & ~1) + ((d.nnzP + 2) & ~1) + 32 * kLdl + 32 + 32 + 32 + 64 + 64 + 2;
  return (size_t)kFastWarps * per * 8 + (((size_t)d.plen * 2 + 15) & ~(size_t)15) + 64;
}
__device__ __forceinline__ void fast_carve(FastWarp &W, const FastDims &d, double *base, int warp) {
  double *p = base + (size_t)warp * fast_warp_doubles(d);
  W.Av = p; p += (d.nnzA + 2) & ~1;  // one zero behind the values: the target of empty ELL slots
  W.Pv = p; p += (d.nnzP + 2) & ~1;
  W.L = p; p += 32 * kLdl;
  W.invd = p; p += 32;
  W.vn0 = p; p += 32;
  W.vn1 = p; p += 32;
  W.vm0 = p; p += 64;
  W.vm1 = p;
}

__device__ __forceinline__ double wmax(double a) {
#pragma unroll
  for (int o = 16; o; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
  return a;
}
__device__ __forceinline__ double wsum(double a) {
#pragma unroll
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

struct FastPat {  // the shared pattern in shared memory
  const short *Acp, *Ari;            // A by columns (the caller's CSC): column pointers, row indices
  const short *Arp, *Acc, *Acq;      // A by rows: row pointers, column indices, position in the CSC value array
  const short *Prp, *Pcc, *Pcq;      // full symmetric P by rows: pointers, columns, position in the triu value array
  const short *EAp, *EAc, *ECp, *ECr, *EPp, *EPc;  // ELL view (FastDims)
  int ell, RA, CA, RP;
};
__device__ __forceinline__ void fast_pat(FastPat &Q, const FastDims &d, const short *p) {
  Q.Acp = p + d.pAcp; Q.Ari = p + d.pAri; Q.Arp = p + d.pArp; Q.Acc = p + d.pAcc; Q.Acq = p + d.pAcq;
  Q.Prp = p + d.pPrp; Q.Pcc = p + d.pPcc; Q.Pcq = p + d.pPcq;
  Q.EAp = p + d.pEAp; Q.EAc = p + d.pEAc; Q.ECp = p + d.pECp; Q.ECr = p + d.pECr; Q.EPp = p + d.pEPp; Q.EPc = p + d.pEPc;
  Q.ell = d.ell; Q.RA = d.RA; Q.CA = d.CA; Q.RP = d.RP;
}

// row i of A times the n-vector in shared memory `v`
__device__ __forceinline__ double arow(const FastPat &Q, const double *Av, const double *v, int i, int m) {
  double a = 0.0;
  if (Q.ell) {  // i < 64 always; rows >= m hold empty slots only
#pragma unroll 4
    for (int t = 0; t < Q.RA; t++) a = fma(Av[Q.EAp[t * 64 + i]], v[Q.EAc[t * 64 + i]], a);
  } else if (i < m) {
#pragma unroll 1
    for (int k = Q.Arp[i]; k < Q.Arp[i + 1]; k++) a = fma(Av[Q.Acq[k]], v[Q.Acc[k]], a);
  }
  return a;
}
// column j of A times the m-vector in shared memory `v`  (= row j of A')
__device__ __forceinline__ double acol(const FastPat &Q, const double *Av, const double *v, int j, int n) {
  double a = 0.0;
  if (Q.ell) {
#pragma unroll 4
    for (int t = 0; t < Q.CA; t++) a = fma(Av[Q.ECp[t * 32 + j]], v[Q.ECr[t * 32 + j]], a);
  } else if (j < n) {
#pragma unroll 1
    for (int k = Q.Acp[j]; k < Q.Acp[j + 1]; k++) a = fma(Av[k], v[Q.Ari[k]], a);
  }
  return a;
}
__device__ __forceinline__ double prow(const FastPat &Q, const double *Pv, const double *v, int j, int n) {
  double a = 0.0;
  if (Q.ell) {
#pragma unroll 4
    for (int t = 0; t < Q.RP; t++) a = fma(Pv[Q.EPp[t * 32 + j]], v[Q.EPc[t * 32 + j]], a);
  } else if (j < n) {
#pragma unroll 1
    for (int k = Q.Prp[j]; k < Q.Prp[j + 1]; k++) a = fma(Pv[Q.Pcq[k]], v[Q.Pcc[k]], a);
  }
  return a;
}

// K = P + sigma I + A' diag(rho) A into W.L (row-major, lower triangle used), then Cholesky; lane j builds column j
// of K on its own (no conflicts, fixed order).  rho of rows lane / lane + 32 in r0 / r1.  Returns false on a
// non-positive pivot (uniform over the warp).
__device__ __noinline__ bool fast_factor(const FastWarp &W, const FastPat &Q, const FastDims &d, double sigma, double r0, double r1) {
  const int n = d.n, m = d.m, lane = threadIdx.x & 31;
  if (lane < m) W.vm0[lane] = r0;
  if (lane + 32 < m) W.vm0[lane + 32] = r1;
  for (int e = lane; e < 32 * kLdl; e += 32) W.L[e] = 0.0;
  __syncwarp();
  if (lane < n) {
    const int j = lane;
    for (int k = Q.Prp[j]; k < Q.Prp[j + 1]; k++) W.L[Q.Pcc[k] * kLdl + j] += W.Pv[Q.Pcq[k]];  // K[c][j], column j
    W.L[j * kLdl + j] += sigma;
    for (int k = Q.Acp[j]; k < Q.Acp[j + 1]; k++) {
      const int row = Q.Ari[k];
      const double s = W.vm0[row] * W.Av[k];
      for (int t = Q.Arp[row]; t < Q.Arp[row + 1]; t++) W.L[Q.Acc[t] * kLdl + j] += s * W.Av[Q.Acq[t]];
    }
  }
  __syncwarp();
  // right-looking Cholesky on the lower triangle: lane i owns row i
  bool ok = true;
  for (int k = 0; k < n; k++) {
    const double dkk = W.L[k * kLdl + k];
    if (!(dkk > 0.0)) { ok = false; break; }
    const double r = sqrt(dkk), rinv = 1.0 / r;
    __syncwarp();
    double lik = 0.0;
    if (lane == k) { W.L[k * kLdl + k] = r; W.invd[k] = rinv; }
    if (lane > k && lane < n) { lik = W.L[lane * kLdl + k] * rinv; W.L[lane * kLdl + k] = lik; }
    __syncwarp();
    if (lane > k && lane < n)
      for (int j = k + 1; j <= lane; j++) W.L[lane * kLdl + j] -= lik * W.L[j * kLdl + k];
    __syncwarp();
  }
  // Explicit inverse of the factor, in place: W.L <- L^{-1} (lower triangular, zeros above the diagonal, rows >= n
  // zero), so that the per-iteration solve is two short matrix-vector products without a dependent chain over the
  // columns (fast_solve).  X = L^{-1} obeys X[i][j] = -(sum_{k=j+1..i} X[i][k] L[k][j]) / L[j][j]; going through the
  // columns from the last to the first, row i (lane i) only needs its own already inverted entries and column j of
  // the original L, which is overwritten after every lane has read it.
  if (ok) {
    if (lane < n) for (int j = lane + 1; j < 32; j++) W.L[lane * kLdl + j] = 0.0;
    __syncwarp();
    for (int j = n - 1; j >= 0; j--) {
      const double dj = W.invd[j];
      double v = 0.0;
      if (lane == j) v = dj;
      else if (lane > j && lane < n) {
        double a = 0.0;
        for (int k = j + 1; k <= lane; k++) a = fma(W.L[lane * kLdl + k], W.L[k * kLdl + j], a);
        v = -a * dj;
      }
      __syncwarp();
      if (lane >= j && lane < n) W.L[lane * kLdl + j] = v;
      __syncwarp();
    }
  }
  __syncwarp();
  return ok;
}

// b (entry `lane` in a register) <- K^{-1} b = L^{-T} (L^{-1} b) with the explicit inverse factor of fast_factor:
// t = L^{-1} b (lane j reads row j: stride kLdl, conflict-free) and x = L^{-T} t (lane j reads column j: consecutive
// words).  The vectors go through `va` / `vb` (shared memory, broadcast reads); entries above the diagonal are stored
// zeros, so both loops are branch-free, and each runs on four independent accumulators -- the critical path of the
// solve is ~2 x (n / 4) fused multiply-adds instead of 2 n dependent shuffle steps of a substitution.
__device__ __forceinline__ double fast_solve(const FastWarp &W, int n, double b, double *va, double *vb) {
  const int lane = threadIdx.x & 31;
  va[lane] = lane < n ? b : 0.0;
  __syncwarp();
  const double *row = W.L + lane * kLdl;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int k = 0;
  for (; k + 4 <= n; k += 4) {
    a0 = fma(row[k], va[k], a0);
    a1 = fma(row[k + 1], va[k + 1], a1);
    a2 = fma(row[k + 2], va[k + 2], a2);
    a3 = fma(row[k + 3], va[k + 3], a3);
  }
  for (; k < n; k++) a0 = fma(row[k], va[k], a0);
  vb[lane] = (a0 + a1) + (a2 + a3);
  __syncwarp();
  const double *col = W.L + lane;
  a0 = a1 = a2 = a3 = 0.0;
  for (k = 0; k + 4 <= n; k += 4) {
    a0 = fma(col[k * kLdl], vb[k], a0);
    a1 = fma(col[(k + 1) * kLdl], vb[k + 1], a1);
    a2 = fma(col[(k + 2) * kLdl], vb[k + 2], a2);
    a3 = fma(col[(k + 3) * kLdl], vb[k + 3], a3);
  }
  for (; k < n; k++) a0 = fma(col[k * kLdl], vb[k], a0);
  return (a0 + a1) + (a2 + a3);
}

__device__ __forceinline__ int ctype_of(double l, double u) {
  return (l < -kInfty * kMinScaling && u > kInfty * kMinScaling) ? -1 : (u - l < kRhoTol ? 1 : 0);
}
__device__ __forceinline__ double rho_of(int t, double rho) {
  return t < 0 ? kRhoMin : (t == 1 ? kRhoEqOverIneq * rho : rho);
}

__global__ void __launch_bounds__(32 * kFastWarps) batch_fast_setup_kernel(
    FastDims d, long long count, double *state, const short *pattern, const double *Px, const double *Ax,
    const double *q, const double *l, const double *u, int scaling_iters, double rho0, double sigma, int *fail) {
  extern __shared__ __align__(16) double smem_b[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = d.n, m = d.m;
  short *pat = reinterpret_cast<short *>(smem_b + (size_t)kFastWarps * fast_warp_doubles(d));
  for (int e = threadIdx.x; e < d.plen; e += blockDim.x) pat[e] = pattern[e];
  __syncthreads();
  const long long b = (long long)blockIdx.x * kFastWarps + warp;
  if (b >= count) return;
  FastWarp W;
  fast_carve(W, d, smem_b, warp);
  FastPat Q;
  fast_pat(Q, d, pat);
  for (int k = lane; k < d.nnzA; k += 32) W.Av[k] = Ax[b * d.nnzA + k];
  for (int k = lane; k < d.nnzP; k += 32) W.Pv[k] = Px[b * d.nnzP + k];
  if (lane == 0) { W.Av[d.nnzA] = 0.0; W.Pv[d.nnzP] = 0.0; }
  double qj = lane < n ? q[b * n + lane] : 0.0, Dj = 1.0;
  double E0 = 1.0, E1 = 1.0, cost = 1.0;
  __syncwarp();
  for (int it = 0; it < scaling_iters; it++) {
    // column norms of [P A'; A 0]: lane j -> column j (P is symmetric: row j of the full pattern), lanes -> rows of A
    double cn = 0.0;
    if (lane < n) {
      for (int k = Q.Prp[lane]; k < Q.Prp[lane + 1]; k++) cn = fmax(cn, fabs(W.Pv[Q.Pcq[k]]));
      for (int k = Q.Acp[lane]; k < Q.Acp[lane + 1]; k++) cn = fmax(cn, fabs(W.Av[k]));
    }
    double rn0 = 0.0, rn1 = 0.0;
    if (lane < m) for (int k = Q.Arp[lane]; k < Q.Arp[lane + 1]; k++) rn0 = fmax(rn0, fabs(W.Av[Q.Acq[k]]));
    if (lane + 32 < m) for (int k = Q.Arp[lane + 32]; k < Q.Arp[lane + 33]; k++) rn1 = fmax(rn1, fabs(W.Av[Q.Acq[k]]));
    const double dt = 1.0 / sqrt(limit_scaling_b(cn)), e0 = 1.0 / sqrt(limit_scaling_b(rn0)), e1 = 1.0 / sqrt(limit_scaling_b(rn1));
    if (lane < n) W.vn0[lane] = dt;
    if (lane < m) W.vm0[lane] = e0;
    if (lane + 32 < m) W.vm0[lane + 32] = e1;
    __syncwarp();
    // P <- Dt P Dt (smaller index first, as the single-QP engine), A <- Et A Dt: lane j scales column j of A and of triu(P)
    if (lane < n) {
      for (int k = Q.Acp[lane]; k < Q.Acp[lane + 1]; k++) W.Av[k] = (W.Av[k] * W.vm0[Q.Ari[k]]) * dt;
    }
    __syncwarp();
    // triu(P) values: entry k of the triu array is reached through the full pattern rows with Pcc >= row (once)
    if (lane < n)
      for (int k = Q.Prp[lane]; k < Q.Prp[lane + 1]; k++) {
        const int c = Q.Pcc[k];
        if (c >= lane) W.Pv[Q.Pcq[k]] = (W.Pv[Q.Pcq[k]] * dt) * W.vn0[c];  // row = lane <= c
      }
    qj *= dt; Dj *= dt; E0 *= e0; E1 *= e1;
    __syncwarp();
    // cost normalisation
    double pm = 0.0;
    if (lane < n) for (int k = Q.Prp[lane]; k < Q.Prp[lane + 1]; k++) pm = fmax(pm, fabs(W.Pv[Q.Pcq[k]]));
    const double mean = wsum(lane < n ? pm : 0.0) / (double)n, qn = wmax(lane < n ? fabs(qj) : 0.0);
    const double ct = 1.0 / limit_scaling_b(fmax(mean, limit_scaling_b(qn)));
    for (int k = lane; k < d.nnzP; k += 32) W.Pv[k] *= ct;
    qj *= ct;
    cost *= ct;
    __syncwarp();
  }
  double l0 = lane < m ? E0 * l[b * m + lane] : 0.0, u0 = lane < m ? E0 * u[b * m + lane] : 0.0;
  double l1 = lane + 32 < m ? E1 * l[b * m + lane + 32] : 0.0, u1 = lane + 32 < m ? E1 * u[b * m + lane + 32] : 0.0;
  const int t0 = ctype_of(l0, u0), t1 = ctype_of(l1, u1);
  const bool ok = fast_factor(W, Q, d, sigma, rho_of(t0, rho0), rho_of(t1, rho0));
  if (!ok && lane == 0) atomicExch(fail, 1);
  double *st = state + b * d.stride;
  for (int k = lane; k < d.nnzA; k += 32) st[d.oAv + k] = W.Av[k];
  for (int k = lane; k < d.nnzP; k += 32) st[d.oPv + k] = W.Pv[k];
  for (int e = lane; e < 32 * kLdl; e += 32) st[d.oL + e] = W.L[e];
  if (lane < n) { st[d.oInvd + lane] = W.invd[lane]; st[d.oQ + lane] = qj; st[d.oD + lane] = Dj; st[d.oX + lane] = 0.0; }
  if (lane < m) { st[d.oLo + lane] = l0; st[d.oUp + lane] = u0; st[d.oE + lane] = E0; st[d.oCt + lane] = (double)t0; st[d.oZ + lane] = 0.0; st[d.oY + lane] = 0.0; }
  if (lane + 32 < m) {
    const int i = lane + 32;
    st[d.oLo + i] = l1; st[d.oUp + i] = u1; st[d.oE + i] = E1; st[d.oCt + i] = (double)t1; st[d.oZ + i] = 0.0; st[d.oY + i] = 0.0;
  }
  if (lane == 0) { st[d.oScal] = cost; st[d.oScal + 1] = rho0; st[d.oScal + 2] = 0.0; }
}

__global__ void __launch_bounds__(32 * kFastWarps) batch_fast_update_kernel(
    FastDims d, long long count, double *state, const short *pattern, const double *q, const double *l, const double *u,
    const double *x, const double *y, int scaling) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = d.n, m = d.m;
  const long long b = (long long)blockIdx.x * kFastWarps + warp;
  if (b >= count) return;
  double *st = state + b * d.stride;
  const double c = st[d.oScal];
  if (q && lane < n) st[d.oQ + lane] = (q[b * n + lane] * st[d.oD + lane]) * c;
  for (int i = lane; i < m; i += 32) {
    if (l) st[d.oLo + i] = st[d.oE + i] * l[b * m + i];
    if (u) st[d.oUp + i] = st[d.oE + i] * u[b * m + i];
    if (y) st[d.oY + i] = scaling ? (y[b * m + i] / st[d.oE + i]) * c : y[b * m + i];
  }
  if (x) {
    if (lane < n) st[d.oX + lane] = scaling ? x[b * n + lane] / st[d.oD + lane] : x[b * n + lane];
    __syncwarp();
    const short *Arp = pattern + d.pArp, *Acc = pattern + d.pAcc, *Acq = pattern + d.pAcq;
    for (int i = lane; i < m; i += 32) {
      double a = 0.0;
      for (int k = Arp[i]; k < Arp[i + 1]; k++) a = fma(st[d.oAv + Acq[k]], st[d.oX + Acc[k]], a);
      st[d.oZ + i] = a;
    }
  }
}

__global__ void __launch_bounds__(32 * kFastWarps, 4) batch_fast_solve_kernel(
    FastDims d, long long count, double *state, const short *pattern, SolveCfg c, long long adaptive_interval,
    int bounds_changed, double *x_out, double *y_out, OSQPB200BatchInfo *info_out) {
  extern __shared__ __align__(16) double smem_b[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = d.n, m = d.m;
  short *pat = reinterpret_cast<short *>(smem_b + (size_t)kFastWarps * fast_warp_doubles(d));
  for (int e = threadIdx.x; e < d.plen; e += blockDim.x) pat[e] = pattern[e];
  __syncthreads();
  const long long b = (long long)blockIdx.x * kFastWarps + warp;
  if (b >= count) return;
  FastWarp W;
  fast_carve(W, d, smem_b, warp);
  FastPat Q;
  fast_pat(Q, d, pat);
  double *st = state + b * d.stride;
  for (int k = lane; k < d.nnzA; k += 32) W.Av[k] = st[d.oAv + k];
  for (int k = lane; k < d.nnzP; k += 32) W.Pv[k] = st[d.oPv + k];
  if (lane == 0) { W.Av[d.nnzA] = 0.0; W.Pv[d.nnzP] = 0.0; }
  for (int e = lane; e < 32 * kLdl; e += 32) W.L[e] = st[d.oL + e];
  const bool hn = lane < n, h0 = lane < m, h1 = lane + 32 < m;
  const int i1 = lane + 32;
  if (hn) W.invd[lane] = st[d.oInvd + lane];
  const double qj = hn ? st[d.oQ + lane] : 0.0, Dj = hn ? st[d.oD + lane] : 1.0, Dinv = 1.0 / Dj;
  double x = (hn && c.warm_start) ? st[d.oX + lane] : 0.0, dx = 0.0;
  const double l0 = h0 ? st[d.oLo + lane] : 0.0, u0 = h0 ? st[d.oUp + lane] : 0.0, E0 = h0 ? st[d.oE + lane] : 1.0;
  const double l1 = h1 ? st[d.oLo + i1] : 0.0, u1 = h1 ? st[d.oUp + i1] : 0.0, E1 = h1 ? st[d.oE + i1] : 1.0;
  const double Ei0 = 1.0 / E0, Ei1 = 1.0 / E1;
  int t0 = h0 ? (int)st[d.oCt + lane] : 0, t1 = h1 ? (int)st[d.oCt + i1] : 0;
  double z0 = (h0 && c.warm_start) ? st[d.oZ + lane] : 0.0, z1 = (h1 && c.warm_start) ? st[d.oZ + i1] : 0.0;
  double y0 = (h0 && c.warm_start) ? st[d.oY + lane] : 0.0, y1 = (h1 && c.warm_start) ? st[d.oY + i1] : 0.0;
  double dy0 = 0.0, dy1 = 0.0;
  const double cost_c = st[d.oScal], cost_cinv = 1.0 / cost_c;
  double rho = st[d.oScal + 1];
  long long rho_updates = (long long)st[d.oScal + 2];
  __syncwarp();
  bool refactor = false;
  if (bounds_changed) {
    int ch = 0;
    if (h0) { const int t = ctype_of(l0, u0); if (t != t0) { t0 = t; ch = 1; } }
    if (h1) { const int t = ctype_of(l1, u1); if (t != t1) { t1 = t; ch = 1; } }
    refactor = __any_sync(0xffffffffu, ch);
  }
  double r0 = rho_of(t0, rho), r1 = rho_of(t1, rho), ri0 = 1.0 / r0, ri1 = 1.0 / r1;
  long long status = ST_UNSOLVED;
  if (refactor && !fast_factor(W, Q, d, c.sigma, r0, r1)) status = ST_NON_CVX;
  __syncwarp();
  const bool unscale = c.scaling && !c.scaled_termination;

  InfoScalars I;
  I.pri_res = I.dua_res = I.obj_val = 0.0;
  I.ndx_t = I.ndy_t = 1.0;
  auto info = [&]() {
    // x, dx -> vn0, vn1 ; y, dy -> vm0, vm1
    if (hn) { W.vn0[lane] = x; W.vn1[lane] = dx; }
    if (h0) { W.vm0[lane] = y0; W.vm1[lane] = dy0; }
    if (h1) { W.vm0[i1] = y1; W.vm1[i1] = dy1; }
    __syncwarp();
    const double Ax0 = arow(Q, W.Av, W.vn0, lane, m), Ax1 = arow(Q, W.Av, W.vn0, i1, m);
    const double Adx0 = arow(Q, W.Av, W.vn1, lane, m), Adx1 = arow(Q, W.Av, W.vn1, i1, m);
    const double Px = prow(Q, W.Pv, W.vn0, lane, n), Pdx = prow(Q, W.Pv, W.vn1, lane, n);
    const double Aty = acol(Q, W.Av, W.vm0, lane, n), Atdy = acol(Q, W.Av, W.vm1, lane, n);
    const double e0 = unscale ? Ei0 : 1.0, e1 = unscale ? Ei1 : 1.0, EE0 = unscale ? E0 : 1.0, EE1 = unscale ? E1 : 1.0;
    const double pr0 = h0 ? fabs(Ax0 - z0) : 0.0, pr1 = h1 ? fabs(Ax1 - z1) : 0.0;
    I.pri_t = wmax(fmax(e0 * pr0, e1 * pr1));
    I.pri_r = wmax(fmax(pr0, pr1));
    I.nz_t = wmax(fmax(h0 ? e0 * fabs(z0) : 0.0, h1 ? e1 * fabs(z1) : 0.0));
    I.nz_r = wmax(fmax(h0 ? fabs(z0) : 0.0, h1 ? fabs(z1) : 0.0));
    I.nAx_t = wmax(fmax(h0 ? e0 * fabs(Ax0) : 0.0, h1 ? e1 * fabs(Ax1) : 0.0));
    I.nAx_r = wmax(fmax(h0 ? fabs(Ax0) : 0.0, h1 ? fabs(Ax1) : 0.0));
    I.ndy_t = wmax(fmax(h0 ? EE0 * fabs(dy0) : 0.0, h1 ? EE1 * fabs(dy1) : 0.0));
    I.lhs = wsum((h0 ? u0 * fmax(dy0, 0.0) + l0 * fmin(dy0, 0.0) : 0.0) + (h1 ? u1 * fmax(dy1, 0.0) + l1 * fmin(dy1, 0.0) : 0.0));
    double mu = -INFINITY, ml = -INFINITY;
    if (h0) { if (u0 < kInfty * kMinScaling) mu = fmax(mu, e0 * Adx0); if (l0 > -kInfty * kMinScaling) ml = fmax(ml, -e0 * Adx0); }
    if (h1) { if (u1 < kInfty * kMinScaling) mu = fmax(mu, e1 * Adx1); if (l1 > -kInfty * kMinScaling) ml = fmax(ml, -e1 * Adx1); }
    I.maxU_t = wmax(mu);
    I.maxNegL_t = wmax(ml);
    const double di = unscale ? Dinv : 1.0, DD = unscale ? Dj : 1.0;
    const double dr = hn ? fabs(qj + Px + Aty) : 0.0;
    I.dua_t = wmax(di * dr); I.dua_r = wmax(dr);
    I.nq_t = wmax(hn ? di * fabs(qj) : 0.0); I.nq_r = wmax(hn ? fabs(qj) : 0.0);
    I.nAty_t = wmax(hn ? di * fabs(Aty) : 0.0); I.nAty_r = wmax(hn ? fabs(Aty) : 0.0);
    I.nPx_t = wmax(hn ? di * fabs(Px) : 0.0); I.nPx_r = wmax(hn ? fabs(Px) : 0.0);
    I.obj = wsum(hn ? x * (0.5 * Px + qj) : 0.0);
    I.ndx_t = wmax(hn ? DD * fabs(dx) : 0.0);
    I.qdx = wsum(hn ? qj * dx : 0.0);
    I.nPdx_t = wmax(hn ? di * fabs(Pdx) : 0.0);
    I.nAtdy_t = wmax(hn ? di * fabs(Atdy) : 0.0);
    I.obj_val = c.scaling ? I.obj * cost_cinv : I.obj;
    I.pri_res = (m == 0) ? 0.0 : I.pri_t;
    I.dua_res = unscale ? cost_cinv * I.dua_t : I.dua_t;
    __syncwarp();
  };

  // Iteration counters are 32-bit countdowns (a 64-bit `it % check_termination` is a software division in the hot
  // loop), and update_info has ONE call site: the last iteration is checked inside the loop like any other, so the
  // (large) body of info() is instantiated once -- the loop body must stay small enough for the instruction cache
  // with 16 warps per SM at unrelated points of it.
  const int max_iter = (int)(c.max_iter < 2000000000LL ? c.max_iter : 2000000000LL);
  const int chk = (int)(c.check_termination < 2000000000LL ? c.check_termination : 2000000000LL);
  const int adp = (c.adaptive_rho && adaptive_interval > 0)
                      ? (int)(adaptive_interval < 2000000000LL ? adaptive_interval : 2000000000LL) : 0;
  int it = 0, info_iter = 0, to_check = chk, to_adapt = adp;
  double rho_est = rho;
  if (status == ST_UNSOLVED)
    for (it = 1; it <= max_iter; it++) {
      // rhs = sigma x - q + A'(rho z - y)
      if (h0) W.vm0[lane] = r0 * z0 - y0;
      if (h1) W.vm0[i1] = r1 * z1 - y1;
      __syncwarp();
      double bj = hn ? c.sigma * x - qj + acol(Q, W.Av, W.vm0, lane, n) : 0.0;
      const double xt = fast_solve(W, n, bj, W.vn0, W.vn1);
      __syncwarp();  // every lane is done with vn0 / vn1
      if (hn) W.vn0[lane] = xt;
      __syncwarp();
      const double zt0 = arow(Q, W.Av, W.vn0, lane, m), zt1 = arow(Q, W.Av, W.vn0, i1, m);
      __syncwarp();
      if (hn) { const double xn = c.alpha * xt + (1.0 - c.alpha) * x; dx = xn - x; x = xn; }
      if (h0) {
        const double zh = c.alpha * zt0 + (1.0 - c.alpha) * z0;
        const double zn = fmin(fmax(zh + ri0 * y0, l0), u0);
        double dd = r0 * (zh - zn);
        y0 += dd; z0 = zn;
        if (u0 > kInfty * kMinScaling) dd = (l0 < -kInfty * kMinScaling) ? 0.0 : fmin(dd, 0.0);
        else if (l0 < -kInfty * kMinScaling) dd = fmax(dd, 0.0);
        dy0 = dd;
      }
      if (h1) {
        const double zh = c.alpha * zt1 + (1.0 - c.alpha) * z1;
        const double zn = fmin(fmax(zh + ri1 * y1, l1), u1);
        double dd = r1 * (zh - zn);
        y1 += dd; z1 = zn;
        if (u1 > kInfty * kMinScaling) dd = (l1 < -kInfty * kMinScaling) ? 0.0 : fmin(dd, 0.0);
        else if (l1 < -kInfty * kMinScaling) dd = fmax(dd, 0.0);
        dy1 = dd;
      }
      const bool checked = chk > 0 && --to_check == 0;
      const bool adapt = adp > 0 && --to_adapt == 0;
      const bool last = it == max_iter;
      if (!(checked || adapt || last)) continue;
      if (checked) to_check = chk;
      if (adapt) to_adapt = adp;
      info();
      info_iter = it;
      if (checked || last) {  // libosqp checks the last iterate after the loop if the loop did not (Appendix A)
        status = check_termination(I, c, m, cost_c, cost_cinv, false);
        if (status != ST_UNSOLVED) break;
      }
      if (adapt) {
        const double rho_new = rho_estimate(I, rho);
        rho_est = rho_new;
        if (rho_new > rho * c.adaptive_rho_tolerance || rho_new < rho / c.adaptive_rho_tolerance) {
          rho = fmin(fmax(rho_new, kRhoMin), kRhoMax);
          rho_updates++;
          r0 = rho_of(t0, rho); r1 = rho_of(t1, rho); ri0 = 1.0 / r0; ri1 = 1.0 / r1;
          if (!fast_factor(W, Q, d, c.sigma, r0, r1)) { status = ST_NON_CVX; break; }
          __syncwarp();
          refactor = true;
        }
      }
      if (last) {  // not solved after max_iter iterations: the x10 tolerances decide between *_inaccurate and Max_iter
        const long long s2 = check_termination(I, c, m, cost_c, cost_cinv, true);
        status = (s2 != ST_UNSOLVED) ? s2 : ST_MAX_ITER;
        break;
      }
    }
  if (status != ST_NON_CVX) rho_est = rho_estimate(I, rho);
  double obj_val = I.obj_val;
  if (status == ST_NON_CVX) obj_val = nan("");
  const bool pinf = (status == ST_PINF || status == ST_PINF_INACC), dinf = (status == ST_DINF || status == ST_DINF_INACC);
  if (pinf) obj_val = kInfty;
  if (dinf) obj_val = -kInfty;
  const bool has_sol = !(pinf || dinf || status == ST_NON_CVX);
  if (hn) {
    double out;
    if (has_sol) out = c.scaling ? Dj * x : x;
    else if (dinf) out = (unscale ? Dj : 1.0) * dx / I.ndx_t;
    else out = nan("");
    x_out[b * n + lane] = out;
    st[d.oX + lane] = has_sol ? x : 0.0;
  }
  if (h0) {
    double out;
    if (has_sol) out = c.scaling ? cost_cinv * E0 * y0 : y0;
    else if (pinf) out = (unscale ? E0 : 1.0) * dy0 / I.ndy_t;
    else out = nan("");
    y_out[b * m + lane] = out;
    st[d.oZ + lane] = has_sol ? z0 : 0.0; st[d.oY + lane] = has_sol ? y0 : 0.0; st[d.oCt + lane] = (double)t0;
  }
  if (h1) {
    double out;
    if (has_sol) out = c.scaling ? cost_cinv * E1 * y1 : y1;
    else if (pinf) out = (unscale ? E1 : 1.0) * dy1 / I.ndy_t;
    else out = nan("");
    y_out[b * m + i1] = out;
    st[d.oZ + i1] = has_sol ? z1 : 0.0; st[d.oY + i1] = has_sol ? y1 : 0.0; st[d.oCt + i1] = (double)t1;
  }
  if (refactor) {
    __syncwarp();
    for (int e = lane; e < 32 * kLdl; e += 32) st[d.oL + e] = W.L[e];
    if (hn) st[d.oInvd + lane] = W.invd[lane];
  }
  if (lane == 0) {
    st[d.oScal + 1] = rho;
    st[d.oScal + 2] = (double)rho_updates;
    OSQPB200BatchInfo &o = info_out[b];
    o.iter = info_iter; o.status_val = status; o.obj_val = obj_val; o.pri_res = I.pri_res; o.dua_res = I.dua_res;
    o.rho_estimate = rho_est; o.rho_updates = rho_updates;
  }
}

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace
}  // namespace osqpb200

using namespace osqpb200;

struct OSQPB200Batch {
  int device = 0;
  c_int count = 0;
  BatchDims d{};
  OSQPSettings st{};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double *state = nullptr;
  long long *Pp = nullptr, *Pi = nullptr, *Ap = nullptr, *Ai = nullptr;
  double *stage = nullptr;  // staging for host inputs / outputs
  size_t stage_doubles = 0;
  OSQPB200BatchInfo *d_info = nullptr;
  // pinned host mirrors of the last solve's x*, y*, info (osqp_batch_solve_view hands them out without a copy)
  double *h_x = nullptr, *h_y = nullptr;
  OSQPB200BatchInfo *h_info = nullptr;
  double *h_in = nullptr;  // pinned input staging q [count][n], l, u [count][m] (osqp_batch_input_view; lazily allocated)
  int *d_fail = nullptr;
  size_t smem = 0;
  int block = 64;
  int bounds_changed = 0;
  double setup_time = 0, solve_ms = 0;
  // fast path (one warp per QP)
  bool fast = false;
  FastDims f{};
  short *d_pattern = nullptr;
};

namespace {

#define BCU(expr)                                                                                      \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      fprintf(stderr, "ERROR in %s: CUDA failure '%s' (%s)\n", __func__, cudaGetErrorString(_e), #expr); \
      return 100 + (c_int)_e;                                                                          \
    }                                                                                                  \
  } while (0)

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

int pad_ld(int k) {  // smallest ld >= k with ld == 1 (mod 16)
  int ld = k;
  while (ld % 16 != 1) ld++;
  return ld;
}

void free_batch(OSQPB200Batch *b) {
  if (!b) return;
  DevGuard g(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  cudaFree(b->state); cudaFree(b->Pp); cudaFree(b->Pi); cudaFree(b->Ap); cudaFree(b->Ai);
  cudaFree(b->stage); cudaFree(b->d_info); cudaFree(b->d_fail); cudaFree(b->d_pattern);
  if (b->h_x) cudaFreeHost(b->h_x);
  if (b->h_y) cudaFreeHost(b->h_y);
  if (b->h_info) cudaFreeHost(b->h_info);
  if (b->h_in) cudaFreeHost(b->h_in);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

SolveCfg make_cfg(const OSQPSettings &s) {
  SolveCfg c;
  memset(&c, 0, sizeof(c));
  c.sigma = s.sigma; c.alpha = s.alpha; c.eps_abs = s.eps_abs; c.eps_rel = s.eps_rel;
  c.eps_prim_inf = s.eps_prim_inf; c.eps_dual_inf = s.eps_dual_inf; c.max_iter = s.max_iter;
  c.check_termination = s.check_termination; c.scaling = s.scaling != 0;
  c.scaled_termination = (int)s.scaled_termination; c.adaptive_rho = (int)s.adaptive_rho;
  c.adaptive_rho_tolerance = s.adaptive_rho_tolerance; c.warm_start = (int)s.warm_start;
  return c;
}

}  // namespace

extern "C" {

c_int osqp_batch_cleanup(OSQPB200Batch *b) {
  free_batch(b);
  return 0;
}

c_int osqp_batch_setup(OSQPB200Batch **out, c_int count, const OSQPData *pattern, const c_float *Px,
                       const c_float *Ax, const c_float *q, const c_float *l, const c_float *u,
                       const OSQPSettings *settings) {
  if (out) *out = nullptr;
  if (!out || !pattern || !pattern->P || !pattern->A || !settings || count <= 0) return 1;
  const double t0 = now_s();
  const c_int n = pattern->n, m = pattern->m;
  if (n <= 0 || m < 0 || n > 256 || m > 256) {
    fprintf(stderr, "ERROR in osqp_batch_setup: the batched engine handles 1 <= n <= 256, 0 <= m <= 256\n");
    return 1;
  }
  if (settings->rho <= 0 || settings->sigma <= 0 || settings->max_iter <= 0 || settings->alpha <= 0 ||
      settings->alpha >= 2 || settings->scaling < 0 || settings->check_termination < 0) {
    fprintf(stderr, "ERROR in osqp_batch_setup: invalid settings\n");
    return 1;
  }
  const c_int nnzP = pattern->P->p[n], nnzA = pattern->A->p[n];
  for (c_int j = 0; j < n; j++)
    for (c_int k = pattern->P->p[j]; k < pattern->P->p[j + 1]; k++)
      if (pattern->P->i[k] > j || pattern->P->i[k] < 0) { fprintf(stderr, "ERROR in osqp_batch_setup: P is not upper triangular\n"); return 1; }
  for (c_int k = 0; k < nnzA; k++)
    if (pattern->A->i[k] < 0 || pattern->A->i[k] >= m) { fprintf(stderr, "ERROR in osqp_batch_setup: row index out of range in A\n"); return 1; }
  for (c_int k = 0; k < count * m; k++)
    if (l[k] > u[k]) { fprintf(stderr, "ERROR in osqp_batch_setup: lower bound greater than upper bound\n"); return 1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "ERROR in osqp_batch_setup: no CUDA device available -- this engine has no CPU fallback\n");
    return 10;
  }
  OSQPB200Batch *b = new OSQPB200Batch();
  struct Fail { OSQPB200Batch *p; ~Fail() { if (p) free_batch(p); } } fail{b};
  int cur = 0;
  cudaGetDevice(&cur);
  const char *envd = getenv("OSQP_B200_DEVICE");
  b->device = (envd && *envd) ? atoi(envd) : cur;
  DevGuard guard(b->device);
  b->count = count;
  b->st = *settings;
  BatchDims &d = b->d;
  d.n = (int)n; d.m = (int)m; d.ldp = pad_ld((int)n); d.lda = pad_ld(std::max<int>((int)m, 1));
  d.nnzP = (int)nnzP; d.nnzA = (int)nnzA;
  int o = 0;
  auto take = [&](int k) { int r = o; o += (k + 1) & ~1; return r; };
  d.oP = take(d.ldp * d.n); d.oA = take(d.lda * d.n); d.oL = take(d.ldp * d.n); d.oInvd = take(d.n);
  d.oQ = take(d.n); d.oL_ = take(d.m); d.oU = take(d.m); d.oD = take(d.n); d.oE = take(d.m); d.oRho = take(d.m);
  d.oX = take(d.n); d.oZ = take(d.m); d.oY = take(d.m); d.oScal = take(4);
  d.stride = o;
  b->smem = batch_smem_bytes(d);
  b->block = std::min(256, std::max(64, ((int)std::max(n, m) + 31) & ~31));
  cudaDeviceProp prop;
  BCU(cudaGetDeviceProperties(&prop, b->device));
  // ---- fast path: one warp per QP with the sparse pattern (n <= 32, m <= 64)
  {
    const char *envf = getenv("OSQP_B200_BATCH_FAST");
    std::vector<short> pat;
    FastDims &f = b->f;
    if (n <= 32 && m <= 64 && nnzA <= 30000 && nnzP <= 15000 && !(envf && atoi(envf) == 0)) {
      f.n = (int)n; f.m = (int)m; f.nnzP = (int)nnzP; f.nnzA = (int)nnzA;
      // A by rows
      std::vector<short> Arp(m + 1, 0), Acc(nnzA), Acq(nnzA);
      for (c_int k = 0; k < nnzA; k++) Arp[pattern->A->i[k] + 1]++;
      for (c_int i = 0; i < m; i++) Arp[i + 1] += Arp[i];
      {
        std::vector<short> w(Arp.begin(), Arp.end() - 1);
        for (c_int j = 0; j < n; j++)
          for (c_int k = pattern->A->p[j]; k < pattern->A->p[j + 1]; k++) {
            const short pos = w[pattern->A->i[k]]++;
            Acc[pos] = (short)j;
            Acq[pos] = (short)k;
          }
      }
      // full symmetric P by rows
      std::vector<short> Prp(n + 1, 0);
      for (c_int j = 0; j < n; j++)
        for (c_int k = pattern->P->p[j]; k < pattern->P->p[j + 1]; k++) {
          const c_int i = pattern->P->i[k];
          Prp[i + 1]++;
          if (i != j) Prp[j + 1]++;
        }
      for (c_int i = 0; i < n; i++) Prp[i + 1] += Prp[i];
      f.nnzPf = Prp[n];
      std::vector<short> Pcc(std::max(1, f.nnzPf)), Pcq(std::max(1, f.nnzPf));
      {
        std::vector<short> w(Prp.begin(), Prp.end() - 1);
        for (c_int j = 0; j < n; j++)
          for (c_int k = pattern->P->p[j]; k < pattern->P->p[j + 1]; k++) {
            const c_int i = pattern->P->i[k];
            short pos = w[i]++;
            Pcc[pos] = (short)j; Pcq[pos] = (short)k;
            if (i != j) { pos = w[j]++; Pcc[pos] = (short)i; Pcq[pos] = (short)k; }
          }
      }
      auto put = [&](const std::vector<short> &v) { int o = (int)pat.size(); pat.insert(pat.end(), v.begin(), v.end()); return o; };
      std::vector<short> Acp(n + 1), Ari(std::max<c_int>(1, nnzA));
      for (c_int j = 0; j <= n; j++) Acp[j] = (short)pattern->A->p[j];
      for (c_int k = 0; k < nnzA; k++) Ari[k] = (short)pattern->A->i[k];
      f.pAcp = put(Acp); f.pAri = put(Ari); f.pArp = put(Arp); f.pAcc = put(Acc); f.pAcq = put(Acq);
      f.pPrp = put(Prp); f.pPcc = put(Pcc); f.pPcq = put(Pcq);
      // ELL view: max row / column lengths and the slot tables (empty slots -> the zero behind the values, index 0)
      f.RA = f.CA = f.RP = 0;
      for (c_int i = 0; i < m; i++) f.RA = std::max<int>(f.RA, Arp[i + 1] - Arp[i]);
      for (c_int j = 0; j < n; j++) {
        f.CA = std::max<int>(f.CA, (int)(pattern->A->p[j + 1] - pattern->A->p[j]));
        f.RP = std::max<int>(f.RP, Prp[j + 1] - Prp[j]);
      }
      // measured on the 8192-QP MPC batch: 1.17 ms with the ELL view against 1.08 ms with the row-pointer loops (two
      // index loads per slot instead of one per entry) -- kept as an option, off by default
      f.ell = f.RA <= kEllMax && f.CA <= kEllMax && f.RP <= kEllMax && getenv("OSQP_B200_BATCH_ELL") && atoi(getenv("OSQP_B200_BATCH_ELL")) == 1;
      if (f.ell) {
        std::vector<short> EAp((size_t)std::max(1, f.RA) * 64, (short)nnzA), EAc((size_t)std::max(1, f.RA) * 64, 0);
        std::vector<short> ECp((size_t)std::max(1, f.CA) * 32, (short)nnzA), ECr((size_t)std::max(1, f.CA) * 32, 0);
        std::vector<short> EPp((size_t)std::max(1, f.RP) * 32, (short)nnzP), EPc((size_t)std::max(1, f.RP) * 32, 0);
        for (c_int i = 0; i < m; i++)
          for (int k = Arp[i], t = 0; k < Arp[i + 1]; k++, t++) { EAp[t * 64 + i] = Acq[k]; EAc[t * 64 + i] = Acc[k]; }
        for (c_int j = 0; j < n; j++) {
          for (c_int k = pattern->A->p[j], t = 0; k < pattern->A->p[j + 1]; k++, t++) { ECp[t * 32 + j] = (short)k; ECr[t * 32 + j] = Ari[k]; }
          for (int k = Prp[j], t = 0; k < Prp[j + 1]; k++, t++) { EPp[t * 32 + j] = Pcq[k]; EPc[t * 32 + j] = Pcc[k]; }
        }
        f.pEAp = put(EAp); f.pEAc = put(EAc); f.pECp = put(ECp); f.pECr = put(ECr); f.pEPp = put(EPp); f.pEPc = put(EPc);
      } else {
        f.pEAp = f.pEAc = f.pECp = f.pECr = f.pEPp = f.pEPc = 0;
      }
      f.plen = (int)pat.size();
      int o = 0;
      auto take = [&](int k) { int r = o; o += (k + 1) & ~1; return r; };
      f.oPv = take(f.nnzP); f.oAv = take(f.nnzA); f.oL = take(32 * kLdl); f.oInvd = take(f.n); f.oQ = take(f.n);
      f.oLo = take(f.m); f.oUp = take(f.m); f.oD = take(f.n); f.oE = take(f.m); f.oCt = take(f.m); f.oX = take(f.n);
      f.oZ = take(f.m); f.oY = take(f.m); f.oScal = take(4);
      f.stride = o;
      b->fast = fast_smem_bytes(f) <= (size_t)prop.sharedMemPerBlockOptin;
    }
    if (b->fast) {
      b->smem = fast_smem_bytes(f);
      BCU(raise_dyn_smem((const void *)batch_fast_setup_kernel, b->smem));
      BCU(raise_dyn_smem((const void *)batch_fast_solve_kernel, b->smem));
      BCU(cudaMalloc(&b->d_pattern, pat.size() * sizeof(short)));
      BCU(cudaMemcpy(b->d_pattern, pat.data(), pat.size() * sizeof(short), cudaMemcpyHostToDevice));
      d.stride = f.stride;  // the state allocation below uses the fast layout
    }
  }
  if (b->smem > (size_t)prop.sharedMemPerBlockOptin) {
    fprintf(stderr, "ERROR in osqp_batch_setup: a QP of this size (%zu B) does not fit in shared memory\n", b->smem);
    return 1;
  }
  BCU(raise_dyn_smem((const void *)batch_setup_kernel, b->smem));
  BCU(raise_dyn_smem((const void *)batch_solve_kernel, b->smem));
  BCU(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  BCU(cudaEventCreate(&b->ev0));
  BCU(cudaEventCreate(&b->ev1));
  BCU(cudaMalloc(&b->state, (size_t)count * d.stride * sizeof(double)));
  BCU(cudaMalloc(&b->Pp, (n + 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->Pi, std::max<c_int>(nnzP, 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->Ap, (n + 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->Ai, std::max<c_int>(nnzA, 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->d_info, (size_t)count * sizeof(OSQPB200BatchInfo)));
  BCU(cudaMallocHost(&b->h_x, (size_t)count * n * sizeof(double)));
  BCU(cudaMallocHost(&b->h_y, std::max<size_t>(1, (size_t)count * m) * sizeof(double)));
  BCU(cudaMallocHost(&b->h_info, (size_t)count * sizeof(OSQPB200BatchInfo)));
  BCU(cudaMalloc(&b->d_fail, sizeof(int)));
  BCU(cudaMemsetAsync(b->d_fail, 0, sizeof(int), b->stream));
  b->stage_doubles = (size_t)count * (size_t)(nnzP + nnzA + 2 * n + 3 * m) + 16;
  BCU(cudaMalloc(&b->stage, b->stage_doubles * sizeof(double)));
  BCU(cudaMemcpyAsync(b->Pp, pattern->P->p, (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemcpyAsync(b->Pi, pattern->P->i, nnzP * sizeof(long long), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemcpyAsync(b->Ap, pattern->A->p, (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemcpyAsync(b->Ai, pattern->A->i, nnzA * sizeof(long long), cudaMemcpyHostToDevice, b->stream));
  double *dPx = b->stage, *dAx = dPx + (size_t)count * nnzP, *dq = dAx + (size_t)count * nnzA;
  double *dl = dq + (size_t)count * n, *du = dl + (size_t)count * m;
  BCU(cudaMemcpyAsync(dPx, Px, (size_t)count * nnzP * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemcpyAsync(dAx, Ax, (size_t)count * nnzA * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  BCU(cudaMemcpyAsync(dq, q, (size_t)count * n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (m > 0) {
    BCU(cudaMemcpyAsync(dl, l, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    BCU(cudaMemcpyAsync(du, u, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  }
  const double rho0 = std::min(std::max(settings->rho, 1e-6), 1e6);
  b->st.rho = rho0;
  if (b->fast)
    batch_fast_setup_kernel<<<(unsigned)((count + kFastWarps - 1) / kFastWarps), 32 * kFastWarps, b->smem, b->stream>>>(
        b->f, count, b->state, b->d_pattern, dPx, dAx, dq, dl, du, (int)settings->scaling, rho0, settings->sigma,
        b->d_fail);
  else
    batch_setup_kernel<<<(unsigned)count, b->block, b->smem, b->stream>>>(d, b->state, b->Pp, b->Pi, b->Ap, b->Ai, dPx,
                                                                          dAx, dq, dl, du, (int)settings->scaling, rho0,
                                                                          settings->sigma, b->d_fail);
  BCU(cudaGetLastError());
  int failed = 0;
  BCU(cudaMemcpyAsync(&failed, b->d_fail, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaStreamSynchronize(b->stream));
  if (failed) {
    fprintf(stderr, "ERROR in osqp_batch_setup: P + sigma*I + A'rho A is not positive definite for at least one QP (non-convex)\n");
    return 7;
  }
  b->setup_time = now_s() - t0;
  fail.p = nullptr;
  *out = b;
  return 0;
}

c_int osqp_batch_update(OSQPB200Batch *b, const c_float *q, const c_float *l, const c_float *u) {
  if (!b) return 1;
  DevGuard guard(b->device);
  const c_int n = b->d.n, m = b->d.m, count = b->count;
  if (l && u)
    for (c_int k = 0; k < count * m; k++)
      if (l[k] > u[k]) { fprintf(stderr, "ERROR in osqp_batch_update: lower bound greater than upper bound\n"); return 1; }
  double *dq = b->stage, *dl = dq + (size_t)count * n, *du = dl + (size_t)count * m;
  if (q) BCU(cudaMemcpyAsync(dq, q, (size_t)count * n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (l && m > 0) BCU(cudaMemcpyAsync(dl, l, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (u && m > 0) BCU(cudaMemcpyAsync(du, u, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (b->fast)
    batch_fast_update_kernel<<<(unsigned)((count + kFastWarps - 1) / kFastWarps), 32 * kFastWarps, 0, b->stream>>>(
        b->f, count, b->state, b->d_pattern, q ? dq : nullptr, l ? dl : nullptr, u ? du : nullptr, nullptr, nullptr,
        b->st.scaling != 0);
  else
    batch_update_kernel<<<(unsigned)count, b->block, 8 * (n + 2), b->stream>>>(b->d, b->state, q ? dq : nullptr,
                                                                             l ? dl : nullptr, u ? du : nullptr, nullptr,
                                                                             nullptr, b->st.scaling != 0);
  BCU(cudaGetLastError());
  BCU(cudaStreamSynchronize(b->stream));  // the inputs are caller-owned
  if (l || u) b->bounds_changed = 1;
  return 0;
}

c_int osqp_batch_warm_start(OSQPB200Batch *b, const c_float *x, const c_float *y) {
  if (!b) return 1;
  DevGuard guard(b->device);
  const c_int n = b->d.n, m = b->d.m, count = b->count;
  double *dx = b->stage, *dy = dx + (size_t)count * n;
  if (x) BCU(cudaMemcpyAsync(dx, x, (size_t)count * n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (y && m > 0) BCU(cudaMemcpyAsync(dy, y, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (b->fast)
    batch_fast_update_kernel<<<(unsigned)((count + kFastWarps - 1) / kFastWarps), 32 * kFastWarps, 0, b->stream>>>(
        b->f, count, b->state, b->d_pattern, nullptr, nullptr, nullptr, x ? dx : nullptr, (y && m > 0) ? dy : nullptr,
        b->st.scaling != 0);
  else
    batch_update_kernel<<<(unsigned)count, b->block, 8 * (n + 2), b->stream>>>(b->d, b->state, nullptr, nullptr, nullptr,
                                                                             x ? dx : nullptr, (y && m > 0) ? dy : nullptr,
                                                                             b->st.scaling != 0);
  BCU(cudaGetLastError());
  BCU(cudaStreamSynchronize(b->stream));
  b->st.warm_start = 1;
  return 0;
}

// One launch for the whole batch; x*, y*, info land in the pinned host mirrors (one D2H each, PCIe speed).
static c_int batch_solve_impl(OSQPB200Batch *b) {
  DevGuard guard(b->device);
  const c_int n = b->d.n, m = b->d.m, count = b->count;
  SolveCfg c = make_cfg(b->st);
  long long interval = b->st.adaptive_rho_interval;
  if (b->st.adaptive_rho && interval == 0) interval = 50;  // no wall-clock rule inside a batch
  double *dx = b->stage, *dy = dx + (size_t)count * n;
  BCU(cudaEventRecord(b->ev0, b->stream));
  if (b->fast)
    batch_fast_solve_kernel<<<(unsigned)((count + kFastWarps - 1) / kFastWarps), 32 * kFastWarps, b->smem, b->stream>>>(
        b->f, count, b->state, b->d_pattern, c, interval, b->bounds_changed, dx, dy, b->d_info);
  else
    batch_solve_kernel<<<(unsigned)count, b->block, b->smem, b->stream>>>(b->d, b->state, c, interval, b->bounds_changed,
                                                                          dx, dy, b->d_info);
  BCU(cudaGetLastError());
  BCU(cudaEventRecord(b->ev1, b->stream));
  BCU(cudaMemcpyAsync(b->h_x, dx, (size_t)count * n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (m > 0) BCU(cudaMemcpyAsync(b->h_y, dy, (size_t)count * m * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaMemcpyAsync(b->h_info, b->d_info, (size_t)count * sizeof(OSQPB200BatchInfo), cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaStreamSynchronize(b->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, b->ev0, b->ev1);
  b->solve_ms = ms;
  b->bounds_changed = 0;
  return 0;
}

c_int osqp_batch_solve(OSQPB200Batch *b, c_float *x_out, c_float *y_out, OSQPB200BatchInfo *info_out) {
  if (!b || !x_out || !info_out) return 1;
  const c_int rc = batch_solve_impl(b);
  if (rc) return rc;
  const size_t count = (size_t)b->count;
  memcpy(x_out, b->h_x, count * b->d.n * sizeof(double));
  if (b->d.m > 0 && y_out) memcpy(y_out, b->h_y, count * b->d.m * sizeof(double));
  memcpy(info_out, b->h_info, count * sizeof(OSQPB200BatchInfo));
  return 0;
}

c_int osqp_batch_solve_view(OSQPB200Batch *b, const c_float **x, const c_float **y, const OSQPB200BatchInfo **info) {
  if (!b) return 1;
  const c_int rc = batch_solve_impl(b);
  if (rc) return rc;
  if (x) *x = b->h_x;
  if (y) *y = b->h_y;
  if (info) *info = b->h_info;
  return 0;
}

c_int osqp_batch_input_view(OSQPB200Batch *b, c_float **q, c_float **l, c_float **u) {
  if (!b) return 1;
  DevGuard guard(b->device);
  const size_t count = (size_t)b->count, n = (size_t)b->d.n, m = (size_t)b->d.m;
  if (!b->h_in) BCU(cudaMallocHost(&b->h_in, std::max<size_t>(1, count * (n + 2 * m)) * sizeof(double)));
  if (q) *q = b->h_in;
  if (l) *l = b->h_in + count * n;
  if (u) *u = b->h_in + count * (n + m);
  return 0;
}

c_int osqp_batch_device_solution(OSQPB200Batch *b, c_float **x_dev, c_float **y_dev) {
  if (!b) return 1;
  if (x_dev) *x_dev = b->stage;
  if (y_dev) *y_dev = b->stage + (size_t)b->count * b->d.n;
  return 0;
}

c_float osqp_batch_last_kernel_ms(const OSQPB200Batch *b) { return b ? b->solve_ms : -1.0; }

c_int osqp_batch_update_setting(OSQPB200Batch *b, const char *name, c_float value) {
  if (!b || !name) return 1;
  OSQPSettings &s = b->st;
  if (!strcmp(name, "max_iter")) { if (value <= 0) return 1; s.max_iter = (c_int)value; }
  else if (!strcmp(name, "eps_abs")) { if (value < 0) return 1; s.eps_abs = value; }
  else if (!strcmp(name, "eps_rel")) { if (value < 0) return 1; s.eps_rel = value; }
  else if (!strcmp(name, "eps_prim_inf")) { if (value < 0) return 1; s.eps_prim_inf = value; }
  else if (!strcmp(name, "eps_dual_inf")) { if (value < 0) return 1; s.eps_dual_inf = value; }
  else if (!strcmp(name, "alpha")) { if (value <= 0 || value >= 2) return 1; s.alpha = value; }
  else if (!strcmp(name, "check_termination")) { if (value < 0) return 1; s.check_termination = (c_int)value; }
  else if (!strcmp(name, "warm_start")) { s.warm_start = value != 0; }
  else if (!strcmp(name, "scaled_termination")) { s.scaled_termination = value != 0; }
  else if (!strcmp(name, "adaptive_rho_interval")) { if (value < 0) return 1; s.adaptive_rho_interval = (c_int)value; }
  else return 1;
  return 0;
}

}  // extern "C"
