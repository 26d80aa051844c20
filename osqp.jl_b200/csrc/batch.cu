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
  int *d_fail = nullptr;
  size_t smem = 0;
  int block = 64;
  int bounds_changed = 0;
  double setup_time = 0, solve_ms = 0;
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
  cudaFree(b->stage); cudaFree(b->d_info); cudaFree(b->d_fail);
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
  if (b->smem > (size_t)prop.sharedMemPerBlockOptin) {
    fprintf(stderr, "ERROR in osqp_batch_setup: a QP of this size (%zu B) does not fit in shared memory\n", b->smem);
    return 1;
  }
  BCU(cudaFuncSetAttribute(batch_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
  BCU(cudaFuncSetAttribute(batch_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
  BCU(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  BCU(cudaEventCreate(&b->ev0));
  BCU(cudaEventCreate(&b->ev1));
  BCU(cudaMalloc(&b->state, (size_t)count * d.stride * sizeof(double)));
  BCU(cudaMalloc(&b->Pp, (n + 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->Pi, std::max<c_int>(nnzP, 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->Ap, (n + 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->Ai, std::max<c_int>(nnzA, 1) * sizeof(long long)));
  BCU(cudaMalloc(&b->d_info, (size_t)count * sizeof(OSQPB200BatchInfo)));
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
  batch_update_kernel<<<(unsigned)count, b->block, 8 * (n + 2), b->stream>>>(b->d, b->state, nullptr, nullptr, nullptr,
                                                                           x ? dx : nullptr, (y && m > 0) ? dy : nullptr,
                                                                           b->st.scaling != 0);
  BCU(cudaGetLastError());
  BCU(cudaStreamSynchronize(b->stream));
  b->st.warm_start = 1;
  return 0;
}

c_int osqp_batch_solve(OSQPB200Batch *b, c_float *x_out, c_float *y_out, OSQPB200BatchInfo *info_out) {
  if (!b || !x_out || !info_out) return 1;
  DevGuard guard(b->device);
  const c_int n = b->d.n, m = b->d.m, count = b->count;
  SolveCfg c = make_cfg(b->st);
  long long interval = b->st.adaptive_rho_interval;
  if (b->st.adaptive_rho && interval == 0) interval = 50;  // no wall-clock rule inside a batch
  double *dx = b->stage, *dy = dx + (size_t)count * n;
  BCU(cudaEventRecord(b->ev0, b->stream));
  batch_solve_kernel<<<(unsigned)count, b->block, b->smem, b->stream>>>(b->d, b->state, c, interval, b->bounds_changed,
                                                                        dx, dy, b->d_info);
  BCU(cudaGetLastError());
  BCU(cudaEventRecord(b->ev1, b->stream));
  BCU(cudaMemcpyAsync(x_out, dx, (size_t)count * n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (m > 0 && y_out) BCU(cudaMemcpyAsync(y_out, dy, (size_t)count * m * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaMemcpyAsync(info_out, b->d_info, (size_t)count * sizeof(OSQPB200BatchInfo), cudaMemcpyDeviceToHost, b->stream));
  BCU(cudaStreamSynchronize(b->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, b->ev0, b->ev1);
  b->solve_ms = ms;
  b->bounds_changed = 0;
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
  else return 1;
  return 0;
}

}  // extern "C"
