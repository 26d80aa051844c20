// admm_rules.cuh -- libosqp 0.6.2 decision rules shared by the single-QP engine (kernels.cu) and the batched engine
// (batch.cu): constants, the scalars update_info produces, check_termination, the rho estimate (SURVEY.md rows
// a9-a11, Appendix A).  Pure scalar code; the kernels differ only in how they produce the scalars.
#pragma once
#include "engine.cuh"

#include <math.h>

namespace osqpb200 {
namespace {

constexpr double kInfty = 1e30;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kDivisionTol = 1e-30;
constexpr int kPrintInterval = 200;

constexpr long long ST_SOLVED = 1, ST_SOLVED_INACC = 2, ST_PINF_INACC = 3, ST_DINF_INACC = 4, ST_MAX_ITER = -2,
                    ST_PINF = -3, ST_DINF = -4, ST_TIME_LIMIT = -6, ST_NON_CVX = -7, ST_UNSOLVED = -10;

// ------------------------------------------------------------------ scaled/infeasibility info scalars
struct InfoScalars {
  double pri_t, pri_r, nz_t, nz_r, nAx_t, nAx_r, ndy_t, lhs, maxU_t, maxNegL_t;
  double dua_t, dua_r, nq_t, nq_r, nAty_t, nAty_r, nPx_t, nPx_r, obj, ndx_t, qdx, nPdx_t, nAtdy_t;
  double pri_res, dua_res, obj_val;  // what update_info publishes
};

// check_termination of libosqp 0.6.2; returns the new status or ST_UNSOLVED.
__device__ __forceinline__ long long check_termination(const InfoScalars &S, const SolveCfg &c, int m, double cost_c,
                                                       double cost_cinv, bool approximate) {
  double eps_abs = c.eps_abs, eps_rel = c.eps_rel, eps_pinf = c.eps_prim_inf, eps_dinf = c.eps_dual_inf;
  if (S.pri_res > kInfty || S.dua_res > kInfty) return ST_NON_CVX;
  if (approximate) {
    eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10;
  }
  const bool unscale = c.scaling && !c.scaled_termination;
  bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
  if (m == 0) prim_ok = true;
  else {
    const double eps_prim = eps_abs + eps_rel * fmax(S.nz_t, S.nAx_t);
    if (S.pri_res < eps_prim) prim_ok = true;
    else if (S.ndy_t > kDivisionTol && S.lhs < -eps_pinf * S.ndy_t) prim_inf = S.nAtdy_t < eps_pinf * S.ndy_t;
  }
  double mx = fmax(S.nq_t, fmax(S.nAty_t, S.nPx_t));
  if (unscale) mx *= cost_cinv;
  const double eps_dual = eps_abs + eps_rel * mx;
  if (S.dua_res < eps_dual) dual_ok = true;
  else {
    const double cs = unscale ? cost_c : 1.0;
    if (S.ndx_t > kDivisionTol && S.qdx < -cs * eps_dinf * S.ndx_t && S.nPdx_t < cs * eps_dinf * S.ndx_t)
      dual_inf = !(S.maxU_t > eps_dinf * S.ndx_t) && !(S.maxNegL_t > eps_dinf * S.ndx_t);
  }
  if (prim_ok && dual_ok) return approximate ? ST_SOLVED_INACC : ST_SOLVED;
  if (prim_inf) return approximate ? ST_PINF_INACC : ST_PINF;
  if (dual_inf) return approximate ? ST_DINF_INACC : ST_DINF;
  return ST_UNSOLVED;
}

__device__ __forceinline__ double rho_estimate(const InfoScalars &S, double rho) {
  double pri = S.pri_r / (fmax(S.nz_r, S.nAx_r) + 1e-10);
  double dua = S.dua_r / (fmax(S.nq_r, fmax(S.nAty_r, S.nPx_r)) + 1e-10);
  double est = rho * sqrt(pri / (dua + 1e-10));
  return fmin(fmax(est, kRhoMin), kRhoMax);
}

}  // namespace
}  // namespace osqpb200
