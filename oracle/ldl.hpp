// oracle/ldl.hpp -- TEST INFRASTRUCTURE (CPU oracle), not part of the product.
//
// Sparse quasi-definite LDL^T used by the oracle's "direct" KKT backend.  It
// stands in for QDLDL + AMD, which libosqp 0.6.2 (the un-vendored OSQP_jll
// binary the reference binds at src/OSQP.jl:7, pinned Project.toml:13,18)
// uses for the per-iteration linear solve.
//
// Algorithms (restated from their published descriptions, no upstream source
// is available in this container):
//   * ordering: minimum external degree on a quotient graph with element
//     absorption (George & Liu); exact degrees, no supervariables.
//   * factorisation: up-looking LDL^T driven by the elimination tree
//     (T. Davis, "Algorithm 849: a concise sparse Cholesky factorization
//     package", ACM TOMS 2005) -- the algorithm QDLDL itself restates.
#pragma once
#include <algorithm>
#include <cstdint>
#include <set>
#include <utility>
#include <vector>

namespace oracle {

typedef long long idx_t;

// Symmetric matrix given by its UPPER triangle in CSC (row <= col).
struct SymCsc {
  idx_t n = 0;
  std::vector<idx_t> p, i;
  std::vector<double> x;
};

// Minimum-degree ordering of the symmetric pattern whose upper triangle is
// (Ap, Ai).  Returns perm with perm[k] = original index of the k-th pivot.
inline std::vector<idx_t> min_degree_order(idx_t n, const std::vector<idx_t> &Ap,
                                           const std::vector<idx_t> &Ai) {
  std::vector<std::vector<idx_t>> adjV(n), adjE(n), elemVars(n);
  for (idx_t j = 0; j < n; j++)
    for (idx_t k = Ap[j]; k < Ap[j + 1]; k++) {
      idx_t r = Ai[k];
      if (r != j) { adjV[r].push_back(j); adjV[j].push_back(r); }
    }
  std::vector<idx_t> degree(n);
  std::vector<char> elim(n, 0), alive(n, 0);
  std::set<std::pair<idx_t, idx_t>> pq;
  for (idx_t v = 0; v < n; v++) {
    std::sort(adjV[v].begin(), adjV[v].end());
    adjV[v].erase(std::unique(adjV[v].begin(), adjV[v].end()), adjV[v].end());
    degree[v] = (idx_t)adjV[v].size();
    pq.insert({degree[v], v});
  }
  std::vector<idx_t> mark(n, -1), mark2(n, -1), perm;
  perm.reserve(n);
  idx_t stamp = 0, stamp2 = 0;
  std::vector<idx_t> Lp;
  while (!pq.empty()) {
    idx_t piv = pq.begin()->second;
    pq.erase(pq.begin());
    perm.push_back(piv);
    elim[piv] = 1;
    stamp++;
    Lp.clear();
    mark[piv] = stamp;
    for (idx_t v : adjV[piv])
      if (!elim[v] && mark[v] != stamp) { mark[v] = stamp; Lp.push_back(v); }
    for (idx_t e : adjE[piv]) {
      if (!alive[e]) continue;
      for (idx_t v : elemVars[e])
        if (!elim[v] && mark[v] != stamp) { mark[v] = stamp; Lp.push_back(v); }
      alive[e] = 0;  // absorbed into the new element
      std::vector<idx_t>().swap(elemVars[e]);
    }
    std::vector<idx_t>().swap(adjV[piv]);
    std::vector<idx_t>().swap(adjE[piv]);
    elemVars[piv] = Lp;
    alive[piv] = 1;
    for (idx_t v : Lp) {
      auto &ev = adjE[v];
      size_t w = 0;
      for (size_t k = 0; k < ev.size(); k++)
        if (alive[ev[k]] && ev[k] != piv) ev[w++] = ev[k];
      ev.resize(w);
      ev.push_back(piv);
      auto &vv = adjV[v];
      w = 0;
      for (size_t k = 0; k < vv.size(); k++)
        if (!elim[vv[k]] && mark[vv[k]] != stamp) vv[w++] = vv[k];
      vv.resize(w);
    }
    for (idx_t v : Lp) {
      stamp2++;
      mark2[v] = stamp2;
      idx_t cnt = 0;
      for (idx_t u : adjV[v])
        if (mark2[u] != stamp2) { mark2[u] = stamp2; cnt++; }
      for (idx_t e : adjE[v])
        for (idx_t u : elemVars[e])
          if (!elim[u] && mark2[u] != stamp2) { mark2[u] = stamp2; cnt++; }
      pq.erase({degree[v], v});
      degree[v] = cnt;
      pq.insert({cnt, v});
    }
  }
  return perm;
}

// Symmetric permutation C = P A P^T of an upper-triangular CSC matrix.
// pinv[old] = new.  AtoC[k] = position in C of A's k-th stored entry.
inline SymCsc sym_permute(const SymCsc &A, const std::vector<idx_t> &pinv,
                          std::vector<idx_t> &AtoC) {
  idx_t n = A.n;
  SymCsc C;
  C.n = n;
  C.p.assign(n + 1, 0);
  std::vector<idx_t> w(n, 0);
  for (idx_t j = 0; j < n; j++)
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) {
      idx_t i2 = pinv[A.i[k]], j2 = pinv[j];
      w[std::max(i2, j2)]++;
    }
  for (idx_t j = 0; j < n; j++) C.p[j + 1] = C.p[j] + w[j];
  for (idx_t j = 0; j < n; j++) w[j] = C.p[j];
  idx_t nz = A.p[n];
  C.i.resize(nz);
  C.x.resize(nz);
  AtoC.resize(nz);
  for (idx_t j = 0; j < n; j++)
    for (idx_t k = A.p[j]; k < A.p[j + 1]; k++) {
      idx_t i2 = pinv[A.i[k]], j2 = pinv[j];
      idx_t q = w[std::max(i2, j2)]++;
      C.i[q] = std::min(i2, j2);
      C.x[q] = A.x[k];
      AtoC[k] = q;
    }
  return C;
}

// LDL^T factor of a symmetric quasi-definite matrix (upper CSC, already permuted).
struct Ldl {
  idx_t n = 0;
  std::vector<idx_t> Lp, Li, parent, Lnz;
  std::vector<double> Lx, D, Dinv;
  // work
  std::vector<idx_t> flag, pattern;
  std::vector<double> y;

  void symbolic(const SymCsc &A) {
    n = A.n;
    parent.assign(n, -1);
    Lnz.assign(n, 0);
    flag.assign(n, -1);
    for (idx_t k = 0; k < n; k++) {
      flag[k] = k;
      for (idx_t p = A.p[k]; p < A.p[k + 1]; p++) {
        idx_t i = A.i[p];
        if (i < k) {
          for (; flag[i] != k; i = parent[i]) {
            if (parent[i] == -1) parent[i] = k;
            Lnz[i]++;
            flag[i] = k;
          }
        }
      }
    }
    Lp.assign(n + 1, 0);
    for (idx_t k = 0; k < n; k++) Lp[k + 1] = Lp[k] + Lnz[k];
    Li.resize(Lp[n]);
    Lx.resize(Lp[n]);
    D.resize(n);
    Dinv.resize(n);
    pattern.resize(n);
    y.assign(n, 0.0);
  }

  // Returns the number of positive pivots, or -1 on a zero pivot.
  idx_t numeric(const SymCsc &A) {
    idx_t npos = 0;
    std::fill(y.begin(), y.end(), 0.0);
    for (idx_t k = 0; k < n; k++) {
      idx_t top = n;
      flag[k] = k;
      Lnz[k] = 0;
      for (idx_t p = A.p[k]; p < A.p[k + 1]; p++) {
        idx_t i = A.i[p];
        if (i <= k) {
          y[i] += A.x[p];
          idx_t len = 0;
          for (; flag[i] != k; i = parent[i]) {
            pattern[len++] = i;
            flag[i] = k;
          }
          while (len > 0) pattern[--top] = pattern[--len];
        }
      }
      D[k] = y[k];
      y[k] = 0.0;
      for (; top < n; top++) {
        idx_t i = pattern[top];
        double yi = y[i];
        y[i] = 0.0;
        idx_t p2 = Lp[i] + Lnz[i];
        idx_t p;
        for (p = Lp[i]; p < p2; p++) y[Li[p]] -= Lx[p] * yi;
        double lki = yi * Dinv[i];
        D[k] -= lki * yi;
        Li[p] = k;
        Lx[p] = lki;
        Lnz[i]++;
      }
      if (D[k] == 0.0) return -1;
      if (D[k] > 0.0) npos++;
      Dinv[k] = 1.0 / D[k];
    }
    return npos;
  }

  void solve(double *x) const {
    for (idx_t j = 0; j < n; j++) {
      double xj = x[j];
      for (idx_t p = Lp[j]; p < Lp[j + 1]; p++) x[Li[p]] -= Lx[p] * xj;
    }
    for (idx_t j = 0; j < n; j++) x[j] *= Dinv[j];
    for (idx_t j = n - 1; j >= 0; j--) {
      double xj = x[j];
      for (idx_t p = Lp[j]; p < Lp[j + 1]; p++) xj -= Lx[p] * x[Li[p]];
      x[j] = xj;
    }
  }
};

}  // namespace oracle
