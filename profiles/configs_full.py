#!/usr/bin/env python
"""BASELINE.json configs 3 and 4 at (or near) their full sizes on the B200 engine: status, iterations, ADMM it/s and the
unscaled KKT residuals of the returned point (no oracle runs at these sizes)."""
import argparse, ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft, problems

ap = argparse.ArgumentParser()
ap.add_argument("--lasso", default="10000,50000,0.15")
ap.add_argument("--portfolio", default="20000,200")
ap.add_argument("--lambdas", type=int, default=11)
args = ap.parse_args()
pkg = graft.load_package(); eng = pkg.load_library(graft.LIB)


def kkt(prob, r):
    P, A, q, l, u = prob["P"], prob["A"], prob["q"], prob["l"], prob["u"]
    Ax = A @ r.x
    pri = np.max(np.abs(Ax - np.clip(Ax, l, u)))
    dua = np.max(np.abs(P @ r.x + q + A.T @ r.y))
    return pri, dua


def prof(mdl):
    p = pkg.types.B200Profile()
    eng.osqp_b200_get_profile(mdl.workspace, C.byref(p))
    return p

nf, ns, dens = args.lasso.split(",")
t0 = time.perf_counter()
prob, lam_max, q_of, n = problems.lasso_c3(int(nf), int(ns), float(dens), 20263)
print(f"lasso generated in {time.perf_counter() - t0:.1f}s: n={n} m={prob['A'].shape[0]} nnz(A)={prob['A'].nnz}")
mdl = pkg.Model(lib=graft.LIB)
t0 = time.perf_counter()
mdl.setup(**dict(prob, q=q_of(2 * lam_max)), verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, max_iter=4000)
print(f"  setup {time.perf_counter() - t0:.2f}s")
tot_it, tot_ms = 0, 0.0
PH = ["stream[A;P]", "barrier", "combine", "reduce+bar", "stream A'", "barrier", "vectors", "reduce+bar"]
for lam in (np.logspace(0, -2, 11) * 2 * lam_max)[:args.lambdas]:
    mdl.update(q=q_of(lam))
    r = mdl.solve()
    p = prof(mdl)
    tot_it += r.info.iter; tot_ms += p.kernel_ms
    pri, dua = kkt(dict(prob, q=q_of(lam)), r)
    print(f"  lambda={lam:9.3f} {r.info.status} iter={r.info.iter} kernel={p.kernel_ms:.1f} ms pcg/admm={p.pcg_iters / max(1, p.admm_iters):.1f} "
          f"nnz(x)={int(np.sum(np.abs(r.x[:int(nf)]) > 1e-4))} pri={pri:.2e} dua={dua:.2e}")
print("  per PCG it (us): " + "  ".join(f"{PH[k]} {p.phase_us[k] / max(1, p.pcg_iters):.1f}" for k in range(8)))
print(f"  lasso sweep: {tot_it} ADMM its in {tot_ms:.1f} ms -> {tot_it / tot_ms * 1e3:.0f} it/s; streams={prof(mdl).streams} groups={prof(mdl).groups_A}/{prof(mdl).groups_At} paired={prof(mdl).paired}")
mdl.clean()

na, k = args.portfolio.split(",")
prob = problems.portfolio_c4(int(na), int(k), 20264)
mdl = pkg.Model(lib=graft.LIB)
t0 = time.perf_counter()
mdl.setup(**prob, verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, max_iter=4000, polish=True)
print(f"portfolio n={prob['P'].shape[0]} m={prob['A'].shape[0]} nnz(A)={prob['A'].nnz}: setup {time.perf_counter() - t0:.2f}s")
for _ in range(2):
    r = mdl.solve()
p = prof(mdl)
pri, dua = kkt(prob, r)
print(f"  {r.info.status} iter={r.info.iter} polish={r.info.status_polish} kernel={p.kernel_ms:.1f} ms polish={p.polish_ms:.1f} ms "
      f"-> {r.info.iter / p.kernel_ms * 1e3:.0f} it/s pcg/admm={p.pcg_iters / max(1, p.admm_iters):.1f} pri={pri:.2e} dua={dua:.2e} "
      f"sum(x)={np.sum(r.x[:int(na)]):.6f} streams={p.streams} groups={p.groups_A}/{p.groups_At} paired={p.paired}")
mdl.clean()
