"""BASELINE.json configs 3 and 4 at test sizes: CUDA engine vs CPU oracle through the reference's API.

Config 3: Lasso as a QP, lambda sweep re-solved with osqp_update_lin_cost + retained iterates (warm start).
Config 4: factor-model portfolio with polishing.  Constructions: problems.py.
"""
import numpy as np
import pytest

import problems

pytestmark = pytest.mark.gpu


def _models(pkg, engine_lib, oracle_lib, prob, opts):
    out = {}
    for name, lib in (("engine", engine_lib), ("oracle", oracle_lib)):
        mdl = pkg.Model(lib=lib)
        mdl.setup(**prob, **opts)
        out[name] = mdl
    return out


@pytest.mark.parametrize("n_feat,n_samp,density", [(200, 1000, 0.15), (1000, 5000, 0.15)])
def test_lasso_lambda_sweep_warm_started(pkg, engine_lib, oracle_lib, n_feat, n_samp, density):
    prob, lam_max, q_of, n = problems.lasso_c3(n_feat, n_samp, density, 20263)
    eps = 1e-4
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=10000, polish=False)
    prob = dict(prob, q=q_of(lam_max))
    mdl = _models(pkg, engine_lib, oracle_lib, prob, opts)
    lams = np.logspace(0, -2, 5) * lam_max
    prev_iters = None
    for lam in lams:
        res = {}
        for name in ("engine", "oracle"):
            mdl[name].update(q=q_of(lam))  # src/interface.jl:240-246; iterates are kept (warm_start = 1)
            res[name] = mdl[name].solve()
        e, o = res["engine"], res["oracle"]
        assert e.info.status == o.info.status == "Solved", (lam, e.info.status, o.info.status)
        # same optimum: objective and solution within the solver's tolerance (rho adapts, so iteration counts of
        # the two linear-system backends may differ by an adaptive-rho interval)
        # the 1000 x 5000 case runs on the tile streams with the slack-elimination preconditioner and follows the oracle to
        # one check interval; the small case runs Jacobi-PCG on the CSR path, whose inexact solves on this badly
        # conditioned K cost up to two intervals (round-1 tolerances)
        xtol, itol = (10, 25) if n_feat >= 1000 else (20, 50)
        assert abs(e.info.obj_val - o.info.obj_val) <= xtol * eps * (1 + abs(o.info.obj_val)), lam
        assert np.max(np.abs(e.x - o.x)) <= xtol * eps * (1 + np.max(np.abs(o.x))), lam
        assert abs(e.info.iter - o.info.iter) <= itol, (lam, e.info.iter, o.info.iter)
        prev_iters = e.info.iter
    # at lambda_max the Lasso solution is x = 0
    assert prev_iters is not None
    for m_ in mdl.values():
        m_.clean()


def test_lasso_at_lambda_max_is_zero(pkg, engine_lib):
    prob, lam_max, q_of, n = problems.lasso_c3(300, 1500, 0.15, 7)
    mdl = pkg.Model(lib=engine_lib)
    # the objective is y'y (no 1/2), so x = 0 is optimal from lambda = 2 |Ad'b|inf on
    mdl.setup(**dict(prob, q=q_of(2.1 * lam_max)), verbose=False, eps_abs=1e-6, eps_rel=1e-6, max_iter=20000,
              adaptive_rho_interval=25)
    r = mdl.solve()
    assert r.info.status == "Solved"
    assert np.max(np.abs(r.x[:300])) <= 1e-4
    mdl.clean()


@pytest.mark.parametrize("n_assets,k", [(500, 10), (4000, 40)])
def test_portfolio_with_polish(pkg, engine_lib, oracle_lib, n_assets, k):
    prob = problems.portfolio_c4(n_assets, k, 20264)
    eps = 1e-4
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=10000, polish=True)
    mdl = _models(pkg, engine_lib, oracle_lib, prob, opts)
    e, o = mdl["engine"].solve(), mdl["oracle"].solve()
    assert e.info.status == o.info.status == "Solved"
    assert abs(e.info.obj_val - o.info.obj_val) <= 20 * eps * (1 + abs(o.info.obj_val))
    x = e.x[:n_assets]
    assert abs(np.sum(x) - 1.0) <= 1e-3 and np.min(x) >= -1e-3 and np.max(x) <= 1 + 1e-3
    # polishing must not make the point worse than the ADMM iterate (libosqp accepts it only if residuals improve)
    assert e.info.status_polish == o.info.status_polish  # 1 at 500 assets, -1 (rejected) at 4000: same on both
    if e.info.status_polish == 1:
        assert e.info.pri_res <= 1e-6 and e.info.dua_res <= 1e-5
        assert np.max(np.abs(e.x - o.x)) <= 1e-3 * (1 + np.max(np.abs(o.x)))
    for m_ in mdl.values():
        m_.clean()


def test_portfolio_woodbury_preconditioner(pkg, engine_lib, oracle_lib):
    # The 41 coupling equality rows of a 4000-asset portfolio (y = F'x, 1'x = 1) form the low-rank part of the PCG
    # preconditioner (engine.cuh WoodDev); the tile streams are forced on at this size so that the oracle's direct
    # LDL' (libosqp's own algorithm, polish included) is still affordable as the reference.  A preconditioner must not
    # change the ADMM trajectory: same status, rho updates, iteration count and polished solution as the oracle and as
    # the engine with plain Jacobi -- at a few PCG iterations per ADMM iteration instead of dozens.
    import ctypes as C
    import os

    n_assets = 4000
    prob = problems.portfolio_c4(n_assets, n_assets // 100, 20264)
    eps = 1e-6
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=20000, polish=True)
    eng = pkg.load_library(engine_lib)
    res, k_per_it, polish_ms = {}, {}, {}
    for name in ("woodbury", "jacobi"):
        os.environ["OSQP_B200_WOODBURY"] = "1" if name == "woodbury" else "0"
        os.environ["OSQP_B200_STREAM_MIN_NNZ"] = "50000"
        try:
            mdl = pkg.Model(lib=engine_lib)
            mdl.setup(**prob, **opts)
        finally:
            os.environ.pop("OSQP_B200_WOODBURY", None)
            os.environ.pop("OSQP_B200_STREAM_MIN_NNZ", None)
        res[name] = mdl.solve()
        prof = pkg.types.B200Profile()
        assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0
        assert int(prof.streams) == 1
        k_per_it[name] = prof.pcg_iters / max(1, prof.admm_iters)
        polish_ms[name] = prof.polish_ms
        mdl.clean()
    mo = pkg.Model(lib=oracle_lib)
    mo.setup(**prob, **opts)
    o = mo.solve()
    mo.clean()
    w, j = res["woodbury"], res["jacobi"]
    assert w.info.status == j.info.status == o.info.status == "Solved"
    assert w.info.rho_updates == o.info.rho_updates == j.info.rho_updates
    assert abs(w.info.iter - o.info.iter) <= 25 and abs(j.info.iter - o.info.iter) <= 25, (w.info.iter, j.info.iter, o.info.iter)
    assert o.info.status_polish == 1 and w.info.status_polish == 1 and j.info.status_polish == 1
    for r in (w, j):
        assert np.max(np.abs(r.x - o.x)) <= 1e-6 * (1 + np.max(np.abs(o.x)))
        assert np.max(np.abs(r.y - o.y)) <= 1e-6 * (1 + np.max(np.abs(o.y)))
        assert abs(r.info.obj_val - o.info.obj_val) <= 1e-8 * (1 + abs(o.info.obj_val))
    assert k_per_it["woodbury"] <= 5.0, k_per_it
    assert k_per_it["jacobi"] >= 4 * k_per_it["woodbury"], k_per_it
    assert polish_ms["woodbury"] < polish_ms["jacobi"], polish_ms


def test_lasso_slack_elimination_preconditioner(pkg, engine_lib, oracle_lib):
    # The 5000 equality rows  Ad x - y = b  of the Lasso QP each own a slack column: eliminating those columns inside
    # the preconditioner (engine.cuh SlackDev) must leave the ADMM trajectory alone -- same status, rho updates,
    # iteration count and solution as the oracle and as plain Jacobi -- at a fraction of the PCG iterations.
    import ctypes as C
    import os

    prob, lam_max, q_of, n = problems.lasso_c3(1000, 5000, 0.15, 20263)
    prob = dict(prob, q=q_of(0.1 * lam_max))
    eps = 1e-5
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=10000, polish=False)
    eng = pkg.load_library(engine_lib)
    res, k_per_it = {}, {}
    for name in ("slack", "jacobi"):
        os.environ["OSQP_B200_SLACK"] = "1" if name == "slack" else "0"
        try:
            mdl = pkg.Model(lib=engine_lib)
            mdl.setup(**prob, **opts)
        finally:
            os.environ.pop("OSQP_B200_SLACK", None)
        res[name] = mdl.solve()
        prof = pkg.types.B200Profile()
        assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0
        assert int(prof.streams) == 1
        k_per_it[name] = prof.pcg_iters / max(1, prof.admm_iters)
        mdl.clean()
    mo = pkg.Model(lib=oracle_lib)
    mo.setup(**prob, **opts)
    o = mo.solve()
    mo.clean()
    s, j = res["slack"], res["jacobi"]
    assert s.info.status == j.info.status == o.info.status == "Solved"
    # with the slack preconditioner (and the energy-norm stopping rule that comes with it) the engine follows the oracle:
    # same rho updates, same check at which the solve ends, solution within a few eps
    assert s.info.rho_updates == o.info.rho_updates
    assert abs(s.info.iter - o.info.iter) <= 25, (s.info.iter, o.info.iter)
    assert np.max(np.abs(s.x - o.x)) <= 5 * eps * (1 + np.max(np.abs(o.x)))
    assert abs(s.info.obj_val - o.info.obj_val) <= 5 * eps * (1 + abs(o.info.obj_val))
    # plain Jacobi with the residual rule reaches the same optimum more slowly on this badly conditioned K (round 1's path)
    assert np.max(np.abs(j.x - o.x)) <= 50 * eps * (1 + np.max(np.abs(o.x)))
    assert j.info.iter >= s.info.iter
    assert k_per_it["slack"] <= 15.0, k_per_it
    assert k_per_it["jacobi"] >= 2.0 * k_per_it["slack"], k_per_it


def _oracle_fresh(pkg, oracle_lib, prob, opts):
    mo = pkg.Model(lib=oracle_lib)
    mo.setup(**prob, **opts)
    r = mo.solve()
    mo.clean()
    return r


def test_woodbury_follows_bound_rho_and_matrix_updates(pkg, engine_lib, oracle_lib):
    # The membership of the Woodbury set is fixed at setup; the data in it is not: bounds that turn coupling equalities
    # into inequalities (their rho class changes), osqp_update_rho and osqp_update_A must all reach the preconditioner
    # (C, C^-1 are rebuilt by the next launch) -- the results must match a fresh oracle setup of the updated problem.
    import os

    n_assets = 4000
    prob = problems.portfolio_c4(n_assets, n_assets // 100, 20264)
    eps = 1e-5
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=20000, polish=False)
    os.environ["OSQP_B200_STREAM_MIN_NNZ"] = "50000"
    try:
        mdl = pkg.Model(lib=engine_lib)
        mdl.setup(**prob, **opts)
    finally:
        os.environ.pop("OSQP_B200_STREAM_MIN_NNZ", None)
    r0 = mdl.solve()
    assert r0.info.status == "Solved"
    k = n_assets // 100
    # (1) relax the factor equalities y = F'x into |y - F'x| <= 0.05, keep the budget row
    l, u = prob["l"].copy(), prob["u"].copy()
    l[:k] -= 0.05
    u[:k] += 0.05
    mdl.update(l=l, u=u)
    mdl.update_settings(rho=0.1)  # the first solve adapted rho: start where a fresh setup starts
    mdl.warm_start(x=np.zeros(prob["P"].shape[0]), y=np.zeros(prob["A"].shape[0]))
    r1 = mdl.solve()
    o1 = _oracle_fresh(pkg, oracle_lib, dict(prob, l=l, u=u), opts)
    assert r1.info.status == o1.info.status == "Solved"
    assert abs(r1.info.iter - o1.info.iter) <= 25, (r1.info.iter, o1.info.iter)
    assert abs(r1.info.obj_val - o1.info.obj_val) <= 10 * eps * (1 + abs(o1.info.obj_val))
    assert np.max(np.abs(r1.x - o1.x)) <= 10 * eps * (1 + np.max(np.abs(o1.x)))
    # (2) a new rho, (3) new values in A (the factor loadings of the first 50 assets doubled)
    mdl.update_settings(rho=0.5)
    A2 = prob["A"].copy().tocsc()
    A2.data = A2.data.copy()
    for j in range(50):
        sl = slice(A2.indptr[j], A2.indptr[j + 1])
        rows = A2.indices[sl]
        A2.data[sl] = np.where(rows < k, 2.0 * A2.data[sl], A2.data[sl])
    mdl.update(Ax=A2.data)
    mdl.warm_start(x=np.zeros(prob["P"].shape[0]), y=np.zeros(prob["A"].shape[0]))
    r2 = mdl.solve()
    o2 = _oracle_fresh(pkg, oracle_lib, dict(prob, A=A2, l=l, u=u), dict(opts, rho=0.5))
    assert r2.info.status == o2.info.status == "Solved"
    assert abs(r2.info.iter - o2.info.iter) <= 25, (r2.info.iter, o2.info.iter)
    assert abs(r2.info.obj_val - o2.info.obj_val) <= 10 * eps * (1 + abs(o2.info.obj_val))
    assert np.max(np.abs(r2.x - o2.x)) <= 10 * eps * (1 + np.max(np.abs(o2.x)))
    mdl.clean()


def test_slack_preconditioner_follows_matrix_and_bound_updates(pkg, engine_lib, oracle_lib):
    # Lasso on the tile streams with the slack-elimination preconditioner: new data values (osqp_update_A on A_d) and a
    # new right-hand side b (osqp_update_bounds) must give what a fresh oracle setup of the updated problem gives
    prob, lam_max, q_of, n = problems.lasso_c3(1000, 5000, 0.15, 20263)
    prob = dict(prob, q=q_of(0.2 * lam_max))
    eps = 1e-5
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=10000, polish=False)
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(**prob, **opts)
    assert mdl.solve().info.status == "Solved"
    rng = np.random.default_rng(3)
    A2 = prob["A"].copy().tocsc()
    A2.data = A2.data * (1.0 + 0.1 * rng.standard_normal(A2.nnz) * (np.abs(A2.data) != 1.0))  # leave the identity blocks
    l, u = prob["l"].copy(), prob["u"].copy()
    shift = 0.05 * rng.standard_normal(5000)
    l[:5000] += shift
    u[:5000] += shift
    mdl.update(Ax=A2.data)
    mdl.update(l=l, u=u)
    mdl.update_settings(rho=0.1)  # the first solve adapted rho: start where a fresh setup starts
    mdl.warm_start(x=np.zeros(n), y=np.zeros(prob["A"].shape[0]))
    r = mdl.solve()
    o = _oracle_fresh(pkg, oracle_lib, dict(prob, A=A2, l=l, u=u), opts)
    assert r.info.status == o.info.status == "Solved"
    assert r.info.rho_updates == o.info.rho_updates
    assert abs(r.info.iter - o.info.iter) <= 25, (r.info.iter, o.info.iter)
    assert np.max(np.abs(r.x - o.x)) <= 10 * eps * (1 + np.max(np.abs(o.x)))
    mdl.clean()
