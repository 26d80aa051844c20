"""Engine vs oracle on the EXACT instances bench.py times (BASELINE.json configs 2, 4, 5 at full size).

Budget (driver limit for `pytest -m gpu`: 1200 s for the whole suite; `--durations` on a 16-core box):
  test_c2_bench_instance ................ ~65 s measured (one oracle solve of the 50k x 100k QP, PCG backend at 1e-12 on all
                                          host threads, termination checked every iteration; two engine solves, < 3 s)
  test_c5_bench_batch_subset ............ ~10 s (engine: the whole 8192-QP batch; oracle: 512 of them incl. every QP the
                                          engine does not report as Solved)
  test_c4_portfolio_polish_at_scale ..... ~10 s measured (4000 assets, eps 1e-6: polish succeeds on both sides)

Contract (north_star): same status; (x*, y*) within the solver's own eps; iteration count within +-1 when rho is held
fixed -- asserted here with termination checked every iteration -- and, under the bench's own settings
(adaptive_rho_interval = check_termination = 25), the same rho updates and the same check at which the solve ends.
The reference itself pins nothing at these sizes: parity here = engine vs our oracle.
"""
import ctypes as C
import os

import numpy as np
import pytest

import bench
import problems

pytestmark = pytest.mark.gpu


def _oracle(pkg, oracle_lib):
    ora = pkg.load_library(oracle_lib)
    ora.osqp_oracle_configure.argtypes = [C.c_longlong, C.c_double, C.c_longlong]
    ora.osqp_oracle_set_num_threads.argtypes = [C.c_longlong]
    ora.osqp_oracle_set_num_threads(len(os.sched_getaffinity(0)))
    return ora


def test_c2_bench_instance(pkg, engine_lib, oracle_lib):
    prob = bench.make_problem(bench.N_VARS, bench.N_CONS, bench.DENSITY, bench.SEED)
    assert prob["A"].nnz == 5_000_000 and abs(prob["P"].nnz - 2.55e6) < 2e4  # nnz(A), nnz(P_full)
    eps = bench.SETTINGS["eps_abs"]
    every = dict(bench.SETTINGS, check_termination=1)  # same trajectory, termination tested at every iteration
    ora = _oracle(pkg, oracle_lib)
    ora.osqp_oracle_configure(1, 1e-12, 0)
    try:
        mo = pkg.Model(lib=oracle_lib)
        mo.setup(**prob, **every)
    finally:
        ora.osqp_oracle_configure(0, 1e-9, 0)
    o = mo.solve()
    mo.clean()
    assert o.info.status == "Solved"
    me = pkg.Model(lib=engine_lib)
    me.setup(**prob, **every)
    e1 = me.solve()
    me.clean()
    # (1) termination checked at every iteration: +-1 on the iteration count
    assert e1.info.status == "Solved"
    assert abs(e1.info.iter - o.info.iter) <= 1, (e1.info.iter, o.info.iter)
    assert e1.info.rho_updates == o.info.rho_updates
    sx, sy = eps * (1 + np.max(np.abs(o.x))), eps * (1 + np.max(np.abs(o.y)))
    assert np.max(np.abs(e1.x - o.x)) <= sx and np.max(np.abs(e1.y - o.y)) <= sy
    assert abs(e1.info.obj_val - o.info.obj_val) <= eps * (1 + abs(o.info.obj_val))
    # (2) the bench's own settings: the solve must end at the first multiple of 25 at or after that iteration
    me = pkg.Model(lib=engine_lib)
    me.setup(**prob, **bench.SETTINGS)
    e2 = me.solve()
    me.clean()
    want = -(-o.info.iter // 25) * 25
    assert e2.info.status == "Solved" and e2.info.rho_updates == o.info.rho_updates
    assert e2.info.iter == want or (o.info.iter % 25 <= 1 and abs(e2.info.iter - want) <= 25), (e2.info.iter, o.info.iter)
    # e2 stops at a multiple of 25, up to 24 iterations after the oracle did: a few eps of drift between the two points
    assert np.max(np.abs(e2.x - o.x)) <= 5 * sx and np.max(np.abs(e2.y - o.y)) <= 5 * sy
    assert abs(e2.info.rho_estimate - o.info.rho_estimate) <= 5e-2 * o.info.rho_estimate or e2.info.iter != o.info.iter


def test_c5_bench_batch_subset(pkg, engine_lib, oracle_lib):
    count = 8192
    Pp, Ap, Px, Ax, q, l, u = problems.mpc_batch_c5(count, bench.SEED + 5)
    bm = pkg.BatchModel(lib=engine_lib)
    bm.setup(Pp, Ap, Px, Ax, q, l, u, **bench.BATCH_SETTINGS)
    br = bm.solve()
    bm.clean()
    rng = np.random.default_rng(5)
    odd = list(np.nonzero(br.status_val != 1)[0])            # whatever the engine did not solve
    slow = list(np.argsort(br.iter)[-16:])                     # the stragglers
    pick = sorted(set(odd + slow + list(rng.choice(count, 512 - len(set(odd + slow)), replace=False))))
    assert len(pick) >= 500
    eps = bench.BATCH_SETTINGS["eps_abs"]
    opts = {k: v for k, v in bench.BATCH_SETTINGS.items()}
    worst = 0
    for k in pick:
        mo = pkg.Model(lib=oracle_lib)
        mo.setup(**problems.batch_instance(Pp, Ap, Px, Ax, q, l, u, k), **opts)
        o = mo.solve()
        mo.clean()
        assert int(br.status_val[k]) == o.info.status_val, (k, int(br.status_val[k]), o.info.status)
        assert int(br.rho_updates[k]) == o.info.rho_updates, (k, int(br.rho_updates[k]), o.info.rho_updates)
        # both check every 25 iterations: the same check ends the solve (one interval of play when the oracle's
        # residual sits within rounding of its tolerance at a check)
        assert abs(int(br.iter[k]) - o.info.iter) <= 25, (k, int(br.iter[k]), o.info.iter)
        worst = max(worst, abs(int(br.iter[k]) - o.info.iter))
        if o.info.status == "Solved":
            assert np.max(np.abs(br.x[k] - o.x)) <= 5 * eps * (1 + np.max(np.abs(o.x))), k
            assert np.max(np.abs(br.y[k] - o.y)) <= 5 * eps * (1 + np.max(np.abs(o.y))), k
    assert worst <= 25


def test_c4_portfolio_polish_at_scale(pkg, engine_lib, oracle_lib):
    # test/polishing.jl:69-93 at BASELINE scale: at eps 1e-6 the active set is identified and polish succeeds
    # (status_polish = 1) on libosqp's algorithm (oracle: delta-regularised KKT + refinement); the engine must agree
    n_assets = 4000
    prob = problems.portfolio_c4(n_assets, n_assets // 100, bench.SEED + 2)
    eps = 1e-6
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=20000, polish=True)
    res = {}
    for name, lib in (("engine", engine_lib), ("oracle", oracle_lib)):
        mdl = pkg.Model(lib=lib)
        mdl.setup(**prob, **opts)
        res[name] = mdl.solve()
        if name == "engine":
            prof = pkg.types.B200Profile()
            assert pkg.load_library(engine_lib).osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0
        mdl.clean()
    e, o = res["engine"], res["oracle"]
    assert e.info.status == o.info.status == "Solved"
    assert o.info.status_polish == 1
    assert e.info.status_polish == 1
    assert e.info.pri_res <= 1e-9 and e.info.dua_res <= 1e-9  # a polished point, not an ADMM iterate
    assert np.max(np.abs(e.x - o.x)) <= 1e-6 * (1 + np.max(np.abs(o.x)))
    assert np.max(np.abs(e.y - o.y)) <= 1e-6 * (1 + np.max(np.abs(o.y)))
    assert abs(e.info.obj_val - o.info.obj_val) <= 1e-8 * (1 + abs(o.info.obj_val))
    assert prof.polish_ms < prof.kernel_ms, (prof.polish_ms, prof.kernel_ms)  # polish must not dominate the solve
