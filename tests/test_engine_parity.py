"""CUDA engine vs CPU oracle on the same seeded inputs (GPU tests; all calls go through the C ABI).

Tolerances (north_star): same status; (x*, y*) within the solver's own eps_abs/eps_rel; iteration
count within +-1 when rho is held fixed.  SpMV: relative 1e-13 (fp64, different summation order).
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from reference_cases import random_qp, sprandn

pytestmark = pytest.mark.gpu

FIXED_RHO = dict(verbose=False, adaptive_rho=False, check_termination=1, max_iter=20000, polish=False)


def solve_both(pkg, engine_lib, oracle_lib, prob, opts, oracle_pcg=False):
    """oracle_pcg: use the oracle's reduced-KKT PCG backend at tolerance 1e-12 instead of the direct
    LDL' (whose fill-in on random patterns makes n > ~2000 take minutes); test_oracle_backends.py
    shows the two oracle backends give identical iteration counts."""
    out = {}
    ora = pkg.load_library(oracle_lib)
    ora.osqp_oracle_configure.argtypes = [C.c_longlong, C.c_double, C.c_longlong]
    for name, lib in (("engine", engine_lib), ("oracle", oracle_lib)):
        if name == "oracle" and oracle_pcg:
            ora.osqp_oracle_configure(1, 1e-12, 0)
        try:
            mdl = pkg.Model(lib=lib)
            mdl.setup(**prob, **opts)
        finally:
            ora.osqp_oracle_configure(0, 1e-9, 0)
        out[name] = (mdl, mdl.solve())
    return out


def assert_parity(e, o, eps, iter_tol=1):
    assert e.info.status == o.info.status, (e.info.status, o.info.status)
    assert abs(e.info.iter - o.info.iter) <= iter_tol, (e.info.iter, o.info.iter)
    sx = eps * (1.0 + np.max(np.abs(o.x)))
    sy = eps * (1.0 + np.max(np.abs(o.y))) if o.y.size else 0.0
    assert np.max(np.abs(e.x - o.x)) <= sx, (np.max(np.abs(e.x - o.x)), sx)
    if o.y.size:
        assert np.max(np.abs(e.y - o.y)) <= sy, (np.max(np.abs(e.y - o.y)), sy)
    assert abs(e.info.obj_val - o.info.obj_val) <= 10 * eps * (1.0 + abs(o.info.obj_val))


@pytest.mark.parametrize("n,m,density,seed,eps,oracle_pcg", [
    (50, 80, 0.2, 1, 1e-4, False), (50, 80, 0.2, 1, 1e-7, False),
    (300, 500, 0.05, 2, 1e-4, False), (300, 500, 0.05, 2, 1e-7, False),
    (2000, 4000, 0.005, 3, 1e-4, False), (2000, 4000, 0.005, 3, 1e-7, True),
    (5000, 10000, 0.002, 4, 1e-4, True), (20000, 40000, 0.0015, 5, 1e-4, True)])
def test_random_qp_fixed_rho(pkg, engine_lib, oracle_lib, n, m, density, seed, eps, oracle_pcg):
    prob = random_qp(n, m, density, seed)
    r = solve_both(pkg, engine_lib, oracle_lib, prob, dict(FIXED_RHO, eps_abs=eps, eps_rel=eps), oracle_pcg)
    assert_parity(r["engine"][1], r["oracle"][1], eps)


@pytest.mark.parametrize("seed", [5, 6])
def test_random_qp_adaptive_rho_fixed_interval(pkg, engine_lib, oracle_lib, seed):
    # adaptive_rho_interval=25 is what the reference's own tests use "for deterministic behavior"
    # (test/MOI_wrapper.jl:41-49)
    prob = random_qp(1000, 2000, 0.01, seed)
    opts = dict(verbose=False, eps_abs=1e-5, eps_rel=1e-5, adaptive_rho_interval=25, max_iter=20000)
    r = solve_both(pkg, engine_lib, oracle_lib, prob, opts)
    e, o = r["engine"][1], r["oracle"][1]
    assert e.info.rho_updates == o.info.rho_updates
    # the final estimate is a ratio of residuals that are ~eps at exit: a few % of play is inherent
    assert abs(e.info.rho_estimate - o.info.rho_estimate) <= 5e-2 * o.info.rho_estimate
    assert_parity(e, o, 1e-5, iter_tol=25)


def test_equality_and_loose_rows(pkg, engine_lib, oracle_lib):
    # mixes the three rho classes of set_rho_vec: equality (1e3 rho), two-sided, loose (1e-6)
    rng = np.random.default_rng(9)
    prob = random_qp(400, 600, 0.03, 9)
    l, u = prob["l"].copy(), prob["u"].copy()
    eq = rng.random(600) < 0.2
    u[eq] = l[eq]
    loose = (~eq) & (rng.random(600) < 0.2)
    l[loose], u[loose] = -np.inf, np.inf
    onesided = (~eq) & (~loose) & (rng.random(600) < 0.3)
    u[onesided] = np.inf
    prob.update(l=l, u=u)
    r = solve_both(pkg, engine_lib, oracle_lib, prob, dict(FIXED_RHO, eps_abs=1e-6, eps_rel=1e-6))
    assert_parity(r["engine"][1], r["oracle"][1], 1e-6, iter_tol=2)


def test_no_scaling_and_scaled_termination(pkg, engine_lib, oracle_lib):
    prob = random_qp(200, 300, 0.05, 12)
    for extra in (dict(scaling=0), dict(scaled_termination=1), dict(alpha=1.0), dict(sigma=1e-3, rho=1.0)):
        r = solve_both(pkg, engine_lib, oracle_lib, prob, dict(FIXED_RHO, eps_abs=1e-6, eps_rel=1e-6, **extra))
        assert_parity(r["engine"][1], r["oracle"][1], 1e-6)


def test_scaling_vectors_match_oracle(pkg, engine_lib, oracle_lib):
    prob = random_qp(500, 700, 0.02, 21)
    r = solve_both(pkg, engine_lib, oracle_lib, prob, dict(FIXED_RHO, eps_abs=1e-3, eps_rel=1e-3))
    n, m = 500, 700
    eng = pkg.load_library(engine_lib)
    ora = pkg.load_library(oracle_lib)
    fp = C.POINTER(C.c_double)
    De, Ee, ce = np.zeros(n), np.zeros(m), C.c_double()
    Do, Eo, co = np.zeros(n), np.zeros(m), C.c_double()
    eng.osqp_b200_get_scaling.restype = C.c_longlong
    assert eng.osqp_b200_get_scaling(r["engine"][0].workspace, De.ctypes.data_as(fp), Ee.ctypes.data_as(fp), C.byref(ce)) == 0
    ora.osqp_oracle_get_scaling(r["oracle"][0].workspace, Do.ctypes.data_as(fp), Eo.ctypes.data_as(fp), C.byref(co))
    # same multiplication order on both sides; only the mean in the cost scaling is summed differently
    assert np.allclose(De, Do, rtol=1e-12, atol=0)
    assert np.allclose(Ee, Eo, rtol=1e-12, atol=0)
    assert abs(ce.value - co.value) <= 1e-12 * co.value


@pytest.mark.parametrize("n,m,density", [(3000, 5000, 0.004), (60000, 70000, 0.0002)])
@pytest.mark.parametrize("which", [0, 1, 2, 10, 11, 12])
def test_spmv_kernels_match_oracle(pkg, engine_lib, oracle_lib, which, n, m, density):
    # 0-2: hot-path format (column blocks staged in shared memory by TMA; 60000 x 70000 needs 3 blocks each
    # way), 10-12: CSR + L1-gather path of the rare phases
    prob = random_qp(n, m, density, 33)
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(**prob, verbose=False, scaling=0, sigma=1e-6)  # scaling off => resident matrices == inputs
    eng = pkg.load_library(engine_lib)
    ora = pkg.load_library(oracle_lib)
    fp = C.POINTER(C.c_double)
    rng = np.random.default_rng(4)
    call = which
    which = which % 10
    vin = rng.standard_normal(m if which == 1 else n)
    out = np.zeros(m if which == 0 else n)
    ms = C.c_double()
    eng.osqp_b200_spmv.restype = C.c_longlong
    rc = eng.osqp_b200_spmv(mdl.workspace, C.c_longlong(call), vin.ctypes.data_as(fp), out.ctypes.data_as(fp),
                            C.c_longlong(3), C.byref(ms))
    assert rc == 0 and ms.value > 0
    A = pkg.ManagedCcsc(prob["A"])
    ref = np.zeros_like(out)
    if which == 0:
        ca = A.ccsc()
        ora.osqp_oracle_mat_vec(C.byref(ca), vin.ctypes.data_as(fp), ref.ctypes.data_as(fp), C.c_longlong(0))
        scale = abs(prob["A"]) @ np.abs(vin)
    elif which == 1:
        ca = A.ccsc()
        ora.osqp_oracle_mat_vec(C.byref(ca), vin.ctypes.data_as(fp), ref.ctypes.data_as(fp), C.c_longlong(1))
        scale = abs(prob["A"]).T @ np.abs(vin)
    else:
        Pm = pkg.ManagedCcsc(prob["P"])
        cp = Pm.ccsc()
        ora.osqp_oracle_mat_vec(C.byref(cp), vin.ctypes.data_as(fp), ref.ctypes.data_as(fp), C.c_longlong(0))
        ref += 1e-6 * vin
        scale = abs(prob["P"]) @ np.abs(vin)
    assert np.all(np.abs(out - ref) <= 1e-13 * (scale + 1.0))


def test_updates_and_warm_start_match_oracle(pkg, engine_lib, oracle_lib):
    # the MOI re-solve path (src/modcaches.jl:166-179): bounds, P, q, A updates then warm-started solve
    rng = np.random.default_rng(17)
    prob = random_qp(300, 450, 0.04, 17)
    opts = dict(FIXED_RHO, eps_abs=1e-6, eps_rel=1e-6)
    r = solve_both(pkg, engine_lib, oracle_lib, prob, opts)
    Pt = sp.triu(prob["P"], format="csc")
    newP = Pt.data * (1.0 + 0.1 * rng.random(Pt.data.size))
    idxA = rng.choice(prob["A"].nnz, size=50, replace=False).astype(np.int64)
    newA = rng.standard_normal(50)
    q2 = prob["q"] + 0.3 * rng.standard_normal(300)
    l2, u2 = prob["l"] - 0.1, prob["u"] + 0.2
    for name in ("engine", "oracle"):
        mdl = r[name][0]
        mdl.update_bounds(l2, u2)
        mdl.update_P(newP, None)
        mdl.update_q(q2)
        mdl.update_A(newA, idxA)
    e, o = r["engine"][0].solve(), r["oracle"][0].solve()
    assert_parity(e, o, 1e-6, iter_tol=2)
    x0, y0 = rng.standard_normal(300), rng.standard_normal(450)
    for name in ("engine", "oracle"):
        r[name][0].warm_start(x=x0, y=y0)
    e, o = r["engine"][0].solve(), r["oracle"][0].solve()
    assert_parity(e, o, 1e-6, iter_tol=2)


def test_resolve_is_deterministic(pkg, engine_lib):
    prob = random_qp(800, 1200, 0.01, 23)
    opts = dict(FIXED_RHO, eps_abs=1e-6, eps_rel=1e-6)
    a = pkg.Model(lib=engine_lib)
    a.setup(**prob, **opts)
    ra = a.solve()
    b = pkg.Model(lib=engine_lib)
    b.setup(**prob, **opts)
    rb = b.solve()
    assert ra.info.iter == rb.info.iter
    assert np.array_equal(ra.x, rb.x) and np.array_equal(ra.y, rb.y)


@pytest.mark.parametrize("n,m,density,seed", [
    (20000, 40000, 0.0015, 41),   # [A;P] stream: 1 column group, A' stream: 2 groups
    (50000, 100000, 0.001, 42),   # BASELINE config 2 size: cluster pairs on [A;P], 4 groups on A'
    (70000, 30000, 0.0004, 43),   # 3 column groups on [A;P] (no pairs), 2 on A'
    (26000, 9000, 0.002, 44),     # 1 group on both streams
])
def test_kkt_optimality_at_scale(pkg, engine_lib, n, m, density, seed):
    # size-independent property: the returned (x*, y*) satisfies the unscaled KKT conditions to eps
    prob = random_qp(n, m, density, seed)
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(**prob, verbose=False, eps_abs=1e-5, eps_rel=1e-5, adaptive_rho_interval=25, max_iter=20000)
    r = mdl.solve()
    assert r.info.status == "Solved"
    P, A, q, l, u = prob["P"], prob["A"], prob["q"], prob["l"], prob["u"]
    Ax = A @ r.x
    z = np.clip(Ax, l, u)
    pri = np.max(np.abs(Ax - z))
    dua = np.max(np.abs(P @ r.x + q + A.T @ r.y))
    eps_pri = 1e-5 + 1e-5 * max(np.max(np.abs(Ax)), np.max(np.abs(z)))
    eps_dua = 1e-5 + 1e-5 * max(np.max(np.abs(P @ r.x)), np.max(np.abs(A.T @ r.y)), np.max(np.abs(q)))
    assert pri <= 2 * eps_pri and dua <= 2 * eps_dua
    # complementary slackness in sign form: y+ only where the upper bound is (nearly) active, y- lower
    tol = 1e-3
    assert np.all((r.y <= tol) | (u - Ax <= 1e-2))
    assert np.all((r.y >= -tol) | (Ax - l <= 1e-2))
    assert abs(r.info.obj_val - (0.5 * r.x @ (P @ r.x) + q @ r.x)) <= 1e-6 * (1 + abs(r.info.obj_val))
    # the stream path and the CSR path are two implementations of the same products: identical iterates up to
    # summation order -> same status and iteration count (+-1 check interval) when the streams are switched off
    mdl.clean()


def test_stream_and_csr_paths_agree(pkg, engine_lib):
    import os
    prob = random_qp(30000, 45000, 0.001, 45)
    opts = dict(FIXED_RHO, eps_abs=1e-5, eps_rel=1e-5, check_termination=25)
    res = {}
    for blocked in ("1", "0"):
        os.environ["OSQP_B200_BLOCKED"] = blocked
        try:
            mdl = pkg.Model(lib=engine_lib)
            mdl.setup(**prob, **opts)
            res[blocked] = mdl.solve()
            mdl.clean()
        finally:
            os.environ.pop("OSQP_B200_BLOCKED", None)
    a, b = res["1"], res["0"]
    assert a.info.status == b.info.status == "Solved"
    assert a.info.iter == b.info.iter
    assert np.max(np.abs(a.x - b.x)) <= 1e-7 * (1 + np.max(np.abs(b.x)))
    assert np.max(np.abs(a.y - b.y)) <= 1e-7 * (1 + np.max(np.abs(b.y)))


def test_unconstrained_large_uses_streams(pkg, engine_lib):
    # m = 0 (test/unconstrained.jl at scale): only the P stream exists
    rng = np.random.default_rng(46)
    n = 40000
    S = sp.triu(sprandn(n, n, 0.0004, rng), k=1)
    S = S + S.T
    d = np.asarray(abs(S).sum(axis=1)).ravel() + rng.uniform(0.5, 1.0, n)
    P = (S + sp.diags(d)).tocsc()
    q = rng.standard_normal(n)
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(P=P, q=q, A=sp.csc_matrix((0, n)), l=np.zeros(0), u=np.zeros(0), verbose=False, eps_abs=1e-8,
              eps_rel=1e-8, max_iter=2000)
    r = mdl.solve()
    assert r.info.status == "Solved"
    assert np.max(np.abs(P @ r.x + q)) <= 1e-6
    mdl.clean()


def test_infeasibility_detected_at_stream_size(pkg, engine_lib, oracle_lib):
    # test/primal_infeasibility.jl and test/dual_infeasibility.jl at a size that runs on the tile streams
    prob = random_qp(4000, 8000, 0.01, 51)
    opts = dict(verbose=False, eps_abs=1e-5, eps_rel=1e-5, eps_prim_inf=1e-5, eps_dual_inf=1e-5, adaptive_rho=False,
                check_termination=5, max_iter=5000)
    # primal infeasible: two rows with the same coefficients and disjoint intervals
    A = prob["A"].tolil()
    A[1, :] = A[0, :]
    l, u = prob["l"].copy(), prob["u"].copy()
    l[0], u[0] = 1.0, 2.0
    l[1], u[1] = -2.0, -1.0
    pinf = dict(prob, A=A.tocsc(), l=l, u=u)
    r = solve_both(pkg, engine_lib, oracle_lib, pinf, opts, oracle_pcg=True)
    e, o = r["engine"][1], r["oracle"][1]
    assert e.info.status == o.info.status == "Primal_infeasible", (e.info.status, o.info.status)
    assert np.all(np.isnan(e.x))
    # the certificate: dy' A ~ 0 and u' dy+ + l' dy- < 0 (workspace.delta_y, src/interface.jl:195-202)
    dy = e.prim_inf_cert
    assert np.max(np.abs(pinf["A"].T @ dy)) <= 1e-3 * np.max(np.abs(dy))
    assert u @ np.maximum(dy, 0) + l @ np.minimum(dy, 0) < 0
    # dual infeasible: a free variable without curvature and a cost that pushes it to infinity
    P = prob["P"].tolil()
    A2 = prob["A"].tolil()
    P[7, :] = 0
    P[:, 7] = 0
    A2[:, 7] = 0
    q = prob["q"].copy()
    q[7] = -1.0
    dinf = dict(prob, P=P.tocsc(), A=A2.tocsc(), q=q)
    r = solve_both(pkg, engine_lib, oracle_lib, dinf, opts, oracle_pcg=True)
    e, o = r["engine"][1], r["oracle"][1]
    assert e.info.status == o.info.status == "Dual_infeasible", (e.info.status, o.info.status)
    dx = e.dual_inf_cert  # workspace.delta_x, src/interface.jl:203-209
    assert abs(dx[7]) == np.max(np.abs(dx)) and q @ dx < 0


def test_grid_reductions_tree_fixed_point_and_fallback(pkg, dev_lib):
    # the two cross-block reductions of the persistent kernel on known data; 148 blocks x 512 threads.  The self-test
    # kernel only exists in the development build (lib/libosqp_dev.so, -DOSQP_B200_DEVTOOLS): same sources as the product
    engine_lib = dev_lib
    lib = pkg.load_library(engine_lib)
    lib.osqp_b200_reduce_selftest.restype = C.c_longlong
    lib.osqp_b200_reduce_selftest.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
    prob = random_qp(30000, 45000, 0.001, 45)  # big enough for the full grid
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(**prob, verbose=False)
    prof = pkg.types.B200Profile()
    assert lib.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0
    nthreads = int(prof.grid) * int(prof.block)
    want_sum, want_max = nthreads * (nthreads + 1) / 2.0, float(nthreads)
    for ref in (want_sum, 1.0, 1e-300, 1e300):  # well scaled | 2^31 off (overflow -> fallback) | absurd | absurd
        out = (C.c_double * 6)()
        assert lib.osqp_b200_reduce_selftest(mdl.workspace, ref, out) == 0
        assert out[0] == want_sum and out[1] == want_max          # fp64 tree
        assert out[3] == want_max and out[5] == want_max          # maxima are exact on either path
        # integer data: exact on the fixed-point path and on the fp64 fallback alike
        assert out[2] == want_sum, (ref, out[2], want_sum)
        assert out[4] == 0.5 * want_sum, (ref, out[4])
    mdl.clean()


def test_update_equals_fresh_setup_at_stream_size(pkg, engine_lib):
    # test/MOI_wrapper.jl:95-205 invariant (update == fresh setup, atol 1e-7) on a problem that runs on the tile
    # streams: value updates must reach the stream copies of A, A' and P, re-equilibrated
    rng = np.random.default_rng(61)
    prob = random_qp(5000, 9000, 0.008, 61)
    opts = dict(FIXED_RHO, eps_abs=1e-7, eps_rel=1e-7, check_termination=5)
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(**prob, **opts)
    mdl.solve()
    Pt = sp.triu(prob["P"], format="csc")
    Pt.sort_indices()
    A = prob["A"].tocsc()
    A.sort_indices()
    newP = Pt.data * (1.0 + 0.2 * rng.random(Pt.nnz))
    idxA = np.sort(rng.choice(A.nnz, size=A.nnz // 3, replace=False)).astype(np.int64)
    newA = A.data[idxA] * (1.0 + 0.05 * rng.standard_normal(idxA.size))
    q2 = prob["q"] + 0.1 * rng.standard_normal(prob["q"].size)
    l2, u2 = prob["l"] - 2.0, prob["u"] + 2.0  # wide enough for the perturbed A to stay feasible
    # q first: like libosqp, a matrix update re-equilibrates with the CURRENT q (the cost scaling depends on it),
    # a later q update would keep the old cost scaling and legitimately follow a different path than a fresh setup
    mdl.update_q(q2)
    mdl.update_bounds(l2, u2)
    mdl.update_P(newP, None)
    mdl.update_A(newA, idxA)
    mdl.warm_start(x=np.zeros(5000), y=np.zeros(9000))
    ru = mdl.solve()
    P2 = sp.csc_matrix((newP, Pt.indices, Pt.indptr), shape=Pt.shape)
    P2 = (P2 + sp.triu(P2, k=1).T).tocsc()
    A2 = A.copy()
    A2.data[idxA] = newA
    fresh = pkg.Model(lib=engine_lib)
    fresh.setup(P=P2, q=q2, A=A2, l=l2, u=u2, **opts)
    rf = fresh.solve()
    assert ru.info.status == rf.info.status == "Solved"
    assert ru.info.iter == rf.info.iter
    assert np.max(np.abs(ru.x - rf.x)) <= 1e-7 and np.max(np.abs(ru.y - rf.y)) <= 1e-7
    prof = pkg.types.B200Profile()
    lib = pkg.load_library(engine_lib)
    assert lib.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0 and int(prof.streams) == 1
    mdl.clean()
    fresh.clean()


@pytest.mark.parametrize("extra", [dict(scaling=0), dict(scaled_termination=True), dict(alpha=1.0),
                                   dict(check_termination=0, max_iter=300), dict(rho=1.0, sigma=1e-3)])
def test_settings_variants_at_stream_size(pkg, engine_lib, oracle_lib, extra):
    prob = random_qp(4000, 7000, 0.01, 62)
    opts = dict(FIXED_RHO, eps_abs=1e-5, eps_rel=1e-5, check_termination=5)
    opts.update(extra)
    r = solve_both(pkg, engine_lib, oracle_lib, prob, opts, oracle_pcg=True)
    e, o = r["engine"][1], r["oracle"][1]
    assert e.info.status == o.info.status
    assert abs(e.info.iter - o.info.iter) <= 1, (e.info.iter, o.info.iter)
    tol = 1e-5 if e.info.status == "Solved" else 1e-3
    assert np.max(np.abs(e.x - o.x)) <= 10 * tol * (1 + np.max(np.abs(o.x)))
    assert np.max(np.abs(e.y - o.y)) <= 10 * tol * (1 + np.max(np.abs(o.y)))


def _shifted_hessian(n, density, seed, lam_min):
    """Sparse symmetric P (not diagonally dominant) whose smallest eigenvalue is `lam_min`."""
    from scipy.sparse.linalg import eigsh

    rng = np.random.default_rng(seed)
    S = sp.random(n, n, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    P0 = (S + S.T).tocsc()
    lo = float(eigsh(P0, k=1, which="SA", return_eigenvectors=False, tol=1e-10)[0])
    return (P0 + (lam_min - lo) * sp.identity(n, format="csc")).tocsc()


@pytest.mark.parametrize("n,density,tier", [(400, 0.03, "dense Cholesky on the host"), (5000, 0.002, "CG probe on the device")])
def test_setup_rejects_mildly_indefinite_hessian(pkg, engine_lib, n, density, tier):
    # test/non_convex.jl:13-21 at sizes where the 2 x 2 example's shortcut (a negative diagonal) does not exist:
    # smallest eigenvalue -1e-3 => setup must fail; the same matrix shifted to +1e-3 must set up and solve
    rng = np.random.default_rng(n)
    A = sp.random(n // 2, n, density=0.01, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    base = dict(q=rng.standard_normal(n), A=A, l=-np.ones(n // 2), u=np.ones(n // 2))
    bad = pkg.Model(lib=engine_lib)
    with pytest.raises(RuntimeError):
        bad.setup(P=_shifted_hessian(n, density, 3, -1e-3), verbose=False, **base)
    good = pkg.Model(lib=engine_lib)
    good.setup(P=_shifted_hessian(n, density, 3, 1e-3), verbose=False, eps_abs=1e-4, eps_rel=1e-4, **base)
    r = good.solve()
    assert r.info.status in ("Solved", "Max_iter_reached"), (tier, r.info.status)
    good.clean()


def test_update_P_to_indefinite_is_refused(pkg, engine_lib):
    # osqp_update_P re-runs the convexity check (src/interface.jl:336-344 raises on a non-zero exit code)
    n = 300
    P = _shifted_hessian(n, 0.05, 5, 1e-2)
    rng = np.random.default_rng(2)
    A = sp.identity(n, format="csc")
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(P=P, q=rng.standard_normal(n), A=A, l=-np.ones(n), u=np.ones(n), verbose=False)
    Pt = sp.triu(P, format="csc")
    Pbad = sp.triu(_shifted_hessian(n, 0.05, 5, -1e-2), format="csc")
    assert (Pt.indices == Pbad.indices).all() and (Pt.indptr == Pbad.indptr).all()
    with pytest.raises(RuntimeError):
        mdl.update(Px=Pbad.data)
    mdl.clean()


def test_two_models_with_different_slices_coexist(pkg, engine_lib):
    # cudaFuncAttributeMaxDynamicSharedMemorySize is per function and process; a second, smaller tile-stream
    # model must not lower the cap under the first one
    big = random_qp(40000, 60000, 0.0008, 71)     # two column groups, paired: close to the full shared memory
    small = random_qp(6000, 9000, 0.006, 72)      # one small slice
    opts = dict(FIXED_RHO, eps_abs=1e-3, eps_rel=1e-3, check_termination=25)
    m1 = pkg.Model(lib=engine_lib)
    m1.setup(**big, **opts)
    r1 = m1.solve()
    m2 = pkg.Model(lib=engine_lib)
    m2.setup(**small, **opts)
    r2 = m2.solve()
    r1b = m1.solve()  # would fail with cudaErrorInvalidValue if the attribute had been lowered
    assert r1.info.status == r2.info.status == r1b.info.status == "Solved"
    assert r1b.info.iter <= r1.info.iter  # warm-started from the solution
    m1.clean()
    m2.clean()


@pytest.mark.parametrize("n,m,density,mode", [(40000, 60000, 0.0008, 1), (12000, 20000, 0.003, 2)])
def test_fixed_mode_kernels_match_plain_kernels(pkg, engine_lib, monkeypatch, n, m, density, mode):
    # csrc/kernels_fast.cu, kernels_fast2.cu: the ADMM and polish kernels compiled with the storage mode of the common
    # large sparse problem fixed (lane rows, fp32 slices, Jacobi; mode 1: [A; P] in cluster pairs, mode 2: no pairs);
    # their PCG phases stream fp32 copies of the matrix values (DevPtrs::mat32), everything else is the same source.
    # Both compilations must end at the same termination check with the same polished point; a workspace that is not
    # in such a mode (here: one with the Woodbury preconditioner switched on by an equality row) must stay on the plain
    # kernels.
    eng = pkg.load_library(engine_lib)
    prob = random_qp(n, m, density, 71)
    opts = dict(FIXED_RHO, eps_abs=1e-4, eps_rel=1e-4, check_termination=25, polish=True)
    out = {}
    for fast in (1, 0):
        monkeypatch.setenv("OSQP_B200_FAST_KERNELS", str(fast))
        mdl = pkg.Model(lib=engine_lib)
        mdl.setup(**prob, **opts)
        r = mdl.solve()
        prof = pkg.types.B200Profile()
        assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0
        assert (int(prof.streams), int(prof.paired), int(prof.fast_kernels)) == (1, 1 if mode == 1 else 0, mode * fast)
        out[fast] = r
        mdl.clean()
    a, b = out[1], out[0]
    assert a.info.status == b.info.status == "Solved" and abs(a.info.iter - b.info.iter) <= 25
    assert a.info.status_polish == b.info.status_polish
    tol = 1e-7 if a.info.status_polish == 1 else 10 * opts["eps_abs"]  # unpolished: two points inside the tolerance
    assert np.max(np.abs(a.x - b.x)) <= tol * (1 + np.max(np.abs(b.x)))
    assert np.max(np.abs(a.y - b.y)) <= tol * (1 + np.max(np.abs(b.y)))
    # tight tolerance: the PCG of the fixed-mode kernels iterates on fp32 copies of the matrix values, the right-hand
    # sides and the periodic residual rebuilds use the fp64 values -- the difference must not leave an error floor
    tight = dict(FIXED_RHO, eps_abs=1e-9, eps_rel=1e-9, check_termination=25, max_iter=40000)
    res = {}
    for fast in (1, 0):
        monkeypatch.setenv("OSQP_B200_FAST_KERNELS", str(fast))
        mdl = pkg.Model(lib=engine_lib)
        mdl.setup(**prob, **tight)
        res[fast] = mdl.solve()
        mdl.clean()
    assert res[1].info.status == res[0].info.status == "Solved"
    assert abs(res[1].info.iter - res[0].info.iter) <= 25
    assert np.max(np.abs(res[1].x - res[0].x)) <= 1e-8 * (1 + np.max(np.abs(res[0].x)))
    assert np.max(np.abs(res[1].y - res[0].y)) <= 1e-8 * (1 + np.max(np.abs(res[0].y)))
    monkeypatch.setenv("OSQP_B200_FAST_KERNELS", "1")
    eq = dict(prob)
    eq["l"] = prob["l"].copy()
    eq["l"][:4] = prob["u"][:4]  # a few equality rows -> Woodbury members
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(**eq, **dict(opts, polish=False))
    r = mdl.solve()
    assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0
    assert r.info.status == "Solved" and int(prof.streams) == 1 and int(prof.fast_kernels) == 0
    mdl.clean()
    # an entry outside the range that is safe for the fp32 copies (after scaling): fp64 values, plain kernels -- at setup
    # and when it arrives through osqp_update_A
    mdl = pkg.Model(lib=engine_lib)
    mdl.setup(**prob, **dict(opts, polish=False, max_iter=2))
    mdl.solve()
    assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0 and int(prof.fast_kernels) == mode
    Ax = prob["A"].tocsc().data.copy()
    Ax[0] = 1e26
    mdl.update(Ax=Ax)
    mdl.solve()
    assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0 and int(prof.fast_kernels) == 0
    mdl.clean()


def test_tiny_mode_settings_variants_match_oracle(pkg, engine_lib, oracle_lib):
    # n, m <= 256: the workspace also lives in the batched engine and is solved there; the settings
    # that change the arithmetic must reach it
    prob = random_qp(60, 90, 0.15, 31)
    eng = pkg.load_library(engine_lib)
    for extra in (dict(), dict(scaling=0), dict(scaled_termination=1), dict(alpha=1.0), dict(sigma=1e-3, rho=1.0),
                  dict(adaptive_rho=True, adaptive_rho_interval=25, check_termination=25)):
        opts = dict(FIXED_RHO, eps_abs=1e-6, eps_rel=1e-6)
        opts.update(extra)
        r = solve_both(pkg, engine_lib, oracle_lib, prob, opts)
        assert_parity(r["engine"][1], r["oracle"][1], 1e-6, iter_tol=25 if extra.get("adaptive_rho") else 1)
        prof = pkg.types.B200Profile()
        assert eng.osqp_b200_get_profile(r["engine"][0].workspace, C.byref(prof)) == 0
        assert int(prof.pcg_iters) == 0 and int(prof.admm_iters) == r["engine"][1].info.iter  # solved by the tiny engine
    # polish needs the general engine: the solve switches over, warm-started from the tiny engine's solution
    mdl = r["engine"][0]
    mdl.update_settings(polish=True)
    rp = mdl.solve()
    assert rp.info.status == "Solved" and rp.info.status_polish in (1, -1)
    assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(prof)) == 0 and int(prof.pcg_iters) > 0
    assert np.max(np.abs(rp.x - r["oracle"][1].x)) <= 1e-5 * (1 + np.max(np.abs(rp.x)))
