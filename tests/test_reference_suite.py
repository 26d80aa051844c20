"""The reference's own test-suite (test/*.jl), restated over the ctypes mirror of src/interface.jl.

Every test runs twice through identical marshalling code:
  backend=oracle  (CPU, default run)  -> pins the oracle against the reference's known answers
  backend=engine  (@gpu)              -> the CUDA engine must pass the very same assertions
Tolerances are the reference's (cited per test).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from reference_cases import basic_problem, polish_random_fixture, sprandn

norm = np.linalg.norm


def make(pkg, backend, prob, opts):
    m = pkg.Model(lib=backend)
    m.setup(**prob, **opts)
    return m


# ------------------------------------------------------------------ test/basic.jl
TOL = 1e-5


def test_basic_QP(pkg, backend):  # test/basic.jl:27-50
    prob, opts = basic_problem()
    res = make(pkg, backend, prob, opts).solve()
    assert norm(res.x - [0.0, 5.0]) <= TOL
    assert norm(res.y - [1.666666666666, 0.0, 1.3333333, 0.0, 0.0]) <= TOL
    assert abs(res.info.obj_val - 20.0) <= TOL
    assert res.info.status == "Solved"


def test_basic_update_q(pkg, backend):  # test/basic.jl:52-76
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    model.update(q=[10.0, 20.0])
    res = model.solve()
    assert norm(res.x - [0.0, 5.0]) <= TOL
    assert norm(res.y - [3.33333333, 0.0, 6.66666666, 0.0, 0.0]) <= TOL
    assert abs(res.info.obj_val - 100.0) <= TOL


def test_basic_update_l(pkg, backend):  # test/basic.jl:78-102
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    model.update(l=-100 * np.ones(5))
    res = model.solve()
    assert norm(res.x - [0.0, 5.0]) <= TOL
    assert norm(res.y - [1.6666666666, 0.0, 1.333333333333, 0.0, 0.0]) <= TOL
    assert abs(res.info.obj_val - 20.0) <= TOL


def test_basic_update_u(pkg, backend):  # test/basic.jl:104-132
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    model.update(u=1000 * np.ones(5))
    res = model.solve()
    assert norm(res.x - [-1.51515152e-01, -3.33282828e02]) <= TOL
    assert norm(res.y - [0.0, 0.0, 1.333333333333, 0.0, 0.0]) <= TOL
    assert abs(res.info.obj_val - (-1333.459595961)) <= TOL


def test_basic_update_max_iter(pkg, backend):  # test/basic.jl:134-152
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    model.update_settings(max_iter=80)
    res = model.solve()
    assert res.info.status == "Max_iter_reached"
    assert not np.isnan(res.x).any()  # Max_iter_reached is in SOLUTION_PRESENT (src/constants.jl:23)


def test_basic_update_check_termination(pkg, backend):  # test/basic.jl:154-172
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    model.update_settings(check_termination=False)
    res = model.solve()
    assert res.info.iter == opts["max_iter"]


def test_basic_update_rho(pkg, backend):  # test/basic.jl:174-208
    prob, opts = basic_problem()
    res_default = make(pkg, backend, prob, opts).solve()
    new_opts = dict(opts)
    new_opts["rho"] = 0.7
    model = make(pkg, backend, prob, new_opts)
    model.update_settings(rho=opts["rho"])
    res_new_rho = model.solve()
    assert res_default.info.iter == res_new_rho.info.iter


def test_basic_time_limit(pkg, backend):  # test/basic.jl:210-240
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    res = model.solve()
    assert res.info.status == "Solved"
    model.update_settings(eps_abs=1e-20, eps_rel=1e-20, time_limit=1e-6, max_iter=1000000, check_termination=0)
    res_tl = model.solve()
    assert res_tl.info.status == "Time_limit_reached"
    # :Time_limit_reached is not in SOLUTION_PRESENT => the Julia layer hands back NaNs (src/interface.jl:194-197)
    assert np.isnan(res_tl.x).all()


def test_update_settings_rejects_unknown(pkg, backend):  # src/interface.jl:447-455
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    with pytest.raises(RuntimeError):
        model.update_settings(sigma=1.0)
    with pytest.raises(RuntimeError):
        model.update_settings(scaled_termination=1)  # not in UPDATABLE_SETTINGS (src/constants.jl:29-44)


# ------------------------------------------------------------------ test/dual_infeasibility.jl
DUAL_OPTS = dict(verbose=False, eps_abs=1e-5, eps_rel=1e-5, eps_prim_inf=1e-15, check_termination=1)


def test_dual_infeasible_lp(pkg, backend):  # test/dual_infeasibility.jl:16-29
    model = make(pkg, backend, dict(P=sp.csc_matrix((2, 2)), q=np.array([2.0, -1.0]), A=sp.eye(2, format="csc"),
                                    u=np.inf * np.ones(2), l=np.zeros(2)), DUAL_OPTS)
    res = model.solve()
    assert res.info.status == "Dual_infeasible"
    assert np.isnan(res.x).all() and np.isnan(res.prim_inf_cert).all()
    cert = res.dual_inf_cert
    assert np.isfinite(cert).all() and abs(np.max(np.abs(cert)) - 1.0) < 1e-12  # normalised certificate
    assert cert @ [2.0, -1.0] < 0  # q'dx < 0: unbounded direction


def test_dual_infeasible_qp(pkg, backend):  # test/dual_infeasibility.jl:31-44
    model = make(pkg, backend, dict(P=sp.diags([4.0, 0.0]).tocsc(), q=np.array([0.0, 2.0]),
                                    A=sp.csc_matrix(np.array([[1.0, 1.0], [-1.0, 1.0]])),
                                    u=np.array([2.0, 3.0]), l=-np.inf * np.ones(2)), DUAL_OPTS)
    assert model.solve().info.status == "Dual_infeasible"


def test_primal_dual_infeasible_as_dual(pkg, backend):  # test/dual_infeasibility.jl:46-62
    model = make(pkg, backend, dict(P=sp.csc_matrix((2, 2)), q=np.array([-1.0, -1.0]),
                                    A=sp.csc_matrix(np.array([[1.0, -1.0], [-1.0, 1.0], [1.0, 0.0], [0.0, 1.0]])),
                                    u=np.inf * np.ones(4), l=np.array([1.0, 1.0, 0.0, 0.0])), DUAL_OPTS)
    model.warm_start(x=np.array([50.0, 30.0]), y=np.array([-2.0, -2.0, -2.0, -2.0]))
    assert model.solve().info.status == "Dual_infeasible"


# ------------------------------------------------------------------ test/primal_infeasibility.jl
PRIM_OPTS = dict(verbose=False, eps_abs=1e-5, eps_rel=1e-5, eps_dual_inf=1e-18, scaling=True)


def test_primal_infeasible_problem(pkg, backend):  # test/primal_infeasibility.jl:15-42 (data regenerated)
    rng = np.random.default_rng(1)
    n, m = 50, 500
    Pt = sprandn(n, n, 0.6, rng)
    P = (Pt.T @ Pt).tocsc()
    q = rng.standard_normal(n)
    A = sprandn(m, n, 0.6, rng).tolil()
    u = 3 + rng.standard_normal(m)
    l = -3 + rng.standard_normal(m)
    k = n // 2
    A[k - 1, :] = A[k, :]
    l[k - 1] = u[k] + 10 * rng.random()
    u[k - 1] = l[k - 1] + 0.5
    model = make(pkg, backend, dict(P=P, q=q, A=A.tocsc(), l=l, u=u), PRIM_OPTS)
    res = model.solve()
    assert res.info.status == "Primal_infeasible"
    assert np.isnan(res.x).all() and np.isnan(res.y).all()
    cert = res.prim_inf_cert
    assert np.isfinite(cert).all() and abs(np.max(np.abs(cert)) - 1.0) < 1e-12


def test_primal_dual_infeasible_as_primal(pkg, backend):  # test/primal_infeasibility.jl:44-59
    model = make(pkg, backend, dict(P=sp.csc_matrix((2, 2)), q=np.array([-1.0, -1.0]),
                                    A=sp.csc_matrix(np.array([[1.0, -1.0], [-1.0, 1.0], [1.0, 0.0], [0.0, 1.0]])),
                                    l=np.array([1.0, 1.0, 0.0, 0.0]), u=np.inf * np.ones(4)), PRIM_OPTS)
    assert model.solve().info.status == "Primal_infeasible"


# ------------------------------------------------------------------ test/non_convex.jl
def nonconvex_problem():
    return dict(P=sp.csc_matrix(np.array([[2.0, 5.0], [5.0, 1.0]])), q=np.array([3.0, 4.0]),
                A=sp.csc_matrix(np.array([[-1.0, 0], [0, -1], [-1, -3], [2, 5], [3, 4]])),
                u=np.array([0.0, 0.0, -15, 100, 80]), l=-np.inf * np.ones(5))


def test_non_convex_small_sigma(pkg, backend):  # test/non_convex.jl:4-22: setup must fail
    with pytest.raises(RuntimeError, match="Error in OSQP setup"):
        make(pkg, backend, nonconvex_problem(), dict(verbose=False, sigma=1e-6))


def test_non_convex_big_sigma(pkg, backend):  # test/non_convex.jl:24-41
    model = make(pkg, backend, nonconvex_problem(), dict(verbose=False, sigma=5.0))
    res = model.solve()
    assert np.isnan(res.info.obj_val)
    assert res.info.status == "Non_convex"


# ------------------------------------------------------------------ test/polishing.jl
POLISH_OPTS = dict(verbose=False, polish=True, eps_abs=1e-3, eps_rel=1e-3, max_iter=5000)
PTOL = 1e-3


def test_polishing_problem(pkg, backend):  # test/polishing.jl:17-38
    prob = dict(P=sp.diags([11.0, 0.0]).tocsc(), q=np.array([3.0, 4.0]),
                A=sp.csc_matrix(np.array([[-1.0, 0.0], [0.0, -1.0], [-1.0, -3], [2.0, 5.0], [3.0, 4.0]])),
                u=np.array([0.0, 0.0, -15.0, 100.0, 80]), l=-np.inf * np.ones(5))
    res = make(pkg, backend, prob, POLISH_OPTS).solve()
    assert np.allclose(res.x, [9.90341e-11, 5.0], atol=PTOL, rtol=0)
    assert np.allclose(res.y, [1.66667, 0.0, 1.33333, 1.20431e-14, 1.49741e-14], atol=PTOL, rtol=0)
    assert abs(res.info.obj_val - 20.0) <= PTOL
    assert res.info.status_polish == 1


def test_polishing_unconstrained(pkg, backend):  # test/polishing.jl:40-67 (data regenerated)
    rng = np.random.default_rng(1)
    n = m = 10
    Pd = rng.random(n) + 0.2
    q = rng.standard_normal(n)
    prob = dict(P=sp.diags(Pd).tocsc(), q=q, A=sp.eye(n, format="csc"), l=-100 * np.ones(m), u=100 * np.ones(m))
    res = make(pkg, backend, prob, POLISH_OPTS).solve()
    x_test = -q / Pd
    assert np.allclose(res.x, x_test, atol=PTOL, rtol=0)
    assert np.allclose(res.y, np.zeros(m), atol=PTOL, rtol=0)
    assert abs(res.info.obj_val - (-0.5 * q @ (q / Pd))) <= PTOL
    assert res.info.status_polish == 1


def test_polish_random(pkg, backend):  # test/polishing.jl:69-93 -- the Mosek golden fixture
    prob, x_test, y_test, obj_test = polish_random_fixture()
    res = make(pkg, backend, prob, POLISH_OPTS).solve()
    assert np.allclose(res.x, x_test, atol=PTOL, rtol=0)
    assert np.allclose(res.y, y_test, atol=PTOL, rtol=0)
    assert abs(res.info.obj_val - obj_test) <= PTOL
    assert res.info.status_polish == 1


# ------------------------------------------------------------------ test/unconstrained.jl
def test_unconstrained_problem(pkg, backend):  # test/unconstrained.jl:15-41 (data regenerated)
    rng = np.random.default_rng(1)
    n = 30
    Pd = rng.random(n) + 0.2
    q = rng.standard_normal(n)
    prob = dict(P=sp.diags(Pd).tocsc(), q=q, A=sp.csc_matrix((0, n)), l=np.zeros(0), u=np.zeros(0))
    res = make(pkg, backend, prob, dict(verbose=False, eps_abs=1e-8, eps_rel=1e-8, eps_dual_inf=1e-18)).solve()
    assert np.allclose(res.x, -q / Pd, atol=1e-5, rtol=0)
    assert res.y.size == 0
    assert abs(res.info.obj_val - (-0.5 * q @ (q / Pd))) <= 1e-5
    assert res.info.status == "Solved"


# ------------------------------------------------------------------ test/warm_start.jl
def test_warm_start_problem(pkg, backend):  # test/warm_start.jl:17-48 (data regenerated)
    rng = np.random.default_rng(1)
    n, m = 100, 200
    Pt = sprandn(n, n, 0.9, rng)
    P = (Pt.T @ Pt).tocsc()
    q = rng.standard_normal(n)
    A = sprandn(m, n, 0.9, rng)
    u = rng.random(m) * 2
    l = -rng.random(m) * 2
    opts = dict(verbose=False, eps_abs=1e-8, eps_rel=1e-8, polish=False, adaptive_rho=False, check_termination=1)
    model = make(pkg, backend, dict(P=P, q=q, A=A, l=l, u=u), opts)
    res = model.solve()
    x_opt, y_opt, tot_iter = res.x.copy(), res.y.copy(), res.info.iter
    assert res.info.status == "Solved"
    model.warm_start(x=np.zeros(n), y=np.zeros(m))
    res = model.solve()
    assert res.info.iter == tot_iter
    model.warm_start(x=x_opt, y=y_opt)
    res = model.solve()
    assert res.info.iter <= 10


# ------------------------------------------------------------------ test/feasibility.jl
def test_feasibility_problem(pkg, backend):  # test/feasibility.jl:15-29 (data regenerated)
    rng = np.random.default_rng(3)
    n = m = 30
    A = sprandn(m, n, 0.8, rng)
    u = rng.standard_normal(m)
    prob = dict(P=sp.csc_matrix((n, n)), q=np.zeros(n), A=A, l=u.copy(), u=u)
    res = make(pkg, backend, prob, dict(verbose=False, eps_abs=1e-6, eps_rel=1e-6, max_iter=5000)).solve()
    assert norm(A @ res.x - u) <= 1e-3


# ------------------------------------------------------------------ test/interface.jl
def test_sparse_matrix_interface_roundtrip(pkg):  # test/interface.jl:4-12
    jl = sp.eye(5, format="csc")
    mc = pkg.ManagedCcsc(jl)
    jl2 = pkg.ccsc_to_scipy(mc.ccsc())
    assert (jl != jl2).nnz == 0


def test_model_error_handling(pkg, backend):  # test/interface.jl:15-18
    with pytest.raises(RuntimeError):
        pkg.Model(lib=backend).solve()


# ------------------------------------------------------------------ test/MOI_wrapper.jl invariants
# (the MOI layer itself is Julia-only; these are its solver-level invariants)
def lp_problem():
    """min -x  s.t.  x + y <= 1, x, y >= 0 in OSQP form -- the LP of test/MOI_wrapper.jl:280-353."""
    P = sp.csc_matrix((2, 2))
    q = np.array([-1.0, 0.0])
    A = sp.csc_matrix(np.array([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0]]))
    l = np.array([-np.inf, 0.0, 0.0])
    u = np.array([1.0, np.inf, np.inf])
    return dict(P=P, q=q, A=A, l=l, u=u)


MOI_OPTS = dict(verbose=False, eps_abs=1e-8, eps_rel=1e-16, max_iter=10000, adaptive_rho_interval=25)


def test_moi_default_warm_start_and_exact_resolve(pkg, backend):  # test/MOI_wrapper.jl:322-353
    prob = lp_problem()
    model = make(pkg, backend, prob, MOI_OPTS)
    r1 = model.solve()
    assert r1.info.status == "Solved"
    assert np.allclose(r1.x, [1.0, 0.0], atol=1e-4)
    assert abs(r1.info.obj_val - (-1.0)) < 1e-4
    x1, y1, it1 = r1.x.copy(), r1.y.copy(), r1.info.iter
    # iterates are retained: solving again takes fewer iterations (:335-338)
    r2 = model.solve()
    assert r2.info.iter < it1
    # zeroed warm start => bitwise the same answer as a fresh model (:345-353, rtol = atol = 0)
    model.warm_start(x=np.zeros(2), y=np.zeros(3))
    r3 = model.solve()
    fresh = make(pkg, backend, prob, MOI_OPTS).solve()
    assert r3.info.iter == fresh.info.iter
    assert np.array_equal(r3.x, fresh.x) and np.array_equal(r3.y, fresh.y)
    assert np.allclose(x1, r3.x, atol=1e-7) and np.allclose(y1, r3.y, atol=1e-6)


def test_moi_update_equals_fresh_setup(pkg, backend):  # test/MOI_wrapper.jl:95-205 (atol 1e-7)
    rng = np.random.default_rng(11)
    n, m = 8, 12
    Pt = sprandn(n, n, 0.7, rng)
    P = sp.triu((Pt.T @ Pt + sp.eye(n)).tocsc(), format="csc")
    A = sprandn(m, n, 0.6, rng)
    q = rng.standard_normal(n)
    l = -1 - rng.random(m)
    u = 1 + rng.random(m)
    opts = dict(verbose=False, eps_abs=1e-9, eps_rel=1e-9, max_iter=20000, adaptive_rho_interval=25)
    # -- q
    model = make(pkg, backend, dict(P=P, q=q, A=A, l=l, u=u), opts)
    q2 = q + rng.standard_normal(n)
    model.update(q=q2)
    model.warm_start(x=np.zeros(n), y=np.zeros(m))
    a = model.solve()
    b = make(pkg, backend, dict(P=P, q=q2, A=A, l=l, u=u), opts).solve()
    assert np.allclose(a.x, b.x, atol=1e-7, rtol=0) and np.allclose(a.y, b.y, atol=1e-7, rtol=0)
    # -- A: one entry via an index vector (0-based here)
    A2 = A.copy()
    A2.data[3] = 2.5
    model = make(pkg, backend, dict(P=P, q=q, A=A, l=l, u=u), opts)
    model.update(Ax=np.array([2.5]), Ax_idx=np.array([3]))
    a = model.solve()
    b = make(pkg, backend, dict(P=P, q=q, A=A2, l=l, u=u), opts).solve()
    assert np.allclose(a.x, b.x, atol=1e-7, rtol=0) and np.allclose(a.y, b.y, atol=1e-7, rtol=0)
    # -- P: all values, no index vector
    P2 = P.copy()
    P2.data = P2.data * 1.5
    model = make(pkg, backend, dict(P=P, q=q, A=A, l=l, u=u), opts)
    model.update(Px=P2.data.copy())
    a = model.solve()
    b = make(pkg, backend, dict(P=P2, q=q, A=A, l=l, u=u), opts).solve()
    assert np.allclose(a.x, b.x, atol=1e-7, rtol=0) and np.allclose(a.y, b.y, atol=1e-7, rtol=0)
    # -- P and A together
    model = make(pkg, backend, dict(P=P, q=q, A=A, l=l, u=u), opts)
    model.update(Px=P2.data.copy(), Ax=A2.data.copy())
    a = model.solve()
    b = make(pkg, backend, dict(P=P2, q=q, A=A2, l=l, u=u), opts).solve()
    assert np.allclose(a.x, b.x, atol=1e-7, rtol=0) and np.allclose(a.y, b.y, atol=1e-7, rtol=0)


def test_moi_equality_constrained_least_squares(pkg, backend):  # test/MOI_wrapper.jl:694-790 (atol 1e-4)
    rng = np.random.default_rng(5)
    n, m = 10, 3
    for _ in range(3):
        F = rng.standard_normal((20, n))
        g = rng.standard_normal(20)
        Ceq = rng.standard_normal((m, n))
        d = rng.standard_normal(m)
        # min ||F x - g||^2 s.t. C x = d  ==  KKT solve
        K = np.block([[2 * F.T @ F, Ceq.T], [Ceq, np.zeros((m, m))]])
        sol = np.linalg.solve(K, np.concatenate([2 * F.T @ g, d]))
        prob = dict(P=sp.csc_matrix(2 * F.T @ F), q=-2 * F.T @ g, A=sp.csc_matrix(Ceq), l=d, u=d)
        res = make(pkg, backend, prob, MOI_OPTS).solve()
        assert res.info.status == "Solved"
        assert np.allclose(res.x, sol[:n], atol=1e-4, rtol=0)


def test_polish_toggle_via_settings(pkg, backend):  # test/MOI_wrapper.jl:514-518
    prob, opts = basic_problem()
    opts = dict(opts, eps_abs=1e-3, eps_rel=1e-3)
    model = make(pkg, backend, prob, opts)
    assert model.solve().info.status_polish == 0
    model.update_settings(polish=True)
    model.warm_start(x=np.zeros(2), y=np.zeros(5))
    assert model.solve().info.status_polish == 1


def test_update_bounds_rejects_crossed(pkg, backend):  # libosqp validate: l > u => non-zero exit => Julia error
    prob, opts = basic_problem()
    model = make(pkg, backend, prob, opts)
    with pytest.raises(RuntimeError):
        model.update(l=np.ones(5), u=np.zeros(5))
    with pytest.raises(RuntimeError):
        bad = dict(prob, l=np.ones(5) * 200)
        make(pkg, backend, bad, opts)
