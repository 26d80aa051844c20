"""Problems and expected values of the reference's own tests, as data.

Each builder cites the Julia test it restates.  Julia's RNG stream is not reproducible
here, so tests that draw random data (seed!(1) + sprandn) keep the *property* and
regenerate the data with numpy.
"""
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def basic_problem():
    """test/basic.jl:1-23."""
    P = sp.csc_matrix(np.array([[11.0, 0.0], [0.0, 0.0]]))
    q = np.array([3.0, 4.0])
    A = sp.csc_matrix(np.array([[-1.0, 0], [0, -1], [-1, -3], [2, 5], [3, 4]]))
    u = np.array([0.0, 0.0, -15, 100, 80])
    l = -np.inf * np.ones(5)
    options = dict(verbose=False, eps_abs=1e-9, eps_rel=1e-9, check_termination=1, polish=False,
                   max_iter=4000, rho=0.1, adaptive_rho=False, warm_start=True)
    return dict(P=P, q=q, A=A, l=l, u=u), options


def polish_random_fixture():
    """test/polishing.jl:69-93 + test/problem_data/random_polish_qp.jld2 (via tests/golden/make_golden.py)."""
    d = np.load(os.path.join(HERE, "golden", "random_polish_qp.npz"))
    P = sp.csc_matrix((d["P_nzval"], d["P_rowval"], d["P_colptr"]), shape=tuple(d["P_shape"]))
    A = sp.csc_matrix((d["A_nzval"], d["A_rowval"], d["A_colptr"]), shape=tuple(d["A_shape"]))
    prob = dict(P=P, q=d["q"], A=A, l=d["l"], u=d["u"])
    return prob, d["x_test"], d["y_test"], float(d["obj_test"])


def sprandn(m, n, density, rng):
    """numpy stand-in for Julia's sprandn(m, n, density)."""
    return sp.random(m, n, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csc")


def random_qp(n, m, density, seed, pd_shift=0.1):
    """Feasible random sparse QP in the style of SURVEY 8d config C2 (small sizes for tests)."""
    rng = np.random.default_rng(seed)
    S = sp.triu(sprandn(n, n, density / 2, rng), k=1)
    S = S + S.T
    d = np.asarray(abs(S).sum(axis=1)).ravel() + rng.uniform(pd_shift, 1.0, n)
    P = (S + sp.diags(d)).tocsc()
    A = sprandn(m, n, density, rng)
    q = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    Ax0 = A @ x0
    l = Ax0 - rng.uniform(0, 1, m)
    u = Ax0 + rng.uniform(0, 1, m)
    return dict(P=P, q=q, A=A, l=l, u=u)
