"""The oracle's two linear-system backends agree (CPU-only test).

direct = sparse LDL' of the KKT matrix (what libosqp 0.6.2 + QDLDL does); pcg = reduced-KKT
Jacobi-PCG.  The GPU parity tests use the pcg backend at tolerance 1e-12 as the comparator where
the direct factorisation's fill-in makes it take minutes; this test is what licenses that.
"""
import ctypes as C

import numpy as np
import pytest

from reference_cases import random_qp


@pytest.mark.parametrize("n,m,density,seed,eps", [(300, 500, 0.05, 2, 1e-6), (1000, 1500, 0.01, 7, 1e-4)])
def test_pcg_backend_matches_direct(pkg, oracle_lib, n, m, density, seed, eps):
    lib = pkg.load_library(oracle_lib)
    lib.osqp_oracle_configure.argtypes = [C.c_longlong, C.c_double, C.c_longlong]
    prob = random_qp(n, m, density, seed)
    opts = dict(verbose=False, adaptive_rho=False, check_termination=1, max_iter=20000, eps_abs=eps, eps_rel=eps)
    res = []
    for mode in (0, 1):
        lib.osqp_oracle_configure(mode, 1e-12, 0)
        try:
            mdl = pkg.Model(lib=oracle_lib)
            mdl.setup(**prob, **opts)
        finally:
            lib.osqp_oracle_configure(0, 1e-9, 0)
        res.append(mdl.solve())
    d, p = res
    assert d.info.status == p.info.status == "Solved"
    assert d.info.iter == p.info.iter
    assert np.max(np.abs(d.x - p.x)) < 1e-9 and np.max(np.abs(d.y - p.y)) < 1e-8


def test_pcg_backend_adaptive_rho(pkg, oracle_lib):
    lib = pkg.load_library(oracle_lib)
    lib.osqp_oracle_configure.argtypes = [C.c_longlong, C.c_double, C.c_longlong]
    prob = random_qp(400, 700, 0.03, 13)
    opts = dict(verbose=False, adaptive_rho_interval=25, max_iter=20000, eps_abs=1e-5, eps_rel=1e-5)
    res = []
    for mode in (0, 1):
        lib.osqp_oracle_configure(mode, 1e-12, 0)
        try:
            mdl = pkg.Model(lib=oracle_lib)
            mdl.setup(**prob, **opts)
        finally:
            lib.osqp_oracle_configure(0, 1e-9, 0)
        res.append(mdl.solve())
    d, p = res
    assert d.info.iter == p.info.iter and d.info.rho_updates == p.info.rho_updates
    assert abs(d.info.rho_estimate - p.info.rho_estimate) < 1e-6 * d.info.rho_estimate
