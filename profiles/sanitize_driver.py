#!/usr/bin/env python
"""A few ADMM iterations on a problem that runs the paired tile streams, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import __graft_entry__ as g
from reference_cases import random_qp
import problems
pkg = g.load_package(); eng = pkg.load_library(g.LIB)
prob = random_qp(30000, 45000, 0.0008, 3)
mdl = pkg.Model(lib=g.LIB)
mdl.setup(**prob, verbose=False, max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 3, check_termination=1, adaptive_rho_interval=2)
r = mdl.solve()
p = pkg.types.B200Profile(); eng.osqp_b200_get_profile(mdl.workspace, C.byref(p))
print("single:", r.info.status, r.info.iter, "streams", p.streams, "paired", p.paired, "fixed mode", p.fast_kernels, "pcg", p.pcg_iters)
mdl.clean()
prob = random_qp(12000, 20000, 0.003, 4)  # one column group, no pairs: fixed mode 2
mdl = pkg.Model(lib=g.LIB)
mdl.setup(**prob, verbose=False, max_iter=3, check_termination=1, adaptive_rho_interval=2, polish=True)
r = mdl.solve()
eng.osqp_b200_get_profile(mdl.workspace, C.byref(p))
print("single:", r.info.status, r.info.iter, "streams", p.streams, "paired", p.paired, "fixed mode", p.fast_kernels, "pcg", p.pcg_iters)
mdl.clean()
lasso, lam_max, q_of, _ = problems.lasso_c3(5000, 1000, 0.15, 5)  # slack-elimination preconditioner: fixed mode 3
mdl = pkg.Model(lib=g.LIB)
mdl.setup(**dict(lasso, q=q_of(0.1 * lam_max)), verbose=False, max_iter=3, check_termination=1, adaptive_rho_interval=2)
r = mdl.solve()
eng.osqp_b200_get_profile(mdl.workspace, C.byref(p))
print("lasso:", r.info.status, r.info.iter, "streams", p.streams, "paired", p.paired, "fixed mode", p.fast_kernels, "pcg", p.pcg_iters)
mdl.clean()
batch = problems.mpc_batch_c5(64, 5)
bm = pkg.BatchModel(lib=g.LIB)
bm.setup(*batch, verbose=False, max_iter=30, adaptive_rho_interval=10, check_termination=5)
rb = bm.solve()
print("batch:", rb.status[:3], rb.iter[:3])
bm.clean()
