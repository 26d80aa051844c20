#!/usr/bin/env python
"""Small driver for ncu on the batched engine: `count` MPC QPs, fixed number of iterations."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g, problems
pkg = g.load_package()
count = int(sys.argv[1]) if len(sys.argv) > 1 else 592
its = int(sys.argv[2]) if len(sys.argv) > 2 else 100
batch = problems.mpc_batch_c5(count, 20267)
bm = pkg.BatchModel(lib=g.LIB)
bm.setup(*batch, verbose=False, eps_abs=1e-12, eps_rel=1e-12, adaptive_rho=False, check_termination=0, warm_start=False, max_iter=its)
r = bm.solve()
print("kernel ms", bm.kernel_ms)
