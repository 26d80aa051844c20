"""Kernel-only throughput of the batched engine on 8192 MPC QPs: fixed iteration count and the realistic setting."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g, problems
pkg = g.load_package()
batch = problems.mpc_batch_c5(8192, 20267)
for mi, chk, ad, eps in ((100, 0, False, 1e-12), (4000, 25, True, 1e-4)):
    bm = pkg.BatchModel(lib=g.LIB)
    bm.setup(*batch, verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho=ad, adaptive_rho_interval=25, check_termination=chk, warm_start=False, max_iter=mi)
    for _ in range(3): r = bm.solve()
    print("max_iter", mi, "check", chk, "adaptive", ad, "kernel ms %.3f" % bm.kernel_ms, "mean iters %.1f" % r.iter.mean(), "QP-it/s %.3g" % (r.iter.sum() / bm.kernel_ms * 1e3))
    bm.clean()
