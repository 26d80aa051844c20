"""Lasso (config 3) at test size: engine with / without slack elimination in the preconditioner against the oracle."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g, problems
pkg = g.load_package(); eng = pkg.load_library(g.LIB)
prob, lam_max, q_of, n = problems.lasso_c3(1000, 5000, 0.15, 20263)
prob = dict(prob, q=q_of(0.1 * lam_max))
for eps in (1e-4, 1e-5):
    opts = dict(verbose=False, eps_abs=eps, eps_rel=eps, adaptive_rho_interval=25, max_iter=10000, polish=False)
    out = {}
    for name, env in (("slack", {"OSQP_B200_SLACK": "1"}), ("slack etaE 1e-1", {"OSQP_B200_PCG_ETA_E": "1e-1"}),
                      ("slack etaE 1e-2", {"OSQP_B200_PCG_ETA_E": "1e-2"}), ("slack etaE 1e-3", {"OSQP_B200_PCG_ETA_E": "1e-3"}),
                      ("slack eta 1e-5", {"OSQP_B200_PCG_ETA": "1e-5"}),
                      ("jacobi", {"OSQP_B200_SLACK": "0"}), ("jacobi etaE 1e-2", {"OSQP_B200_SLACK": "0", "OSQP_B200_PCG_ETA_E": "1e-2"})):
        os.environ.update(env)
        mdl = pkg.Model(lib=g.LIB); mdl.setup(**prob, **opts)
        for k in env: os.environ.pop(k)
        r = mdl.solve(); p = pkg.types.B200Profile(); eng.osqp_b200_get_profile(mdl.workspace, C.byref(p))
        out[name] = r
        print(eps, name, r.info.status, r.info.iter, r.info.rho_updates, "%.4g" % r.info.rho_estimate, "obj %.8f" % r.info.obj_val, "k %.2f" % (p.pcg_iters / max(1, p.admm_iters)), "ms %.1f" % p.kernel_ms, flush=True)
        mdl.clean()
    mo = pkg.Model(lib=g.ORACLE_LIB); mo.setup(**prob, **opts); o = mo.solve(); mo.clean()
    print(eps, "oracle", o.info.status, o.info.iter, o.info.rho_updates, "%.4g" % o.info.rho_estimate, "obj %.8f" % o.info.obj_val)
    for name, r in out.items():
        print("   ", name, "|x - x_oracle| %.2e" % np.max(np.abs(r.x - o.x)), "|y - y_oracle| %.2e" % np.max(np.abs(r.y - o.y)))
