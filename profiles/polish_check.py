"""Polish outcome (status_polish, residuals, time) of the engine next to the oracle on a few problems."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ctypes as C
import __graft_entry__ as g, problems
from reference_cases import random_qp
pkg = g.load_package(); eng = pkg.load_library(g.LIB)
def run(name, prob, lib, **kw):
    mdl = pkg.Model(lib=lib)
    mdl.setup(**prob, verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, max_iter=10000, polish=True, **kw)
    r = mdl.solve()
    extra = ""
    if lib == g.LIB:
        p = pkg.types.B200Profile(); eng.osqp_b200_get_profile(mdl.workspace, C.byref(p)); extra = f"polish_ms={p.polish_ms:.1f}"
    print(f"{name:28s} {'engine' if lib == g.LIB else 'oracle'}: {r.info.status} iter={r.info.iter} status_polish={r.info.status_polish} pri={r.info.pri_res:.2e} dua={r.info.dua_res:.2e} obj={r.info.obj_val:.8f} {extra}")
    mdl.clean()
for name, prob in (("portfolio 500/10", problems.portfolio_c4(500, 10, 20264)), ("portfolio 4000/40", problems.portfolio_c4(4000, 40, 20264)),
                   ("random 300x500", random_qp(300, 500, 0.05, 2)), ("random 3000x5000", random_qp(3000, 5000, 0.01, 3))):
    for lib in (g.LIB, g.ORACLE_LIB):
        if lib == g.ORACLE_LIB and prob["P"].shape[0] > 3500: continue
        run(name, prob, lib)
