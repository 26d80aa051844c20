"""Latency of small problems on the engine (CSR path): per-solve launch cost, update cost, us per ADMM iteration."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import __graft_entry__ as g
from reference_cases import basic_problem, random_qp
pkg = g.load_package()
prob, opts = basic_problem()
t0 = time.perf_counter(); mdl = pkg.Model(lib=g.LIB); mdl.setup(**prob, **opts); t1 = time.perf_counter()
print("basic QP setup ms", (t1 - t0) * 1e3)
for _ in range(3): r = mdl.solve()
t0 = time.perf_counter()
for _ in range(50): r = mdl.solve()
dt = (time.perf_counter() - t0) / 50
print("basic QP: iter", r.info.iter, "solve ms %.3f" % (dt * 1e3), "-> us per ADMM it %.2f" % (dt * 1e6 / r.info.iter), "solve_time field ms %.3f" % (r.info.solve_time * 1e3))
mdl.update_settings(max_iter=1, check_termination=0)
t0 = time.perf_counter()
for _ in range(200): r = mdl.solve()
print("1-iteration solve (launch + copies) us %.1f" % ((time.perf_counter() - t0) / 200 * 1e6))
t0 = time.perf_counter()
for _ in range(200): mdl.update(q=prob["q"])
print("update_q us %.1f" % ((time.perf_counter() - t0) / 200 * 1e6))
for n, m, d in ((50, 80, 0.2), (300, 500, 0.05), (2000, 4000, 0.005)):
    p2 = random_qp(n, m, d, 1)
    t0 = time.perf_counter(); m2 = pkg.Model(lib=g.LIB); m2.setup(**p2, verbose=False, eps_abs=1e-4, eps_rel=1e-4); ts = time.perf_counter() - t0
    for _ in range(2): r = m2.solve()
    t0 = time.perf_counter()
    for _ in range(10): r = m2.solve()
    dt = (time.perf_counter() - t0) / 10
    print(f"n={n} m={m}: setup {ts*1e3:.1f} ms, solve {dt*1e3:.2f} ms, iter {r.info.iter}, us/it {dt*1e6/r.info.iter:.1f}")
