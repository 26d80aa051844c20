#!/usr/bin/env python
"""Small driver for ncu: set up the bench workload, run `--solves` solves and the standalone SpMVs.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python profiles/profile_driver.py --solves 2
  ncu --set full --clock-control none --import-source on -k regex:'admm_kernel|spmv_kernel' \
      -o gpurun_out/prof python profiles/profile_driver.py --solves 1 --max-iter 50
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=bench.N_VARS)
ap.add_argument("--m", type=int, default=bench.N_CONS)
ap.add_argument("--density", type=float, default=bench.DENSITY)
ap.add_argument("--solves", type=int, default=1)
ap.add_argument("--max-iter", type=int, default=4000)
ap.add_argument("--spmv-reps", type=int, default=2)
ap.add_argument("--lib", default=graft.LIB, help="engine library (a compile-time variant from profiles/variants.py)")
args = ap.parse_args()

pkg = graft.load_package()
eng = pkg.load_library(args.lib)
prob = bench.make_problem(args.n, args.m, args.density, bench.SEED)
mdl = pkg.Model(lib=args.lib)
mdl.setup(**prob, **dict(bench.SETTINGS, max_iter=args.max_iter, warm_start=False))
PHASES = ["stream[A;P]", "barrier", "combine", "reduce+bar", "stream A'", "barrier", "vectors", "reduce+bar",
          "admm z/y/x+bar", "admm A'rhs+bar", "admm rhs+red", "refresh", "update_info", "rho upd", "epilogue", "-"]
for _ in range(args.solves):
    r = mdl.solve()
    print("solve:", r.info.status, r.info.iter, f"{r.info.solve_time * 1e3:.1f} ms")
    p = pkg.types.B200Profile()
    eng.osqp_b200_get_profile(mdl.workspace, C.byref(p))
    tot = sum(p.phase_us) or 1.0
    print(f"  kernel {p.kernel_ms:.1f} ms, admm {p.admm_iters}, pcg {p.pcg_iters}, info {p.info_evals}, refresh {p.refreshes};"
          f" phase total {tot / 1e3:.1f} ms")
    print("  per PCG it (us):  " + "  ".join(f"{PHASES[k]} {p.phase_us[k] / max(1, p.pcg_iters):.2f}" for k in range(8)))
    print("  per ADMM it (us): " + "  ".join(f"{PHASES[k]} {p.phase_us[k] / max(1, p.admm_iters):.2f}" for k in range(8, 15)))
fp = C.POINTER(C.c_double)
eng.osqp_b200_spmv.restype = C.c_longlong
rng = np.random.default_rng(1)
ap_which = ((0, args.n), (1, args.m), (2, args.n), (10, args.n), (11, args.m), (12, args.n))
for which, ilen in ap_which:
    vin = rng.standard_normal(ilen)
    ms = C.c_double()
    eng.osqp_b200_spmv(mdl.workspace, C.c_longlong(which), vin.ctypes.data_as(fp), None, C.c_longlong(args.spmv_reps),
                       C.byref(ms))
    print("spmv", which, f"{ms.value * 1e3:.1f} us")
mdl.clean()
