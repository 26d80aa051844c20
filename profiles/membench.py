#!/usr/bin/env python
"""Memory-system ceiling for the access shape of the tile-stream phase (osqp_b200_membench)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); eng = pkg.load_library(graft.DEV_LIB)
eng.osqp_b200_membench.restype = C.c_double
eng.osqp_b200_membench.argtypes = [C.c_longlong] * 4
for mb in (75, 1000):
    for pattern in (0, 1, 2):
        row = []
        for depth in (2, 4, 6, 8):
            row.append("%7.0f" % eng.osqp_b200_membench(mb, pattern, depth, 20))
        print(f"{mb:5d} MB pattern {pattern}: depth 2/4/6/8 -> GB/s", " ".join(row))

import bench
prob = bench.make_problem(bench.N_VARS, bench.N_CONS, bench.DENSITY, bench.SEED)
mdl = pkg.Model(lib=graft.DEV_LIB); mdl.setup(**prob, **bench.SETTINGS)
eng.osqp_b200_barrier_bench.restype = C.c_double
eng.osqp_b200_barrier_bench.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong]
for mode, name in ((0, "bare grid barrier"), (1, "reduce_and_barrier<2>"), (2, "barrier after scattered stores"),
                   (3, "reduce_and_barrier_fx<2>")):
    print(f"{name:32s} {eng.osqp_b200_barrier_bench(mdl.workspace, 2000, mode):8.0f} ns")
eng.osqp_b200_cluster_probe.restype = C.c_longlong
eng.osqp_b200_cluster_probe.argtypes = [C.c_void_p, C.c_longlong]
print("co-resident clusters of 2/4/8/16 blocks:", [int(eng.osqp_b200_cluster_probe(mdl.workspace, k)) for k in (2, 4, 8, 16)])
