#!/usr/bin/env python
"""Memory-system ceiling for the access shape of the tile-stream phase (osqp_b200_membench)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); eng = pkg.load_library(graft.LIB)
eng.osqp_b200_membench.restype = C.c_double
eng.osqp_b200_membench.argtypes = [C.c_longlong] * 4
for mb in (75, 1000):
    for pattern in (0, 1, 2):
        row = []
        for depth in (2, 4, 6, 8):
            row.append("%7.0f" % eng.osqp_b200_membench(mb, pattern, depth, 20))
        print(f"{mb:5d} MB pattern {pattern}: depth 2/4/6/8 -> GB/s", " ".join(row))
