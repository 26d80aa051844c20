#!/usr/bin/env python
"""Per-block phase timing of the tile-stream SpMV (globaltimer probes written by spmv_stream_kernel)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft, bench
pkg = graft.load_package(); eng = pkg.load_library(graft.LIB)
prob = bench.make_problem(bench.N_VARS, bench.N_CONS, bench.DENSITY, bench.SEED)
mdl = pkg.Model(lib=graft.LIB); mdl.setup(**prob, **bench.SETTINGS)
fp = C.POINTER(C.c_double); eng.osqp_b200_spmv.restype = C.c_longlong
rng = np.random.default_rng(1)
for which, ilen in ((0, bench.N_VARS), (1, bench.N_CONS)):
    vin = rng.standard_normal(ilen); ms = C.c_double()
    eng.osqp_b200_spmv(mdl.workspace, C.c_longlong(which), vin.ctypes.data_as(fp), None, C.c_longlong(5), C.byref(ms))
    buf = np.zeros(16 * 148, dtype=np.uint64)
    eng.osqp_b200_debug_read(mdl.workspace, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), C.c_longlong(buf.size))
    t = buf.reshape(148, 16).astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t[:, :7] - t0) / 1e3
    print("which", which, "event ms/launch", ms.value)
    print("  probes (us after the earliest block start): start, barrier1, first warp done, last warp done, block done, barrier2, slice landed")
    print("  mean", np.round(rel.mean(0), 2), "\n  min ", np.round(rel.min(0), 2), "\n  max ", np.round(rel.max(0), 2))
    dur = rel[:, 3] - rel[:, 1]
    order = np.argsort(dur)
    print("  stream duration per block (us): min %.2f  median %.2f  max %.2f; slowest blocks %s" %
          (dur.min(), np.median(dur), dur.max(), order[-5:].tolist()))
