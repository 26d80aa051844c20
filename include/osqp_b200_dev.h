/*
 * osqp_b200_dev.h -- measurement and self-test entry points of the DEVELOPMENT build of the engine
 * (osqp.jl_b200/lib/libosqp_dev.so, compiled with -DOSQP_B200_DEVTOOLS).  They are not part of the product:
 * lib/libosqp.so neither contains the kernels behind them nor exports these names.  Used by
 * tests/test_engine_parity.py (reduction self-test) and profiles/membench.py only.
 */
#ifndef OSQP_B200_DEV_H
#define OSQP_B200_DEV_H
#include "osqp_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Stream micro-benchmark: GB/s of reading `mbytes` MB with the load shape of the tile-stream phase (per lane and
 * chunk 2 x 16 B + 8 B, `depth` chunks in flight).  pattern 0: one contiguous share per warp; 1: the 16 warps of a
 * block interleave chunk by chunk; 2: as 1 with values and columns of a chunk in one 1280 B record.  < 0 on error. */
c_float osqp_b200_membench(c_int mbytes, c_int pattern, c_int depth, c_int reps);

/* ns per grid barrier on the workspace's persistent grid.  mode 0: bare barrier; 1: barrier with a 2-slot reduction;
 * 2: barrier after ~32 scattered 8 B stores per block (write drain). */
c_float osqp_b200_barrier_bench(OSQPWorkspace *work, c_int iters, c_int mode);

/* Co-resident thread-block clusters of size `csize` for the workspace's persistent kernel (occupancy query). */
c_int osqp_b200_cluster_probe(OSQPWorkspace *work, c_int csize);

/* Self-test of the grid-wide reductions: every thread of the persistent grid contributes (global index + 1);
 * out[6] = {sum, max} by the fp64 tree, by the fixed-point atomics scaled with `ref`, and by a second fixed-point
 * call (0.5 * index as the summand).  A tiny `ref` forces the overflow fallback. */
c_int osqp_b200_reduce_selftest(OSQPWorkspace *work, c_float ref, c_float *out);

#ifdef __cplusplus
}
#endif
#endif
