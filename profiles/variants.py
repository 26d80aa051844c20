#!/usr/bin/env python
"""Builds compile-time variants of the engine (warps per block, stream pipeline depth, scan interleave) into
osqp.jl_b200/lib/variants/ for A/B measurements with profiles/profile_driver.py --lib.  Not part of the product."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

VARIANTS = {
    "w16d4i1": dict(OSQP_B200_WARPS=16, OSQP_B200_DEPTH=4, OSQP_B200_ILP=1),
    "w16d4i2": dict(OSQP_B200_WARPS=16, OSQP_B200_DEPTH=4, OSQP_B200_ILP=2),
    "w32d2i1": dict(OSQP_B200_WARPS=32, OSQP_B200_DEPTH=2, OSQP_B200_ILP=1),
    "w32d2i2": dict(OSQP_B200_WARPS=32, OSQP_B200_DEPTH=2, OSQP_B200_ILP=2),
    "w32d4i2": dict(OSQP_B200_WARPS=32, OSQP_B200_DEPTH=4, OSQP_B200_ILP=2),
    "lr6": dict(OSQP_B200_DEPTH_LR=6),
    "lr8": dict(OSQP_B200_DEPTH_LR=8),
    "lr2": dict(OSQP_B200_DEPTH_LR=2),
    "m32d6": dict(OSQP_B200_DEPTH_LR32=6),
    "m32d8": dict(OSQP_B200_DEPTH_LR32=8),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    os.makedirs(os.path.join(g.PKG_DIR, "lib", "variants"), exist_ok=True)
    for name in names:
        out = os.path.join(g.PKG_DIR, "lib", "variants", f"libosqp_{name}.so")
        g.build_engine(force=True, defines=VARIANTS[name], out=out, verbose="-v" in os.environ.get("VARIANT_FLAGS", ""))
        print("built", out)
