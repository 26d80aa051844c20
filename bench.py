#!/usr/bin/env python
"""bench.py -- ADMM iterations/sec of the B200 OSQP engine on BASELINE.json's config 2.

Workload (SURVEY.md 8d "C2", frozen here): random sparse QP, n=50,000, m=100,000, density 1e-3
(nnz(A)=5.0e6, nnz(P_full)~2.5e6), fp64, seed 20262 (+rank).  One *step* = one osqp_solve from a cold
start to eps_abs=eps_rel=1e-4 (adaptive_rho_interval=25, check_termination=25, polish off).

  value : total ADMM iterations of the K timed steps / wall time, problem data resident in HBM
          (settings.warm_start=0 makes the kernel cold-start itself; no input crosses PCIe)
  e2e   : same metric through the public API with HOST buffers every step:
          Model.update(q,l,u) + Model.warm_start(x0,y0) (H2D) + Model.solve() (D2H of x*, y*, info)
  N > 1 : one process per GPU (torchrun), one independent copy of the QP per rank (same seed, so the work per
          GPU is exactly that of N = 1), no collective in the loop; value = sum of iterations over ranks / max
          time over ranks ("weak" scaling).

`--impl reference` times the CPU stand-in for the reference's libosqp path (the oracle port, reduced-KKT
PCG backend on all host cores -- the direct LDL' of a 150k KKT with this pattern does not fit)
on the same workload, each step a bounded sample (max_iter capped), rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import problems  # noqa: E402

SEED = 20262
N_VARS, N_CONS, DENSITY = 50_000, 100_000, 1e-3
SETTINGS = dict(verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, check_termination=25,
                polish=False, max_iter=4000)


def make_problem(n, m, density, seed):
    """SURVEY.md 8d config C2 construction (problems.py)."""
    return problems.random_qp_c2(n, m, density, seed)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_lib(pkg):
    lib = pkg.load_library(graft.ORACLE_LIB)
    lib.osqp_oracle_configure.argtypes = [C.c_longlong, C.c_double, C.c_longlong]
    lib.osqp_oracle_num_threads.restype = C.c_longlong
    lib.osqp_oracle_set_num_threads.argtypes = [C.c_longlong]
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    lib.osqp_oracle_set_num_threads(cores)  # torchrun sets OMP_NUM_THREADS=1: use every host core we may run on
    return lib


def run_cpu(pkg, prob, steps, warmup, max_iter):
    """Oracle port, reduced-KKT PCG backend (same stopping rule as the engine), all host threads."""
    lib = oracle_lib(pkg)
    threads = int(lib.osqp_oracle_num_threads())
    lib.osqp_oracle_configure(3, 1e-3, 0)
    try:
        mdl = pkg.Model(lib=graft.ORACLE_LIB)
        t0 = time.perf_counter()
        mdl.setup(**prob, **dict(SETTINGS, max_iter=max_iter, warm_start=False))
        setup_s = time.perf_counter() - t0
    finally:
        lib.osqp_oracle_configure(0, 1e-9, 0)
    iters, secs = 0, 0.0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        r = mdl.solve()
        dt = time.perf_counter() - t0
        if s >= warmup:
            iters += r.info.iter
            secs += dt
    mdl.clean()
    return dict(value=iters / secs, iters=iters, secs=secs, threads=threads, setup_s=setup_s,
                sample=f"{steps} solve(s) capped at max_iter={max_iter} ADMM iterations each (cold start), "
                       f"oracle PCG backend, {threads} OpenMP threads")


_JSON_FD = None


def emit(line):
    """The one JSON line of the contract goes to the process's original stdout; everything else that anybody prints
    (NCCL's version banner, verbose solvers, warnings of C libraries) was re-routed to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_VARS)
    ap.add_argument("--m", type=int, default=N_CONS)
    ap.add_argument("--density", type=float, default=DENSITY)
    ap.add_argument("--cpu-iters", type=int, default=40, help="ADMM iterations per CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=8192, help="QPs in the batched leg (config 5); 0 disables it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = (f"random sparse QP n={args.n} m={args.m} density={args.density:g} fp64 (SURVEY 8d C2), "
                f"eps=1e-4, adaptive_rho_interval=25, cold-start solve per step")
    config = {"workload": workload, "n": args.n, "m": args.m, "seed": SEED,
              "parallelism": f"{world} independent replica(s) of the QP, one per GPU, no collectives in the loop"}

    if rank == 0 or not os.path.exists(graft.LIB):
        graft.build()  # under torchrun the other ranks use what rank 0 built (or what travelled with the snapshot)
    pkg = graft.load_package()

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        prob = make_problem(args.n, args.m, args.density, SEED)
        cpu = run_cpu(pkg, prob, max(1, args.steps), min(args.warmup, 1), args.cpu_iters)
        line = {
            "impl": "reference", "metric": "admm_iterations_per_sec", "value": cpu["value"], "unit": "iter/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cpu["secs"] / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": cpu["value"], "unit": "iter/s", "cores": cpu["threads"], "kind": "port",
                             "sample": cpu["sample"]},
            "e2e": {"value": cpu["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "libosqp 0.6.2 (OSQP_jll) is an un-vendored binary: the CPU arm is the oracle port in its "
                    "reduced-KKT PCG mode; the direct LDL' mode does not fit this pattern in memory",
        }
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()  # rank 0 has finished building before anybody loads the library
    eng = pkg.load_library(graft.LIB)
    def profile(mdl):
        p = pkg.types.B200Profile()
        assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(p)) == 0
        return p

    prob = make_problem(args.n, args.m, args.density, SEED)  # every rank solves the same QP: per-GPU work is fixed
    n, m = args.n, args.m
    mat_mb = (10.6 * (2 * prob["A"].nnz + prob["P"].nnz)) / 1e6  # 10 B per stored entry, ~6 % quad padding
    config["l2"] = (f"no flush: one K-apply streams {mat_mb:.0f} MB of matrix data (L2 = 126 MB) and every termination "
                    f"check reads another {12.0 * (2 * prob['A'].nnz + prob['P'].nnz) / 1e6:.0f} MB of CSR copies; the "
                    "solver's own reuse across iterations is part of the workload (ncu: 58 % L2 hit rate, DRAM traffic "
                    "in roofline.traffic)")
    mdl = pkg.Model(lib=graft.LIB)
    t0 = time.perf_counter()
    mdl.setup(**prob, **SETTINGS)
    setup_s = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value)
    mdl.update_settings(warm_start=False)
    for _ in range(args.warmup):
        res = mdl.solve()
    p0 = profile(mdl)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    iters = 0
    kern_ms = alg_bytes = pcg = 0.0
    for _ in range(args.steps):
        res = mdl.solve()
        iters += res.info.iter
        p = profile(mdl)
        kern_ms += p.kernel_ms
        alg_bytes += p.alg_bytes
        pcg += p.pcg_iters
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    p1 = profile(mdl)
    launches = int(p1.launches - p0.launches)
    status = res.info.status
    # SpMV phases as they run INSIDE the fused launch (block 0's device-side phase clock of the last step): the
    # [A; P] stream incl. the vector-slice staging and the pair combine, and the A' stream
    in_loop = {}
    if p1.pcg_iters > 0 and p1.streams:
        for name, k, nbytes in (("A_and_P", 0, p1.spmv_bytes_A + p1.spmv_bytes_P), ("At", 4, p1.spmv_bytes_At)):
            us = p1.phase_us[k] / p1.pcg_iters
            if us > 0:
                in_loop[name] = {"us": us, "alg_GBs": nbytes / us / 1e3, "frac_of_8TBs": nbytes / us / 1e3 / 8000.0}
        in_loop["us_per_pcg_iteration"] = sum(p1.phase_us[k] for k in range(8)) / p1.pcg_iters

    # ---- end to end through the public API with host buffers (e2e)
    mdl.update_settings(warm_start=True)
    x0, y0 = np.zeros(n), np.zeros(m)
    q, l, u = prob["q"].copy(), prob["l"].copy(), prob["u"].copy()
    for _ in range(2):
        mdl.update(q=q, l=l, u=u); mdl.warm_start(x=x0, y=y0); mdl.solve()
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        mdl.update(q=q, l=l, u=u)
        mdl.warm_start(x=x0, y=y0)
        r2 = mdl.solve()
        e2e_iters += r2.info.iter
    barrier()
    dt_e2e = time.perf_counter() - t0

    # ---- standalone SpMV kernels (same device code / work split as the ADMM kernel)
    spmv = {}
    if rank == 0:
        fp = C.POINTER(C.c_double)
        eng.osqp_b200_spmv.restype = C.c_longlong
        rng = np.random.default_rng(1)
        for which, name, ilen, nbytes in ((0, "A", n, p1.spmv_bytes_A), (1, "At", m, p1.spmv_bytes_At),
                                          (2, "P", n, p1.spmv_bytes_P)):
            vin = rng.standard_normal(ilen)
            ms = C.c_double()
            rc = eng.osqp_b200_spmv(mdl.workspace, C.c_longlong(which), vin.ctypes.data_as(fp), None,
                                    C.c_longlong(50), C.byref(ms))
            if rc == 0 and ms.value > 0:
                spmv[name] = {"ms": ms.value, "alg_GBs": nbytes / ms.value / 1e6}

    # ---- batched leg: BASELINE config 5, 8192 MPC QPs (n=30, m=60) sharded over the ranks (strong scaling)
    batch_line = None
    if args.batch > 0:
        lo, hi = pkg.shard_range(args.batch, world, rank)
        Pp, Ap, Px, Ax, bq, bl, bu = problems.mpc_batch_c5(args.batch, SEED + 5)
        bm = pkg.BatchModel(lib=graft.LIB)
        bsettings = dict(verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, check_termination=25,
                         warm_start=False, max_iter=4000)
        t0 = time.perf_counter()
        bm.setup(Pp, Ap, Px[lo:hi], Ax[lo:hi], bq[lo:hi], bl[lo:hi], bu[lo:hi], **bsettings)
        b_setup = time.perf_counter() - t0
        for _ in range(3):
            br = bm.solve()
        barrier()
        t0 = time.perf_counter()
        b_iters, b_kern = 0, 0.0
        for _ in range(args.steps):
            br = bm.solve()  # host buffers in and out: x*, y*, info of every QP come back each step
            b_iters += int(br.iter.sum())
            b_kern += bm.kernel_ms
        barrier()
        b_dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        x_all = pkg.batch.gather_sharded(br.x, args.batch, world, rank, device="cuda")  # the only collective
        barrier()
        b_gather = time.perf_counter() - t0
        solved = int(np.sum(br.status_val == 1))
        bt = torch.tensor([b_dt, b_kern], device="cuda", dtype=torch.float64)
        bc = torch.tensor([float(b_iters), float(solved)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bc, op=dist.ReduceOp.SUM)
        batch_line = {
            "workload": f"{args.batch} MPC QPs n=30 m=60, one pattern, cold start, eps=1e-4 (SURVEY 8d C5)",
            "scaling": "strong", "qps_per_gpu": hi - lo,
            "qp_iterations_per_sec": float(bc[0]) / float(bt[0]), "solves_per_sec": args.batch * args.steps / float(bt[0]),
            "kernel_ms_per_step": float(bt[1]) / args.steps, "ms_per_step": 1e3 * float(bt[0]) / args.steps,
            "mean_iters": float(bc[0]) / (args.batch * args.steps), "solved": int(bc[1]), "setup_s": b_setup,
            "gather_ms": 1e3 * b_gather, "gathered_rows": int(x_all.shape[0]),
        }
        bm.clean()

    # ---- aggregate over ranks: sum of iterations, max of time
    if world > 1:
        t = torch.tensor([dt, dt_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([float(iters), float(e2e_iters), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        dt, dt_e2e = float(t[0]), float(t[1])
        iters_all, e2e_iters_all, launches_all = float(c[0]), float(c[1]), int(c[2])
    else:
        iters_all, e2e_iters_all, launches_all = float(iters), float(e2e_iters), launches

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("n") == n and tj.get("m") == m:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            c0 = run_cpu(pkg, prob, 1, 0, args.cpu_iters)
            cpu = {"value": c0["value"], "unit": "iter/s", "cores": c0["threads"], "kind": "port",
                   "sample": c0["sample"]}
        line = {
            "metric": "admm_iterations_per_sec", "value": iters_all / dt, "unit": "iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_iters_all / dt_e2e, "unit": "iter/s",
                    "h2d_bytes_per_step": 8 * (n + 2 * m) + 8 * (n + m), "d2h_bytes_per_step": 8 * (n + m) + 136},
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "kernel": "admm_kernel (persistent; one launch per solve)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "peak_source": peak_src, "traffic": traffic,
                         "alg_bytes_per_launch": alg_bytes / args.steps, "launch_ms": kern_ms / args.steps,
                         "pcg_iters_per_admm_iter": pcg / max(1.0, float(iters)), "spmv_in_loop": in_loop,
                         "spmv_standalone_launch": spmv},
            "cpu_baseline": cpu,
            "batch": batch_line,
            "clocks": clocks,
            "solve": {"status": status, "admm_iters_per_solve": iters / args.steps, "setup_s": setup_s,
                      "grid": int(p1.grid), "block": int(p1.block), "lanes": [int(p1.lanes_A), int(p1.lanes_N)]},
        }
        emit(line)
    mdl.clean()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
