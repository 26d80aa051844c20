#!/usr/bin/env python
"""bench.py -- ADMM iterations/sec of the B200 OSQP engine on BASELINE.json's configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config {2,3,4,5}] [--impl reference]

Default workload (BASELINE.json: "ADMM iterations/sec on random sparse QP ... at 1 GPU and batched at 2/4/8 B200"):
  N = 1 : config 2 -- random sparse QP n=50,000 m=100,000 density 1e-3 (nnz(A)=5.0e6, nnz(P_full)=2.55e6), fp64, seed
          20262.  One *step* = one osqp_solve from a cold start to eps_abs=eps_rel=1e-4 (adaptive_rho_interval=25,
          check_termination=25, polish off).
  N > 1 : config 5 -- the batch of 8192 MPC QPs (n=30, m=60) sharded over the N ranks (contiguous blocks, data
          scattered once at setup, no collective in the loop); one step = one cold-start solve of the whole batch;
          value = ADMM iterations of all QPs on all ranks / max time over ranks ("strong": the batch is fixed).  The
          weak-scaling figure (8192 QPs per GPU) and the single-QP replica figure ride along as extra keys.
  --config 3 / 4 : the Lasso lambda-sweep (warm-started re-solves) and the portfolio QP with polishing, one GPU.

  value : the metric with everything resident in HBM (no input crosses PCIe inside the timed region)
  e2e   : the same metric through the public API with HOST buffers every step (H2D of the step's inputs, D2H of
          x*, y*, info inside the timed region)

`--impl reference` times the CPU stand-in for the reference's libosqp path on the box's host cores, on the same
config / metric / unit (rank 0 only): the oracle port (oracle/), which is the only CPU implementation of libosqp 0.6.2
that exists here -- OSQP_jll is an un-vendored binary and nothing under /root/reference compiles.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import problems  # noqa: E402

SEED = 20262
N_VARS, N_CONS, DENSITY = 50_000, 100_000, 1e-3
SETTINGS = dict(verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, check_termination=25,
                polish=False, max_iter=4000)
BATCH_SETTINGS = dict(verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, check_termination=25,
                      warm_start=False, max_iter=4000)
METRIC, UNIT = "admm_iterations_per_sec", "iter/s"


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


# ---------------------------------------------------------------------------------------------- CPU legs (oracle)
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_lib(pkg):
    """The CPU checker, used here ONLY as the timed CPU baseline (cpu_baseline / --impl reference)."""
    graft.build_oracle()
    lib = pkg.load_library(graft.ORACLE_LIB)
    lib.osqp_oracle_configure.argtypes = [C.c_longlong, C.c_double, C.c_longlong]
    lib.osqp_oracle_num_threads.restype = C.c_longlong
    lib.osqp_oracle_set_num_threads.argtypes = [C.c_longlong]
    lib.osqp_oracle_set_num_threads(host_cores())  # torchrun exports OMP_NUM_THREADS=1: use every core we may run on
    return lib


def cpu_single_qp(pkg, prob, settings, steps, warmup, pcg, label):
    """`steps` whole cold-start solves of one QP on the oracle.  pcg=True: reduced-KKT PCG backend with the engine's
    stopping rule on all host threads (OpenMP SpMV); pcg=False: libosqp's own algorithm, direct sparse LDL', 1 thread."""
    lib = oracle_lib(pkg)
    threads = int(lib.osqp_oracle_num_threads()) if pcg else 1
    lib.osqp_oracle_configure(3 if pcg else 0, 1e-3 if pcg else 1e-9, 0)
    try:
        mdl = pkg.Model(lib=graft.ORACLE_LIB)
        t0 = time.perf_counter()
        mdl.setup(**prob, **dict(settings, warm_start=False))
        setup_s = time.perf_counter() - t0
    finally:
        lib.osqp_oracle_configure(0, 1e-9, 0)
    iters, secs, status = 0, 0.0, None
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        r = mdl.solve()
        dt = time.perf_counter() - t0
        if s >= warmup:
            iters += r.info.iter
            secs += dt
            status = r.info.status
    mdl.clean()
    backend = f"reduced-KKT PCG backend (eta 1e-3), {threads} OpenMP threads" if pcg else "direct LDL' backend, 1 thread"
    return dict(value=iters / secs, iters=iters, secs=secs, threads=threads, setup_s=setup_s, status=status,
                sample=f"{steps} whole cold-start solve(s) of {label} ({iters // max(1, steps)} ADMM iterations each, "
                       f"status {status}), oracle port, {backend}")


def cpu_direct_attempt(n, m, density, seed, mem_gb=48, timeout_s=45):
    """BASELINE.md comparator A: libosqp's own path -- single-threaded direct LDL' of the full KKT -- attempted for
    real on the C2 instance in a child process under an address-space cap and a time limit."""
    code = (
        "import sys, resource, time\n"
        f"resource.setrlimit(resource.RLIMIT_AS, ({mem_gb} << 30, {mem_gb} << 30))\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import __graft_entry__ as g, problems\n"
        "pkg = g.load_package()\n"
        f"prob = problems.random_qp_c2({n}, {m}, {density}, {seed})\n"
        "mdl = pkg.Model(lib=g.ORACLE_LIB)\n"
        "t0 = time.perf_counter()\n"
        "mdl.setup(**prob, verbose=False, eps_abs=1e-4, eps_rel=1e-4, adaptive_rho_interval=25, max_iter=50)\n"
        "t1 = time.perf_counter(); r = mdl.solve(); t2 = time.perf_counter()\n"
        "print('OK', t1 - t0, r.info.iter / (t2 - t1))\n")
    t0 = time.perf_counter()
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout_s,
                           env=dict(os.environ, OMP_NUM_THREADS="1"))
        out = (r.stdout or "").strip().splitlines()
        if r.returncode == 0 and out and out[-1].startswith("OK"):
            _, fs, ips = out[-1].split()
            return {"outcome": "fit", "factor_s": float(fs), "iter_per_s": float(ips)}
        tail = ((r.stderr or "").strip().splitlines() or ["?"])[-1][:160]
        return {"outcome": f"did not fit under a {mem_gb} GB address-space cap", "after_s": time.perf_counter() - t0,
                "error": tail}
    except subprocess.TimeoutExpired:
        return {"outcome": f"factorisation not finished after {timeout_s} s (1 thread, {mem_gb} GB cap)",
                "after_s": time.perf_counter() - t0}


def cpu_batch(pkg, count_sample, steps, warmup):
    """Config 5 on the CPU: a sample of the batch, every QP a libosqp-style solve (oracle, direct LDL', single-threaded
    like libosqp), the QPs spread over all host threads by an OpenMP loop inside the oracle (osqp_oracle_solve_many) --
    no interpreter in the timed region."""
    lib = oracle_lib(pkg)
    lib.osqp_oracle_configure(0, 1e-9, 0)
    cores = host_cores()
    lib.osqp_oracle_set_num_threads(cores)
    lib.osqp_oracle_solve_many.restype = C.c_longlong
    lib.osqp_oracle_solve_many.argtypes = [C.POINTER(C.c_void_p), C.c_longlong]
    Pp, Ap, Px, Ax, q, l, u = problems.mpc_batch_c5(count_sample, SEED + 5)
    opts = dict(BATCH_SETTINGS)
    mdls = []
    for k in range(count_sample):
        mdl = pkg.Model(lib=graft.ORACLE_LIB)
        mdl.setup(**problems.batch_instance(Pp, Ap, Px, Ax, q, l, u, k), **opts)
        mdls.append(mdl)
    works = (C.c_void_p * count_sample)(*[C.cast(m.workspace, C.c_void_p) for m in mdls])
    iters, secs = 0, 0.0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        bad = lib.osqp_oracle_solve_many(works, count_sample)
        dt = time.perf_counter() - t0
        assert bad == 0
        if s >= warmup:
            iters += sum(int(m.workspace.contents.info.contents.iter) for m in mdls)
            secs += dt
    for mdl in mdls:
        mdl.clean()
    return dict(value=iters / secs, iters=iters, secs=secs, threads=cores,
                sample=f"{steps} cold-start solve(s) of the first {count_sample} of the 8192 MPC QPs, oracle port with "
                       f"the direct LDL' backend, one QP per OpenMP task on {cores} host threads")


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


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def measured_traffic(n, m, nnz):
    """DRAM bytes per launch from the round's own `ncu --set full` capture of admm_kernel on this workload
    (profiles/r2_traffic.json, written by profiles/ncu_traffic.py); None when no capture of this exact instance exists."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if tj.get("n") == n and tj.get("m") == m and tj.get("nnz") == nnz:
            return tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------- GPU legs
class Dist:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.torch = None
        self.dist = None

    def init(self):
        import torch
        import torch.distributed as dist

        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.torch, self.dist = torch, dist
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            dist.barrier()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor([float(v) for v in vals], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def profile_of(pkg, eng, mdl):
    p = pkg.types.B200Profile()
    assert eng.osqp_b200_get_profile(mdl.workspace, C.byref(p)) == 0
    return p


def run_single_qp(pkg, eng, D, prob, settings, steps, warmup, label, e2e_inputs=True):
    """value / e2e / roofline of cold-start solves of ONE resident QP (configs 2 and 4; the C2 replicas at N > 1)."""
    n, m = prob["P"].shape[0], prob["A"].shape[0]
    mdl = pkg.Model(lib=graft.LIB)
    t0 = time.perf_counter()
    mdl.setup(**prob, **settings)
    setup_s = time.perf_counter() - t0
    mdl.update_settings(warm_start=False)
    for _ in range(warmup):
        res = mdl.solve()
    p0 = profile_of(pkg, eng, mdl)
    D.barrier()
    t0 = time.perf_counter()
    iters = 0
    kern_ms = alg_bytes = pcg = polish_ms = 0.0
    for _ in range(steps):
        res = mdl.solve()
        iters += res.info.iter
        p = profile_of(pkg, eng, mdl)
        kern_ms += p.kernel_ms
        polish_ms += p.polish_ms
        alg_bytes += p.alg_bytes
        pcg += p.pcg_iters
    D.barrier()
    dt = time.perf_counter() - t0
    p1 = profile_of(pkg, eng, mdl)
    out = dict(iters=iters, dt=dt, kern_ms=kern_ms, polish_ms=polish_ms, alg_bytes=alg_bytes, pcg=pcg, setup_s=setup_s,
               launches=int(p1.launches - p0.launches), status=res.info.status, status_polish=res.info.status_polish,
               prof=p1, last=res)
    in_loop = {}
    if p1.pcg_iters > 0 and p1.streams:
        for name, k, nbytes in (("A_and_P", 0, p1.spmv_bytes_A + p1.spmv_bytes_P), ("At", 4, p1.spmv_bytes_At)):
            us = p1.phase_us[k] / p1.pcg_iters
            if us > 0:
                in_loop[name] = {"us": us, "alg_GBs": nbytes / us / 1e3, "frac_of_8TBs": nbytes / us / 1e3 / 8000.0}
        in_loop["us_per_pcg_iteration"] = sum(p1.phase_us[k] for k in range(8)) / p1.pcg_iters
    out["in_loop"] = in_loop
    # end to end through the public API with host buffers: update(q,l,u) + warm_start(x0,y0) (H2D), solve (D2H)
    mdl.update_settings(warm_start=True)
    x0, y0 = np.zeros(n), np.zeros(m)
    q, l, u = prob["q"].copy(), prob["l"].copy(), prob["u"].copy()
    for _ in range(2):
        mdl.update(q=q, l=l, u=u); mdl.warm_start(x=x0, y=y0); mdl.solve()
    D.barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(steps):
        mdl.update(q=q, l=l, u=u)
        mdl.warm_start(x=x0, y=y0)
        e2e_iters += mdl.solve().info.iter
    D.barrier()
    out["dt_e2e"], out["e2e_iters"] = time.perf_counter() - t0, e2e_iters
    out["h2d"], out["d2h"] = 8 * (n + 2 * m) + 8 * (n + m), 8 * (n + m) + 136
    out["mdl"] = mdl
    return out


def run_batch(pkg, D, total, steps, warmup, weak):
    """Config 5.  weak=False: `total` QPs sharded over the ranks; weak=True: `total` QPs on EVERY rank."""
    lo, hi = (0, total) if weak else pkg.shard_range(total, D.world, D.rank)
    Pp, Ap, Px, Ax, bq, bl, bu = problems.mpc_batch_c5(total, SEED + 5)
    bm = pkg.BatchModel(lib=graft.LIB)
    t0 = time.perf_counter()
    bm.setup(Pp, Ap, Px[lo:hi], Ax[lo:hi], bq[lo:hi], bl[lo:hi], bu[lo:hi], **BATCH_SETTINGS)
    setup_s = time.perf_counter() - t0
    for _ in range(max(3, warmup)):
        br = bm.solve(copy=False)
    D.barrier()
    t0 = time.perf_counter()
    iters, kern = 0, 0.0
    for _ in range(steps):
        br = bm.solve(copy=False)  # cold start in the kernel; x*, y*, info of every QP land in pinned host memory
        iters += int(br.iter.sum())
        kern += bm.kernel_ms
    D.barrier()
    dt = time.perf_counter() - t0
    # e2e: every step uploads the step's q, l, u from (pinned) host memory first
    q_, l_, u_ = bm.input_views()
    q_[:], l_[:], u_[:] = bq[lo:hi], np.clip(bl[lo:hi], -1e30, 1e30), np.clip(bu[lo:hi], -1e30, 1e30)
    for _ in range(2):
        bm.update(q=q_, l=l_, u=u_); bm.solve(copy=False)
    D.barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(steps):
        bm.update(q=q_, l=l_, u=u_)
        e2e_iters += int(bm.solve(copy=False).iter.sum())
    D.barrier()
    dt_e2e = time.perf_counter() - t0
    # the only collective of the path, after the loop: every rank receives all solutions
    gather_ms, gathered = None, None
    if not weak:
        pkg.batch.gather_sharded_device(bm, total, D.world, D.rank)  # warm-up (communicator, buffers)
        D.barrier()
        t0 = time.perf_counter()
        x_all = pkg.batch.gather_sharded_device(bm, total, D.world, D.rank)  # [total, n] on the device of every rank
        D.barrier()
        gather_ms, gathered = 1e3 * (time.perf_counter() - t0), int(x_all.shape[0])
    status = br.status_val
    unsolved = [(int(lo + k), int(status[k]), int(br.iter[k])) for k in np.nonzero(status != 1)[0][:8]]
    n, m = bm.n, bm.m
    cnt = hi - lo
    tmax = D.reduce([dt, kern, dt_e2e], "max")
    csum = D.reduce([iters, e2e_iters, int(np.sum(status == 1)), cnt * steps], "sum")
    bm.clean()
    return dict(dt=tmax[0], kern_ms=tmax[1], dt_e2e=tmax[2], iters=csum[0], e2e_iters=csum[1], solved=int(csum[2]),
                solves=int(csum[3]), setup_s=setup_s, qps_per_gpu=cnt, gather_ms=gather_ms, gathered_rows=gathered,
                unsolved=unsolved, h2d=8 * cnt * (n + 2 * m), d2h=8 * cnt * (n + m) + 56 * cnt,
                max_iter_qp=int(br.iter.max()), mean_iter_qp=float(br.iter.mean()))


def batch_summary(b, steps, scaling):
    return {"scaling": scaling, "qps_per_gpu": b["qps_per_gpu"], "value": b["iters"] / b["dt"], "unit": UNIT,
            "e2e": b["e2e_iters"] / b["dt_e2e"], "solves_per_sec": b["solves"] / b["dt"],
            "ms_per_step": 1e3 * b["dt"] / steps, "kernel_ms_per_step": b["kern_ms"] / steps,
            "e2e_ms_per_step": 1e3 * b["dt_e2e"] / steps, "mean_iters": b["mean_iter_qp"], "max_iters": b["max_iter_qp"],
            "solved": b["solved"], "of": b["solves"] // steps * 1, "not_solved_sample": b["unsolved"],
            "gather_ms": b["gather_ms"], "setup_s": b["setup_s"]}


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
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5],
                    help="BASELINE.json config; 0 = the default of the contract (2 at one GPU, 5 at several)")
    ap.add_argument("--n", type=int, default=N_VARS)
    ap.add_argument("--m", type=int, default=N_CONS)
    ap.add_argument("--density", type=float, default=DENSITY)
    ap.add_argument("--batch", type=int, default=8192, help="QPs of config 5")
    ap.add_argument("--cpu-iters", type=int, default=250,
                    help="reference arm, config 2: ADMM iterations per timed step (a bounded sample of the cold-start "
                         "solve; 0 = whole solves).  One whole solve is always run and reported first.")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (batch at N=1, replicas/weak at N>1)")
    ap.add_argument("--lasso", default="10000x50000x0.15", help="config 3: features x samples x density")
    ap.add_argument("--assets", type=int, default=20000, help="config 4: assets (factors = assets / 100)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    D = Dist()
    cfg = args.config or (2 if max(D.world, args.gpus) == 1 else 5)
    if cfg != 5 and D.world > 1 and args.config:
        raise SystemExit("bench.py: configs 2-4 are single-QP workloads (replicas only); run them with --gpus 1")

    c2_label = (f"random sparse QP n={args.n} m={args.m} density={args.density:g} fp64 (SURVEY 8d C2: nnz(A)=5.0e6, "
                f"nnz(P_full)=2.55e6), eps=1e-4, adaptive_rho_interval=25, cold-start solve per step")
    c5_label = (f"{args.batch} MPC QPs n=30 m=60, one sparsity pattern (SURVEY 8d C5), eps=1e-4, cold-start solve of the "
                f"whole batch per step")
    nf, ns, ld = args.lasso.split("x")
    nf, ns, ld = int(nf), int(ns), float(ld)
    workloads = {
        2: c2_label,
        3: f"Lasso as QP, {nf} features x {ns} samples, density {ld:g} (SURVEY 8d C3), eps=1e-4, one step = 11-point "
           f"lambda sweep of warm-started re-solves (osqp_update_lin_cost + osqp_solve)",
        4: f"portfolio QP, {args.assets} assets, {args.assets // 100} factors (SURVEY 8d C4), eps=1e-4, polish on, "
           f"cold-start solve + polish per step",
        5: c5_label,
    }
    config = {"workload": workloads[cfg], "baseline_config": cfg, "seed": SEED,
              "parallelism": ("one QP on one GPU" if cfg != 5 else
                              f"contiguous blocks of ceil({args.batch}/{D.world}) QPs per GPU, no collectives in the loop")}
    if cfg == 2:
        config.update(n=args.n, m=args.m)

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if D.rank != 0:
            return 0
        pkg = graft.load_package()  # the product library is neither built nor mapped on this arm
        steps, warm = max(1, args.steps), min(args.warmup, 1)
        extra = {}
        if cfg == 2:
            prob = make_problem(args.n, args.m, args.density, SEED)
            whole = cpu_single_qp(pkg, prob, SETTINGS, 1, 0, True, "C2")
            extra["whole_solve"] = {"value": whole["value"], "unit": UNIT, "iters": whole["iters"], "secs": whole["secs"],
                                    "status": whole["status"], "setup_s": whole["setup_s"]}
            if args.cpu_iters > 0 and args.cpu_iters < whole["iters"]:
                cpu = cpu_single_qp(pkg, prob, dict(SETTINGS, max_iter=args.cpu_iters), steps, 0, True,
                                    f"C2, first {args.cpu_iters} ADMM iterations of the cold-start solve")
            else:
                cpu = whole if steps == 1 else cpu_single_qp(pkg, prob, SETTINGS, steps, 0, True, "C2")
            note = ("libosqp 0.6.2 (OSQP_jll) is an un-vendored binary: the CPU arm is the oracle port in its "
                    "reduced-KKT PCG mode on all host threads; libosqp's own single-threaded direct LDL' of this "
                    "150k KKT does not fit (cpu_baseline.direct_ldl_attempt of the B200 arm records a real attempt)")
        elif cfg == 3:
            prob, lam_max, q_of, _ = problems.lasso_c3(nf, ns, ld, SEED + 1)
            cpu = cpu_single_qp(pkg, dict(prob, q=q_of(lam_max)), dict(SETTINGS, max_iter=100), steps, 0, True,
                                "C3 at lambda_max (first 100 ADMM iterations)")
            note = "bounded sample: the first point of the lambda sweep, 100 ADMM iterations, oracle PCG backend"
        elif cfg == 4:
            prob = problems.portfolio_c4(args.assets, args.assets // 100, SEED + 2)
            cpu = cpu_single_qp(pkg, prob, dict(SETTINGS, polish=False, max_iter=20), steps, 0, True,
                                "C4, first 20 ADMM iterations")
            note = ("bounded sample: 20 ADMM iterations on the oracle's PCG backend (its min-degree ordering makes the direct "
                    "LDL' of this 40k KKT take > 10 min at setup)")
        else:
            cpu = cpu_batch(pkg, min(args.batch, 512), steps, warm)
            note = "libosqp-style solves (oracle, direct LDL') of a 512-QP sample, one QP per task on all host threads"
        emit({
            "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cpu["secs"] / steps,
            "higher_is_better": True, "scaling": "strong" if cfg == 5 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": cpu["threads"], "kind": "port",
                             "sample": cpu["sample"]},
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "product_library_mapped": any("libosqp" in ln for ln in open("/proc/self/maps")),
            "note": note, **extra})
        return 0

    # ------------------------------------------------------------------ B200 arm
    if D.rank == 0 or not os.path.exists(graft.LIB):
        graft.build()  # under torchrun the other ranks use what rank 0 built (or what travelled with the snapshot)
    pkg = graft.load_package()
    D.init()
    eng = pkg.load_library(graft.LIB)
    peak, peak_src = peak_hbm()
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config}

    if cfg in (2, 4):
        if cfg == 2:
            prob, settings = make_problem(args.n, args.m, args.density, SEED), SETTINGS
        else:
            prob, settings = problems.portfolio_c4(args.assets, args.assets // 100, SEED + 2), dict(SETTINGS, polish=True, max_iter=10000)
        nnz = int(2 * prob["A"].nnz + prob["P"].nnz)
        r = run_single_qp(pkg, eng, D, prob, settings, args.steps, args.warmup, workloads[cfg])
        clocks = sampler.stop() if D.rank == 0 else None
        p1 = r["prof"]
        n, m = prob["P"].shape[0], prob["A"].shape[0]
        total_ms = r["kern_ms"] + r["polish_ms"]
        achieved = r["alg_bytes"] / (r["kern_ms"] * 1e-3) / 1e9 if r["kern_ms"] > 0 else 0.0
        config["l2"] = (f"no flush between steps: one K-apply streams {10.6 * nnz / 1e6:.0f} MB of matrix data against a "
                        f"126 MB L2, so a step's {r['iters'] // args.steps} ADMM iterations x "
                        f"{r['pcg'] / max(1, r['iters']):.1f} K-applies evict each other continuously; the solver's own "
                        "reuse across iterations is part of the workload")
        spmv = {}
        if D.rank == 0 and cfg == 2:
            fp = C.POINTER(C.c_double)
            eng.osqp_b200_spmv.restype = C.c_longlong
            rng = np.random.default_rng(1)
            for which, name, ilen, nbytes in ((0, "A", n, p1.spmv_bytes_A), (1, "At", m, p1.spmv_bytes_At),
                                              (2, "P", n, p1.spmv_bytes_P)):
                vin = rng.standard_normal(ilen)
                ms = C.c_double()
                rc = eng.osqp_b200_spmv(r["mdl"].workspace, C.c_longlong(which), vin.ctypes.data_as(fp), None,
                                        C.c_longlong(50), C.byref(ms))
                if rc == 0 and ms.value > 0:
                    spmv[name] = {"ms": ms.value, "alg_GBs": nbytes / ms.value / 1e6}
        r["mdl"].clean()
        cpu = None
        if D.rank == 0 and not args.no_cpu_baseline:
            if cfg == 2:
                c0 = cpu_single_qp(pkg, prob, settings, 1, 0, True, "C2")
                cpu = {"value": c0["value"], "unit": UNIT, "cores": c0["threads"], "kind": "port", "sample": c0["sample"],
                       "direct_ldl_attempt": cpu_direct_attempt(args.n, args.m, args.density, SEED)}
            else:
                c0 = cpu_single_qp(pkg, prob, dict(settings, polish=False, max_iter=20), 1, 0, True,
                                   "C4, first 20 ADMM iterations")
                cpu = {"value": c0["value"], "unit": UNIT, "cores": c0["threads"], "kind": "port", "sample": c0["sample"]}
        batch_line = None
        if cfg == 2 and not args.no_extras and args.batch > 0:
            batch_line = batch_summary(run_batch(pkg, D, args.batch, args.steps, args.warmup, False), args.steps, "strong")
            batch_line["workload"] = c5_label
        line.update({
            "value": r["iters"] / r["dt"], "ms_per_step": 1e3 * r["dt"] / args.steps, "scaling": "weak",
            "e2e": {"value": r["e2e_iters"] / r["dt_e2e"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
                    "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": r["launches"],
            "roofline": {"bound": "hbm", "kernel": "admm_kernel (persistent; one launch per solve)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "peak_source": peak_src,
                         "traffic": measured_traffic(n, m, nnz), "alg_bytes_per_launch": r["alg_bytes"] / args.steps,
                         "launch_ms": r["kern_ms"] / args.steps,
                         "pcg_iters_per_admm_iter": r["pcg"] / max(1.0, float(r["iters"])),
                         "spmv_in_loop": r["in_loop"], "spmv_standalone_launch": spmv},
            "cpu_baseline": cpu, "batch": batch_line, "clocks": clocks,
            "solve": {"status": r["status"], "status_polish": r["status_polish"],
                      "admm_iters_per_solve": r["iters"] / args.steps, "setup_s": r["setup_s"],
                      "polish_ms_per_solve": r["polish_ms"] / args.steps, "admm_ms_per_solve": r["kern_ms"] / args.steps,
                      "grid": int(p1.grid), "block": int(p1.block), "lanes": [int(p1.lanes_A), int(p1.lanes_N)]}})
        del total_ms
    elif cfg == 3:
        prob, lam_max, q_of, n = problems.lasso_c3(nf, ns, ld, SEED + 1)
        lams = np.logspace(0, -2, 11) * lam_max
        mdl = pkg.Model(lib=graft.LIB)
        t0 = time.perf_counter()
        mdl.setup(**dict(prob, q=q_of(lam_max)), **dict(SETTINGS, max_iter=10000))
        setup_s = time.perf_counter() - t0
        p0 = None
        tot = dict(iters=0, dt=0.0, dt_e2e=0.0, kern_ms=0.0, alg=0.0, pcg=0.0)
        x0, y0 = np.zeros(n), np.zeros(prob["A"].shape[0])
        statuses = set()
        for s in range(args.warmup + args.steps):
            timed = s >= args.warmup
            if timed and p0 is None:
                p0 = profile_of(pkg, eng, mdl)
            mdl.warm_start(x=x0, y=y0)  # every sweep starts from zero
            D.barrier()
            t_sweep = time.perf_counter()
            for lam in lams:
                mdl.update(q=q_of(lam))  # src/interface.jl:240-246; iterates are kept (warm_start = 1)
                t1 = time.perf_counter()
                res = mdl.solve()
                dt1 = time.perf_counter() - t1
                statuses.add(res.info.status)
                if timed:
                    p = profile_of(pkg, eng, mdl)
                    tot["iters"] += res.info.iter; tot["dt"] += dt1; tot["kern_ms"] += p.kernel_ms
                    tot["alg"] += p.alg_bytes; tot["pcg"] += p.pcg_iters
            if timed:
                tot["dt_e2e"] += time.perf_counter() - t_sweep
        p1 = profile_of(pkg, eng, mdl)
        clocks = sampler.stop() if D.rank == 0 else None
        mdl.clean()
        cpu = None
        if not args.no_cpu_baseline:
            c0 = cpu_single_qp(pkg, dict(prob, q=q_of(lam_max)), dict(SETTINGS, max_iter=100), 1, 0, True,
                               "C3 at lambda_max (first 100 ADMM iterations)")
            cpu = {"value": c0["value"], "unit": UNIT, "cores": c0["threads"], "kind": "port", "sample": c0["sample"]}
        achieved = tot["alg"] / (tot["kern_ms"] * 1e-3) / 1e9 if tot["kern_ms"] > 0 else 0.0
        config["l2"] = "inputs larger than L2: the matrix streams are ~1.6 GB per K-apply"
        line.update({
            "value": tot["iters"] / tot["dt"], "ms_per_step": 1e3 * tot["dt_e2e"] / args.steps, "scaling": "weak",
            "e2e": {"value": tot["iters"] / tot["dt_e2e"], "unit": UNIT, "h2d_bytes_per_step": 11 * 8 * n,
                    "d2h_bytes_per_step": 11 * (8 * (n + prob["A"].shape[0]) + 136)},
            "gpu_launches": int(p1.launches - p0.launches),
            "roofline": {"bound": "hbm", "kernel": "admm_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                         "pcg_iters_per_admm_iter": tot["pcg"] / max(1, tot["iters"]),
                         "launch_ms": tot["kern_ms"] / (11 * args.steps)},
            "cpu_baseline": cpu, "clocks": clocks,
            "solve": {"statuses": sorted(statuses), "admm_iters_per_sweep": tot["iters"] / args.steps, "setup_s": setup_s}})
    else:
        # config 5, the multi-GPU workload of BASELINE.json: strong scaling over the ranks
        b = run_batch(pkg, D, args.batch, args.steps, args.warmup, False)
        extras = {}
        if not args.no_extras and D.world > 1:
            extras["batch_weak"] = batch_summary(run_batch(pkg, D, args.batch, args.steps, args.warmup, True), args.steps,
                                                 "weak")
            extras["batch_weak"]["workload"] = f"{args.batch} MPC QPs on EVERY GPU"
            prob = make_problem(args.n, args.m, args.density, SEED)
            r = run_single_qp(pkg, eng, D, prob, SETTINGS, min(args.steps, 5), 3, c2_label)
            r["mdl"].clean()
            agg = D.reduce([r["iters"]], "sum")[0] / D.reduce([r["dt"]], "max")[0]
            extras["replicas_c2"] = {"workload": c2_label + f" -- {D.world} independent replicas, one per GPU",
                                     "value": agg, "unit": UNIT, "scaling": "weak"}
        clocks = sampler.stop() if D.rank == 0 else None
        cpu = None
        if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
            c0 = cpu_batch(pkg, min(args.batch, 512), 1, 1)
            cpu = {"value": c0["value"], "unit": UNIT, "cores": c0["threads"], "kind": "port", "sample": c0["sample"]}
        s = batch_summary(b, args.steps, "strong")
        config["l2"] = ("not applicable: every QP's working set lives in shared memory / registers for the whole solve; "
                        "HBM is touched once per solve (state in, solution out)")
        line.update({
            "value": s["value"], "ms_per_step": s["ms_per_step"], "scaling": "strong",
            "e2e": {"value": s["e2e"], "unit": UNIT, "h2d_bytes_per_step": b["h2d"], "d2h_bytes_per_step": b["d2h"]},
            "gpu_launches": 2 * args.steps,
            # HBM is touched once per solve: the per-QP state comes in (values, inverse factor, vectors: 12.8 KB for the
            # 30 x 60 MPC QP) and x*, y*, info and the iterates go out (1.1 KB) -- 13.9 KB per QP measured by ncu
            # (profiles/r2_ncu_batch.md: 104.7 MB read + 8.9 MB written per 8192-QP launch)
            "roofline": {"bound": "hbm", "kernel": "batch_fast_solve_kernel",
                         "achieved": 13.9e3 * b["qps_per_gpu"] / (b["kern_ms"] / args.steps * 1e-3) / 1e9 if b["kern_ms"] > 0 else None,
                         "peak": peak, "unit": "GB/s",
                         "frac": 13.9e3 * b["qps_per_gpu"] / (b["kern_ms"] / args.steps * 1e-3) / 1e9 / peak if b["kern_ms"] > 0 else None,
                         "peak_source": peak_src, "traffic": 113.6e6 * b["qps_per_gpu"] / 8192.0,
                         "note": "per GPU; latency / issue bound, not HBM bound: the per-QP working set lives in shared "
                                 "memory and registers for the whole solve (DESIGN.md 7)"},
            "cpu_baseline": cpu, "clocks": clocks, "batch": s, **extras})
    if D.rank == 0:
        emit(line)
    D.finish()
    return 0


if __name__ == "__main__":
    sys.exit(main())
