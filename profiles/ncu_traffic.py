#!/usr/bin/env python
"""DRAM traffic of one full `admm_kernel` launch on the bench workload, from an `ncu --set full` report.

  OSQP_B200_PAIRS=0 ncu --set full --clock-control none -k regex:admm_kernel -c 1 -o gpurun_out/admm_full \
      python profiles/profile_driver.py --solves 1 --spmv-reps 1          # on the GPU box (gpurun)
  python profiles/ncu_traffic.py gpurun_out/admm_full.ncu-rep             # here -> profiles/r2_traffic.json

(ncu cannot intercept the clustered cooperative launch, so the capture runs the unpaired variant of the [A;P] stream:
same matrix stream, the pair combine goes through `part` and one more grid barrier.)  bench.py reads the JSON for
`roofline.traffic` when n, m and nnz match the workload it times.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

    def nbytes(key):
        u, v = d[key]
        return float(v) * scale[u]

    import bench
    prob = bench.make_problem(bench.N_VARS, bench.N_CONS, bench.DENSITY, bench.SEED)
    out = {
        "n": bench.N_VARS, "m": bench.N_CONS, "nnz": int(2 * prob["A"].nnz + prob["P"].nnz),
        "dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
        "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
        "launch_ms_under_ncu": float(d["gpu__time_duration.sum"][1]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[d["gpu__time_duration.sum"][0]],
        "lts_hit_rate_pct": float(d["lts__t_sector_hit_rate.pct"][1]),
        "variant": "unpaired [A;P] stream (OSQP_B200_PAIRS=0: ncu cannot intercept the clustered cooperative launch)",
        "source": os.path.basename(rep), "command": "profiles/profile_driver.py --solves 1 (one full cold-start solve of the bench workload)",
    }
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
