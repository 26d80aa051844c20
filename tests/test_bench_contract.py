"""bench.py's reference arm runs on the CPU: check the JSON line it prints against the driver contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "0", "--n", "1500", "--m", "3000", "--density", "0.01", "--cpu-iters", "10"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # exactly one line on stdout; everything else goes to stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "admm_iterations_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["product_library_mapped"] is False  # the CPU arm neither builds nor maps lib/libosqp.so
    assert d["whole_solve"]["status"] == "Solved" and d["whole_solve"]["iters"] >= 10


def test_reference_arm_batched_config():
    # N > 1 measures BASELINE config 5 (the sharded batch): the CPU arm solves a sample of it on all host threads
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--batch", "24"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert d["impl"] == "reference" and d["config"]["baseline_config"] == 5 and d["scaling"] == "strong"
    assert d["metric"] == "admm_iterations_per_sec" and d["value"] > 0 and d["product_library_mapped"] is False


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--n", "500", "--m", "800"], capture_output=True, text=True, timeout=300,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
