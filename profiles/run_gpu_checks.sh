# GPU validation recipe of round 1 (run under gpurun from the repo root)
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -2 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err); python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['spmv_in_loop'], d['batch']['qp_iterations_per_sec'], d['solve']['setup_s'])"
(timeout 120 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv" | tail -4)
