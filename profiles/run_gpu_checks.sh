# GPU validation recipe (run under gpurun from the repo root): smoke, the GPU test-suite, the default bench line,
# per-phase device timing of the bench workload
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
(timeout 1800 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1); tail -30 gpurun_out/pytest_gpu.log
(timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err); tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
(OSQP_B200_DEBUG=1 timeout 200 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv" | tail -30)
