mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -2 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err); cat gpurun_out/bench.json | cut -c1-2600
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err); cut -c1-200 gpurun_out/bench_ref.json
(timeout 120 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv" | tail -4)
