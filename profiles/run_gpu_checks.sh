mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4)
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -3 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err); cat gpurun_out/bench.json | cut -c1-3000
