mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_configs.py -m gpu -q 2>&1 | tail -12)
(timeout 600 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err); tail -c 1800 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
