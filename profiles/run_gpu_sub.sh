mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_configs.py -m gpu -q --durations=4 2>&1 | tail -15)
(timeout 900 python bench.py --config 3 --steps 1 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err); tail -c 1300 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
