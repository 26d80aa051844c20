mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 --deselect tests/test_bench_parity.py::test_c2_bench_instance > gpurun_out/pytest_gpu.log 2>&1); tail -25 gpurun_out/pytest_gpu.log
(timeout 100 python profiles/batch_bench.py 2>&1 | tail -3)
(timeout 120 python profiles/latency_small.py 2>&1 | tail -7)
(OSQP_B200_TINY=0 timeout 120 python profiles/latency_small.py 2>&1 | tail -7)
(OSQP_B200_DEBUG=1 timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extras > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err); grep "osqp_b200\] setup" gpurun_out/bench_dbg.err | head; python -c "
import json; d=json.load(open('gpurun_out/bench_dbg.json')); print(d['value'], d['solve']['setup_s'], d['roofline']['frac'])"
