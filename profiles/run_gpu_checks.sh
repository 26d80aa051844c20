mkdir -p gpurun_out
(timeout 300 python profiles/latency_small.py 2>&1 | tail -8)
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -3 gpurun_out/pytest_gpu.log
