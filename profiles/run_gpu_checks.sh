mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -3 gpurun_out/pytest_gpu.log
(timeout 900 python profiles/configs_full.py > gpurun_out/configs_full.log 2>&1); tail -17 gpurun_out/configs_full.log
