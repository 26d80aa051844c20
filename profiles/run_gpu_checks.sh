mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -3 gpurun_out/pytest_gpu.log
(timeout 120 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv" | tail -4)
