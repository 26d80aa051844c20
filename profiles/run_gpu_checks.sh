mkdir -p gpurun_out
(OSQP_B200_DEBUG=1 timeout 120 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv 1[0-2]" | tail -10)
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -3 gpurun_out/pytest_gpu.log
