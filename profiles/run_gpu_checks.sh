mkdir -p gpurun_out
(OSQP_B200_DEBUG=1 timeout 120 python profiles/profile_driver.py --solves 2 --spmv-reps 1 2>&1 | grep -E "setup|^solve" )
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -2 gpurun_out/pytest_gpu.log
