mkdir -p gpurun_out
for v in 1 0 1; do echo "INFO_STREAMS=$v"; (OSQP_B200_INFO_STREAMS=$v timeout 120 python profiles/profile_driver.py --solves 3 2>&1 | grep -A3 "^solve" | tail -4 | grep -E "^solve|ADMM"); done
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -2 gpurun_out/pytest_gpu.log
