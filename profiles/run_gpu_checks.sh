mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -3 gpurun_out/pytest_gpu.log
(timeout 300 python profiles/probe_spmv.py 2>&1 | grep -E "which|mean|stream duration")
(timeout 300 python profiles/profile_driver.py --solves 2 2>&1 | tail -10)
(timeout 300 python profiles/membench.py 2>&1 | tail -6)
