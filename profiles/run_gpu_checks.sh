mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1); tail -5 gpurun_out/pytest_gpu.log
(timeout 300 python profiles/probe_spmv.py 2>&1 | tail -14)
(timeout 300 python profiles/profile_driver.py --solves 2 2>&1 | tail -8)
for pf in 0 16; do echo "PF=$pf"; (OSQP_B200_PF=$pf timeout 300 python profiles/profile_driver.py --solves 2 2>&1 | tail -8 | head -5); done
echo "GROUPS=4"; (OSQP_B200_GROUPS=4 timeout 300 python profiles/profile_driver.py --solves 2 2>&1 | tail -8 | head -5)
