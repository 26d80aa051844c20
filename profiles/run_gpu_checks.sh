mkdir -p gpurun_out
(OSQP_B200_DEBUG=1 timeout 120 python profiles/profile_driver.py --solves 1 --spmv-reps 1 2>&1 | grep "osqp_b200\]" )
