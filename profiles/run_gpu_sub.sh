mkdir -p gpurun_out
(OSQP_B200_F32_SLICES=0 timeout 200 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv" | tail -4)
(timeout 100 python profiles/batch_bench.py 2>&1 | tail -3)
(timeout 300 ncu --set full --import-source on --clock-control none -k regex:batch_fast_solve -c 1 -o gpurun_out/batch_prof python profiles/batch_driver.py 8192 100 2>&1 | tail -3)
