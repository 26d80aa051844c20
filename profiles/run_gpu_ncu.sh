mkdir -p gpurun_out
(OSQP_B200_PAIRS=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:admm_kernel -c 1 -o gpurun_out/admm_prof python profiles/profile_driver.py --solves 1 --max-iter 30 --spmv-reps 1 --lib osqp.jl_b200/lib/variants/libosqp_w16d4i1.so 2>&1 | tail -12)
