mkdir -p gpurun_out
(OSQP_B200_PAIRS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1); tail -c 200 gpurun_out/ncu_bench.log
(OSQP_B200_PAIRS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"admm_kernel" -s 1 -c 1 -o gpurun_out/r1h_admm python profiles/profile_driver.py --solves 2 --spmv-reps 1 > gpurun_out/ncu_admm.log 2>&1); tail -2 gpurun_out/ncu_admm.log
