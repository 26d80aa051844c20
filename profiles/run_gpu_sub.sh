mkdir -p gpurun_out
(OSQP_B200_PAIRS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-extras --cpu-iters 5 > gpurun_out/r2_bench_under_ncu.log 2>&1); wc -l gpurun_out/r2_launches_final.csv
(OSQP_B200_PAIRS=0 timeout 900 ncu --set full --clock-control none -k regex:admm_kernel -c 1 -o gpurun_out/admm_full2 python profiles/profile_driver.py --solves 1 --spmv-reps 1 2>&1 | tail -2)
ls -la gpurun_out/
