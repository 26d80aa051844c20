# Round-2 measurement pass (one B200): bench line, reference arm, ncu launch list of the bench command, full ncu capture
# of one whole admm_kernel launch (DRAM traffic), per-SpMV ncu rows.
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err); tail -c 400 gpurun_out/r2_bench.json
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err); cut -c1-300 gpurun_out/r2_bench_reference.json
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1); wc -l gpurun_out/r2_launches.csv
(OSQP_B200_PAIRS=0 timeout 600 ncu --set full --clock-control none -k regex:admm_kernel -c 1 -o gpurun_out/admm_full python profiles/profile_driver.py --solves 1 --spmv-reps 1 2>&1 | tail -2)
(timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:spmv_stream_kernel --csv --log-file gpurun_out/r2_spmv_rows.csv python profiles/profile_driver.py --solves 0 --spmv-reps 3 2>&1 | tail -2); wc -l gpurun_out/r2_spmv_rows.csv
