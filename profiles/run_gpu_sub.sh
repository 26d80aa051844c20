mkdir -p gpurun_out
for v in w16d4i1 w16d4i2 w32d2i1 w32d2i2 w32d4i2; do echo "== variant $v"; (timeout 120 python profiles/profile_driver.py --solves 2 --lib osqp.jl_b200/lib/variants/libosqp_$v.so 2>&1 | grep -v "^spmv" | tail -3); done
(timeout 100 python profiles/batch_bench.py 2>&1 | tail -3)
(timeout 300 python -m pytest tests/test_batch.py tests/test_bench_parity.py::test_c5_bench_batch_subset -m gpu -q -x 2>&1 | tail -3)
for B in 64 128 512; do echo "block $B"; (OSQP_B200_BLOCK=$B timeout 120 python profiles/latency_small.py 2>&1 | tail -7); done
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:admm_kernel -c 1 -o gpurun_out/admm_prof python profiles/profile_driver.py --solves 1 --max-iter 30 --spmv-reps 1 --lib osqp.jl_b200/lib/variants/libosqp_w16d4i1.so 2>&1 | tail -3)
