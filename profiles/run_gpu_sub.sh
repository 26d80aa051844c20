mkdir -p gpurun_out
for rep in 1 2; do
echo "== head"; (timeout 200 python profiles/profile_driver.py --solves 2 --lib osqp.jl_b200/lib/variants/libosqp_head.so 2>&1 | grep -v "^spmv\|stream build" | tail -3)
echo "== new"; (timeout 200 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv\|stream build" | tail -3)
done
(timeout 900 python -m pytest tests/test_engine_parity.py tests/test_bench_parity.py -m gpu -q 2>&1 | tail -5)
