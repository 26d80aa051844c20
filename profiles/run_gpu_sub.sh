mkdir -p gpurun_out
for F in 1 0 1 0; do echo "== f32 slices $F"; (OSQP_B200_F32_SLICES=$F timeout 120 python profiles/profile_driver.py --solves 2 2>&1 | grep -v "^spmv" | tail -3); done
