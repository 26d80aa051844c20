mkdir -p gpurun_out
for pf in 4 12; do echo "PF=$pf"; (OSQP_B200_PF=$pf timeout 600 python profiles/configs_full.py --lambdas 2 2>&1 | grep -E "per PCG|lasso sweep|Solved iter=25"); done
