mkdir -p gpurun_out
for LR in 1 0; do echo "== C3 (5000x25000) lane_rows $LR"; (OSQP_B200_LANE_ROWS=$LR timeout 400 python bench.py --config 3 --lasso 5000x25000x0.15 --steps 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['pcg_iters_per_admm_iter'], d['roofline']['frac'], d['solve'])"); done
for LR in 1 0; do echo "== C4 lane_rows $LR"; (OSQP_B200_LANE_ROWS=$LR timeout 400 python bench.py --config 4 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['pcg_iters_per_admm_iter'], d['solve'])"); done
(timeout 300 python -m pytest tests/test_configs.py tests/test_engine_parity.py -m gpu -q -x 2>&1 | tail -3)
