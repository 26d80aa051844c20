mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_batch.py tests/test_configs.py -m gpu -q > gpurun_out/pytest_new.log 2>&1); tail -4 gpurun_out/pytest_new.log
(timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err); python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print(d['value'], d['batch'])"; tail -3 gpurun_out/bench_b.err
