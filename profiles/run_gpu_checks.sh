mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_engine_parity.py -m gpu -q -x -k "infeasib" > gpurun_out/pytest_inf.log 2>&1); tail -15 gpurun_out/pytest_inf.log
