mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_engine_parity.py -m gpu -q -x -k "fresh_setup_at_stream or settings_variants" > gpurun_out/pytest_new.log 2>&1); tail -15 gpurun_out/pytest_new.log
