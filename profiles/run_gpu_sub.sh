mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_configs.py -m gpu -q -x -k "follows" --durations=3 2>&1 | tail -25)
