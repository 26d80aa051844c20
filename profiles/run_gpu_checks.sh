mkdir -p gpurun_out
(timeout 300 python profiles/batch_bench.py 2>&1 | tail -3)
(timeout 600 python -m pytest tests/test_batch.py -m gpu -q -x 2>&1 | tail -2)
