mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_engine_parity.py tests/test_bench_parity.py -m gpu -q 2>&1 | tail -12)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
