mkdir -p gpurun_out
(OSQP_B200_DEBUG=1 timeout 900 python -m pytest tests/test_engine_parity.py -m gpu -q -x -k "kkt or agree or unconstrained" > gpurun_out/pytest_scale.log 2>&1); grep -E "osqp_b200\] tile|passed|failed|Error|error" gpurun_out/pytest_scale.log | tail -20
