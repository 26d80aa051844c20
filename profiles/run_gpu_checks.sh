mkdir -p gpurun_out
(timeout 600 python profiles/polish_check.py 2>&1 | tail -9)
