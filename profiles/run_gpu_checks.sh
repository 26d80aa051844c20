mkdir -p gpurun_out
(timeout 900 python profiles/configs_full.py > gpurun_out/configs_full.log 2>&1); tail -22 gpurun_out/configs_full.log
