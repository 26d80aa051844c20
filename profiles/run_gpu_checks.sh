mkdir -p gpurun_out
(timeout 900 python profiles/configs_full.py --lambdas 3 > gpurun_out/configs_full.log 2>&1); tail -9 gpurun_out/configs_full.log
(timeout 120 python profiles/profile_driver.py --solves 2 2>&1 | grep "^solve" | tail -1)
