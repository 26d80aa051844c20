mkdir -p gpurun_out
(timeout 100 python profiles/batch_driver.py 1024 250 2>&1 | tail -1)
(timeout 100 python profiles/batch_driver.py 148 250 2>&1 | tail -1)
(timeout 300 ncu --set full --import-source on --clock-control none -k regex:batch_fast_solve -c 1 -o gpurun_out/batch_lat python profiles/batch_driver.py 1024 250 2>&1 | tail -2)
