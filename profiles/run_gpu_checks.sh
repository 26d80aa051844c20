mkdir -p gpurun_out
(timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_driver.py 1 > gpurun_out/racecheck.log 2>&1); echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
(timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python profiles/sanitize_driver.py 1 > gpurun_out/synccheck.log 2>&1); echo "synccheck rc=$?"; tail -3 gpurun_out/synccheck.log
(timeout 120 python profiles/profile_driver.py --solves 3 2>&1 | grep "^solve" | tail -2)
