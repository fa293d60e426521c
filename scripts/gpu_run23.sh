mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/stress_corr.py 12 41 mma ) > gpurun_out/sanitizer_mem.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|cases" gpurun_out/sanitizer_mem.log | head -8
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/stress_corr.py 6 42 mma ) > gpurun_out/sanitizer_race.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|hazard|cases" gpurun_out/sanitizer_race.log | head -8
