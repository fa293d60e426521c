mkdir -p gpurun_out
timeout 900 python scripts/stress_corr.py 200 6 > gpurun_out/stress_corr_sad.log 2>&1; echo "stress exit $?"; tail -5 gpurun_out/stress_corr_sad.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python scripts/run_configs.py --only SAD-colour > gpurun_out/configs_sadc.jsonl 2> gpurun_out/configs_sadc.err; cat gpurun_out/configs_sadc.jsonl; tail -3 gpurun_out/configs_sadc.err
# compute-sanitizer over small random configurations of every sliding kernel (memcheck, racecheck, initcheck, synccheck)
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/stress_dense.py 6 21 > gpurun_out/sanitizer_${tool}_dense.log 2>&1; echo "$tool dense exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|cases" gpurun_out/sanitizer_${tool}_dense.log | tail -2
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/stress_corr.py 6 22 > gpurun_out/sanitizer_${tool}_corr.log 2>&1; echo "$tool corr exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|cases" gpurun_out/sanitizer_${tool}_corr.log | tail -2
done
