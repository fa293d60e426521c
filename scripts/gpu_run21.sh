mkdir -p gpurun_out
( timeout 900 python scripts/stress_corr.py 120 31 mma ) > gpurun_out/stress_mma7.log 2>&1; echo "stress mma exit $?"; tail -8 gpurun_out/stress_mma7.log
( timeout 900 python scripts/stress_corr.py 60 32 ) > gpurun_out/stress_corr8.log 2>&1; echo "stress corr exit $?"; tail -3 gpurun_out/stress_corr8.log
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r21.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r21.log
( timeout 300 python scripts/run_configs.py --only ZNCC --c3-pairs 16 ) > gpurun_out/configs_mma_v6.log 2>&1; tail -1 gpurun_out/configs_mma_v6.log | cut -c240-420
