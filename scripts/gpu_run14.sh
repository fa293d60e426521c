mkdir -p gpurun_out
( timeout 600 python scripts/stress_corr.py 60 11 mma ) > gpurun_out/stress_mma.log 2>&1; echo "stress mma exit $?"; tail -12 gpurun_out/stress_mma.log
( timeout 300 python scripts/run_configs.py --only ZNCC --c3-pairs 8 ) > gpurun_out/configs_mma.log 2>&1; echo "configs exit $?"; tail -3 gpurun_out/configs_mma.log
( USV_CORR_MMA=0 timeout 300 python scripts/run_configs.py --only ZNCC --c3-pairs 8 ) > gpurun_out/configs_alu.log 2>&1; echo "configs(alu) exit $?"; tail -3 gpurun_out/configs_alu.log
