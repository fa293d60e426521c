mkdir -p gpurun_out
( timeout 600 python scripts/stress_corr.py 60 16 mma ) > gpurun_out/stress_mma6.log 2>&1; echo "stress mma exit $?"; tail -3 gpurun_out/stress_mma6.log
( timeout 300 python scripts/run_configs.py --only ZNCC --c3-pairs 16 ) > gpurun_out/configs_mma_v5.log 2>&1; echo "configs exit $?"; tail -1 gpurun_out/configs_mma_v5.log | cut -c240-420
ncu --set full --clock-control none --import-source on -k regex:dense_corr_mma -s 1 -c 1 -f -o gpurun_out/corr_mma_v5 python scripts/prof_corr.py > gpurun_out/ncu_corr_mma_v5.log 2>&1; tail -1 gpurun_out/ncu_corr_mma_v5.log
