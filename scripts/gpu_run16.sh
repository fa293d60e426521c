mkdir -p gpurun_out
( timeout 600 python scripts/stress_corr.py 80 12 mma ) > gpurun_out/stress_mma2.log 2>&1; echo "stress mma exit $?"; tail -6 gpurun_out/stress_mma2.log
( USV_MMA_CTAS=3 timeout 600 python scripts/stress_corr.py 30 13 mma ) > gpurun_out/stress_mma3.log 2>&1; echo "stress mma(3 ctas) exit $?"; tail -3 gpurun_out/stress_mma3.log
for c in 2 3; do
( USV_MMA_CTAS=$c timeout 300 python scripts/run_configs.py --only ZNCC --c3-pairs 16 ) > gpurun_out/configs_mma_c$c.log 2>&1; echo "configs ctas=$c exit $?"; tail -1 gpurun_out/configs_mma_c$c.log | cut -c1-420
done
