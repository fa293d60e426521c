mkdir -p gpurun_out
timeout 900 python scripts/stress_dense.py 1500 9001 > gpurun_out/soak_dense.log 2>&1; echo "dense exit $?"; tail -2 gpurun_out/soak_dense.log
timeout 1200 python scripts/stress_corr.py 1200 9002 mma > gpurun_out/soak_corr_mma.log 2>&1; echo "corr mma exit $?"; tail -2 gpurun_out/soak_corr_mma.log
timeout 900 python scripts/stress_corr.py 500 9003 > gpurun_out/soak_corr.log 2>&1; echo "corr exit $?"; tail -2 gpurun_out/soak_corr.log
USV_CORR_MMA=0 timeout 900 python scripts/stress_corr.py 300 9004 > gpurun_out/soak_corr_alu.log 2>&1; echo "corr alu exit $?"; tail -1 gpurun_out/soak_corr_alu.log
