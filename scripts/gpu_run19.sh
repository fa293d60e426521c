mkdir -p gpurun_out
( timeout 600 python scripts/stress_corr.py 60 15 mma ) > gpurun_out/stress_mma5.log 2>&1; echo "stress mma exit $?"; tail -3 gpurun_out/stress_mma5.log
( timeout 300 python scripts/run_configs.py --only ZNCC --c3-pairs 16 ) > gpurun_out/configs_mma_v4.log 2>&1; echo "configs exit $?"; tail -1 gpurun_out/configs_mma_v4.log | cut -c1-420
( timeout 300 python scripts/run_configs.py --only ZNCC --c3-pairs 1 ) > gpurun_out/configs_mma_v4_1pair.log 2>&1; tail -1 gpurun_out/configs_mma_v4_1pair.log | cut -c240-420
( timeout 300 python scripts/run_configs.py --only SSD ) > gpurun_out/configs_mma_v4_ssd.log 2>&1; cut -c1-60,200-460 gpurun_out/configs_mma_v4_ssd.log
( USV_CORR_MMA=0 timeout 300 python scripts/run_configs.py --only SSD ) > gpurun_out/configs_alu_ssd.log 2>&1; cut -c1-60,200-460 gpurun_out/configs_alu_ssd.log
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r19.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r19.log
