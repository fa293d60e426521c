mkdir -p gpurun_out
timeout 600 python scripts/quick_dense.py 64 > gpurun_out/quick_dense.log 2>&1; echo "quick exit $?"; tail -40 gpurun_out/quick_dense.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu_2.log; tail -30 gpurun_out/pytest_gpu_2.log
