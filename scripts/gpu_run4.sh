mkdir -p gpurun_out
timeout 600 python scripts/quick_dense.py 64 > gpurun_out/quick_dense.log 2>&1; echo "quick exit $?"; tail -12 gpurun_out/quick_dense.log
