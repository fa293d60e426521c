mkdir -p gpurun_out
timeout 300 python scripts/quick_dense.py 64 > gpurun_out/quick_dense.log 2>&1; echo "quick exit $?"; tail -14 gpurun_out/quick_dense.log
