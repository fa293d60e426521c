mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r13.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r13.log
