mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_3.log; tail -15 gpurun_out/pytest_gpu_3.log
python scripts/prof_dense.py 16 > gpurun_out/prof_plain.log 2>&1 && cat gpurun_out/prof_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:dense_sad -s 2 -c 1 -f -o gpurun_out/dense_v1 python scripts/prof_dense.py 16 > gpurun_out/ncu_dense_v1.log 2>&1; tail -5 gpurun_out/ncu_dense_v1.log
