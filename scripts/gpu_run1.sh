set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
./unsynchronized_stereo_vision_proj325_b200/usv_microbench > gpurun_out/microbench_r1.jsonl 2>&1; echo "microbench exit $?"
python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu_1.log; tail -30 gpurun_out/pytest_gpu_1.log
