# usage: bash scripts/gpu_prof.sh <tag>   -> launch list + full ncu capture of the dense kernel
mkdir -p gpurun_out
TAG=$1
python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
python scripts/prof_dense.py 64 > gpurun_out/prof_plain_$TAG.log 2>&1 && cat gpurun_out/prof_plain_$TAG.log && \
ncu --set full --clock-control none --import-source on -k regex:dense_ -s 2 -c 1 -f -o gpurun_out/dense_$TAG python scripts/prof_dense.py 64 > gpurun_out/ncu_dense_$TAG.log 2>&1; tail -3 gpurun_out/ncu_dense_$TAG.log
