mkdir -p gpurun_out
timeout 900 python scripts/run_c5_streams.py > gpurun_out/c5_n1.json 2> gpurun_out/c5_n1.err; echo "c5 exit $?"; cat gpurun_out/c5_n1.json; tail -3 gpurun_out/c5_n1.err
python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/plain_traffic.log 2>&1 && \
ncu --set full --clock-control none -k regex:dense_ -s 2 -c 1 -f -o gpurun_out/dense_bench256 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_bench256.log 2>&1; tail -2 gpurun_out/ncu_bench256.log
