mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dense_sad -s 3 -c 1 -f -o gpurun_out/dense_final_bench256 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dense_final.log 2>&1; tail -1 gpurun_out/ncu_dense_final.log
timeout 900 python scripts/run_c4_sharded.py > gpurun_out/c4_full_n1_final.json 2> gpurun_out/c4_full_n1_final.err; echo "c4 exit $?"; cut -c1-400 gpurun_out/c4_full_n1_final.json
timeout 900 python scripts/run_c5_streams.py --slots 8 > gpurun_out/c5_gather_n1_final.json 2> gpurun_out/c5_gather_n1_final.err; echo "c5 exit $?"; cut -c1-700 gpurun_out/c5_gather_n1_final.json
