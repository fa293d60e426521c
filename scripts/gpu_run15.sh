mkdir -p gpurun_out
python scripts/prof_corr.py > gpurun_out/prof_corr_plain.log 2>&1 && cat gpurun_out/prof_corr_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:dense_corr_mma -s 1 -c 1 -f -o gpurun_out/corr_mma_v1 python scripts/prof_corr.py > gpurun_out/ncu_corr_mma_v1.log 2>&1; tail -3 gpurun_out/ncu_corr_mma_v1.log
