mkdir -p gpurun_out
python scripts/prof_corr.py > gpurun_out/prof_corr_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/corr_mma_launches.csv python scripts/prof_corr.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_corr_mma -s 1 -c 1 -f -o gpurun_out/corr_mma_v2 python scripts/prof_corr.py > gpurun_out/ncu_corr_mma_v2.log 2>&1; tail -2 gpurun_out/ncu_corr_mma_v2.log
