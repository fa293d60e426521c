# The round-end evidence set on one B200: GPU suite, smoke, bench + reference arm, ncu launch list, tensor-pipe kernel capture, per-config table.
mkdir -p gpurun_out
nvidia-smi -L | head -1; nproc
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r31.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r31.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r31.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r31.log
( time timeout 900 python bench.py > gpurun_out/bench_r31.json 2> gpurun_out/bench_r31.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_r31.err; cut -c1-200 gpurun_out/bench_r31.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r31.json 2> gpurun_out/bench_ref_r31.err; echo "ref exit $?"; cut -c1-160 gpurun_out/bench_ref_r31.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r31.csv python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu-baseline > gpurun_out/ncu_launches_r31.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:dense_corr_mma -s 1 -c 1 -f -o gpurun_out/corr_mma_final python scripts/prof_corr.py > gpurun_out/ncu_corr_mma_final.log 2>&1; tail -1 gpurun_out/ncu_corr_mma_final.log
rm -f gpurun_out/configs_r31.jsonl
timeout 600 python scripts/run_configs.py --c3-pairs 16 > gpurun_out/configs_r31.jsonl 2> gpurun_out/configs_r31.err; echo "configs exit $?"; wc -l gpurun_out/configs_r31.jsonl
