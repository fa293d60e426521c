# The round-end evidence set on one B200: GPU suite, smoke, bench + reference arm, ncu launch list, ncu captures of the
# headline kernel and the row-resolve kernel, per-config table.   usage: bash scripts/gpu_final_evidence.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -L | head -1; nproc
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_$TAG.log
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_$TAG.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; cut -c1-160 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:dense_sad -s 2 -c 1 -f -o gpurun_out/dense_sad_$TAG python scripts/prof_dense.py 256 > gpurun_out/ncu_dense_sad_$TAG.log 2>&1; tail -1 gpurun_out/ncu_dense_sad_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:dense_resolve_rows -s 1 -c 1 -f -o gpurun_out/resolve_rows_$TAG python scripts/time_resolve_rows.py > gpurun_out/ncu_resolve_rows_$TAG.log 2>&1; tail -1 gpurun_out/ncu_resolve_rows_$TAG.log
rm -f gpurun_out/configs_$TAG.jsonl
timeout 600 python scripts/run_configs.py --c3-pairs 16 > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "configs exit $?"; wc -l gpurun_out/configs_$TAG.jsonl
timeout 200 python scripts/time_resolve_rows.py > gpurun_out/resolve_rows_timing_$TAG.jsonl 2>&1; tail -2 gpurun_out/resolve_rows_timing_$TAG.jsonl
