mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r9.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r9.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r9.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_r9.log
timeout 900 python bench.py > gpurun_out/bench_r9.json 2> gpurun_out/bench_r9.err; echo "bench exit $?"; tail -2 gpurun_out/bench_r9.err; cut -c1-400 gpurun_out/bench_r9.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r9.json 2> gpurun_out/bench_ref_r9.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref_r9.json
timeout 900 python scripts/run_c4_sharded.py > gpurun_out/c4_full_n1.json 2> gpurun_out/c4_full_n1.err; echo "c4 exit $?"; cat gpurun_out/c4_full_n1.json; tail -3 gpurun_out/c4_full_n1.err
