mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r25.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r25.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r25.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_r25.log
( time timeout 900 python bench.py > gpurun_out/bench_r25.json 2> gpurun_out/bench_r25.err ) 2>&1 | grep real; echo "bench exit $?"; tail -2 gpurun_out/bench_r25.err; cut -c1-300 gpurun_out/bench_r25.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r25.json 2> gpurun_out/bench_ref_r25.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref_r25.json
