mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print({k:d[k] for k in ('value','ms_per_step','cand_evals_per_s','parity_vs_oracle','clocks')}); print(d['e2e']); print(d['roofline'])"
