mkdir -p gpurun_out
timeout 300 python scripts/quick_dense.py 64 > gpurun_out/quick_dense.log 2>&1; echo "quick exit $?"; grep -E "MISMATCH|C2|ALL|FAIL|Error" gpurun_out/quick_dense.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c.json')); print({k:d[k] for k in ('value','ms_per_step','cand_evals_per_s','parity_vs_oracle')}); print(d['e2e']['value'], d['roofline']['frac'])"
