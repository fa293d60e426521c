mkdir -p gpurun_out
timeout 300 python scripts/stress_dense.py 40 151 | tail -1
timeout 600 python bench.py --no-cpu-baseline --steps 6 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_vs_oracle'])"
timeout 300 python scripts/run_configs.py --only "C4" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['pairs_per_s'],1))"
timeout 300 python scripts/run_configs.py --only "C2 D=128" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['pairs_per_s'],1))"
