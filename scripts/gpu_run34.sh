mkdir -p gpurun_out
timeout 300 python scripts/stress_dense.py 90 161 | tail -1
timeout 300 python scripts/stress_corr.py 60 162 mma | tail -1
timeout 600 python bench.py --no-cpu-baseline --steps 6 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity_vs_oracle'])"
for sel in "C4" "colour 32x32 ZNCC"; do timeout 300 python scripts/run_configs.py --only "$sel" --c3-pairs 16 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['pairs_per_s'],1))"; done
ncu --set full --clock-control none --import-source on -k regex:dense_sad -s 3 -c 1 -f -o gpurun_out/dense_final_bench256 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dense_final.log 2>&1; tail -1 gpurun_out/ncu_dense_final.log
