mkdir -p gpurun_out
for m in 1 0; do
  for sel in "gray 16x16" "colour 32x32 ZNCC" "colour 32x32 NCC" "colour 32x32 SSD"; do
    USV_CORR_MMA=$m timeout 300 python scripts/run_configs.py --only "$sel" --c3-pairs 16 2>> gpurun_out/configs_corr_ab.err | sed "s/^{/{\"USV_CORR_MMA\": $m, /" >> gpurun_out/configs_corr_ab.jsonl
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/configs_corr_ab.jsonl'):
    d=json.loads(l); print(d['USV_CORR_MMA'], d['config'], d['pairs'], round(d['pairs_per_s'],1), round(d['cand_evals_per_s']/1e12,3), d['kernel'])
PY
