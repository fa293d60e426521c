mkdir -p gpurun_out
timeout 900 python scripts/run_c4_sharded.py > gpurun_out/c4_full_n1_final.json 2> gpurun_out/c4_full_n1_final.err; echo "c4 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/c4_full_n1_final.json').read().strip().splitlines()[-1]); print(d['pairs_per_s'], d['cand_evals_per_s'])"
timeout 900 python scripts/run_c5_streams.py --slots 8 > gpurun_out/c5_gather_n1_final.json 2> gpurun_out/c5_gather_n1_final.err; echo "c5 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/c5_gather_n1_final.json').read().strip().splitlines()[-1]); print(d['e2e_pairs_per_s'], d['pairing_pairs_per_s'])"
timeout 600 python scripts/run_configs.py --c3-pairs 16 > gpurun_out/configs_r35.jsonl 2> gpurun_out/configs_r35.err; echo "configs exit $?"
python -c "
import json
for l in open('gpurun_out/configs_r35.jsonl'):
    c=json.loads(l); print(c['config'], c.get('pairs'), round(c.get('pairs_per_s',0),1), round(c.get('cand_evals_per_s',0)/1e12,3), c.get('kernel'), c.get('us_per_launch'))"
