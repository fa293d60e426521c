# Parity stress + C3 / gray timings of the opt-in tcgen05 correlation kernel (USV_CORR_UMMA=1)
USV_CORR_UMMA=1 timeout 400 python scripts/stress_corr.py 120 306 mma 2>&1 | tail -1
for sel in "colour 32x32 ZNCC" "zncc gray 16x16"; do USV_CORR_UMMA=1 timeout 200 python scripts/run_configs.py --only "$sel" --c3-pairs 16 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['pairs_per_s'],1), d['kernel'])"; done
USV_CORR_UMMA=1 timeout 400 python scripts/stress_corr.py 200 401 mma | tail -1
USV_CORR_UMMA=1 timeout 200 python scripts/run_configs.py --only "ncc gray" 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['pairs_per_s'],1), d['kernel'])"
