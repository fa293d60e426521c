mkdir -p gpurun_out
P=unsynchronized_stereo_vision_proj325_b200
run() { for sel in "colour 32x32 ZNCC" "zncc gray 16x16"; do timeout 300 python scripts/run_configs.py --only "$sel" --c3-pairs 16 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$1', d['config'], round(d['pairs_per_s'],1), d['kernel'])"; done; }
run base
cp $P/libusv_b200.so /tmp/orig.so; cp scripts/dev/libusv_flat.so $P/libusv_b200.so
timeout 300 python scripts/stress_corr.py 30 51 mma | tail -1
run flat
cp /tmp/orig.so $P/libusv_b200.so
