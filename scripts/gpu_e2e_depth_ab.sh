# bench.py at N GPUs with the e2e pipeline depth pinned to 1, 2 and calibrated (usage: gpu_e2e_depth_ab.sh N)
mkdir -p gpurun_out
N=$1
for d in 1 2 auto; do
  if [ "$d" = auto ]; then unset USV_BENCH_E2E_DEPTH; else export USV_BENCH_E2E_DEPTH=$d; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r37_n${N}_$d.json 2> gpurun_out/bench_r37_n${N}_$d.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_r37_n${N}_$d.json').read().strip().splitlines()[-1]); print('$d', round(d['value']), round(d['e2e']['value']), round(d['e2e']['compact_results']['value']), d['e2e']['api'][-55:-30], d['e2e']['calibration_ms_by_depth'])"
done
