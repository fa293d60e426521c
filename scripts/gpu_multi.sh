mkdir -p gpurun_out
N=$1
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_pipeline.py -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/bench_n$N.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','cand_evals_per_s','parity_vs_oracle','gpu_launches')}); print(d['e2e']['value'], d['roofline']['frac'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N exit $?"; cut -c1-300 gpurun_out/bench_ref_n$N.json
