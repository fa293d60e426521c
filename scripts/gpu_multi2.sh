# usage: gpu_multi2.sh N [bench|all]
mkdir -p gpurun_out
N=$1; WHAT=${2:-all}
nvidia-smi -L | wc -l; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r1_n$N.json 2> gpurun_out/bench_r1_n$N.err; echo "bench N=$N exit $?"; tail -2 gpurun_out/bench_r1_n$N.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r1_n$N.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','cand_evals_per_s','parity_vs_oracle','gpu_launches')}); print('e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])"
if [ "$WHAT" = "all" ]; then
timeout 900 $TR --master-port 29513 scripts/run_c4_sharded.py > gpurun_out/c4_full_n$N.json 2> gpurun_out/c4_full_n$N.err; echo "c4 N=$N exit $?"; tail -1 gpurun_out/c4_full_n$N.json; tail -2 gpurun_out/c4_full_n$N.err
timeout 900 $TR --master-port 29514 scripts/run_c5_streams.py --slots 8 > gpurun_out/c5_gather_n$N.json 2> gpurun_out/c5_gather_n$N.err; echo "c5 N=$N exit $?"; tail -1 gpurun_out/c5_gather_n$N.json; tail -2 gpurun_out/c5_gather_n$N.err
fi
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_r1_n$N.json 2> gpurun_out/bench_ref_r1_n$N.err; echo "ref N=$N exit $?"; cut -c1-200 gpurun_out/bench_ref_r1_n$N.json
