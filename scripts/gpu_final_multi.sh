# Round-end bench on N GPUs of one box, launched the way the driver does it.   usage: bash scripts/gpu_final_multi.sh N
mkdir -p gpurun_out
N=$1
nvidia-smi -L | wc -l; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err ) 2>&1 | grep real; echo "bench N=$N exit $?"; tail -2 gpurun_out/r2_final_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_final_bench_n$N.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','parity_vs_oracle','gpu_launches')}); print('e2e', d['e2e']['value'], 'cpp', d['e2e'].get('cpp_dropin'), 'frac', d['roofline']['frac'], d['clocks']); print({k:(v.get('value'), v.get('e2e_value', None), v.get('parity_vs_oracle')) for k,v in d['configs'].items()})"
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref_n$N.json 2> gpurun_out/r2_final_bench_ref_n$N.err; echo "ref N=$N exit $?"; cut -c1-200 gpurun_out/r2_final_bench_ref_n$N.json
