# usage: bash scripts/gpu_bench.sh <tag>  -> smoke, bench (ours), bench (reference arm)
mkdir -p gpurun_out
TAG=$1
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; cat gpurun_out/bench_ref_$TAG.json
