# Randomised parity soak against the CPU oracle on one B200 (bit-exact or it prints MISMATCH).   usage: bash scripts/gpu_soak.sh [scale]
S=${1:-1}
mkdir -p gpurun_out
( time timeout 1500 python scripts/stress_dense.py $((1000 * S)) 9101 ) > gpurun_out/soak_dense.log 2>&1; echo "dense exit $?"; tail -5 gpurun_out/soak_dense.log
( time timeout 1200 python scripts/stress_corr.py $((600 * S)) 9102 mma ) > gpurun_out/soak_corr_mma.log 2>&1; echo "corr mma exit $?"; tail -5 gpurun_out/soak_corr_mma.log
( time timeout 900 python scripts/stress_corr.py $((300 * S)) 9103 ) > gpurun_out/soak_corr.log 2>&1; echo "corr exit $?"; tail -5 gpurun_out/soak_corr.log
( time timeout 900 python scripts/stress_corr.py $((200 * S)) 9105 mma tcgen05 ) > gpurun_out/soak_corr_tcgen05.log 2>&1; echo "corr tcgen05 exit $?"; tail -5 gpurun_out/soak_corr_tcgen05.log
( time timeout 600 python scripts/stress_corr.py $((150 * S)) 9104 all alu ) > gpurun_out/soak_corr_alu.log 2>&1; echo "corr alu exit $?"; tail -5 gpurun_out/soak_corr_alu.log
