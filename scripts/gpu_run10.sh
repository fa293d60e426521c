mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class) local_cpus=$(cat $d/local_cpulist)"; fi; done; cat /sys/devices/system/node/node*/cpulist; cat /proc/self/status | grep -i allowed; numactl -H ) > gpurun_out/topo_n1.txt 2>&1
timeout 900 python scripts/stress_corr.py 150 5 > gpurun_out/stress_corr_ssd.log 2>&1; echo "stress exit $?"; tail -5 gpurun_out/stress_corr_ssd.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python scripts/run_configs.py --only SSD > gpurun_out/configs_ssd.jsonl 2> gpurun_out/configs_ssd.err; python scripts/run_configs.py --only ssd >> gpurun_out/configs_ssd.jsonl 2>> gpurun_out/configs_ssd.err; cat gpurun_out/configs_ssd.jsonl; tail -3 gpurun_out/configs_ssd.err
