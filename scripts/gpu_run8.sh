mkdir -p gpurun_out
timeout 300 python scripts/quick_dense.py 256 > gpurun_out/quick_dense.log 2>&1; echo "quick exit $?"; grep -E "MISMATCH|C2|ALL|FAIL|Error|got|first" gpurun_out/quick_dense.log | head -40
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
