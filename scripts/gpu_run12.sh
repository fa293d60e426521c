mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python scripts/run_c5_streams.py > gpurun_out/c5_gather_n1.json 2> gpurun_out/c5_gather_n1.err; echo "c5 gather exit $?"; cat gpurun_out/c5_gather_n1.json; tail -3 gpurun_out/c5_gather_n1.err
timeout 900 python scripts/run_c5_streams.py --staged > gpurun_out/c5_staged_n1.json 2> gpurun_out/c5_staged_n1.err; echo "c5 staged exit $?"; cat gpurun_out/c5_staged_n1.json
timeout 900 python scripts/run_c5_streams.py --slots 8 > gpurun_out/c5_gather8_n1.json 2> gpurun_out/c5_gather8_n1.err; echo "c5 gather8 exit $?"; cat gpurun_out/c5_gather8_n1.json
