#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_e2e.json
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q --no-header -rA 2>&1 | tail -150 > gpurun_out/c2_t_e2e.log; echo "e2e rc=${PIPESTATUS[0]}"
tail -n 15 gpurun_out/c2_t_e2e.log
cat gpurun_out/parity_e2e.json
