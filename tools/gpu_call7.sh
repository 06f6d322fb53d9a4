#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_router_trunk.py -q --no-header -rA 2>&1 | tail -60 > gpurun_out/c7_t_trunk.log; echo "trunk rc=${PIPESTATUS[0]}"
grep -E "PASSED|FAILED|passed|failed|Error|error" gpurun_out/c7_t_trunk.log | head -30
