#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_router_trunk.py -q --no-header -rA 2>&1 | tail -60 > gpurun_out/c6_t_trunk.log; echo "trunk rc=${PIPESTATUS[0]}"
grep -E "PASSED|FAILED|passed|failed|Error|error" gpurun_out/c6_t_trunk.log | head -30
timeout 900 python -m pytest tests/test_gpu_gconv.py -q -k "gwgrad" 2>&1 | tail -3
timeout 600 python tools/prof_step.py > gpurun_out/c6_prof_step.txt 2>&1; echo "prof rc=$?"
