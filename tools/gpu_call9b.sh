#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -q --no-header -x 2>&1 | tail -45
timeout 600 python -m pytest tests/test_gpu_glue.py -q --no-header 2>&1 | tail -15
