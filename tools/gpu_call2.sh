#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_e2e.json
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q --no-header -rA 2>&1 | tail -60 > gpurun_out/c2_t_e2e.log; echo "e2e rc=${PIPESTATUS[0]}"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_e2e.py > gpurun_out/c2_t_all.log 2>&1; echo "suite rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c2_bench.log 2> gpurun_out/c2_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c2_bench_ref.log 2>&1; echo "ref rc=$?"
tail -n 5 gpurun_out/c2_t_e2e.log; tail -n 3 gpurun_out/c2_t_all.log
cat gpurun_out/parity_e2e.json
