#!/bin/bash
# Working trip on one B200: full GPU suite, smoke and the default bench line at HEAD.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_full.txt
timeout 1200 python -m pytest tests -q -m gpu -x -s 2>&1 | tail -70 > gpurun_out/final_tests.log; tail -2 gpurun_out/final_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/final_smoke.log
timeout 700 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.log; cut -c1-200 gpurun_out/final_bench.log
