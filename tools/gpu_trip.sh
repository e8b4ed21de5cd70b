#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -m gpu -x -k "attention or stage_parity or batched" 2>&1 | tail -3 > gpurun_out/tests.log; cat gpurun_out/tests.log
timeout 300 python tools/prof_kernels.py attn_timeline 2>&1 | tail -24 | tee gpurun_out/attn_timeline.log
