#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention" 2>&1 | tail -4 > gpurun_out/tests.log; cat gpurun_out/tests.log
timeout 300 python tools/attn_poly_sweep.py 2>&1 | tee gpurun_out/attn_poly_sweep.log
