#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -m gpu -x -k "gemm or stage_parity" 2>&1 | tail -2 > gpurun_out/tests.log
timeout 300 python tools/gemm_epi_bench.py 2>&1 | grep -E "GELU|RoPE|LayerScale" > gpurun_out/gemm_epi_bench_short.log
cat gpurun_out/tests.log gpurun_out/gemm_epi_bench_short.log
date
