#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -m gpu -x -k "gemm or stage_parity" 2>&1 | tail -3 | tee gpurun_out/tests.log
timeout 300 python tools/gemm_epi_bench.py 2>&1 | tee gpurun_out/gemm_epi_bench.log
