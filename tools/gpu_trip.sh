#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py tests/test_loader_auto_map.py tests/test_request_builders.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/tests.log
