#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/attn_poly_sweep.py 2>&1 | tee gpurun_out/attn_poly_sweep.log
