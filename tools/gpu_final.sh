#!/bin/bash
# Round-end verification on one B200: full GPU test suite (incl. the full-size parity gate), smoke(), the default bench
# line (with cpu_baseline and gpu_lib_baseline), the reference arm, and the driver-runnable lines of BASELINE configs 3/4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_full.txt
timeout 1200 python -m pytest tests -q -m gpu -x -s 2>&1 | tail -70 > gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 700 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.log; cut -c1-260 gpurun_out/final_bench.log
if [ -n "$GAR_FINAL_REFERENCE" ]; then timeout 500 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/final_bench_reference.log; cut -c1-300 gpurun_out/final_bench_reference.log; fi
timeout 600 python bench.py --masks 4 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/final_bench_config3.log; cut -c1-200 gpurun_out/final_bench_config3.log
timeout 900 python bench.py --model gar-8b --batch 2 --new-tokens 256 --max-num-tiles 16 --steps 2 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/final_bench_config4.log; cut -c1-200 gpurun_out/final_bench_config4.log
