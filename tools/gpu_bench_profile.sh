#!/bin/bash
# GPU-box script: full bench line, then the ncu launch list and one full capture of the top kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench (gar-1b, config 2)" | tee gpurun_out/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -25 | tee -a gpurun_out/bench.log
echo "== ncu launch list" | tee -a gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --batch 1 --new-tokens 8 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
echo "== ncu full capture of the GEMM" | tee -a gpurun_out/bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 30 -c 3 -f -o gpurun_out/prof_gemm \
  python bench.py --steps 1 --warmup 0 --batch 1 --new-tokens 4 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
