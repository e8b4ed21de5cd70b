#!/bin/bash
# Final numbers after the GEMM-epilogue / prepare-overlap changes: full GPU suite, smoke, full bench line, GEMM full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_full.txt
timeout 1200 python -m pytest tests -q -m gpu -x -s 2>&1 | tail -70 > gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 700 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.log; cut -c1-260 gpurun_out/final_bench.log
timeout 600 python bench.py --masks 4 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/final_bench_config3.log; cut -c1-200 gpurun_out/final_bench_config3.log
timeout 900 python bench.py --model gar-8b --batch 2 --new-tokens 256 --max-num-tiles 16 --steps 2 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/final_bench_config4.log; cut -c1-200 gpurun_out/final_bench_config4.log
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 8 --no-e2e --no-cpu-baseline --no-lib-baseline --no-profile"
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16_tn -c 400 --csv --log-file gpurun_out/gemm_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 12 -c 3 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
ls -la gpurun_out | tail -8
