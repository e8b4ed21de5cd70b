#!/bin/bash
# Round-end verification on one B200: full GPU test suite, smoke(), the default bench line, the reference arm, the
# assemble/replay kernels under ncu, and the GAR-8B secondary configuration.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/final_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 500 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.log; cut -c1-260 gpurun_out/final_bench.log
grep -o '"cpu_baseline": {[^}]*}' gpurun_out/final_bench.log | cut -c1-400
grep -o '"stages": {.*}}, "phases' gpurun_out/final_bench.log
timeout 400 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/final_bench_reference.log; cut -c1-300 gpurun_out/final_bench_reference.log
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:plan_rows|gather_rows|replay_rows" -c 6 --csv --log-file gpurun_out/assemble_list.csv python bench.py --steps 1 --warmup 0 --new-tokens 4 --no-e2e --no-cpu-baseline --no-profile > /dev/null 2>&1
grep -E "plan_rows|gather_rows|replay_rows" gpurun_out/assemble_list.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -12
timeout 240 python bench.py --model gar-8b --batch 2 --new-tokens 256 --max-num-tiles 8 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/final_bench_8b.log; cut -c1-400 gpurun_out/final_bench_8b.log
