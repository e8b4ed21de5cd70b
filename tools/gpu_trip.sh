#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "assemble" 2>&1 | tail -3 | tee gpurun_out/tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/bench.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["clocks"])
print("phases", d["phases"]); print("stages", d["stages"])
for r in d["rooflines"]: print(r["kernel"][:40], round(r["frac"], 3))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:plan_rows|gather_rows|replay_rows" -c 9 --csv --log-file gpurun_out/r02_assemble_list.csv python bench.py --steps 1 --warmup 0 --new-tokens 4 --no-e2e --no-cpu-baseline --no-lib-baseline --no-profile > /dev/null 2>&1
grep -E "plan_rows|gather_rows|replay_rows" gpurun_out/r02_assemble_list.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -30 | tee gpurun_out/r02_assemble_kernels_ncu.txt | grep duration
