#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "raw_uint8 or pipelined" 2>&1 | tail -3 | tee gpurun_out/tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline --no-profile 2>&1 | tail -1 > gpurun_out/bench.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
PY
