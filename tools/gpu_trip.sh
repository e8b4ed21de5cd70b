#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -m gpu -x -k "decode or long_k or engine or eos" 2>&1 | tail -4 | tee gpurun_out/tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -k "8b" 2>&1 | tail -3 | tee -a gpurun_out/tests.log
timeout 900 python bench.py --model gar-8b --batch 2 --new-tokens 256 --max-num-tiles 16 --steps 2 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/bench_8b.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_8b.log").read().strip().splitlines()[-1])
print("8B value", d["value"], "e2e", d["e2e"]["value"], d["clocks"]); print("phases", d["phases"])
PY
