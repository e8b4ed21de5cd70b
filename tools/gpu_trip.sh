mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -4 | tee gpurun_out/tests.log
timeout 300 python tools/prof_kernels.py attn_timeline 2>&1 | tail -25 | tee gpurun_out/attn_timeline.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["clocks"])
print("phases", d["phases"]); print("stages", d["stages"])
PY
