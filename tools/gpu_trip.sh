# Generic short GPU trip (edit per experiment; outputs -> gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 | tee gpurun_out/tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"])
print("phases", d["phases"]); print("stages", d["stages"])
PY
