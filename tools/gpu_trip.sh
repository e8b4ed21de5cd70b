#!/bin/bash
# Working trip on one B200: decode-path tests + A/B of decode-step options through bench.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -m gpu -x -k "decode or gemv or engine" 2>&1 | tail -3 | tee gpurun_out/tests.log
run() {
  timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-lib-baseline --no-e2e --no-profile "$@" 2>&1 | tail -1 > gpurun_out/ab.log
  python - "$*" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
    print(sys.argv[1] or "default", "| captions/s %.2f" % d["value"], "| decode ms/step %.4f" % d["phases"]["decode_ms_per_token_step"], "| enc ms/req %.2f" % d["phases"]["vit_assemble_prefill_ms_per_request"], d["clocks"]["sm_mhz"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex, open("gpurun_out/ab.log").read()[-400:])
PY
}
run
run --opt gemv_norm_smem=0
run --opt decode_splits=4
run --opt decode_splits=8
run
