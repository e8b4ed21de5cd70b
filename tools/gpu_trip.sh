#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
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
run --opt gemv_wide=0
run
run --opt gemv_wide=0
