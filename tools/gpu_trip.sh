#!/bin/bash
# In-step A/B of two library builds on the same box: (old) before the attention P-in-TMEM change, (new) HEAD.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/ab.log
  python - "$1" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
print(sys.argv[1], "| captions/s %.2f" % d["value"], "| enc ms/req %.2f" % d["phases"]["vit_assemble_prefill_ms_per_request"], "| attn ms %.1f (%.0f TF/s)" % (d["stages"]["attn_tcgen05"]["ms"], d["stages"]["attn_tcgen05"]["TFLOP/s"]), "| gemm ms %.1f" % d["roofline"]["ms_per_step"], d["clocks"]["sm_mhz"])
PY
}
GAR_B200_LIB=$PWD/tools/micro/libgar_b200_before_attn.so run old
run new
GAR_B200_LIB=$PWD/tools/micro/libgar_b200_before_attn.so run old
run new
