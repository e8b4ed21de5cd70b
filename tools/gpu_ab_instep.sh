#!/bin/bash
# Same box, back to back, interleaved: the bench step with the round's two scheduling changes switched off
# (per-tile attention CTAs, single-group GEMM raster) and on (defaults).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/instep_ab.txt
echo "# bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline --no-e2e, same box, interleaved" > $out
for round in 1 2; do
  for cfg in "old --opt attn_persist=0 --opt gemm_group_kb=0" "attn_only --opt gemm_group_kb=0" "new"; do
    set -- $cfg; name=$1; shift
    timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline --no-e2e "$@" 2>/dev/null | tail -1 > gpurun_out/_ab.json
    python - "$name" <<'PY' >> $out
import json, sys
d = json.load(open("gpurun_out/_ab.json"))
r = {x["kernel"].split(" ")[0]: x for x in d["rooflines"]}
g, a = r["gemm_bf16_tn_kernel"], r["attn_fwd_kernel"]
print(f"{sys.argv[1]:9s} | captions/s {d['value']:.2f} | ms/step {d['ms_per_step']:.1f} | enc ms/req {d['phases']['vit_assemble_prefill_ms_per_request']:.2f} | "
      f"attn ms {a['ms_per_step']:.1f} ({a['achieved']:.0f} TF/s) | gemm ms {g['ms_per_step']:.1f} ({g['achieved']:.0f} TF/s) | "
      f"decode ms/step {d['phases']['decode_ms_per_token_step']:.3f} | sm {d['clocks']['sm_mhz']} MHz")
PY
  done
done
cat $out
