#!/bin/bash
# Reduced round-end profile pass (GPU-minute budget): ncu launch list of one bench step (8 decode tokens, one decode
# group: ncu serialises launches anyway), DRAM traffic of every tcgen05 GEMM launch of that step, and one full capture
# of the persistent attention kernel (ViT shape and causal prefill shape).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 8 --decode-groups 1 --no-e2e --no-cpu-baseline --no-lib-baseline --no-profile"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 2 -f -o gpurun_out/prof_attn_persist python tools/prof_attn_one.py > gpurun_out/ncu_full_attn.log 2>&1
tail -1 gpurun_out/ncu_full_attn.log
ls -la gpurun_out | grep -E "launches.csv|prof_attn_persist"
