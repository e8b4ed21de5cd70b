#!/bin/bash
# Round-end evidence: ncu launch list of the bench command (reduced decode length so it finishes), DRAM traffic of
# every tcgen05 GEMM launch of one step, one full capture of the top kernels, and the decode kernels of the GAR-8B
# configuration.  Outputs -> gpurun_out/ (summaries are copied to profiles/ with tools/ncu_summary.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 8 --no-e2e --no-cpu-baseline --no-lib-baseline --no-profile"
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16_tn -c 400 --csv --log-file gpurun_out/gemm_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 12 -c 3 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_full_attn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemv_decode -s 20 -c 4 -f -o gpurun_out/prof_gemv $CMD > gpurun_out/ncu_full_gemv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_attn -s 8 -c 1 -f -o gpurun_out/prof_decode_attn $CMD > gpurun_out/ncu_full_dattn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:gather_rows|replay_rows" -c 2 -f -o gpurun_out/prof_assemble $CMD > gpurun_out/ncu_full_asm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k "regex:decode_attn|gemv_decode|argmax|embed_gather|gemm_skinny" -c 700 --csv --log-file gpurun_out/decode_list_8b.csv python bench.py --model gar-8b --batch 2 --new-tokens 4 --max-num-tiles 4 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-lib-baseline --no-profile > gpurun_out/ncu_8b.log 2>&1
tail -1 gpurun_out/ncu_8b.log | cut -c1-200
ls -la gpurun_out | tail -16
