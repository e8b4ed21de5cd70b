#!/bin/bash
# Round-2 evidence on one B200: assemble kernels under ncu, compute-sanitizer over a kernel subset, the driver-runnable
# bench lines of BASELINE configs 3 and 4.  Outputs -> gpurun_out/ (summaries are copied to profiles/ afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -m gpu -x -k "assemble or raw_uint8 or malformed" 2>&1 | tail -3 | tee gpurun_out/ev_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:plan_rows|gather_rows|replay_rows" -c 9 --csv --log-file gpurun_out/r02_assemble_list.csv python bench.py --steps 1 --warmup 0 --new-tokens 4 --no-e2e --no-cpu-baseline --no-lib-baseline --no-profile > /dev/null 2>&1
grep -E "plan_rows|gather_rows|replay_rows" gpurun_out/r02_assemble_list.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -30 | tee gpurun_out/r02_assemble_kernels_ncu.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -p no:cacheprovider -k "assemble_replay or device_preprocessing or (gemm_plain and 64) or decode_attention or gemv_decode_chain or (attention_vit and not 128)" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/r02_sanitizer_memcheck.log
tail -5 gpurun_out/r02_sanitizer_memcheck.log
timeout 600 python bench.py --masks 4 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/r02_bench_config3_4masks_batch4.log; cut -c1-330 gpurun_out/r02_bench_config3_4masks_batch4.log
timeout 900 python bench.py --model gar-8b --batch 2 --new-tokens 256 --max-num-tiles 16 --steps 2 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 > gpurun_out/r02_bench_config4_gar8b_batch2_256tok_T17.log; cut -c1-330 gpurun_out/r02_bench_config4_gar8b_batch2_256tok_T17.log
