#!/bin/bash
# compute-sanitizer racecheck of this repo's kernels (PyTorch's own reference kernels excluded: its fmha_cutlassF
# reports shared-memory hazards of its own and exhausts the hazard limit)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --kernel-regex-exclude kns=fmha_cutlass python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -p no:cacheprovider -k "gemv_decode or decode_attention_fused or decode_attention_long or (assemble_plm and True-1)" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/r02_sanitizer_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer_racecheck.log | tail -3
grep -E "Race reported|hazard" gpurun_out/r02_sanitizer_racecheck.log | cut -c1-220 | sort | uniq -c | sort -rn | head -12
