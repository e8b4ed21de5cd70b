#!/bin/bash
# Runs every GPU test function in its own process (a device-side trap poisons the CUDA context)
# with a hard timeout, and writes a compact report to gpurun_out/selftest.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/selftest.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
FILES="${@:-tests/test_gpu_kernels.py tests/test_gpu_engine.py}"
FUNCS=$(python -m pytest $FILES --collect-only -q -m gpu 2>/dev/null | grep '::' | sed 's/\[.*//' | sort -u)
for f in $FUNCS; do
  echo "=== $f" >> $LOG
  timeout 150 python -m pytest "$f" -q -m gpu --tb=short -s -p no:cacheprovider 2>&1 | tail -60 >> $LOG
  echo "=== exit ${PIPESTATUS[0]}" >> $LOG
done
grep -E "^=== |passed|failed|error" $LOG | tail -80
