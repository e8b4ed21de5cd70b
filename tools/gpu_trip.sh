mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "megakernel or batched_decode" 2>&1 | tail -5 | tee gpurun_out/engine_tests.log
for o in "mega_unit_warps=2" "mega_unit_warps=4"; do
echo "== $o"; timeout 300 python tools/mega_timeline.py $o 2>&1 | tail -14 | tee gpurun_out/mega_timeline_$o.log
done
