mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "batched_encode or batched_decode or pipelined" 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
