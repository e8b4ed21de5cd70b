mkdir -p gpurun_out
bash tools/gpu_selftest.sh tests/test_gpu_engine.py
echo "== bench"; timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
