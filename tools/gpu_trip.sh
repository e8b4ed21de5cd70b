mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "decode_attention" 2>&1 | tail -3 | tee gpurun_out/dec_tests.log
timeout 400 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/engine_tests.log
echo "== bench fused rope"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_a.log | cut -c1-300
grep -o '"phases": {[^}]*}' gpurun_out/bench_a.log
echo "== bench unfused rope"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --opt fuse_decode_rope=0 2>&1 | tail -1 | tee gpurun_out/bench_b.log | cut -c1-300
grep -o '"phases": {[^}]*}' gpurun_out/bench_b.log
