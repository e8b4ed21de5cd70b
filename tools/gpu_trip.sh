mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/kernel_tests.log
timeout 400 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/engine_tests.log
timeout 200 python tools/prof_kernels.py attn_vit 3 2>&1 | tail -2 | tee gpurun_out/attn_timing.log
timeout 200 python tools/prof_kernels.py gemm 3 2>&1 | grep "epi=2\|N=4096" | tee gpurun_out/gemm_timing.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_a.log | cut -c1-200
grep -o '"phases": {[^}]*}' gpurun_out/bench_a.log; grep -o '"stages": {.*}}, "phases' gpurun_out/bench_a.log; grep -o '"roofline": {[^}]*}' gpurun_out/bench_a.log
