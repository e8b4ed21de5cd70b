mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm" 2>&1 | tail -3 | tee gpurun_out/gemm_tests.log
timeout 300 python tools/prof_kernels.py gemm_big 3 2>&1 | tee gpurun_out/kernels_timing.log
timeout 300 python tools/prof_kernels.py gemm 3 2>&1 | tee -a gpurun_out/kernels_timing.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
