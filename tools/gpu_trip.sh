mkdir -p gpurun_out
python tools/prof_kernels.py attn_timeline 2>&1 | tee gpurun_out/attn_timeline.log
python tools/prof_kernels.py gemm 5 2>&1 | tee gpurun_out/kernels_timing.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm_epilogues or inplace" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
