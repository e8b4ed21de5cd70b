mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" 2>&1 | tail -5
python tools/prof_kernels.py attn_vit 5 2>&1 | tee gpurun_out/kernels_timing.log
python tools/prof_kernels.py attn_causal 5 2>&1 | tee -a gpurun_out/kernels_timing.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
