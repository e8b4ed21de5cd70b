mkdir -p gpurun_out
bash tools/gpu_selftest.sh
python tools/prof_kernels.py attn_vit 5 2>&1 | tee gpurun_out/kernels_timing.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
