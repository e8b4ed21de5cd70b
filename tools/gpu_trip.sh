mkdir -p gpurun_out
bash tools/gpu_selftest.sh
python tools/prof_kernels.py attn_vit 5 2>&1 | tee gpurun_out/kernels_timing.log
python tools/prof_kernels.py attn_causal 5 2>&1 | tee -a gpurun_out/kernels_timing.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
echo "== bench gemv off"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-profile --opt gemv_decode=0 2>&1 | tail -3 | tee gpurun_out/bench_b.log
