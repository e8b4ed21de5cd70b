mkdir -p gpurun_out
bash tools/gpu_selftest.sh
for k in attn_vit attn_causal decode_attn gemm; do python tools/prof_kernels.py $k 5 2>&1 | tee -a gpurun_out/kernels_timing.log; done
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_a.log
