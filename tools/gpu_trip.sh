mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention or assemble" 2>&1 | tail -3 | tee gpurun_out/kernel_tests.log
timeout 200 python tools/prof_kernels.py attn_vit 3 2>&1 | tail -4 | tee gpurun_out/attn_timing.log
timeout 200 python tools/prof_kernels.py attn_causal 3 2>&1 | tail -4 | tee -a gpurun_out/attn_timing.log
timeout 400 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/engine_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -f -o gpurun_out/prof_attn2 python tools/prof_attn_one.py > gpurun_out/ncu_attn.log 2>&1
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_a.log | cut -c1-300
grep -o '"phases": {[^}]*}' gpurun_out/bench_a.log; grep -o '"stages": {.*}}, "phases' gpurun_out/bench_a.log
