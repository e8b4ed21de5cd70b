bash tools/gpu_selftest.sh
echo "== bench pdl on"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_pdl1.log
echo "== bench pdl off"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --opt pdl=0 2>&1 | tail -3 | tee gpurun_out/bench_pdl0.log
