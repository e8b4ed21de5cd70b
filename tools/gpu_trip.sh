mkdir -p gpurun_out
echo "== bench full (ours)"; timeout 1200 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_full.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_reference.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
