# Generic short GPU trip: full GPU tests + one bench line (edit per experiment; outputs -> gpurun_out/).
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_full.txt
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 | tee gpurun_out/tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s 2>&1 | tail -60 | tee gpurun_out/tests_fullsize.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-600
