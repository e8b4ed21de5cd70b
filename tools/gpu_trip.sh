# Generic short GPU trip: full GPU tests + one bench line (edit per experiment; outputs -> gpurun_out/).
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/tests.log
timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
