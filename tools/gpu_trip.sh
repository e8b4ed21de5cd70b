# Generic short GPU trip (edit per experiment; outputs -> gpurun_out/).
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_full.txt
timeout 120 tools/micro/green_ctx_probe 32 2>&1 | tee gpurun_out/green_probe.log
timeout 120 tools/micro/green_ctx_probe 48 2>&1 | tail -4 | tee -a gpurun_out/green_probe.log
timeout 300 python tools/gemv_timeline.py 2>&1 | tee gpurun_out/gemv_timeline.log
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu -x -k "assemble or pipelined or malformed or stage_parity" -s 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s 2>&1 | tail -40 | tee gpurun_out/tests_fullsize.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lib-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-lib-baseline --no-e2e --pipeline 2>&1 | tail -1 | tee gpurun_out/bench_pipe.log | cut -c1-300
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-lib-baseline --no-e2e --pipeline --decode-sms 32 2>&1 | tail -1 | tee gpurun_out/bench_pipe32.log | cut -c1-300
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-lib-baseline --no-e2e --pipeline --decode-sms 48 2>&1 | tail -1 | tee gpurun_out/bench_pipe48.log | cut -c1-300
