mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "non_square" 2>&1 | tail -12 | tee gpurun_out/ns_test.log
