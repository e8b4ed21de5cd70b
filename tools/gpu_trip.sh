mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "forward_loss" -s 2>&1 | tail -8 | tee gpurun_out/forward_loss.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
