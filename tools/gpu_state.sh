#!/bin/bash
# Re-entry check on one B200: full GPU suite (incl. full-size parity), smoke(), default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -s 2>&1 | tail -60 > gpurun_out/state_tests.log; tail -5 gpurun_out/state_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/state_smoke.log
timeout 700 python bench.py 2>&1 | tail -1 > gpurun_out/state_bench.log; cut -c1-300 gpurun_out/state_bench.log
