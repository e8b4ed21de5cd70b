#!/bin/bash
# N=2 check of the torchrun path, every leg under its own short timeout (a wedged NCCL must not eat the GPU budget).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo.log 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== probe (eager init)"
NCCL_DEBUG=WARN timeout 150 $RUN tools/nccl_probe.py > gpurun_out/probe_eager.log 2>&1; echo "rc=$?"
grep -c "done" gpurun_out/probe_eager.log
if ! grep -q "done" gpurun_out/probe_eager.log; then
  echo "== probe (lazy init)"
  PROBE_EAGER=0 NCCL_DEBUG=INFO timeout 150 $RUN tools/nccl_probe.py > gpurun_out/probe_lazy.log 2>&1; echo "rc=$?"
  if ! grep -q "done" gpurun_out/probe_lazy.log; then
    echo "== probe (P2P disabled)"
    NCCL_P2P_DISABLE=1 NCCL_DEBUG=INFO timeout 150 $RUN tools/nccl_probe.py > gpurun_out/probe_nop2p.log 2>&1; echo "rc=$?"
    tail -5 gpurun_out/probe_nop2p.log
  fi
fi
echo "== bench N=$N"
NCCL_DEBUG=WARN timeout 420 $RUN bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -2 gpurun_out/bench_n$N.log | cut -c1-600
grep "bench rank" gpurun_out/bench_n$N.err | tail -12
tail -5 gpurun_out/probe_eager.log
