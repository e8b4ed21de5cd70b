"""Minimal NCCL sanity probe (one rank per GPU, launched with torchrun): init, all-reduce, all-gather, barrier.
Separates 'the box's NCCL/NVLink set-up is broken' from 'bench.py's N>1 path is broken'.  Prints progress to stderr."""
import datetime
import os
import sys
import time

import torch
import torch.distributed as dist


def note(msg):
    print(f"[probe rank {os.environ.get('RANK')}] {time.time() % 1000:8.2f} {msg}", file=sys.stderr, flush=True)


def main():
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    note(f"device {torch.cuda.get_device_name(lr)}")
    eager = os.environ.get("PROBE_EAGER", "1") == "1"
    kw = {"device_id": torch.device(f"cuda:{lr}")} if eager else {}
    dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=60), **kw)
    note("process group up")
    x = torch.ones(1 << 20, device=f"cuda:{lr}")
    dist.all_reduce(x)
    torch.cuda.synchronize()
    note(f"all_reduce ok {x[0].item()}")
    t = torch.full((8, 128), dist.get_rank(), dtype=torch.int32, device=f"cuda:{lr}")
    buf = torch.empty((8 * dist.get_world_size(), 128), dtype=torch.int32, device=f"cuda:{lr}")
    dist.all_gather_into_tensor(buf, t)
    torch.cuda.synchronize()
    note(f"all_gather ok {buf[:, 0].tolist()}")
    dist.barrier()
    torch.cuda.synchronize()
    note("barrier ok")
    dist.destroy_process_group()
    note("done")


if __name__ == "__main__":
    main()
