"""Probe: do two independent decode loops (two engines = two KV groups, two streams, two host threads) overlap on the GPU?

The decode step is latency-bound (0.58 of HBM peak, ~80 dependent launches per step): a second, independent chain of
kernels can fill the gaps.  Prints ms for one group decoded alone and for two groups decoded concurrently.
"""
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from grasp_any_region_b200 import GARConfig, GARModel  # noqa: E402
from grasp_any_region_b200.config import gar_1b  # noqa: E402
from grasp_any_region_b200.engine import Engine  # noqa: E402
from grasp_any_region_b200.synth import make_request  # noqa: E402

dev = torch.device("cuda:0")
spec = gar_1b()
B, NT = 8, 128
model = GARModel(GARConfig.from_spec(spec)).random_init_(0).to(dev)
reqs = [make_request(spec, seed=i, n_masks=1, max_num_tiles=16, dtype=torch.bfloat16) for i in range(B)]
S, T = reqs[0]["input_ids"].shape[1], reqs[0]["pixel_values"].shape[0]
for r in reqs:
    for k in ("pixel_values", "global_mask_values", "input_ids"):
        r[k] = r[k].to(dev)
caps = dict(max_tiles=B * T, max_seq=((S + NT + 127) // 128) * 128, max_batch=B, max_new_tokens=NT, max_prefill_rows=B * S)
model.configure_engine(**caps)
e1 = model.engine()
ways = int(sys.argv[1]) if len(sys.argv) > 1 else 2
engines = [e1] + [Engine(spec, model.state_dict(), device=str(dev), **caps) for _ in range(ways - 1)]
print(torch.cuda.get_device_name(0), "S", S, "T", T, flush=True)


def stage(e):
    model._stage(e, 0, reqs)
    e.stream.synchronize()


def decode(e, out, i):
    out[i] = e.generate(B, NT, None)


res = [None] * ways
for e in engines:  # warm-up: kernel attributes, decode graphs
    stage(e)
    decode(e, res, 0)
ref = res[0]
for rep in range(3):
    for e in engines:
        stage(e)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); decode(e1, res, 0); torch.cuda.synchronize(); t1 = time.perf_counter()
    alone = (t1 - t0) * 1e3
    same1 = res[0] == ref
    for e in engines:
        stage(e)
    torch.cuda.synchronize()
    th = [threading.Thread(target=decode, args=(e, res, i)) for i, e in enumerate(engines)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dual = (time.perf_counter() - t0) * 1e3
    print(f"decode of one group of {B} alone: {alone:7.1f} ms ({alone / (NT - 1):.3f} ms/step) | {ways} groups concurrently: {dual:7.1f} ms "
          f"= {dual / ways:7.1f} ms per group ({dual / ways / (NT - 1):.3f} ms/step) | tokens identical: {same1 and all(r == ref for r in res)}",
          flush=True)
