"""Cost of preparing a batch of raw (uint8) requests (host enqueue / GPU time) and end-to-end ms per step of
generate_batches(overlap="decode") over raw requests for ways = 3 / 1 and 3 / 6 steps (warm allocator)."""
import sys, time, torch
sys.path.insert(0, ".")
from grasp_any_region_b200 import GARConfig, GARModel
from grasp_any_region_b200.config import gar_1b
from grasp_any_region_b200.synth import make_raw_request
dev = torch.device("cuda:0")
spec = gar_1b()
B, NT = 8, 128
model = GARModel(GARConfig.from_spec(spec)).random_init_(0).to(dev)
reqs = [make_raw_request(spec, seed=i, n_masks=1, max_num_tiles=16) for i in range(B)]
S, T = reqs[0]["input_ids"].shape[1], 17
model.configure_engine(max_tiles=B * T, max_seq=((S + NT + 127) // 128) * 128, max_batch=B, max_new_tokens=NT, max_prefill_rows=B * S)
eng = model.engine()
for it in range(4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    p = model._prepare(eng, reqs)
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"prepare of {B} raw requests: host enqueue {1e3*(t1-t0):.1f} ms, GPU {a.elapsed_time(b):.1f} ms, until idle {1e3*(t2-t0):.1f} ms", flush=True)
for ways in (3, 1):
    for K in (3, 6):
        model.generate_batches([reqs] * K, max_new_tokens=NT, overlap="decode", ways=ways)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        model.generate_batches([reqs] * K, max_new_tokens=NT, overlap="decode", ways=ways)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"ways={ways} K={K} raw e2e: {1e3*(t1-t0)/K:.1f} ms/step", flush=True)
