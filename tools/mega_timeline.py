"""Per-phase timeline of the persistent decode-step kernel at the bench shape (CTA 0's globaltimer stamps)."""
import sys
import torch
sys.path.insert(0, ".")
from grasp_any_region_b200 import GARConfig, GARModel, _lib
from grasp_any_region_b200.config import PRESETS
from grasp_any_region_b200.synth import make_request

dev, B, NT = "cuda:0", 8, 6
spec = PRESETS["gar-1b"]()
model = GARModel(GARConfig.from_spec(spec)).random_init_(0).to(dev)
reqs = [make_request(spec, seed=i, n_masks=1, max_num_tiles=16, dtype=torch.bfloat16) for i in range(B)]
S = reqs[0]["input_ids"].shape[1]
T = reqs[0]["pixel_values"].shape[0]
model.configure_engine(max_tiles=B * T, max_seq=((S + 128 + 127) // 128) * 128, max_batch=B, max_new_tokens=128,
                       max_prefill_rows=B * S)
eng = model.engine()
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
L = spec.text.num_layers
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
_lib.load().gar_decode_mega_debug_timeline(buf.data_ptr())
model.generate_batch(reqs, max_new_tokens=NT, eos_token_id=None)
torch.cuda.synchronize()
_lib.load().gar_decode_mega_debug_timeline(None)
t = buf.cpu().tolist()
full = t
n = 2 + 2 * (1 + 5 * L) 
t = t[:n]
print("step total us:", (t[-1] - t[0]) / 1e3)
names = ["qkv", "attn", "o", "gu", "down"]
print("embed work %.1f bar %.1f" % ((t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3))
work = {k: 0.0 for k in names}; bar = {k: 0.0 for k in names}
for l in range(L):
    for p, k in enumerate(names):
        i = 2 + 2 * (l * 5 + p)
        work[k] += (t[i + 1] - t[i]) / 1e3
        bar[k] += (t[i + 2] - t[i + 1]) / 1e3
for k in names:
    print(f"{k:5s} work/layer {work[k] / L:7.2f} us   barrier wait/layer {bar[k] / L:7.2f} us")
print("lm_head us:", (t[-1] - t[-2]) / 1e3)
l = 5
print("layer 5 detail:", [round((t[2 + 2 * (l * 5) + j + 1] - t[2 + 2 * (l * 5) + j]) / 1e3, 1) for j in range(10)])

for name, off, ph in (("attn", 512, 26), ("gu", 768, 28)):
    start = full[2 + 2 * ph]  # CTA 0: end of the barrier before this phase
    arr = sorted((full[off + i] - start) / 1e3 for i in range(148))
    print(name, "CTA arrival us after phase start: min %.1f p25 %.1f p50 %.1f p75 %.1f p90 %.1f max %.1f" % (
        arr[0], arr[37], arr[74], arr[111], arr[133], arr[-1]))
    slow = sorted(range(148), key=lambda i: -full[off + i])[:8]
    print("  slowest CTAs:", slow)
