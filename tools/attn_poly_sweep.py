"""Attention throughput vs the share of softmax exponentials moved to the FMA pipe (global option attn_poly)."""
import sys

import torch

sys.path.insert(0, ".")
from grasp_any_region_b200 import _lib, ops  # noqa: E402

dev, bf = "cuda:0", torch.bfloat16
lib = _lib.load()
T, ntok, H, hd = 136, 1025, 16, 64
qkv = torch.randn(T * ntok, 3 * H * hd, device=dev).to(bf)
S, hq, hkv, ms_ = 4672, 32, 8, 4864
q = torch.randn(S, (hq + 2 * hkv) * hd, device=dev).to(bf)
kc = torch.randn(hkv, ms_, hd, device=dev).to(bf)
vc = torch.randn(hkv, ms_, hd, device=dev).to(bf)


def ms(fn, n=8):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


fl_v = 4.0 * T * ntok * ntok * H * hd / 1e9
fl_p = 2.0 * S * S * hd * hq / 1e9
print(torch.cuda.get_device_name(0))
ref = None
for poly, label in ((0, "all MUFU.EX2"), (1, "2/8 on the FMA pipe"), (2, "3/8 on the FMA pipe"), (3, "4/8 on the FMA pipe")):
    lib.gar_set_global_option(b"attn_poly", poly)
    out = ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd).float()
    if ref is None:
        ref = out
    tv = ms(lambda: ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd))
    tp = ms(lambda: ops.attn_fwd_cache(q, q.stride(0), kc, vc, hq, hkv, hd, S, ms_))
    print(f"attn_poly={poly} ({label:20s}) ViT {tv:7.3f} ms {fl_v / tv:6.0f} TF/s | causal prefill {tp:7.3f} ms {fl_p / tp:6.0f} TF/s"
          f" | max|out - out(poly=0)| = {(out - ref).abs().max().item():.2e}")
lib.gar_set_global_option(b"attn_poly", 0)
