"""Prefill / ViT attention: persistent kernel (device-side tile counter) vs the per-tile kernel, on the bench's shapes.

Outputs must be bit-identical (same tiles, same arithmetic); prints ms per launch and TFLOP/s for both.
"""
import sys

import torch

sys.path.insert(0, ".")
from grasp_any_region_b200 import _lib, ops  # noqa: E402

dev, bf = "cuda:0", torch.bfloat16
lib = _lib.load()


def ms(fn, n=10):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print(torch.cuda.get_device_name(0))
for hd, T, H, S, hq, hkv in ((64, 136, 16, 4672, 32, 8), (128, 40, 12, 4672, 32, 8)):
    ntok, ms_ = 1025, 4864
    qkv = torch.randn(T * ntok, 3 * H * hd, device=dev).to(bf)
    q = torch.randn(S, (hq + 2 * hkv) * hd, device=dev).to(bf)
    kc = torch.randn(hkv, ms_, hd, device=dev).to(bf)
    vc = torch.randn(hkv, ms_, hd, device=dev).to(bf)
    fl_v = 4.0 * T * ntok * ntok * H * hd / 1e9
    fl_p = 2.0 * S * S * hd * hq / 1e9
    res = {}
    for persist in (0, 1, 0, 1):
        lib.gar_set_global_option(b"attn_persist", persist)
        ov = ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd)
        op = ops.attn_fwd_cache(q, q.stride(0), kc, vc, hq, hkv, hd, S, ms_)
        tv = ms(lambda: ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd))
        tp = ms(lambda: ops.attn_fwd_cache(q, q.stride(0), kc, vc, hq, hkv, hd, S, ms_))
        if persist in res:
            same = torch.equal(res[persist][0], ov) and torch.equal(res[persist][1], op)
        else:
            res[persist] = (ov, op)
            same = True
        print(f"hd {hd:3d} persist={persist}: ViT ({T}x{H}x{ntok}) {tv:7.3f} ms {fl_v / tv:6.0f} TF/s | causal prefill S={S} "
              f"{tp:7.3f} ms {fl_p / tp:6.0f} TF/s | repeatable={same}", flush=True)
    print(f"hd {hd:3d} persistent == per-tile bit for bit: ViT {torch.equal(res[0][0], res[1][0])}, "
          f"prefill {torch.equal(res[0][1], res[1][1])}", flush=True)
lib.gar_set_global_option(b"attn_persist", 1)
