"""One ViT-shaped and one causal-prefill-shaped attention launch (for ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
from grasp_any_region_b200 import ops

dev, bf = "cuda:0", torch.bfloat16
T, ntok, H, hd = 34, 1025, 16, 64
qkv = torch.randn(T * ntok, 3 * H * hd, device=dev).to(bf)
ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd)
S, hq, hkv, ms_ = 4672, 32, 8, 4864
q = torch.randn(S, (hq + 2 * hkv) * hd, device=dev).to(bf)
kc = torch.randn(hkv, ms_, hd, device=dev).to(bf); vc = torch.randn(hkv, ms_, hd, device=dev).to(bf)
ops.attn_fwd_cache(q, q.stride(0), kc, vc, hq, hkv, hd, S, ms_)
torch.cuda.synchronize()
print("ok")
