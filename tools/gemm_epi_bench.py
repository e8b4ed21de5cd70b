"""tcgen05 GEMM with each fused epilogue at the production shapes of one bench step (B200): what the epilogue costs on
top of the plain GEMM, next to torch.matmul.   python tools/gemm_epi_bench.py"""
import sys

import torch

sys.path.insert(0, ".")
from grasp_any_region_b200 import ops  # noqa: E402
from grasp_any_region_b200.config import gar_1b  # noqa: E402
from grasp_any_region_b200.rope import rope_table_chunk_major, vit_rope_table  # noqa: E402

dev, bf = "cuda:0", torch.bfloat16
spec = gar_1b()
v = spec.vision
Mv, Ml = 8 * 17 * v.tokens, 8 * 4672


def ms(fn, iters=6, warm=2):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def rnd(*shape, scale=1.0):
    return (torch.randn(shape, device=dev) * scale).to(bf)


print(torch.cuda.get_device_name(0))
rope = rope_table_chunk_major(vit_rope_table(v).to(bf).to(dev), v.head_dim)
cases = [
    ("vit.fc1  + bias only", Mv, v.mlp_dim, v.embed_dim, ops.EPI_BIAS),
    ("vit.proj + bias only", Mv, v.embed_dim, v.embed_dim, ops.EPI_BIAS),
    ("vit.proj + residual only", Mv, v.embed_dim, v.embed_dim, ops.EPI_RES),
    ("vit.qkv  + bias + 2-D RoPE", Mv, 3 * v.embed_dim, v.embed_dim, "rope"),
    ("vit.proj + bias + LayerScale + residual", Mv, v.embed_dim, v.embed_dim, ops.EPI_BIAS_LS_RES),
    ("vit.fc1  + bias + GELU", Mv, v.mlp_dim, v.embed_dim, ops.EPI_BIAS_GELU),
    ("vit.fc2  + bias + LayerScale + residual", Mv, v.embed_dim, v.mlp_dim, ops.EPI_BIAS_LS_RES),
    ("llm.gate_up + SwiGLU", Ml, 16384, 2048, ops.EPI_SWIGLU),
    ("llm.down + residual", Ml, 2048, 8192, ops.EPI_RES),
]
for name, M, N, K, epi in cases:
    a, w = rnd(M, K, scale=0.5), rnd(N, K, scale=0.03)
    bias, gamma = rnd(N, scale=0.1), rnd(N, scale=0.2)
    n_out = N // 2 if epi == ops.EPI_SWIGLU else N
    res = rnd(M, n_out) if epi in (ops.EPI_BIAS_LS_RES, ops.EPI_RES) else None
    out = torch.empty((M, n_out), device=dev, dtype=bf)
    plain_out = torch.empty((M, N), device=dev, dtype=bf)
    t_plain = ms(lambda: ops.gemm(a, w, out=plain_out))
    if epi == "rope":
        t_epi = ms(lambda: ops.gemm_rope2d(a, w, bias, rope, v.tokens, 1 if v.use_cls_token else 0, v.head_dim, 2 * v.embed_dim, chunk_major=True))
    else:
        t_epi = ms(lambda: ops.gemm(a, w, epi=epi, bias=bias, gamma=gamma, res=res, out=out))
    t_lib = ms(lambda: torch.matmul(a, w.t(), out=plain_out))
    fl = 2.0 * M * N * K / 1e9
    print(f"{name:42s} M={M:6d} N={N:5d} K={K:4d}  plain {t_plain:6.3f} ms ({fl / t_plain:5.0f} TF/s)  fused {t_epi:6.3f} ms "
          f"({fl / t_epi:5.0f} TF/s, +{(t_epi / t_plain - 1) * 100:4.1f} %)  torch.matmul {t_lib:6.3f} ms")
    del a, w, out, plain_out, res
