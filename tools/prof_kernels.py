"""Standalone launches of the hot kernels at GAR-1B shapes (for ncu captures and quick CUDA-event timings)."""
import sys
import torch
sys.path.insert(0, ".")
from grasp_any_region_b200 import ops

dev, bf = "cuda:0", torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3


def timeit(name, fn, work=None, unit="TFLOP/s"):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    extra = ""
    if work:
        extra = f"  {work / (ms * 1e-3) / (1e12 if unit == 'TFLOP/s' else 1e9):8.1f} {unit}"
    print(f"{name:40s} {ms * 1e3:9.1f} us{extra}", flush=True)


if which in ("all", "attn_vit", "attn_causal"):
    from grasp_any_region_b200 import _lib as _l
    for bkv in (64, 128):
        _l.load().gar_set_global_option(b"attn_bkv", bkv)
        if which in ("all", "attn_vit"):
            T, ntok, H, hd = 17, 1025, 16, 64
            qkv = torch.randn(T * ntok, 3 * H * hd, device=dev).to(bf)
            timeit(f"attn vit T=17 N=1025 H=16 bkv={bkv}", lambda: ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd), 4.0 * T * ntok * ntok * H * hd)
        if which in ("all", "attn_causal"):
            S, hq, hkv, hd, ms_ = 4672, 32, 8, 64, 4864
            q = torch.randn(S, (hq + 2 * hkv) * hd, device=dev).to(bf)
            kc = torch.randn(hkv, ms_, hd, device=dev).to(bf); vc = torch.randn(hkv, ms_, hd, device=dev).to(bf)
            timeit(f"attn causal S=4672 32/8 heads bkv={bkv}", lambda: ops.attn_fwd_cache(q, q.stride(0), kc, vc, hq, hkv, hd, S, ms_), 2.0 * S * S * hq * hd)
    _l.load().gar_set_global_option(b"attn_bkv", 128)
if which in ("all", "decode_attn"):
    B, hq, hkv, hd, ms_ = 8, 32, 8, 64, 4864
    qkv = torch.randn(B, (hq + 2 * hkv) * hd, device=dev).to(bf)
    kc = torch.randn(B, hkv, ms_, hd, device=dev).to(bf); vc = torch.randn(B, hkv, ms_, hd, device=dev).to(bf)
    pos = torch.full((B,), 4700, dtype=torch.int32, device=dev)
    from grasp_any_region_b200 import _lib
    for impl in (1, 0):
        _lib.load().gar_set_global_option(b"decode_attn_mma", impl)
        for sp in (4, 8, 10, 16):
            timeit(f"decode attn mma={impl} B=8 len=4701 splits={sp}",
                   lambda: ops.decode_attn(qkv, kc, vc, pos, hq, hkv, hd, ms_, splits=sp), B * hkv * 4701 * hd * 2 * 2, "GB/s")
    _lib.load().gar_set_global_option(b"decode_attn_mma", 1)
if which in ("all", "skinny"):
    for (N, K) in ((3072, 2048), (2048, 2048), (16384, 2048), (2048, 8192), (128512, 2048)):
        x = torch.randn(8, K, device=dev).to(bf); w = (torch.randn(N, K, device=dev) * 0.05).to(bf)
        timeit(f"skinny M=8 N={N} K={K}", lambda: ops.gemm_skinny(x, w), N * K * 2, "GB/s")
if which in ("all", "gemm"):
    for (M, N, K, epi) in ((17425, 3072, 1024, 1), (17425, 1024, 1024, 3), (17425, 4096, 1024, 2), (17425, 1024, 4096, 3),
                           (4672, 3072, 2048, 0), (4672, 2048, 2048, 4), (4672, 16384, 2048, 6), (4672, 2048, 8192, 4)):
        a = torch.randn(M, K, device=dev).to(bf); w = (torch.randn(N, K, device=dev) * 0.05).to(bf)
        bias = torch.randn(N, device=dev).to(bf); gamma = torch.randn(N, device=dev).to(bf)
        res = torch.randn(M, N, device=dev).to(bf)
        timeit(f"gemm M={M} N={N} K={K} epi={epi}", lambda: ops.gemm(a, w, epi=epi, bias=bias, gamma=gamma, res=res), 2.0 * M * N * K)

if which == "attn_timeline":
    from grasp_any_region_b200 import _lib
    T, ntok, H, hd = 136, 1025, 16, 64
    qkv = torch.randn(T * ntok, 3 * H * hd, device=dev).to(bf)
    buf = torch.zeros(1024, dtype=torch.int64, device=dev)
    ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd); torch.cuda.synchronize()
    _lib.load().gar_attn_debug_timeline(buf.data_ptr())
    ops.attn_fwd_packed_qkv(qkv, T, ntok, H, hd); torch.cuda.synchronize()
    _lib.load().gar_attn_debug_timeline(None)
    t = buf[:128].view(16, 8).cpu()
    m = buf[512:640].view(16, 8).cpu()
    print("softmax thread, per KV block (cycles): wait_S | tmem_ld | max | wait_PV | exp+store | fence+arrive | block total | start-to-start")
    for j in range(9):
        r = [int(x) for x in t[j]]
        nxt = int(t[j + 1, 0]) if j < 8 else r[6]
        wpv = (r[4] - r[3]) if j > 0 else 0
        exp0 = r[4] if j > 0 else r[3]
        print(f"j={j}: {r[1]-r[0]:6d} | {r[2]-r[1]:6d} | {r[3]-r[2]:6d} | {wpv:6d} | {r[5]-exp0:6d} | {r[6]-r[5]:6d} | {r[6]-r[0]:6d} | {nxt-r[0]:6d}")
    print("MMA thread, per KV block (cycles): wait_SCONS | issue_S(j+1) (incl. wait K) | wait_PFULL | wait_V | issue_PV | iteration")
    for j in range(9):
        r = [int(x) for x in m[j]]
        if j < 8:
            print(f"j={j}: {r[1]-r[0]:6d} | {r[2]-r[1]:6d} | {r[3]-r[2]:6d} | {r[4]-r[3]:6d} | {r[5]-r[4]:6d} | {r[5]-r[0]:6d}")
        else:
            print(f"j={j}:      - |      - | {r[3]-r[2]:6d} | {r[4]-r[3]:6d} | {r[5]-r[4]:6d} | {r[5]-r[0]:6d}")
    print("softmax first stamp -> last stamp:", int(t[8, 6]) - int(t[0, 0]), "cycles; MMA thread:", int(m[8, 5]) - int(m[0, 0]))

if which == "gemm_big":
    from grasp_any_region_b200 import _lib as _l2
    for (M, N, K, epi) in ((139400, 3072, 1024, 1), (139400, 4096, 1024, 2), (139400, 1024, 4096, 3), (37376, 16384, 2048, 6),
                           (37376, 2048, 8192, 4)):
        a = torch.randn(M, K, device=dev).to(bf); w = (torch.randn(N, K, device=dev) * 0.05).to(bf)
        bias = torch.randn(N, device=dev).to(bf); gamma = torch.randn(N, device=dev).to(bf)
        res = torch.randn(M, N, device=dev).to(bf)
        for two in (1, 0):
            _l2.load().gar_set_global_option(b"gemm_2cta", two)
            timeit(f"gemm 2cta={two} M={M} N={N} K={K} epi={epi}", lambda: ops.gemm(a, w, epi=epi, bias=bias, gamma=gamma, res=res), 2.0 * M * N * K)
        del a, w, res
    _l2.load().gar_set_global_option(b"gemm_2cta", 1)
