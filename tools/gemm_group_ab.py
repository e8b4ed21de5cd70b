"""tcgen05 GEMM raster: column panels walked in L2-sized groups (global option gemm_group_kb) vs one group (0).

Shapes: the SwiGLU gate/up and down projections of the GAR-1B / GAR-8B prefill (the only launches of the path whose
weight matrix does not stay L2-resident when it is walked whole).  Settings are interleaved round by round (the GPU
drifts under its power cap, so back-to-back blocks per setting are not comparable); the median over rounds is printed.
Outputs must be bit-identical across settings.  `--once` launches every setting once (for an ncu traffic pass).
"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from grasp_any_region_b200 import _lib, ops  # noqa: E402

dev, bf = "cuda:0", torch.bfloat16
lib = _lib.load()
once = "--once" in sys.argv
settings = (0, 8192, 16384, 32768)


def ms(fn, n=3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print(torch.cuda.get_device_name(0))
shapes = [("gar-1b gate/up+SwiGLU", 37376, 16384, 2048, ops.EPI_SWIGLU), ("gar-1b down+res", 37376, 2048, 8192, ops.EPI_RES),
          ("gar-8b gate/up+SwiGLU", 9400, 28672, 4096, ops.EPI_SWIGLU), ("gar-8b down+res", 9400, 4096, 14336, ops.EPI_RES)]
for name, M, N, K, epi in shapes:
    a = (torch.randn((M, K), device=dev) * 0.5).to(bf)
    w = (torch.randn((N, K), device=dev) * 0.02).to(bf)
    res = torch.randn((M, N), device=dev).to(bf) if epi == ops.EPI_RES else None
    fl = 2.0 * M * N * K / 1e9
    ref, same = None, True
    times = {kb: [] for kb in settings}
    for r in range(1 if once else 8):
        for kb in settings:
            lib.gar_set_global_option(b"gemm_group_kb", kb)
            out = ops.gemm(a, w, epi=epi, res=res)
            torch.cuda.synchronize()
            if r == 0:
                ref = out if ref is None else ref
                same = same and torch.equal(ref, out)
            if not once:
                times[kb].append(ms(lambda: ops.gemm(a, w, epi=epi, res=res)))
    if once:
        print(f"{name}: launched {len(settings)} settings {settings}, bit-identical={same}")
    else:
        print(f"{name:22s} M={M} N={N} K={K}: " + " | ".join(
            f"{kb >> 10:2d} MB {statistics.median(t):6.3f} ms {fl / statistics.median(t):5.0f} TF/s" for kb, t in times.items())
            + f" | bit-identical={same}", flush=True)
    del a, w, res, ref, out
lib.gar_set_global_option(b"gemm_group_kb", 32768)
