"""Where does a decode-step GEMV launch spend its time?  Per-CTA globaltimer stamps (gar_gemv_debug_timeline) of the
GAR-1B decode shapes, cold L2 (a 512 MB write in front of every launch), alone and inside a PDL chain.
    python tools/gemv_timeline.py            -> gpurun_out/gemv_timeline.log"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from grasp_any_region_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
bf = torch.bfloat16
D, I, QKV, B = 2048, 8192, 3072, 8
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
names = ["entry", "w_issued", "dep_wait", "norm_ready", "stream_done", "partials", "stored"]


def report(tag, buf, grid):
    t = buf[: grid * 8].view(grid, 8).cpu()
    t0 = int(t[:, 0].min())
    rel = (t[:, :7] - t0).double() / 1e3
    med = rel.median(0).values.tolist()
    mx = rel.max(0).values.tolist()
    mn = rel.min(0).values.tolist()
    print(f"{tag}: grid {grid}  kernel span {mx[6]:.2f} us (first CTA entry -> last CTA stored)")
    print("   stamp       " + "  ".join(f"{n:>11s}" for n in names))
    print("   min    us   " + "  ".join(f"{v:11.2f}" for v in mn))
    print("   median us   " + "  ".join(f"{v:11.2f}" for v in med))
    print("   max    us   " + "  ".join(f"{v:11.2f}" for v in mx))
    d = (t[:, 1:7] - t[:, 0:6]).double() / 1e3
    print("   per-CTA phase medians us: " + "  ".join(f"{names[i]}->{names[i + 1]} {d[:, i].median().item():.2f}" for i in range(6)))


def run(tag, fn, grid, chain=None):
    buf = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
    for _ in range(3):
        flush.fill_(1)
        if chain:
            chain()
        lib.gar_gemv_debug_timeline(C.c_void_p(buf.data_ptr()))
        fn()
        lib.gar_gemv_debug_timeline(None)
        torch.cuda.synchronize()
    report(tag, buf, grid)


x = torch.randn((B, D), device=dev, dtype=bf)
xi = torch.randn((B, I), device=dev, dtype=bf)
res = torch.randn((B, D), device=dev, dtype=bf)
nw = torch.ones((D,), device=dev, dtype=bf)
wqkv = torch.randn((QKV, D), device=dev, dtype=bf) * 0.02
wo = torch.randn((D, D), device=dev, dtype=bf) * 0.02
wgu = torch.randn((2 * I, D), device=dev, dtype=bf) * 0.02
wdn = torch.randn((D, I), device=dev, dtype=bf) * 0.02
sq = torch.rand((128, 8), device=dev, dtype=torch.float32)
oq = torch.empty((B, QKV), device=dev, dtype=bf)
og = torch.empty((B, I), device=dev, dtype=bf)
oo = torch.empty((B, D), device=dev, dtype=bf)

print(torch.cuda.get_device_name(0))
run("qkv  (N=3072,K=2048, norm prologue)", lambda: ops.gemv_decode(x, wqkv, 0, norm_w=nw, sumsq_in=sq, out=oq), QKV // 16)
run("o    (N=2048,K=2048, residual)", lambda: ops.gemv_decode(x, wo, 1, res=res, want_sumsq=True, out=oo), D // 16)
run("gu   (N=16384,K=2048, norm+swiglu)", lambda: ops.gemv_decode(x, wgu, 2, norm_w=nw, sumsq_in=sq, out=og), I // 32)
run("down (N=2048,K=8192, residual)", lambda: ops.gemv_decode(xi, wdn, 1, res=res, want_sumsq=True, out=oo), D // 16)
# inside a PDL chain: the predecessor is another weight stream (gu), as in the decode step
run("down after gu (PDL chain)", lambda: ops.gemv_decode(xi, wdn, 1, res=res, want_sumsq=True, out=oo), D // 16,
    chain=lambda: ops.gemv_decode(x, wgu, 2, norm_w=nw, sumsq_in=sq, out=og))
run("qkv after down (PDL chain)", lambda: ops.gemv_decode(x, wqkv, 0, norm_w=nw, sumsq_in=sq, out=oq), QKV // 16,
    chain=lambda: ops.gemv_decode(xi, wdn, 1, res=res, want_sumsq=True, out=oo))

# plain event timing of each kernel in a loop over 8 distinct weight copies (cold-ish L2), for reference
def ev_ms(fn, n=64):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(0); torch.cuda.synchronize(); a.record()
    for i in range(n):
        fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


wq8 = [wqkv.clone() for _ in range(12)]
wg8 = [wgu.clone() for _ in range(4)]
wd8 = [wdn.clone() for _ in range(6)]
print("loop over distinct weights, us/launch: qkv %.2f  gu %.2f  down %.2f" % (
    ev_ms(lambda i: ops.gemv_decode(x, wq8[i % 12], 0, norm_w=nw, sumsq_in=sq, out=oq)),
    ev_ms(lambda i: ops.gemv_decode(x, wg8[i % 4], 2, norm_w=nw, sumsq_in=sq, out=og)),
    ev_ms(lambda i: ops.gemv_decode(xi, wd8[i % 6], 1, res=res, want_sumsq=True, out=oo))))
