"""Decode attention in isolation at the bench's shape (batch 8, 32 q / 8 kv heads x 64, ~4.7k keys): one CUDA graph of
16 launches over 16 distinct KV caches (as the 16 layers of a decode step: every launch reads cold HBM), replayed.
    python tools/decode_attn_bench.py [path/to/lib.so ...]      (default: the in-tree library)"""
import ctypes as C
import math
import sys

import torch

sys.path.insert(0, ".")
from grasp_any_region_b200 import _lib  # noqa: E402
from grasp_any_region_b200.config import TextSpec  # noqa: E402
from grasp_any_region_b200.rope import llama_cos_sin  # noqa: E402

dev = "cuda:0"
bf = torch.bfloat16
B, hq, hkv, hd, max_seq, L = 8, 32, 8, 64, 4864, 16
pos_val = 4736
cos, sin = llama_cos_sin(TextSpec(), max_seq)
cos, sin = cos.to(bf).to(dev), sin.to(bf).to(dev)
qkv = (torch.randn((B, (hq + 2 * hkv) * hd), device=dev) * 0.5).to(bf)
kc = (torch.randn((L, B, hkv, max_seq, hd), device=dev) * 0.5).to(bf)
vc = (torch.randn((L, B, hkv, max_seq, hd), device=dev) * 0.5).to(bf)
pos = torch.full((B,), pos_val, dtype=torch.int32, device=dev)
out = torch.empty((B, hq * hd), dtype=bf, device=dev)
bytes_per_launch = 2.0 * B * hkv * (pos_val + 1) * hd * 2
p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

print(torch.cuda.get_device_name(0), f"| {bytes_per_launch / 1e6:.1f} MB of K+V per launch")
for path in (sys.argv[1:] or [_lib.LIB_PATH]):
    lib = C.CDLL(path)
    fn = lib.gar_decode_attn_rope
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_float] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
    for splits in (4, 6, 8, 9, 11, 13, 16):
        tickets = torch.zeros((B * hkv,), dtype=torch.int32, device=dev)
        o_part = torch.empty((B, hq, splits, hd), dtype=torch.float32, device=dev)
        ml = torch.empty((B, hq, splits, 2), dtype=torch.float32, device=dev)
        st = torch.cuda.Stream()

        def launch_all():
            for layer in range(L):
                rc = fn(p(qkv), qkv.stride(0), p(kc[layer]), p(vc[layer]), p(pos), None, B, hq, hkv, hd, max_seq, splits,
                        1.0 / math.sqrt(hd), p(cos), p(sin), p(o_part), p(ml), p(tickets), p(out), hq * hd,
                        C.c_void_p(st.cuda_stream))
                assert rc == 0, rc

        with torch.cuda.stream(st):
            launch_all()
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                launch_all()
            for _ in range(3):
                g.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(20):
                g.replay()
            b.record(st)
            st.synchronize()
        us = a.elapsed_time(b) / 20 / L * 1e3
        print(f"{path.split('/')[-1]:24s} splits {splits:2d}: {us:6.2f} us/launch  {bytes_per_launch / us / 1e3:7.1f} GB/s")
