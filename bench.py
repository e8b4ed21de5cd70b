#!/usr/bin/env python
"""Headline benchmark: region-captions/sec (128 decode tokens), GAR-1B bf16, on N B200s.

Workload (BASELINE.json configs[1]): 8 independent requests per GPU, each a synthetic 1024x1024 image
(17 tiles of 448^2) + 1 mask, S = 4672 prompt positions, 128 greedy tokens, random-init GAR-1B.
One "step" = that batch through the whole hot path: one ViT pass and one prefill pass over the 8 equal-shaped
requests (row plan -> patchify -> PE-ViT -> projector -> assemble/RoI replay -> Llama prefill), then one batched
128-token decode.

  value : captions/s with the requests already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : same metric through GARModel.generate_batch() from pinned HOST buffers (H2D of every
          request and D2H of the tokens inside the timed region)
  roofline : dominant kernel = the tcgen05 GEMM (gemm_bf16_tn_kernel); achieved = summed algorithmic
          2*M*N*K of every GEMM launch of one step / summed CUDA-event durations of those launches.
          `rooflines` carries the same object for the other kernels north_star names: tcgen05 attention (tensor),
          the decode step in production mode (CUDA graph; HBM: weights + live KV per step / measured step time) and
          the fused assemble + RoI replay (HBM).
  cpu_baseline : the oracle port (fp32, all host cores) on ONE full request of the same workload, no extrapolation
  gpu_lib_baseline : the oracle modules in bf16 on the same B200 (cuBLAS + SDPA + torchvision roi_align, batch 1 per
          request like the reference) and per-kernel A/B against torch.matmul / F.scaled_dot_product_attention

`--impl reference` times the reference's CPU implementation of the path (the oracle port; the reference itself is
pure Python over timm/transformers and cannot be installed offline): min(steps, 2) REAL full requests of config 1
(T=17, S=4672, 128 tokens, fp32, all host cores), one request per step as the reference runs them.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# The pipelined mode keeps a decode stream and an encode stream busy at the same time; with the default 8 hardware
# work queues the two can land on the same queue and serialise behind each other's pre-enqueued launches.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="gar-1b")
    ap.add_argument("--batch", type=int, default=8, help="requests per GPU per step")
    ap.add_argument("--masks", type=int, default=1)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--max-num-tiles", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-inputs", default="raw", choices=["raw", "tiles"],
                    help="what the e2e arm ships from pinned host memory: raw = uint8 image + id map (preprocessing on "
                         "the device), tiles = host-preprocessed bf16 tile tensors")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel-class event timing step")
    ap.add_argument("--no-lib-baseline", action="store_true", help="skip the same-box torch/cuBLAS/SDPA baseline")
    ap.add_argument("--pipeline", action="store_true",
                    help="overlap decode of batch i with ViT+prefill of batch i+1 (pays off only for many steps)")
    ap.add_argument("--decode-sms", type=int, default=0,
                    help="with --pipeline: SMs of the decode partition (green contexts); 0 = no SM partitioning")
    ap.add_argument("--decode-groups", type=int, default=3,
                    help="consecutive steps whose decode loops run concurrently (GARModel.generate_batches(overlap='decode', "
                         "ways=N): each step is still one batch of --batch requests through its own ViT / prefill / KV cache; "
                         "1 = every step strictly after the previous one")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (pdl, skinny_gemm, decode_splits)")
    return ap.parse_args()


def gemm_traffic_bytes():
    """Mean DRAM bytes (read+write) per tcgen05 GEMM launch of this same command, from the committed ncu pass
    (profiles/r02_gemm_traffic.json; ncu cannot run inside the timed bench).  None if the file is absent."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_gemm_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def kernel_traffic():
    """Per-launch / per-step DRAM bytes of the attention kernel, the decode step and the assemble call, from the committed
    ncu launch list of this same workload (profiles/r02_kernel_traffic.json); {} if the file is absent."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_kernel_traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port): REAL full requests, no extrapolation
# ----------------------------------------------------------------------------------------------
def cpu_full_requests(spec, args, n_requests, warmup=0):
    """Runs the oracle (fp32, all host cores) on `warmup + n_requests` complete requests of the bench workload --
    mask decode, both patch convs, the whole PE-ViT over all T tiles, projector, assemble + RoI replay, prefill over all
    S positions, args.new_tokens greedy steps with a KV cache -- and returns the timed requests' phase seconds."""
    from grasp_any_region_b200.synth import make_request
    from oracle.gar_oracle import build_oracle

    torch.set_num_threads(os.cpu_count())
    oracle = build_oracle(spec, seed=0, dtype=torch.float32)
    out = []
    for i in range(warmup + n_requests):
        req = make_request(spec, seed=i, n_masks=args.masks, max_num_tiles=args.max_num_tiles, dtype=torch.float32)
        with torch.no_grad():
            t0 = time.perf_counter()
            taps = oracle.prefill_embeds(req["pixel_values"], req["global_mask_values"], req["aspect_ratios"],
                                         req["bboxes"], req["input_ids"])
            t1 = time.perf_counter()
            # greedy() = prefill + (new_tokens - 1) decode steps; time the prefill separately through a 1-token call
            lm = oracle.mllm.model.language_model
            S = taps["inputs_embeds"].shape[1]
            caches = [dict() for _ in lm.layers]
            h = lm(taps["inputs_embeds"], torch.arange(S), caches)
            logits = oracle.mllm.lm_head(h[:, -1:, :])
            t2 = time.perf_counter()
            toks = []
            for j in range(args.new_tokens):
                nxt = logits[:, -1].argmax(-1)
                toks.append(int(nxt))
                if j == args.new_tokens - 1:
                    break
                e = lm.embed_tokens(nxt)[:, None, :]
                h = lm(e, torch.tensor([S + j]), caches)
                logits = oracle.mllm.lm_head(h[:, -1:, :])
            t3 = time.perf_counter()
        if i >= warmup:
            out.append({"vit_assemble_s": t1 - t0, "prefill_s": t2 - t1, "decode_s": t3 - t2, "total_s": t3 - t0,
                        "T": int(req["pixel_values"].shape[0]), "S": int(S), "tokens": len(toks)})
    return out


def cpu_sample_text(rows, args):
    r = rows[0]
    return (f"oracle port fp32, {len(rows)} full request(s), no extrapolation: ViT over all {r['T']} tiles + assemble, "
            f"prefill of all {r['S']} positions, {r['tokens']} greedy tokens with KV cache "
            f"(mean ViT {sum(x['vit_assemble_s'] for x in rows) / len(rows):.1f} s, prefill "
            f"{sum(x['prefill_s'] for x in rows) / len(rows):.1f} s, decode {sum(x['decode_s'] for x in rows) / len(rows):.1f} s)")


def run_reference(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a full request is 45-90 s of fp32 CPU work: two timed requests, no warm-up (nothing is lazily compiled on this arm)
    steps, warmup = max(1, min(args.steps, 2)), 0
    rows = cpu_full_requests(spec, args, steps, warmup)
    per_req = sum(r["total_s"] for r in rows) / len(rows)
    val = 1.0 / per_req
    line = {
        "impl": "reference", "metric": metric_name(args), "value": val,
        "unit": "captions/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": per_req * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "requests_per_step": 1,
                   "note": (f"CPU port of the reference path, one full request per step (the reference is batch 1); "
                            f"{steps} timed + {warmup} warm-up real requests (requested --steps {args.steps} --warmup "
                            f"{args.warmup}, bounded so the run ends within minutes)")},
        "cpu_baseline": {"value": val, "unit": "captions/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": cpu_sample_text(rows, args)},
        "e2e": {"value": val, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(args):
    """BASELINE.json's metric for the default workload; secondary configurations name their model / length."""
    return f"region-captions/sec ({args.new_tokens} decode tok) {args.model.upper()}"


def workload_name(args):
    from grasp_any_region_b200.synth import tile_grid_for

    tw, th = tile_grid_for((1024, 1024), args.max_num_tiles)
    return (f"{args.model} bf16, batch={args.batch} x (1024x1024 image -> {tw * th}+1 tiles, "
            f"{args.masks} mask), {args.new_tokens}-token greedy decode")


# ----------------------------------------------------------------------------------------------
# Same-box GPU library baseline (BASELINE.md §3 "B-gpu-lib"): what every hand-written kernel must beat
# ----------------------------------------------------------------------------------------------
def _cuda_ms(fn, iters=5, warm=2):
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


def gpu_lib_baseline(spec, args, model, reqs_dev, dev, n_requests=2):
    """(1) The oracle modules with the bench's weights in bf16 on this GPU -- cuBLAS GEMMs, SDPA (flash) attention,
    torchvision roi_align, torch.cat KV cache, one request at a time like the reference -- timed per phase.
    (2) Kernel A/B at the production shapes: our tcgen05 GEMM vs torch.matmul, our attention vs
    F.scaled_dot_product_attention (flash and cuDNN backends), our batched decode step vs one oracle decode step at
    batch 8 (cuBLAS + SDPA over a torch.cat cache)."""
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel

    from grasp_any_region_b200 import ops
    from oracle.gar_oracle import GAROracle

    bf = torch.bfloat16
    with torch.device(dev):
        o = GAROracle(spec)
    o.load_state_dict({k: v.float() for k, v in model.state_dict().items()})
    o = o.to(bf).eval()
    lm = o.mllm.model.language_model
    NT = args.new_tokens
    rows = []
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    with torch.no_grad():
        for i in range(1 + n_requests):
            r = reqs_dev[i % len(reqs_dev)]
            e = [ev() for _ in range(4)]
            torch.cuda.synchronize()
            e[0].record()
            taps = o.prefill_embeds(r["pixel_values"], r["global_mask_values"], r["aspect_ratios"], r["bboxes"],
                                    r["input_ids"])
            e[1].record()
            S = taps["inputs_embeds"].shape[1]
            caches = [dict() for _ in lm.layers]
            h = lm(taps["inputs_embeds"], torch.arange(S, device=dev), caches)
            logits = o.mllm.lm_head(h[:, -1:, :])
            e[2].record()
            for j in range(NT - 1):
                nxt = logits[:, -1].argmax(-1)
                h = lm(lm.embed_tokens(nxt)[:, None, :], torch.tensor([S + j], device=dev), caches)
                logits = o.mllm.lm_head(h[:, -1:, :])
            e[3].record()
            torch.cuda.synchronize()
            if i > 0:
                rows.append([e[k].elapsed_time(e[k + 1]) for k in range(3)])
    vit, pre, dec = (sum(r[k] for r in rows) / len(rows) for k in range(3))
    out = {"what": "oracle modules, bf16, this GPU: cuBLAS + SDPA(flash) + torchvision roi_align + torch.cat KV cache, "
                   "batch 1 per request (as the reference runs), eager PyTorch",
           "requests_timed": len(rows), "vit_assemble_ms_per_request": vit, "prefill_ms_per_request": pre,
           "decode_ms_per_token": dec / max(1, NT - 1), "captions_per_s": 1e3 / (vit + pre + dec)}

    # ---- kernel A/B at production shapes
    v, t = spec.vision, spec.text
    B = len(reqs_dev)
    T = reqs_dev[0]["pixel_values"].shape[0]
    S = reqs_dev[0]["input_ids"].shape[1]
    Mv, Ml = B * T * v.tokens, B * S
    dq, qkvw = v.embed_dim, (t.num_heads + 2 * t.num_kv_heads) * t.head_dim
    shapes = [("vit.qkv", Mv, 3 * dq, v.embed_dim), ("vit.proj", Mv, v.embed_dim, dq), ("vit.fc1", Mv, v.mlp_dim, v.embed_dim),
              ("vit.fc2", Mv, v.embed_dim, v.mlp_dim), ("llm.qkv", Ml, qkvw, t.hidden_size),
              ("llm.o", Ml, t.hidden_size, t.num_heads * t.head_dim), ("llm.gate_up", Ml, 2 * t.intermediate_size, t.hidden_size),
              ("llm.down", Ml, t.hidden_size, t.intermediate_size)]
    gemms = {}
    for name, M, N, K in shapes:
        if N % 32 or K % 64:
            continue
        a = torch.randn((M, K), device=dev, dtype=bf) * 0.05
        w = torch.randn((N, K), device=dev, dtype=bf) * 0.05
        c = torch.empty((M, N), device=dev, dtype=bf)
        ours = _cuda_ms(lambda: ops.gemm(a, w, out=c), 4, 1)
        lib = _cuda_ms(lambda: torch.matmul(a, w.t(), out=c), 4, 1)
        fl = 2.0 * M * N * K / 1e9
        gemms[name] = {"M": M, "N": N, "K": K, "ours_ms": ours, "torch_matmul_ms": lib, "ours_TFLOPs": fl / ours,
                       "torch_TFLOPs": fl / lib}
        del a, w, c
    out["gemm_vs_torch_matmul"] = gemms

    def sdpa_ms(q, k, vv, causal, backend):
        try:
            with sdpa_kernel(backend):
                return _cuda_ms(lambda: F.scaled_dot_product_attention(q, k, vv, is_causal=causal), 4, 1)
        except Exception as ex:  # a backend that refuses the shape is reported, not fatal
            return f"unavailable: {type(ex).__name__}"

    att = {}
    if v.head_dim == 64:
        nt, H = v.tokens, v.num_heads
        qkv = torch.randn((B * T * nt, 3 * dq), device=dev, dtype=bf)
        ours = _cuda_ms(lambda: ops.attn_fwd_packed_qkv(qkv, B * T, nt, H, v.head_dim), 4, 1)
        q, k, vv = (x.contiguous() for x in qkv.view(B * T, nt, 3, H, v.head_dim).permute(2, 0, 3, 1, 4).unbind(0))
        fl = 4.0 * B * T * nt * nt * dq / 1e9
        att["vit"] = {"shape": [B * T, H, nt, v.head_dim], "ours_ms": ours, "ours_TFLOPs": fl / ours,
                      "sdpa_flash_ms": sdpa_ms(q, k, vv, False, SDPBackend.FLASH_ATTENTION),
                      "sdpa_cudnn_ms": sdpa_ms(q, k, vv, False, SDPBackend.CUDNN_ATTENTION)}
        del qkv, q, k, vv
    hq, hkv, hd = t.num_heads, t.num_kv_heads, t.head_dim
    msq = ((S + 127) // 128) * 128
    qb = torch.randn((S, qkvw), device=dev, dtype=bf)
    kc = torch.randn((hkv, msq, hd), device=dev, dtype=bf)
    vc = torch.randn((hkv, msq, hd), device=dev, dtype=bf)
    ours = _cuda_ms(lambda: ops.attn_fwd_cache(qb, qkvw, kc, vc, hq, hkv, hd, S, msq), 4, 1)
    q = qb[:, : hq * hd].reshape(S, hq, hd).transpose(0, 1)[None].contiguous()
    k = kc[:, :S].repeat_interleave(hq // hkv, 0)[None].contiguous()
    vv = vc[:, :S].repeat_interleave(hq // hkv, 0)[None].contiguous()
    fl = 2.0 * S * S * hd * hq / 1e9
    att["prefill_causal_one_request"] = {"shape": [1, hq, S, hd], "ours_ms": ours, "ours_TFLOPs": fl / ours,
                                         "sdpa_flash_ms": sdpa_ms(q, k, vv, True, SDPBackend.FLASH_ATTENTION),
                                         "sdpa_cudnn_ms": sdpa_ms(q, k, vv, True, SDPBackend.CUDNN_ATTENTION)}
    del qb, kc, vc, q, k, vv
    out["attention_vs_sdpa"] = att

    # one oracle decode step at batch B over a context of S (torch.cat cache growth included, as DynamicCache does)
    with torch.no_grad():
        caches = [{"k": torch.randn((B, hkv, S, hd), device=dev, dtype=bf), "v": torch.randn((B, hkv, S, hd), device=dev, dtype=bf)}
                  for _ in lm.layers]
        tok = torch.randint(0, 1000, (B,), device=dev)

        def step():
            for c in caches:  # keep the context length fixed across timed iterations
                c["k"], c["v"] = c["k"][:, :, :S], c["v"][:, :, :S]
            h = lm(lm.embed_tokens(tok)[:, None, :], torch.tensor([S], device=dev), caches)
            return o.mllm.lm_head(h[:, -1:, :]).argmax(-1)

        out["decode_step_batch%d_ms" % B] = {"torch_cublas_sdpa_eager": _cuda_ms(step, 5, 2)}
    del o, caches
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from grasp_any_region_b200.config import PRESETS

    spec = PRESETS[args.model]()
    if args.impl == "reference":
        run_reference(args, spec)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    dist = None
    if world > 1:
        import datetime

        import torch.distributed as dist

        # short collective timeout: a rank that dies or wedges must fail the run in minutes, not hang it
        dist.init_process_group("nccl", device_id=torch.device(dev), timeout=datetime.timedelta(seconds=240))

    from grasp_any_region_b200 import GARConfig, GARModel
    from grasp_any_region_b200.synth import make_raw_request, make_request

    def note(msg):
        if os.environ.get("GAR_BENCH_VERBOSE") or world > 1:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    peaks = load_peaks()
    B, NT = args.batch, args.new_tokens
    model = GARModel(GARConfig.from_spec(spec)).random_init_(0).to(dev)
    reqs_host = [make_request(spec, seed=1000 * rank + i, n_masks=args.masks, max_num_tiles=args.max_num_tiles,
                              dtype=torch.bfloat16) for i in range(B)]
    S = max(r["input_ids"].shape[1] for r in reqs_host)
    T = reqs_host[0]["pixel_values"].shape[0]
    model.configure_engine(max_tiles=B * T, max_seq=((S + NT + 127) // 128) * 128, max_batch=2 * B if args.pipeline else B,
                           max_new_tokens=NT, max_prefill_rows=B * S)
    note("model initialised")
    eng = model.engine()
    note("engine ready")
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    partition = None
    if args.pipeline and args.decode_sms > 0:
        partition = ({"encode_sms": eng.partition_sms[0], "decode_sms": eng.partition_sms[1]}
                     if eng.enable_partition(args.decode_sms) else {"unavailable": getattr(eng, "partition_error", "?")})

    def pin(r):
        o = dict(r)
        for k in ("pixel_values", "global_mask_values", "input_ids", "attention_mask"):
            o[k] = r[k].pin_memory()
        return o

    if args.e2e_inputs == "raw":
        reqs_pinned = [make_raw_request(spec, seed=1000 * rank + i, n_masks=args.masks, max_num_tiles=args.max_num_tiles)
                       for i in range(B)]
        assert reqs_pinned[0]["input_ids"].shape[1] == S
        e2e_keys = ("image_u8", "idmap_u8", "input_ids")
    else:
        reqs_pinned = [pin(r) for r in reqs_host]
        e2e_keys = ("pixel_values", "global_mask_values", "input_ids")
    reqs_dev = []
    for r in reqs_host:
        o = dict(r)
        for k in ("pixel_values", "global_mask_values", "input_ids"):
            o[k] = r[k].to(dev)
        reqs_dev.append(o)
    h2d = sum(r[k].numel() * r[k].element_size() for r in reqs_pinned for k in e2e_keys)
    d2h = B * NT * 4

    def run_steps(batch_list, steps, gather=True):
        """`steps` batches through the public API (GARModel.generate_batch; --pipeline: generate_batches).
        gather=False for the rank-0-only diagnostic passes: a collective there would wait for ranks that never come."""
        raw = "image_u8" in batch_list[0]
        if not args.pipeline and args.decode_groups > 1 and steps > 1:
            # `decode_groups` consecutive steps: encodes back to back, decode loops concurrently (one engine twin per
            # step in flight, shared weights); the next round's copies / device preprocessing under the current decode
            toks = model.generate_batches([batch_list] * steps, max_new_tokens=NT, eos_token_id=None, overlap="decode",
                                          ways=args.decode_groups)
            if world > 1 and gather:  # the path's one collective: all-gather of the decoded ids (north_star)
                from grasp_any_region_b200.parallel import gather_tokens, pack_tokens

                for step_toks in toks:
                    gather_tokens(pack_tokens(step_toks, B, NT, dev), world * B)
        elif args.pipeline or (raw and steps > 1):
            # raw (uint8) requests: the next step's host -> device copies and device preprocessing run under the current
            # step's decode (generate_batches(overlap="prepare")); every step's copies stay inside the timed region
            toks = model.generate_batches([batch_list] * steps, max_new_tokens=NT, eos_token_id=None,
                                          overlap="encode" if args.pipeline else "prepare")
            if world > 1 and gather:  # the path's one collective: all-gather of the decoded ids (north_star)
                from grasp_any_region_b200.parallel import gather_tokens, pack_tokens

                for step_toks in toks:
                    gather_tokens(pack_tokens(step_toks, B, NT, dev), world * B)
        elif world > 1 and gather:
            # public sharded API: this rank's shard in, every rank's captions out (one NCCL all-gather per step)
            toks = [model.generate_sharded(batch_list, max_new_tokens=NT, eos_token_id=None, already_sharded=True)
                    for _ in range(steps)]
        else:
            toks = [model.generate_batch(batch_list, max_new_tokens=NT, eos_token_id=None) for _ in range(steps)]
        return toks

    def step_resident(steps=1):
        return run_steps(reqs_dev, steps)

    def step_e2e(steps=1):
        return run_steps(reqs_pinned, steps)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA events on the engine stream (where every kernel is launched), max over ranks."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(eng.stream)
        fn(steps)
        eng.stream.wait_stream(eng.decode_stream)
        b.record(eng.stream)
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # >= 2 so both KV-slot groups (--pipeline) / every engine twin (--decode-groups) capture their decode graph untimed
    step_resident(max(args.warmup, 2, 1 if args.pipeline else args.decode_groups))
    note("warm-up done")
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    ms = timed(step_resident, args.steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    pipe_stats = getattr(eng, "pipeline_stats", None) if args.pipeline else None
    if os.environ.get("GAR_BENCH_TRACE"):
        eng.trace = []
        timed(step_resident, 3)
        pipe_stats = dict(pipe_stats or {}, trace=[(l, round(v, 1)) for l, v in eng.trace_ms()],
                          pipeline_stats_traced=getattr(eng, "pipeline_stats", None))
        eng.trace = None
    note(f"timed region done: {ms:.1f} ms")
    value = world * B * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        # untimed: as many steps as one round holds, so that every engine twin has seen raw requests and the caching
        # allocator already owns a round's worth of preprocessing buffers (a cold round pays ~150 ms of cudaMalloc)
        step_e2e(max(2, 1 if args.pipeline else args.decode_groups))
        ms_e = timed(step_e2e, args.steps)
        e2e = {"value": world * B * args.steps / (ms_e / 1e3), "unit": "captions/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / args.steps,
               "inputs": ("uint8 image + uint8 region-id map from pinned host memory; resize / tiling / normalisation on "
                          "the device inside the timed region; a round's copies and preprocessing run under the previous "
                          "round's decode (GARModel.generate_batches(overlap='decode' / 'prepare')), the first round's "
                          "in front of its first encode"
                          if args.e2e_inputs == "raw" else
                          "host-preprocessed bf16 tile tensors from pinned host memory")}

    # ---- phase split of one more step (events on the engine stream between the phases)
    phases = None
    if rank == 0:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record(eng.stream)
        model._stage(eng, 0, reqs_dev)
        ev[1].record(eng.stream)
        eng.generate(B, NT, None)
        ev[2].record(eng.stream)
        torch.cuda.synchronize()
        enc, dec = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        phases = {"vit_assemble_prefill_ms_per_request": enc / B, "decode_ms_total": dec,
                  "decode_ms_per_token_step": dec / max(1, NT - 1), "batch": B}
        ways = args.decode_groups if (not args.pipeline and args.steps > 1) else 1
        if ways > 1:
            # the production decode: `ways` groups (one engine twin each) decoded concurrently, all encoded beforehand
            import threading

            engines = [eng] + eng.twins(ways - 1)
            for e in engines:
                model._stage(e, 0, reqs_dev)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(eng.stream)
            ths = [threading.Thread(target=e.generate, args=(B, NT, None)) for e in engines[1:]]
            for th in ths:
                th.start()
            eng.generate(B, NT, None)
            for th in ths:
                th.join()
            for e in engines[1:]:
                eng.stream.wait_stream(e.stream)
            c1.record(eng.stream)
            torch.cuda.synchronize()
            cms = c0.elapsed_time(c1)
            phases["decode_concurrent"] = {"groups": ways, "ms_total": cms, "ms_per_group": cms / ways,
                                           "ms_per_token_step_per_group": cms / ways / max(1, NT - 1)}

    # ---- per-kernel-class timing of ONE more step (CUDA events around every launch of the class)
    roof, rooflines, stages = None, None, None
    if rank == 0 and not args.no_profile:
        eng.profile(True)
        toks_prof = model.generate_batch(reqs_dev, max_new_tokens=NT, eos_token_id=None)
        prof = eng.profile_read()
        eng.profile(False)
        del toks_prof
        g_ms, g_fl, g_n = prof["gemm"]
        a_ms, a_fl, a_n = prof["attn"]
        s_ms, s_by, s_n = prof["assemble"]
        step_ms = ms / args.steps
        tf = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05)", "achieved": tf,
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"],
                "peak_source": f"{peaks['source']} sustained bf16",
                # the committed ncu pass is of the default workload only
                "traffic": gemm_traffic_bytes() if (args.model, B, args.max_num_tiles) == ("gar-1b", 8, 16) else None,
                "launches_per_step": g_n, "ms_per_step": g_ms, "share_of_step": g_ms / step_ms}
        atf = a_fl / (a_ms * 1e-3) / 1e12 if a_ms > 0 else 0.0
        # decode step in PRODUCTION mode (CUDA graph replay): algorithmic bytes = every weight once per step + the live
        # K and V prefixes of every sequence; time = the event-timed decode phase above / steps
        t = spec.text
        w_bytes = 2.0 * (t.num_layers * ((t.num_heads + 2 * t.num_kv_heads) * t.head_dim * t.hidden_size
                                         + t.hidden_size * t.num_heads * t.head_dim
                                         + 3 * t.intermediate_size * t.hidden_size) + eng.vocab_padded * t.hidden_size)
        kv_bytes = 2.0 * 2 * t.num_kv_heads * t.head_dim * t.num_layers * B * (S + NT / 2.0)
        d_alone = phases["decode_ms_per_token_step"]
        conc = phases.get("decode_concurrent")
        # production mode: `groups` decode chains in flight -> per group-step time = concurrent total / groups / steps
        d_ms = conc["ms_per_token_step_per_group"] if conc else d_alone
        d_gbs = (w_bytes + kv_bytes) / (d_ms * 1e-3) / 1e9
        asm_gbs = s_by / (s_ms * 1e-3) / 1e9 if s_ms > 0 else 0.0
        default_workload = (args.model, B, args.max_num_tiles, args.masks) == ("gar-1b", 8, 16, 1)
        kt = kernel_traffic() if default_workload else {}
        rooflines = [
            roof,
            {"bound": "tensor", "kernel": "attn_fwd_persist_kernel (tcgen05, ViT + causal prefill)", "achieved": atf,
             "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": atf / peaks["tf_sustained"],
             "launches_per_step": a_n, "ms_per_step": a_ms, "share_of_step": a_ms / step_ms,
             "traffic": kt.get("attention", {}).get("dram_bytes_per_launch")},
            {"bound": "hbm", "kernel": "decode step (CUDA graph: gemv_decode + decode_attn + argmax kernels)",
             "achieved": d_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": d_gbs / peaks["hbm"],
             "bytes_per_step": w_bytes + kv_bytes, "ms_per_launch": d_ms, "launches_per_step": NT - 1,
             "share_of_step": d_ms * (NT - 1) / step_ms,
             "traffic": kt.get("decode_step", {}).get("dram_bytes_per_step"),  # per decode step of ONE group, like bytes_per_step
             "alone": {"ms_per_launch": d_alone, "GB/s": (w_bytes + kv_bytes) / (d_alone * 1e-3) / 1e9,
                       "frac": (w_bytes + kv_bytes) / (d_alone * 1e-3) / 1e9 / peaks["hbm"]},
             "mode": (f"production (graph replay): {conc['groups']} independent groups of {B} sequences decoded concurrently, "
                      "every group streams the weights and its own K/V once per step (bytes_per_step is ONE group's); "
                      "ms_per_launch = concurrent total / groups / steps (phases.decode_concurrent); `alone` = one group "
                      "by itself (phases.decode_ms_per_token_step)") if conc else
                     "production (graph replay), from phases.decode_ms_per_token_step"},
            {"bound": "hbm", "kernel": "replay_rows + gather_rows (embed gather + image scatter + RoI replay; one call, "
                                       "the two grids side by side)",
             "achieved": asm_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": asm_gbs / peaks["hbm"],
             "bytes_per_step": s_by, "ms_per_step": s_ms, "launches_per_step": s_n, "share_of_step": s_ms / step_ms,
             "traffic": kt.get("assemble", {}).get("dram_bytes_per_call"),
             "note": "the row plan (plan_rows_kernel: ids only, no feature bytes) runs before the ViT so that a malformed "
                     f"prompt raises first; timed separately: {prof['plan'][0] * 1e3:.1f} us per step"},
        ]
        stages = {
            "attn_tcgen05": {"ms": a_ms, "TFLOP/s": atf, "launches": a_n},
            "assemble_replay": {"ms": s_ms, "GB/s": asm_gbs, "frac_hbm": asm_gbs / peaks["hbm"], "launches": s_n,
                                "plan_ms": prof["plan"][0]},
            "decode_attn_ungraphed": {"ms": prof["decode_attn"][0], "launches": prof["decode_attn"][2]},
            "decode_gemv_ungraphed": {"ms": prof["gemv"][0], "launches": prof["gemv"][2]},
        }

    lib = None
    if rank == 0 and world == 1 and not args.no_lib_baseline:
        try:
            lib = gpu_lib_baseline(spec, args, model, reqs_dev, dev)
            if phases is not None:
                lib["decode_step_batch%d_ms" % B]["ours_graph"] = phases["decode_ms_per_token_step"]
                lib["ours_vit_assemble_prefill_ms_per_request"] = phases["vit_assemble_prefill_ms_per_request"]
        except Exception as ex:  # a baseline must not take the bench line down
            lib = {"failed": f"{type(ex).__name__}: {ex}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rows = cpu_full_requests(spec, args, 1)
            cpu = {"value": 1.0 / rows[0]["total_s"], "unit": "captions/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": cpu_sample_text(rows, args)}
        except Exception as ex:  # the baseline must not take the bench line down
            cpu = {"value": None, "unit": "captions/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": "captions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "requests_per_gpu_per_step": B, "tiles": T, "prompt_len": S,
                       "l2": "inputs+weights per step (3.2 GB) exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"dp{world} (independent requests per rank, one all-gather of token ids)",
                       "pipeline": bool(args.pipeline), "sm_partition": partition,
                       "decode_groups": (f"{args.decode_groups}: the steps are taken {args.decode_groups} at a time -- encoded "
                                         "back to back, then their decode loops run concurrently, each step's batch in its own "
                                         "KV cache / workspaces / stream / CUDA graph (engine twins, one shared copy of the "
                                         "weights); per-request tokens identical to the one-at-a-time run"
                                         if (not args.pipeline and args.decode_groups > 1 and args.steps > 1) else 1),
                       "encode": "the B requests of a step share one ViT pass and one prefill pass (equal shapes); decode batched"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "rooflines": rooflines,
            "stages": stages, "phases": phases, "pipeline_stats": pipe_stats, "cpu_baseline": cpu, "gpu_lib_baseline": lib,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
