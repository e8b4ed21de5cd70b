#!/usr/bin/env python
"""Headline benchmark: region-captions/sec (128 decode tokens), GAR-1B bf16, on N B200s.

Workload (BASELINE.json configs[1]): 8 independent requests per GPU, each a synthetic 1024x1024 image
(17 tiles of 448^2) + 1 mask, S = 4672 prompt positions, 128 greedy tokens, random-init GAR-1B.
One "step" = that batch through the whole hot path: per request ViT encode -> assemble/RoI replay ->
Llama prefill; then one batched 128-token decode.

  value : captions/s with the requests already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : same metric through GARModel.generate_batch() from pinned HOST buffers (H2D of every
          request and D2H of the tokens inside the timed region)
  roofline : dominant kernel = the tcgen05 GEMM (gemm_bf16_tn_kernel); achieved = summed algorithmic
          2*M*N*K of every GEMM launch of one step / summed CUDA-event durations of those launches
  cpu_baseline : the oracle port (fp32, all host cores) on a bounded sample of the same request

`--impl reference` times the reference's CPU implementation of the path (the oracle port; the
reference itself is pure Python over timm/transformers and cannot be installed offline) on the same
config and prints the same JSON line with "impl": "reference".
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
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel-class event timing step")
    ap.add_argument("--pipeline", action="store_true",
                    help="overlap decode of batch i with ViT+prefill of batch i+1 (pays off only for many steps)")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (pdl, skinny_gemm, decode_splits)")
    return ap.parse_args()


def gemm_traffic_bytes():
    """Mean DRAM bytes (read+write) per tcgen05 GEMM launch of this same command, from the committed ncu pass
    (profiles/r01_gemm_traffic.json; ncu cannot run inside the timed bench).  None if the file is absent."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_gemm_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


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
# CPU arm (oracle port): bounded sample of one request, extrapolated to the full request
# ----------------------------------------------------------------------------------------------
def cpu_request_seconds(spec, args, vit_tiles, prefill_tokens, decode_steps, repeat=1):
    """Times the oracle (fp32, all cores) on: `vit_tiles` of the T tiles, the first `prefill_tokens` of the
    S prompt positions, `decode_steps` decode steps; scales each linearly to the full request."""
    from grasp_any_region_b200.synth import make_request
    from oracle.gar_oracle import build_oracle

    torch.set_num_threads(os.cpu_count())
    oracle = build_oracle(spec, seed=0, dtype=torch.float32)
    req = make_request(spec, seed=0, n_masks=args.masks, max_num_tiles=args.max_num_tiles, dtype=torch.float32)
    T = req["pixel_values"].shape[0]
    S = req["input_ids"].shape[1]
    lm = oracle.mllm.model.language_model
    out = []
    for _ in range(repeat):
        with torch.no_grad():
            t0 = time.perf_counter()
            me = oracle.mask_embeds(req["global_mask_values"][:vit_tiles])
            oracle.image_features(req["pixel_values"][:vit_tiles], me)
            t_vit = (time.perf_counter() - t0) * T / vit_tiles
            emb = lm.embed_tokens(req["input_ids"][:, :prefill_tokens])
            caches = [dict() for _ in lm.layers]
            t0 = time.perf_counter()
            h = lm(emb, torch.arange(prefill_tokens), caches)
            oracle.mllm.lm_head(h[:, -1:])
            # linear layers scale with S, causal attention with S^2: extrapolate both parts
            t_pref_sample = time.perf_counter() - t0
            t_prefill = t_pref_sample * (S / prefill_tokens)
            t0 = time.perf_counter()
            for i in range(decode_steps):
                e = lm.embed_tokens(torch.tensor([[1 + i]]))
                h = lm(e, torch.tensor([prefill_tokens + i]), caches)
                oracle.mllm.lm_head(h[:, -1:])
            t_dec = (time.perf_counter() - t0) / decode_steps * args.new_tokens
        out.append(t_vit + t_prefill + t_dec)
    sample = (f"oracle port fp32: ViT on {vit_tiles}/{T} tiles, prefill of {prefill_tokens}/{S} positions "
              f"(linear extrapolation; understates the S^2 attention share), {decode_steps} decode steps at "
              f"context {prefill_tokens} -> scaled to one full request of {args.new_tokens} tokens")
    return out, sample


def run_reference(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, sample = cpu_request_seconds(spec, args, vit_tiles=1, prefill_tokens=512, decode_steps=2,
                                        repeat=args.warmup + args.steps)
    timed = times[args.warmup:]
    per_req = sum(timed) / len(timed)
    val = 1.0 / per_req
    line = {
        "impl": "reference", "metric": metric_name(args), "value": val,
        "unit": "captions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_req * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "model": args.model, "note": "CPU port of the reference path; each step = one bounded-sample request"},
        "cpu_baseline": {"value": val, "unit": "captions/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(args):
    """BASELINE.json's metric for the default workload; secondary configurations name their model / length."""
    return f"region-captions/sec ({args.new_tokens} decode tok) {args.model.upper()}"


def workload_name(args):
    return (f"{args.model} bf16, batch={args.batch} x (1024x1024 image -> {args.max_num_tiles}+1 tiles, "
            f"{args.masks} mask), {args.new_tokens}-token greedy decode")


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
    from grasp_any_region_b200.synth import make_request

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

    def pin(r):
        o = dict(r)
        for k in ("pixel_values", "global_mask_values", "input_ids", "attention_mask"):
            o[k] = r[k].pin_memory()
        return o

    reqs_pinned = [pin(r) for r in reqs_host]
    reqs_dev = []
    for r in reqs_host:
        o = dict(r)
        for k in ("pixel_values", "global_mask_values", "input_ids"):
            o[k] = r[k].to(dev)
        reqs_dev.append(o)
    h2d = sum(r[k].numel() * r[k].element_size() for r in reqs_host
              for k in ("pixel_values", "global_mask_values", "input_ids"))
    d2h = B * NT * 4

    def run_steps(batch_list, steps, gather=True):
        """`steps` batches through the public API (GARModel.generate_batch; --pipeline: generate_batches).
        gather=False for the rank-0-only diagnostic passes: a collective there would wait for ranks that never come."""
        if args.pipeline:
            toks = model.generate_batches([batch_list] * steps, max_new_tokens=NT, eos_token_id=None)
        else:
            toks = [model.generate_batch(batch_list, max_new_tokens=NT, eos_token_id=None) for _ in range(steps)]
        if world > 1 and gather:  # the path's one collective: all-gather of the decoded ids (north_star)
            flat = [t + [-1] * (NT - len(t)) for step_toks in toks for t in step_toks]
            t = torch.tensor(flat, dtype=torch.int32, device=dev)
            buf = torch.empty((world * t.shape[0], NT), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(buf, t)
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

    step_resident(max(args.warmup, 2))  # >= 2 so both KV-slot groups capture their decode graph
    note("warm-up done")
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    ms = timed(step_resident, args.steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    note(f"timed region done: {ms:.1f} ms")
    value = world * B * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        step_e2e(2)
        ms_e = timed(step_e2e, args.steps)
        e2e = {"value": world * B * args.steps / (ms_e / 1e3), "unit": "captions/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / args.steps}

    # ---- phase split of one more step (events on the engine stream between the phases)
    phases = None
    if rank == 0:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record(eng.stream)
        if not model._stage_requests_batched(eng, 0, reqs_dev):
            with torch.cuda.stream(eng.stream):
                for slot, r in enumerate(reqs_dev):
                    eng.request(slot, r["pixel_values"], r["global_mask_values"], r["input_ids"][0], r["bboxes"][0],
                                r["aspect_ratios"][0].tolist())
        ev[1].record(eng.stream)
        eng.generate(B, NT, None)
        ev[2].record(eng.stream)
        torch.cuda.synchronize()
        enc, dec = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        phases = {"vit_assemble_prefill_ms_per_request": enc / B, "decode_ms_total": dec,
                  "decode_ms_per_token_step": dec / max(1, NT - 1), "batch": B}

    # ---- per-kernel-class timing of ONE more step (CUDA events around every launch of the class)
    roof, stages = None, None
    if rank == 0 and not args.no_profile:
        eng.profile(True)
        run_steps(reqs_dev, 1, gather=False)
        prof = eng.profile_read()
        eng.profile(False)
        g_ms, g_fl, g_n = prof["gemm"]
        a_ms, a_fl, a_n = prof["attn"]
        s_ms, s_by, s_n = prof["assemble"]
        tf = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05)", "achieved": tf,
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"],
                "peak_source": f"{peaks['source']} sustained bf16",
                # the committed ncu pass is of the default workload only
                "traffic": gemm_traffic_bytes() if (args.model, B, args.max_num_tiles) == ("gar-1b", 8, 16) else None,
                "launches_per_step": g_n, "ms_per_step": g_ms, "share_of_step": g_ms / (ms / args.steps)}
        stages = {
            "attn_tcgen05": {"ms": a_ms, "TFLOP/s": a_fl / (a_ms * 1e-3) / 1e12 if a_ms > 0 else 0.0, "launches": a_n},
            "assemble_replay": {"ms": s_ms, "GB/s": s_by / (s_ms * 1e-3) / 1e9 if s_ms > 0 else 0.0,
                                "frac_hbm": (s_by / (s_ms * 1e-3) / 1e9) / peaks["hbm"] if s_ms > 0 else 0.0,
                                "launches": s_n},
            "decode_attn": {"ms": prof["decode_attn"][0], "launches": prof["decode_attn"][2]},
            "decode_gemv": {"ms": prof["gemv"][0], "launches": prof["gemv"][2],
                            "GB/s": prof["gemv"][1] / (prof["gemv"][0] * 1e-3) / 1e9 if prof["gemv"][0] > 0 else 0.0,
                            "frac_hbm": (prof["gemv"][1] / (prof["gemv"][0] * 1e-3) / 1e9) / peaks["hbm"] if prof["gemv"][0] > 0 else 0.0},
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            times, sample = cpu_request_seconds(spec, args, vit_tiles=2, prefill_tokens=1024, decode_steps=3)
            cpu = {"value": 1.0 / times[0], "unit": "captions/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": sample}
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
                       "pipeline": bool(args.pipeline),
                       "encode": "the B requests of a step share one ViT pass and one prefill pass (equal shapes); decode batched"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "stages": stages, "phases": phases,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
