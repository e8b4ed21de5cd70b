"""Full-size GPU parity gate (BASELINE.json configs 2-4): GAR-1B (T=17, 1 and 4 masks, S ~ 4.7k / 5.6k) and GAR-8B
(T=5 and T=17) through the engine (C ABI) against the oracle restatement run on the SAME B200 in fp32 ("truth", on the
same bf16-rounded weights) and in bf16 (cuBLAS + SDPA: the reference's own GPU arithmetic).

north_star asks for "logits within 1e-3 abs bf16".  One bf16 ulp is 2^-8 = 3.9e-3 for |logit| in [0.5, 1) and 7.8e-3
in [1, 2), so two correct bf16 pipelines with different fp32 summation orders cannot agree to 1e-3; what CAN be
demanded, and is gated here, is that our logits are as close to the fp32 truth as the reference's bf16 path is:
    max|ours - fp32|  <=  1.5 * max|oracle_bf16 - fp32| + 2e-3     (image features: same, relative to their scale)
plus, reported for the record (gpurun_out/r02_parity_full.txt -> profiles/): max|ours - oracle_bf16|, the fraction of
logits that differ from the bf16 oracle by more than one bf16 ulp, top-5 overlap, and the first greedy step at which
the token streams part (legitimate only where the fp32 top-2 margin is below the measured bf16 noise).

Integer / index work is bit-exact: inputs_embeds (embed gather, image-token scatter, RoI replay rows) `torch.equal`.
Weights: random init with a unit-scale embedding table so that the next token depends on the current one and greedy
sequences do not collapse to a constant (a 0.02-scale table makes h_final ~ independent of the input token).
EOS: stop ids that fire mid-sequence, a list of stop ids, and one that never fires; early stop is asserted through the
number of decode steps the engine actually launched.
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from grasp_any_region_b200.config import gar_1b, gar_8b  # noqa: E402
from grasp_any_region_b200.engine import Engine  # noqa: E402
from grasp_any_region_b200.synth import make_request  # noqa: E402
from oracle.gar_oracle import GAROracle  # noqa: E402

DEV = "cuda:0"
bf = torch.bfloat16
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_parity_full.txt")


def _report(line):
    print(line)
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(line + "\n")


def _gpu_oracles(spec, seed=0, embed_std=1.0):
    """fp32 oracle initialised ON the GPU (an 8B fp32 init on the host would take minutes), weights rounded to bf16;
    returns (oracle_fp32, oracle_bf16, bf16 state dict)."""
    with torch.device(DEV):
        o32 = GAROracle(spec)
    g = torch.Generator(device=DEV).manual_seed(seed)
    with torch.no_grad():
        for name, p in o32.named_parameters():
            if name.endswith("gamma_1") or name.endswith("gamma_2"):
                p.fill_(0.1)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g, device=DEV))
            elif name.endswith("embed_tokens.weight"):
                p.copy_(embed_std * torch.randn(p.shape, generator=g, device=DEV))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=DEV))
            p.copy_(p.to(bf).float())
    o32 = o32.eval()
    sd = {k: v.to(bf) for k, v in o32.state_dict().items()}
    with torch.device(DEV):
        o16 = GAROracle(spec)
    o16.load_state_dict({k: v.float() for k, v in sd.items()})
    o16 = o16.to(bf).eval()
    return o32, o16, sd


def _ulp_bf16(x):
    return torch.exp2(torch.floor(torch.log2(x.abs().clamp_min(1e-30))) - 7.0)


def _compare(name, ours, ref16, ref32):
    ours, ref16, ref32 = ours.float(), ref16.float(), ref32.float()
    e_our = (ours - ref32).abs().max().item()
    e_ref = (ref16 - ref32).abs().max().item()
    d = (ours - ref16).abs()
    frac = (d > _ulp_bf16(torch.maximum(ours.abs(), ref16.abs()))).float().mean().item()
    _report(f"  {name}: max|ours-fp32|={e_our:.4g}  max|oracle_bf16-fp32|={e_ref:.4g}  max|ours-oracle_bf16|={d.max().item():.4g}"
            f"  frac(>1 bf16 ulp vs oracle_bf16)={frac:.4f}  scale={ref32.abs().max().item():.3g}")
    return e_our, e_ref


def _first_divergence(a, b):
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return i
    return None if len(a) == len(b) else min(len(a), len(b))


def _run_config(tag, spec, o32, o16, sd, n_masks, max_num_tiles, new_tokens, seed):
    req = make_request(spec, seed=seed, n_masks=n_masks, max_num_tiles=max_num_tiles, dtype=bf, device=DEV)
    ids = req["input_ids"]
    S, T, D, V = ids.shape[1], req["pixel_values"].shape[0], spec.text.hidden_size, spec.text.vocab_size
    _report(f"[{tag}] {spec.name} T={T} masks={n_masks} S={S} new_tokens={new_tokens}")
    eng = Engine(spec, sd, device=DEV, max_tiles=T, max_seq=((S + new_tokens + 127) // 128) * 128, max_batch=1,
                 max_new_tokens=new_tokens)
    with torch.cuda.stream(eng.stream):
        eng.request(0, req["pixel_values"], req["global_mask_values"], ids[0].contiguous(), req["bboxes"][0],
                    req["aspect_ratios"][0].tolist())
    eng.stream.synchronize()
    assert eng.status(0) == 0

    # ---- R2-R5 image features at full depth / width / tile count
    feats = eng.buffer("image_features", (T, spec.tokens_per_tile, D))
    with torch.no_grad():
        f32 = o32.image_features(req["pixel_values"].float(), o32.mask_embeds(req["global_mask_values"]))
        f16 = o16.image_features(req["pixel_values"], o16.mask_embeds(req["global_mask_values"]))
    e_our, e_ref = _compare("image_features", feats, f16, f32)
    assert e_our <= 1.5 * e_ref + 2e-3 * f32.abs().max().item()

    # ---- R6-R7 inputs_embeds: bit-exact on the engine's own features
    emb = eng.buffer("inputs_embeds", (S, D))
    want = o16.assemble(ids, feats, req["aspect_ratios"], req["bboxes"])[0]
    crop_rows = torch.isin(ids[0], torch.tensor(spec.crop_tokens_ids, device=DEV))
    assert int(crop_rows.sum()) == n_masks * spec.tokens_per_tile
    assert torch.equal(emb[~crop_rows], want[~crop_rows]), "embed gather / image scatter rows must be bit-exact"
    # replay rows: the oracle evaluates torchvision's CUDA roi_align here (different FMA contraction than the CPU
    # operator the kernel is pinned to bit-for-bit in test_gpu_kernels / test_roi_replay_ref): allow one bf16 ulp
    # (where the four bilinear terms nearly cancel, an fp32-ulp difference of the sum is many ulps of the tiny bf16
    # result: hence the absolute floor of 1e-5 of the feature scale)
    ours_r, want_r = emb[crop_rows].float(), want[crop_rows].float()
    diff = (ours_r - want_r).abs()
    tol = torch.maximum(_ulp_bf16(torch.maximum(ours_r.abs(), want_r.abs())),
                        torch.full_like(diff, 1e-5 * want_r.abs().max().item()))
    n_off = int((diff > 0).sum())
    _report(f"  replay rows vs torchvision CUDA roi_align: {n_off} of {diff.numel()} elements differ, max {diff.max().item():.3g},"
            f" beyond 1 bf16 ulp: {int((diff > tol).sum())}")
    assert bool((diff <= tol).all())
    want_cpu = o16.to("cpu").assemble(ids.cpu(), feats.cpu(), req["aspect_ratios"].cpu(), req["bboxes"])[0]
    o16.to(DEV)
    assert torch.equal(emb.cpu(), want_cpu), "inputs_embeds must be bit-exact vs the CPU roi_align the reference's CPU path runs"

    # ---- R8 prefill logits (same inputs_embeds for every arm)
    logits = eng.buffer("logits", (eng.vocab_padded,))[:V].float()
    seq32, lg32, all32 = o32.greedy(emb.float()[None], new_tokens, None, return_logits=True)
    seq16, lg16 = o16.greedy(emb[None], new_tokens, None)
    e_our, e_ref = _compare("prefill logits", logits, lg16[0], lg32[0])
    top5 = lambda x: set(x.topk(5).indices.tolist())  # noqa: E731
    _report(f"  top-5 overlap with fp32: ours {len(top5(logits) & top5(lg32[0]))}/5, oracle_bf16 {len(top5(lg16[0].float()) & top5(lg32[0]))}/5;"
            f" argmax ours={int(logits.argmax())} fp32={int(lg32[0].argmax())} bf16={int(lg16[0].argmax())}")
    assert e_our <= 1.5 * e_ref + 2e-3
    noise = max(e_our, e_ref)

    # ---- R9 greedy decode, no stop id
    toks = eng.generate(1, new_tokens, None)[0]
    t32, t16 = seq32[0].tolist(), seq16[0].tolist()
    assert len(toks) == new_tokens and eng.last_steps_run == new_tokens - 1
    d32, d16, d_ref = _first_divergence(toks, t32), _first_divergence(toks, t16), _first_divergence(t16, t32)
    _report(f"  greedy {new_tokens} tokens: distinct ours={len(set(toks))}; first divergence ours/fp32={d32} ours/oracle_bf16={d16}"
            f" oracle_bf16/fp32={d_ref}")
    assert len(set(toks)) >= min(16, new_tokens // 4), "degenerate token stream: the init does not discriminate"
    if d32 is not None:
        top2 = all32[d32][0].topk(2).values
        margin = (top2[0] - top2[1]).item()
        _report(f"  fp32 top-2 margin at step {d32}: {margin:.4g} (bf16 noise {noise:.4g})")
        assert margin <= 4 * noise, f"token {d32}: ours {toks[d32]} != fp32 {t32[d32]} with margin {margin} > 4 x noise {noise}"

    # ---- R9 with stop ids: eos = the token ours emits mid-sequence -> truncated there (inclusive), decode stops early
    def restage():
        with torch.cuda.stream(eng.stream):
            eng.request(0, req["pixel_values"], req["global_mask_values"], ids[0].contiguous(), req["bboxes"][0],
                        req["aspect_ratios"][0].tolist())

    k = new_tokens // 3
    eos = toks[k]
    first = toks.index(eos)
    restage()
    got = eng.generate(1, new_tokens, eos)[0]
    assert got == toks[: first + 1], "stop id: sequence must end AT the first stop token (HF semantics)"
    assert eng.last_steps_run <= first + 2 * 8 + 1 < new_tokens - 1, (eng.last_steps_run, first)
    _report(f"  eos={eos} fires at step {first}: {len(got)} tokens returned, {eng.last_steps_run} of {new_tokens - 1} decode steps launched")
    restage()
    absent = next(t for t in range(1000, 2000) if t not in toks)
    got = eng.generate(1, new_tokens, [absent, toks[k + 5], eos])[0]      # list of stop ids: first hit of ANY of them
    cut = min(toks.index(toks[k + 5]), first)
    assert got == toks[: cut + 1]
    restage()
    got = eng.generate(1, new_tokens, [absent])[0]                        # never fires
    assert got == toks and eng.last_steps_run == new_tokens - 1
    eng.close()
    del eng
    torch.cuda.empty_cache()


def test_gar_1b_full_size_parity():
    """BASELINE configs 2 and 3: GAR-1B, 1024^2 image -> 16+1 tiles, 1 mask (S=4672) and 4 masks, 128 greedy tokens."""
    spec = gar_1b()
    o32, o16, sd = _gpu_oracles(spec)
    _run_config("config 2", spec, o32, o16, sd, n_masks=1, max_num_tiles=16, new_tokens=128, seed=3)
    _run_config("config 3", spec, o32, o16, sd, n_masks=4, max_num_tiles=16, new_tokens=128, seed=4)


def test_gar_8b_full_size_parity():
    """BASELINE config 4: GAR-8B (PE-G depth 47 / d 1536 / hd 96 no cls; Llama-3.1-8B hd 128), T=5 (the reference's own
    max_num_tiles=8 training setting) and T=17; 64 tokens keep the fp32 oracle's 8B decode loop to a few seconds."""
    spec = gar_8b()
    o32, o16, sd = _gpu_oracles(spec)
    _run_config("config 4, T=5", spec, o32, o16, sd, n_masks=2, max_num_tiles=4, new_tokens=64, seed=5)
    _run_config("config 4, T=17", spec, o32, o16, sd, n_masks=1, max_num_tiles=16, new_tokens=48, seed=6)
