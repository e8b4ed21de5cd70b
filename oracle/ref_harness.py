"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  It executes the
reference's own ``GARModel.generate`` (hf_models/modeling_gar.py:295-417) and the vendored
``PerceptionLMForConditionalGeneration`` (models/modeling/modeling_perception_lm.py) on CPU
with the shims listed in SURVEY.md §8(c):

  (i)   ``timm.models._manipulate.checkpoint`` stubbed (timm is not installed);
  (ii)  the vendored PLM files are loaded by file path (importing the package pulls mmengine);
  (iii) ``AutoModel.from_config`` returns a TimmWrapperModel-shaped stand-in whose
        ``.timm_model`` is the PE-ViT restatement (``oracle.gar_oracle.PEViT``) -- the
        reference then monkey-patches ITS OWN ``custom_forward_features`` onto it (:219);
  (iv)  ``mllm.generate`` is intercepted to capture ``inputs_embeds`` (HF generate is
        API-incompatible with transformers 5.5) and a manual greedy loop drives the vendored
        forward with ``past_key_values``.
  (v)   for the PLM replay fixtures the vendored ``prepare_inputs_for_generation`` is called directly;
        transformers 5.5's ``GenerationMixin.prepare_inputs_for_generation`` omits the ``inputs_embeds`` key the
        vendored code indexes (:609; the pinned 4.56.2 returns it as None), so the key is defaulted to None.

``python -m oracle.ref_harness [case ...]`` regenerates tests/golden/ref_*.npz (all, or the named cases).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GAR_REFERENCE_ROOT", "/root/reference")


def _load_reference():
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree not present at {REF}")
    # (i) timm stub
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        tm = types.ModuleType("timm.models")
        tmm = types.ModuleType("timm.models._manipulate")
        tmm.checkpoint = lambda fn, *a, **k: fn(*a, **k)
        timm.models = tm
        tm._manipulate = tmm
        sys.modules.update({"timm": timm, "timm.models": tm, "timm.models._manipulate": tmm})
    # (ii) vendored PLM by file path, as a synthetic package so the relative import resolves
    mdir = os.path.join(REF, "projects/grasp_any_region/models/modeling")
    pkg = types.ModuleType("ref_plm")
    pkg.__path__ = [mdir]
    sys.modules["ref_plm"] = pkg
    mods = {}
    for name in ("configuration_perception_lm", "modeling_perception_lm"):
        spec = importlib.util.spec_from_file_location(f"ref_plm.{name}", os.path.join(mdir, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"ref_plm.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from projects.grasp_any_region.hf_models import configuration_gar, modeling_gar

    return mods["configuration_perception_lm"], mods["modeling_perception_lm"], configuration_gar, modeling_gar


class _TowerOut:
    def __init__(self, x):
        self.last_hidden_state = x


class _TimmWrapperStandIn(torch.nn.Module):
    """TimmWrapperModel.forward contract (site-packages/transformers/models/timm_wrapper/
    modeling_timm_wrapper.py:259): last_hidden_state = timm_model.forward_features(px, **kwargs)."""

    def __init__(self, vit):
        super().__init__()
        self.timm_model = vit

    def forward(self, pixel_values, **kwargs):
        return _TowerOut(self.timm_model.forward_features(pixel_values, **kwargs))


def build_reference_model(spec, oracle_model):
    """Instantiate the reference GARModel around the vendored PLM and load the oracle's weights."""
    import transformers
    from transformers import AutoModel

    from oracle.gar_oracle import PEViT

    cfg_plm, mod_plm, cfg_gar, mod_gar = _load_reference()
    t, v = spec.text, spec.vision
    text_cfg = dict(
        model_type="llama", hidden_size=t.hidden_size, num_hidden_layers=t.num_layers,
        num_attention_heads=t.num_heads, num_key_value_heads=t.num_kv_heads, head_dim=t.head_dim,
        intermediate_size=t.intermediate_size, vocab_size=t.vocab_size, rms_norm_eps=t.rms_eps,
        max_position_embeddings=t.max_position_embeddings, hidden_act="silu", attention_bias=False,
        mlp_bias=False, tie_word_embeddings=False,
        rope_parameters=dict(rope_type="llama3", rope_theta=t.rope_theta, factor=t.rope_factor,
                             low_freq_factor=t.rope_low_freq_factor,
                             high_freq_factor=t.rope_high_freq_factor,
                             original_max_position_embeddings=t.rope_original_max_pos),
    )
    vision_cfg = dict(architecture="vit_pe_core_large_patch14_336",
                      model_args=dict(img_size=[v.img_size] * 2, ref_feat_shape=[v.ref_feat_shape] * 2,
                                      embed_dim=v.embed_dim, depth=v.depth, num_heads=v.num_heads))
    plm_cfg = cfg_plm.PerceptionLMConfig(
        vision_config=vision_cfg, text_config=text_cfg, vision_use_cls_token=v.use_cls_token,
        projector_pooling_ratio=spec.projector_pooling_ratio, image_token_id=spec.image_token_id)
    plm_cfg._attn_implementation = "eager"
    plm_cfg.text_config._attn_implementation = "sdpa"

    orig = AutoModel.from_config

    def patched(config, *a, **k):
        if type(config).__name__ == "TimmWrapperConfig":
            return _TimmWrapperStandIn(PEViT(v))
        k.pop("attn_implementation", None)
        return orig(config, *a, **k)

    AutoModel.from_config = staticmethod(patched)
    try:
        mllm = mod_plm.PerceptionLMForConditionalGeneration(plm_cfg)
    finally:
        AutoModel.from_config = orig

    gar_cfg = cfg_gar.GARConfig(
        mllm_config=dict(vision_config=vision_cfg, text_config=dict(text_cfg),
                         vision_use_cls_token=v.use_cls_token,
                         projector_pooling_ratio=spec.projector_pooling_ratio),
        prompt_numbers=spec.prompt_numbers, crop_tokens_ids=list(spec.crop_tokens_ids))
    mask_conv = torch.nn.Conv2d(3, v.embed_dim, v.patch_size, v.patch_size, bias=False)
    model = mod_gar.GARModel(gar_cfg, mllm=mllm, mask_patch_embedding=mask_conv, use_flash_attn=False)
    missing, unexpected = model.load_state_dict(oracle_model.state_dict(), strict=False)
    assert not unexpected, unexpected
    assert all("rotary_emb" in m or "rope_embed" in m for m in missing), missing
    return model.eval()


@torch.no_grad()
def run_reference(model, request, max_new_tokens, eos_token_id=None):
    """Execute reference generate() up to the mllm.generate call, then greedy-decode with the
    vendored forward.  Returns the same taps as GAROracle.generate."""
    taps = {}

    class _Stop(Exception):
        pass

    def fake_generate(inputs_embeds=None, **kw):
        taps["inputs_embeds"] = inputs_embeds
        raise _Stop()

    mllm = model.mllm
    real_gif = mllm.get_image_features

    def tap_gif(pixel_values=None, mask_embeds=None, **kw):
        taps["mask_embeds"] = mask_embeds
        out = real_gif(pixel_values=pixel_values, mask_embeds=mask_embeds, **kw)
        taps["image_features"] = out
        return out

    mllm.get_image_features = tap_gif
    mllm.generate = fake_generate
    try:
        model.generate(**request)
    except _Stop:
        pass
    finally:
        del mllm.generate
        del mllm.get_image_features

    emb = taps["inputs_embeds"]
    S = emb.shape[1]
    out = mllm(inputs_embeds=emb, attention_mask=torch.ones(1, S, dtype=torch.long), use_cache=True,
               logits_to_keep=1)
    logits = out.logits[:, -1]
    taps["logits"] = logits.clone()
    toks = []
    emb_fn = mllm.get_input_embeddings()
    for i in range(max_new_tokens):
        nxt = logits.argmax(-1)
        toks.append(nxt)
        if (eos_token_id is not None and int(nxt[0]) == eos_token_id) or i == max_new_tokens - 1:
            break
        out = mllm(inputs_embeds=emb_fn(nxt)[:, None, :], past_key_values=out.past_key_values,
                   use_cache=True)
        logits = out.logits[:, -1]
    taps["sequences"] = torch.stack(toks, 1)
    return taps


def compact(taps, request, spec):
    """Reduce taps to a small fixture: strided samples + the exact rows that matter."""
    ids = request["input_ids"][0]
    crop_rows = torch.cat([(ids == c).nonzero().flatten() for c in spec.crop_tokens_ids])
    text_rows = ((ids != spec.image_token_id) & ~torch.isin(ids, torch.tensor(spec.crop_tokens_ids))).nonzero().flatten()
    img_rows = (ids == spec.image_token_id).nonzero().flatten()[::37]
    rows = torch.cat([crop_rows, text_rows, img_rows]).sort().values
    lg = taps["logits"][0].float()
    top = lg.topk(64)
    return dict(
        mask_embeds=taps["mask_embeds"][:, ::4, ::2, ::2].float().numpy(),
        image_features=taps["image_features"][:, ::3, ::2].float().numpy(),
        embed_rows=rows.numpy(),
        inputs_embeds=taps["inputs_embeds"][0, rows].float().numpy(),
        logits_top_idx=top.indices.numpy(), logits_top_val=top.values.numpy(),
        logits_strided=lg[::16].numpy(),
        sequences=taps["sequences"].numpy(),
    )


def frame_crop_token_id(frame_idx: int) -> int:
    """Llama-3 tokenizer id of ``<|reserved_special_token_{2 + frame_idx}|>`` (modeling_perception_lm.py:778-780):
    the reserved tokens are numbered around the named specials 128006/128007/128009."""
    k = 2 + frame_idx
    return {0: 128002, 1: 128003, 2: 128004, 3: 128005, 4: 128008}.get(k, 128005 + k)


class _FrameTokenizer:
    def convert_tokens_to_ids(self, tok: str) -> int:
        return frame_crop_token_id(int(tok.split("_")[-1].rstrip("|>")) - 2)


class _FrameProcessor:
    tokenizer = _FrameTokenizer()


@torch.no_grad()
def run_reference_plm_replay(model, request, video: bool, crop_tokens=None):
    """inputs_embeds produced by the UNMODIFIED vendored ``prepare_inputs_for_generation`` replay branches
    (feature_replay / feature_replay_video) for one request, plus the image features they were built from."""
    from transformers.generation.utils import GenerationMixin

    mllm = model.mllm
    orig = GenerationMixin.prepare_inputs_for_generation

    def shim(self, *a, **k):
        out = orig(self, *a, **k)
        out.setdefault("inputs_embeds", None)
        return out

    taps = {}
    real_gif = mllm.get_image_features

    def tap_gif(pixel_values=None, mask_embeds=None, **kw):
        out = real_gif(pixel_values=pixel_values, mask_embeds=mask_embeds, **kw)
        taps["image_features"] = out
        return out

    ids = request["input_ids"]
    S = ids.shape[1]
    GenerationMixin.prepare_inputs_for_generation = shim
    mllm.get_image_features = tap_gif
    try:
        kw = dict(feature_replay_video=True, processor=_FrameProcessor()) if video else \
            dict(feature_replay=True, crop_tokens=list(crop_tokens))
        mi = mllm.prepare_inputs_for_generation(
            ids, past_key_values=None, inputs_embeds=None, pixel_values=request["pixel_values"], mask_embeds=None,
            attention_mask=torch.ones(1, S, dtype=torch.long), cache_position=torch.arange(S), logits_to_keep=1,
            bboxes=request["bboxes"], aspect_ratios=request["aspect_ratios"].tolist(), **kw)
    finally:
        GenerationMixin.prepare_inputs_for_generation = orig
        del mllm.get_image_features
    taps["inputs_embeds"] = mi["inputs_embeds"]
    return taps


def compact_plm(taps, request, spec):
    ids = request["input_ids"][0]
    crop_ids = torch.tensor([frame_crop_token_id(f) for f in range(8)])
    crop_rows = torch.isin(ids, crop_ids).nonzero().flatten()
    text_rows = ((ids != spec.image_token_id) & ~torch.isin(ids, crop_ids)).nonzero().flatten()
    img_rows = (ids == spec.image_token_id).nonzero().flatten()[::37]
    rows = torch.cat([crop_rows, text_rows, img_rows]).sort().values
    return dict(image_features=taps["image_features"][:, ::3, ::2].float().numpy(), embed_rows=rows.numpy(),
                inputs_embeds=taps["inputs_embeds"][0, rows].float().numpy())


PLM_CASES = {
    # name: (video, n_masks, max_num_tiles, seed) -- 3 regions on a 2x2+1 tile request: the image branch keeps only the
    # last crop token's replay; the video branch treats the 5 tiles as 5 frames and replays frames 0..2 from their own map
    "ref_tiny_plm_replay": (False, 3, 4, 21),
    "ref_tiny_plm_video": (True, 3, 4, 22),
}


CASES = {
    # name: (n_masks, max_num_tiles, image_size, new_tokens, seed)
    "ref_tiny_1mask": (1, 4, 1024, 12, 11),
    "ref_tiny_3mask": (3, 4, 1024, 8, 12),
}


def main():
    from grasp_any_region_b200.config import gar_tiny
    from grasp_any_region_b200.synth import make_request
    from oracle.gar_oracle import build_oracle

    torch.set_num_threads(os.cpu_count())
    spec = gar_tiny()
    oracle = build_oracle(spec, seed=0, dtype=torch.float32)
    ref = build_reference_model(spec, oracle)
    outdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[1:])  # `python -m oracle.ref_harness [case names]`: regenerate just those fixtures
    for name, (n_masks, mnt, isz, ntok, seed) in CASES.items():
        if only and name not in only:
            continue
        req = make_request(spec, seed=seed, n_masks=n_masks, image_size=isz, max_num_tiles=mnt,
                           dtype=torch.float32)
        taps = run_reference(ref, req, ntok)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **compact(taps, req, spec))
        print(name, "S =", req["input_ids"].shape[1], "tokens:", taps["sequences"].tolist())
    for name, (video, n_masks, mnt, seed) in PLM_CASES.items():
        if only and name not in only:
            continue
        req = make_request(spec, seed=seed, n_masks=n_masks, image_size=1024, max_num_tiles=mnt, dtype=torch.float32)
        taps = run_reference_plm_replay(ref, req, video, crop_tokens=spec.crop_tokens_ids)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **compact_plm(taps, req, spec))
        print(name, "S =", req["input_ids"].shape[1])


if __name__ == "__main__":
    main()
