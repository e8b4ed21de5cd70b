"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain-torch restatement of the reference's region-captioning inference path
``GARModel.generate`` (reference: projects/grasp_any_region/hf_models/modeling_gar.py:295-428).
Runs on CPU (fp32 = the timed CPU baseline / golden generator) or, as the same-box checker
for the bf16 CUDA path, on a GPU in bf16.  Module/parameter names follow the reference's HF
state-dict layout so ONE state dict feeds this oracle, the real reference
modules (oracle/ref_harness.py) and the B200 engine.

Every function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from grasp_any_region_b200.config import GARSpec, TextSpec, VisionSpec


# --------------------------------------------------------------------------------------
# PE ViT (timm Eva) -- upstream timm==1.0.19 timm/models/eva.py, timm/layers/attention.py,
# timm/layers/pos_embed_sincos.py (NOT under /root/reference).
# The reference reaches it only through modeling_perception_lm.py:181-217.
# --------------------------------------------------------------------------------------
def build_vit_rope_table(spec: VisionSpec) -> torch.Tensor:
    """cat(sin, cos) of shape (grid*grid, 2*head_dim), fp32 (SURVEY.md Appendix A 'rope table')."""
    hd = spec.head_dim
    nb = hd // 4
    bands = 1.0 / (spec.rope_theta ** (torch.arange(0, nb, dtype=torch.float32) / nb))
    g = spec.grid
    t = (torch.arange(g, dtype=torch.float32) + 1.0) / g * spec.ref_feat_shape  # grid_offset=1
    # 'xy' indexing: first coordinate of the pair is x (varies along w), second is y
    yy, xx = torch.meshgrid(t, t, indexing="ij")
    grid = torch.stack([xx, yy], dim=-1)  # (h, w, 2)
    pos = grid.unsqueeze(-1) * bands  # (h, w, 2, nb)
    pos = pos.reshape(g * g, 2 * nb)  # (N, hd/2)
    sin = pos.sin().repeat_interleave(2, -1)
    cos = pos.cos().repeat_interleave(2, -1)
    return torch.cat([sin, cos], dim=-1)  # (N, 2*hd)


def _rot_interleaved(x: torch.Tensor) -> torch.Tensor:
    return torch.stack([-x[..., 1::2], x[..., ::2]], -1).reshape(x.shape)


def apply_rot_embed_cat(x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    sin_emb, cos_emb = emb.tensor_split(2, -1)
    return x * cos_emb + _rot_interleaved(x) * sin_emb


class PatchEmbed(nn.Module):
    def __init__(self, spec: VisionSpec):
        super().__init__()
        self.proj = nn.Conv2d(3, spec.embed_dim, spec.patch_size, spec.patch_size, bias=False)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class AttentionRope(nn.Module):
    def __init__(self, spec: VisionSpec):
        super().__init__()
        d = spec.embed_dim
        self.num_heads = spec.num_heads
        self.head_dim = spec.head_dim
        self.num_prefix_tokens = spec.num_prefix
        self.qkv = nn.Linear(d, 3 * d, bias=True)
        self.proj = nn.Linear(d, d, bias=True)

    def forward(self, x, rope=None):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        if rope is not None:
            npt = self.num_prefix_tokens
            q = torch.cat([q[:, :, :npt], apply_rot_embed_cat(q[:, :, npt:], rope)], 2).type_as(v)
            k = torch.cat([k[:, :, :npt], apply_rot_embed_cat(k[:, :, npt:], rope)], 2).type_as(v)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(B, N, C)
        return self.proj(o)


class Mlp(nn.Module):
    def __init__(self, d, hidden):
        super().__init__()
        self.fc1 = nn.Linear(d, hidden, bias=True)
        self.fc2 = nn.Linear(hidden, d, bias=True)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class EvaBlock(nn.Module):
    def __init__(self, spec: VisionSpec):
        super().__init__()
        d = spec.embed_dim
        self.norm1 = nn.LayerNorm(d, eps=spec.ln_eps)
        self.attn = AttentionRope(spec)
        self.gamma_1 = nn.Parameter(spec.layer_scale_init * torch.ones(d))
        self.norm2 = nn.LayerNorm(d, eps=spec.ln_eps)
        self.mlp = Mlp(d, spec.mlp_dim)
        self.gamma_2 = nn.Parameter(spec.layer_scale_init * torch.ones(d))

    def forward(self, x, rope=None):
        x = x + self.gamma_1 * self.attn(self.norm1(x), rope=rope)
        x = x + self.gamma_2 * self.mlp(self.norm2(x))
        return x


class PEViT(nn.Module):
    """Exposes exactly the attribute protocol the reference's ``custom_forward_features``
    uses (modeling_perception_lm.py:181-217): patch_embed, _pos_embed, norm_pre, blocks, norm,
    grad_checkpointing."""

    def __init__(self, spec: VisionSpec):
        super().__init__()
        self.spec = spec
        d = spec.embed_dim
        self.patch_embed = PatchEmbed(spec)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, d)) if spec.use_cls_token else None
        self.pos_embed = nn.Parameter(torch.zeros(1, spec.tokens, d))
        self.norm_pre = nn.LayerNorm(d, eps=spec.ln_eps)
        self.blocks = nn.ModuleList([EvaBlock(spec) for _ in range(spec.depth)])
        self.norm = nn.Identity()
        self.grad_checkpointing = False
        # non-persistent float buffer => follows the module dtype under .to(bf16), like timm
        self.register_buffer("rope_embed", build_vit_rope_table(spec), persistent=False)

    def _pos_embed(self, x):
        if self.cls_token is not None:
            x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = x + self.pos_embed
        return x, self.rope_embed

    def forward_features(self, x, mask_embeds=None):
        # == reference custom_forward_features, modeling_perception_lm.py:194-217
        x = self.patch_embed(x)
        if mask_embeds is not None:
            x = x + mask_embeds.flatten(2).transpose(1, 2)
        x, rope = self._pos_embed(x)
        x = self.norm_pre(x)
        for blk in self.blocks:
            x = blk(x, rope=rope)
        return self.norm(x)


# --------------------------------------------------------------------------------------
# Projector: modeling_perception_lm.py:42-92
# --------------------------------------------------------------------------------------
class Projector(nn.Module):
    def __init__(self, spec: GARSpec):
        super().__init__()
        self.linear_1 = nn.Linear(spec.vision.embed_dim, spec.text.hidden_size, bias=True)
        self.linear_2 = nn.Linear(spec.text.hidden_size, spec.text.hidden_size, bias=True)
        self.ratio = spec.projector_pooling_ratio

    def forward(self, x):  # (T, N, C)
        x = self.linear_2(F.gelu(self.linear_1(x)))
        if self.ratio > 1:
            b, n, c = x.shape
            h = int(math.sqrt(n))
            assert h * h == n
            x = x.permute(0, 2, 1).reshape(b, -1, h, h)
            x = F.adaptive_avg_pool2d(x, (h // self.ratio, h // self.ratio))
            x = x.flatten(2).transpose(1, 2)
        return x


# --------------------------------------------------------------------------------------
# Llama: site-packages/transformers/models/llama/modeling_llama.py
# --------------------------------------------------------------------------------------
def llama3_inv_freq(t: TextSpec) -> torch.Tensor:
    """transformers.modeling_rope_utils._compute_llama3_parameters restated."""
    dim = t.head_dim
    inv_freq = 1.0 / (t.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    low_w = t.rope_original_max_pos / t.rope_low_freq_factor
    high_w = t.rope_original_max_pos / t.rope_high_freq_factor
    wavelen = 2 * math.pi / inv_freq
    inv_l = torch.where(wavelen > low_w, inv_freq / t.rope_factor, inv_freq)
    smooth = (t.rope_original_max_pos / wavelen - t.rope_low_freq_factor) / (
        t.rope_high_freq_factor - t.rope_low_freq_factor)
    smoothed = (1 - smooth) * inv_l / t.rope_factor + smooth * inv_l
    is_med = ~(wavelen < high_w) * ~(wavelen > low_w)
    return torch.where(is_med, smoothed, inv_l)


def llama_rope_cos_sin(t: TextSpec, positions: torch.Tensor, dtype) -> (torch.Tensor, torch.Tensor):
    """LlamaRotaryEmbedding.forward: fp32 table, cast to the activation dtype. positions (S,)"""
    inv = llama3_inv_freq(t).to(positions.device)
    freqs = positions.float()[:, None] * inv[None, :]
    emb = torch.cat([freqs, freqs], -1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], -1)


class RMSNorm(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.eps = eps

    def forward(self, x):
        dt = x.dtype
        x32 = x.float()
        x32 = x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + self.eps)
        return self.weight * x32.to(dt)


class LlamaAttention(nn.Module):
    def __init__(self, t: TextSpec):
        super().__init__()
        self.t = t
        d = t.hidden_size
        self.q_proj = nn.Linear(d, t.num_heads * t.head_dim, bias=False)
        self.k_proj = nn.Linear(d, t.num_kv_heads * t.head_dim, bias=False)
        self.v_proj = nn.Linear(d, t.num_kv_heads * t.head_dim, bias=False)
        self.o_proj = nn.Linear(t.num_heads * t.head_dim, d, bias=False)

    def forward(self, x, cos, sin, cache, causal):
        t = self.t
        B, S, _ = x.shape
        q = self.q_proj(x).view(B, S, t.num_heads, t.head_dim).transpose(1, 2)
        k = self.k_proj(x).view(B, S, t.num_kv_heads, t.head_dim).transpose(1, 2)
        v = self.v_proj(x).view(B, S, t.num_kv_heads, t.head_dim).transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        if cache is not None:
            if "k" in cache:
                k = torch.cat([cache["k"], k], 2)
                v = torch.cat([cache["v"], v], 2)
            cache["k"], cache["v"] = k, v
        rep = t.num_heads // t.num_kv_heads
        kk = k.repeat_interleave(rep, 1)
        vv = v.repeat_interleave(rep, 1)
        o = F.scaled_dot_product_attention(q, kk, vv, is_causal=causal and S > 1)
        o = o.transpose(1, 2).reshape(B, S, t.num_heads * t.head_dim)
        return self.o_proj(o)


class LlamaMLP(nn.Module):
    def __init__(self, t: TextSpec):
        super().__init__()
        self.gate_proj = nn.Linear(t.hidden_size, t.intermediate_size, bias=False)
        self.up_proj = nn.Linear(t.hidden_size, t.intermediate_size, bias=False)
        self.down_proj = nn.Linear(t.intermediate_size, t.hidden_size, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class LlamaLayer(nn.Module):
    def __init__(self, t: TextSpec):
        super().__init__()
        self.self_attn = LlamaAttention(t)
        self.mlp = LlamaMLP(t)
        self.input_layernorm = RMSNorm(t.hidden_size, t.rms_eps)
        self.post_attention_layernorm = RMSNorm(t.hidden_size, t.rms_eps)

    def forward(self, h, cos, sin, cache, causal):
        h = h + self.self_attn(self.input_layernorm(h), cos, sin, cache, causal)
        h = h + self.mlp(self.post_attention_layernorm(h))
        return h


class Llama(nn.Module):
    def __init__(self, t: TextSpec):
        super().__init__()
        self.t = t
        self.embed_tokens = nn.Embedding(t.vocab_size, t.hidden_size)
        self.layers = nn.ModuleList([LlamaLayer(t) for _ in range(t.num_layers)])
        self.norm = RMSNorm(t.hidden_size, t.rms_eps)

    def forward(self, inputs_embeds, positions, caches=None):
        """inputs_embeds (B,S,D); positions (S,) ; caches: list of dicts or None."""
        cos, sin = llama_rope_cos_sin(self.t, positions, inputs_embeds.dtype)
        h = inputs_embeds
        for i, layer in enumerate(self.layers):
            h = layer(h, cos, sin, None if caches is None else caches[i], True)
        return self.norm(h)


# --------------------------------------------------------------------------------------
# GAR glue -- modeling_gar.py:309-417 + modeling_perception_lm.py:239-331
# --------------------------------------------------------------------------------------
class _VisionTower(nn.Module):
    def __init__(self, spec):
        super().__init__()
        self.timm_model = PEViT(spec)


class _PLMModel(nn.Module):
    def __init__(self, spec: GARSpec):
        super().__init__()
        self.vision_tower = _VisionTower(spec.vision)
        self.multi_modal_projector = Projector(spec)
        self.language_model = Llama(spec.text)


class _PLM(nn.Module):
    def __init__(self, spec: GARSpec):
        super().__init__()
        self.model = _PLMModel(spec)
        self.lm_head = nn.Linear(spec.text.hidden_size, spec.text.vocab_size, bias=False)


def decode_mask_ids(global_mask_values: torch.Tensor, prompt_numbers: int) -> torch.Tensor:
    """modeling_gar.py:315-325 -- evaluated in the tensor's own dtype (quirk 7)."""
    ids = torch.round((global_mask_values + 1.0) / 2.0 * 255.0).long()
    ids = torch.clamp(ids, min=0, max=prompt_numbers)
    assert ids.max() < prompt_numbers + 1 and ids.min() >= 0
    return ids


def roi_box_f32(bbox, feat_h: int, feat_w: int):
    """modeling_gar.py:367-387: python-double arithmetic, then rounded to fp32 by torch.tensor.
    Returns (roi tensor (5,) fp32, spatial_scale python float)."""
    x1, y1, x2, y2 = [float(v) for v in bbox]
    orig_h, orig_w = feat_h * 28, feat_w * 28
    spatial_scale = feat_w / orig_w
    roi = torch.tensor(
        [0, x1 * orig_w * spatial_scale, y1 * orig_h * spatial_scale,
         x2 * orig_w * spatial_scale, y2 * orig_h * spatial_scale], dtype=torch.float32)
    return roi, spatial_scale


class GAROracle(nn.Module):
    def __init__(self, spec: GARSpec):
        super().__init__()
        self.spec = spec
        self.mllm = _PLM(spec)
        self.mask_patch_embedding = nn.Conv2d(
            3, spec.mask_path_embedding_out_channels, spec.vision.patch_size, spec.vision.patch_size,
            bias=False)

    # -- taps ---------------------------------------------------------------------
    def mask_embeds(self, global_mask_values):
        ids = decode_mask_ids(global_mask_values, self.spec.prompt_numbers)
        dt = self.mask_patch_embedding.weight.dtype
        return self.mask_patch_embedding((ids != self.spec.prompt_numbers).to(dt))

    def image_features(self, pixel_values, mask_embeds):
        # modeling_perception_lm.py:239-269
        x = self.mllm.model.vision_tower.timm_model.forward_features(pixel_values, mask_embeds)
        if self.spec.vision.use_cls_token:
            x = x[:, 1:, :]
        return self.mllm.model.multi_modal_projector(x)

    def assemble(self, input_ids, image_features, aspect_ratios, bboxes):
        """modeling_gar.py:332-414 for batch 1.  Returns inputs_embeds (1,S,D)."""
        spec = self.spec
        emb = self.mllm.model.language_model.embed_tokens(input_ids)
        img_mask = input_ids == spec.image_token_id
        if int(img_mask.sum()) * emb.shape[-1] != image_features.numel():
            raise ValueError(
                f"Image features and image tokens do not match: tokens: {int(img_mask.sum())}, "
                f"features {image_features.shape[:-1].numel()}")
        emb = emb.masked_scatter(img_mask.unsqueeze(-1).expand_as(emb), image_features.to(emb.dtype))
        tpt = int(math.isqrt(spec.tokens_per_tile))
        T1, _, D = image_features[1:].shape
        out = []
        for b in range(emb.shape[0]):
            cur = emb[b]
            for crop_token in spec.crop_tokens_ids:
                pos = (input_ids[b] == crop_token).nonzero().flatten()
                if pos.numel() == 0:
                    continue
                head, tail = int(pos.min()), int(pos.max())
                ncw, nch = int(aspect_ratios[b][0]), int(aspect_ratios[b][1])
                assert T1 == ncw * nch, f"{ncw * nch} != {T1}"
                # _merge (modeling_gar.py:248-260): tiles row-major over (nch, ncw)
                tiles = image_features[1:].reshape(nch, ncw, tpt, tpt, D)
                fmap = tiles.permute(4, 0, 2, 1, 3).reshape(1, D, nch * tpt, ncw * tpt)
                roi, scale = roi_box_f32(bboxes[b][str(crop_token)], nch * tpt, ncw * tpt)
                import torchvision

                feats = torchvision.ops.roi_align(
                    fmap.float(), roi.unsqueeze(0).to(fmap.device), (tpt, tpt), spatial_scale=scale,
                    sampling_ratio=2, aligned=True)
                replay = feats.permute(0, 2, 3, 1).flatten(1, 2).to(fmap.dtype).squeeze(0)
                cur = torch.cat([cur[:head], replay, cur[tail + 1:]])
            out.append(cur.unsqueeze(0))
        return torch.cat(out, 0)

    def assemble_plm(self, input_ids, image_features, aspect_ratios, bboxes, crop_tokens=(128004,), video=False,
                     frame_crop_ids=None):
        """The replay branches of the vendored PLM's ``prepare_inputs_for_generation``
        (models/modeling/modeling_perception_lm.py:608-858), batch 1.  Returns inputs_embeds (1,S,D).

        video=False (``feature_replay``, :648-763): same merged-map RoI replay as GARModel.generate, over ``crop_tokens``
        in list order -- but every replay is spliced into the ORIGINAL scattered embeddings (``inputs_embeds[batch_idx]``,
        :736-742), so only the LAST crop token present in the prompt keeps its replay; the others keep the embedding of
        their crop-token id.
        video=True (``feature_replay_video``, :765-852): frame f's crop token (``frame_crop_ids[f]`` = the id of
        ``<|reserved_special_token_{2+f}|>``) is replayed from frame f's OWN 16x16 map (no thumbnail drop, no merge),
        on the running tensor."""
        spec = self.spec
        emb = self.mllm.model.language_model.embed_tokens(input_ids)
        img_mask = input_ids == spec.image_token_id
        if int(img_mask.sum()) * emb.shape[-1] != image_features.numel():
            raise ValueError(
                f"Image features and image tokens do not match: tokens: {int(img_mask.sum())}, "
                f"features {image_features.shape[:-1].numel()}")
        emb = emb.masked_scatter(img_mask.unsqueeze(-1).expand_as(emb), image_features.to(emb.dtype))
        import torchvision

        tpt = int(math.isqrt(spec.tokens_per_tile))
        D = image_features.shape[-1]
        assert emb.shape[0] == 1, "Currently only support batch_size=1 for feature replay"
        b = 0
        cur = emb[b]

        def replay_rows(fmap, bbox):
            roi, scale = roi_box_f32(bbox, fmap.shape[2], fmap.shape[3])
            feats = torchvision.ops.roi_align(fmap.float(), roi.unsqueeze(0).to(fmap.device), (tpt, tpt),
                                              spatial_scale=scale, sampling_ratio=2, aligned=True)
            return feats.permute(0, 2, 3, 1).flatten(1, 2).to(fmap.dtype).squeeze(0)

        if not video:
            ncw, nch = int(aspect_ratios[b][0]), int(aspect_ratios[b][1])
            for crop_token in crop_tokens:
                pos = (input_ids[b] == crop_token).nonzero().flatten()
                if pos.numel() == 0:
                    continue
                head, tail = int(pos.min()), int(pos.max())
                T1 = image_features.shape[0] - 1
                assert T1 == ncw * nch, f"{ncw * nch} != {T1}"
                tiles = image_features[1:].reshape(nch, ncw, tpt, tpt, D)
                fmap = tiles.permute(4, 0, 2, 1, 3).reshape(1, D, nch * tpt, ncw * tpt)
                # :736-742 -- spliced into emb[b], NOT into `cur`: earlier replays are discarded
                cur = torch.cat([emb[b][:head], replay_rows(fmap, bboxes[b][str(crop_token)]), emb[b][tail + 1:]])
        else:
            for f in range(image_features.shape[0]):
                crop_token = frame_crop_ids[f]
                pos = (input_ids[b] == crop_token).nonzero().flatten()
                if pos.numel() == 0:
                    continue
                head, tail = int(pos.min()), int(pos.max())
                fmap = image_features[f].reshape(tpt, tpt, D).permute(2, 0, 1).unsqueeze(0)
                cur = torch.cat([cur[:head], replay_rows(fmap, bboxes[b][str(crop_token)]), cur[tail + 1:]])
        return cur.unsqueeze(0)

    # -- full path -----------------------------------------------------------------
    @torch.no_grad()
    def prefill_embeds(self, pixel_values, global_mask_values, aspect_ratios, bboxes, input_ids):
        dt = self.mask_patch_embedding.weight.dtype
        pixel_values = pixel_values.to(dt)
        me = self.mask_embeds(global_mask_values) if global_mask_values is not None else None
        feats = self.image_features(pixel_values, me)
        emb = self.assemble(input_ids, feats, aspect_ratios, bboxes)
        return dict(mask_embeds=me, image_features=feats, inputs_embeds=emb)

    @torch.no_grad()
    def greedy(self, inputs_embeds, max_new_tokens: int, eos_token_id: Optional[int] = None,
               return_logits: bool = False):
        """Manual greedy loop (HF GenerationMixin greedy semantics):
        returns only the NEW tokens (quirk 9)."""
        lm = self.mllm.model.language_model
        S = inputs_embeds.shape[1]
        caches = [dict() for _ in lm.layers]
        dev = inputs_embeds.device
        h = lm(inputs_embeds, torch.arange(S, device=dev), caches)
        logits = self.mllm.lm_head(h[:, -1:, :])
        first_logits = logits[:, -1].clone()
        toks = []
        all_logits = [first_logits]
        for i in range(max_new_tokens):
            nxt = logits[:, -1].argmax(-1)
            toks.append(nxt)
            if eos_token_id is not None and int(nxt[0]) == eos_token_id:
                break
            if i == max_new_tokens - 1:
                break
            e = lm.embed_tokens(nxt)[:, None, :]
            h = lm(e, torch.tensor([S + i], device=dev), caches)
            logits = self.mllm.lm_head(h[:, -1:, :])
            if return_logits:
                all_logits.append(logits[:, -1].clone())
        seq = torch.stack(toks, 1)
        if return_logits:
            return seq, first_logits, all_logits
        return seq, first_logits

    @torch.no_grad()
    def generate(self, pixel_values, global_mask_values, aspect_ratios, bboxes, input_ids,
                 max_new_tokens=128, eos_token_id=None):
        taps = self.prefill_embeds(pixel_values, global_mask_values, aspect_ratios, bboxes, input_ids)
        seq, logits = self.greedy(taps["inputs_embeds"], max_new_tokens, eos_token_id)
        taps["sequences"] = seq
        taps["logits"] = logits
        return taps


def init_weights_(model: nn.Module, seed: int = 0) -> None:
    """Deterministic random init (SURVEY.md §8d 'Weights'): normal(0, 0.02) for Linear / Embedding /
    conv / pos_embed / cls; norm weights 1, biases small-random (so bias paths are exercised);
    LayerScale 0.1; mask conv random too."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("gamma_1") or name.endswith("gamma_2"):
                p.fill_(0.1)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))


def build_oracle(spec: GARSpec, seed: int = 0, dtype=torch.float32, device="cpu") -> GAROracle:
    m = GAROracle(spec)
    init_weights_(m, seed)
    return m.to(device=device, dtype=dtype).eval()
