"""Host-side constant tables handed to the engine as read-only buffers (computed once at load).

* PE-ViT 2-D axial RoPE table cat(sin, cos) (timm RotaryEmbeddingCat with feat_shape given; the
  reference reaches it through Eva._pos_embed at modeling_perception_lm.py:197).
* Llama-3 scaled rotary cos/sin (HF LlamaRotaryEmbedding.forward + _compute_llama3_parameters).
Both are produced in fp32 and cast to the model dtype, as the reference's modules do.
"""
from __future__ import annotations

import math

import torch

from .config import TextSpec, VisionSpec


def vit_rope_table(v: VisionSpec) -> torch.Tensor:
    hd, g = v.head_dim, v.grid
    nb = hd // 4
    bands = 1.0 / (v.rope_theta ** (torch.arange(nb, dtype=torch.float32) / nb))
    coords = (torch.arange(g, dtype=torch.float32) + 1.0) / g * v.ref_feat_shape
    ax_y, ax_x = torch.meshgrid(coords, coords, indexing="ij")
    ang = torch.cat([ax_x.reshape(-1, 1) * bands, ax_y.reshape(-1, 1) * bands], dim=1)  # (g*g, hd/2): x bands | y bands
    sin = ang.sin().repeat_interleave(2, dim=1)
    cos = ang.cos().repeat_interleave(2, dim=1)
    return torch.cat([sin, cos], dim=1)  # (g*g, 2*hd)


def llama_cos_sin(t: TextSpec, max_pos: int):
    dim = t.head_dim
    inv = 1.0 / (t.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    lo_w = t.rope_original_max_pos / t.rope_low_freq_factor
    hi_w = t.rope_original_max_pos / t.rope_high_freq_factor
    wl = 2 * math.pi / inv
    scaled = torch.where(wl > lo_w, inv / t.rope_factor, inv)
    sm = (t.rope_original_max_pos / wl - t.rope_low_freq_factor) / (t.rope_high_freq_factor - t.rope_low_freq_factor)
    blend = (1 - sm) * scaled / t.rope_factor + sm * scaled
    mid = ~(wl < hi_w) & ~(wl > lo_w)
    inv = torch.where(mid, blend, scaled)
    ang = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv[None, :]  # (max_pos, hd/2)
    return ang.cos(), ang.sin()


def rope_table_chunk_major(table: torch.Tensor, hd: int) -> torch.Tensor:
    """(n_tokens, 2 * hd) = cat(sin, cos)  ->  [sin | cos][hd / 8][n_tokens][8], the layout gar_gemm_bf16_rope2d reads
    (include/gar_b200.h): consecutive tokens are contiguous for every 8-wide piece of the head dimension."""
    n = table.shape[0]
    return table.reshape(n, 2, hd // 8, 8).permute(1, 2, 0, 3).contiguous()
