"""Architecture constants for the GAR region-captioning path.

Mirrors the fields the reference reads from ``GARConfig`` / ``PerceptionLMConfig``
(reference: projects/grasp_any_region/hf_models/configuration_gar.py:14-64,
models/modeling/configuration_perception_lm.py:56-86) plus the timm PE-ViT and Llama
hyper-parameters that live in the HF-hub ``config.json`` (SURVEY.md §8 model constants).
Everything is a config input; kernels never hard-code these.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class VisionSpec:
    embed_dim: int = 1024
    depth: int = 23
    num_heads: int = 16
    mlp_dim: int = 4096
    use_cls_token: bool = True
    img_size: int = 448
    patch_size: int = 14
    ref_feat_shape: int = 32  # img_size // patch_size
    ln_eps: float = 1e-5
    layer_scale_init: float = 0.1
    rope_theta: float = 10000.0

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_prefix(self) -> int:
        return 1 if self.use_cls_token else 0

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + self.num_prefix


@dataclass
class TextSpec:
    hidden_size: int = 2048
    num_layers: int = 16
    num_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 64
    intermediate_size: int = 8192
    vocab_size: int = 128262
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_factor: float = 32.0
    rope_low_freq_factor: float = 1.0
    rope_high_freq_factor: float = 4.0
    rope_original_max_pos: int = 8192
    max_position_embeddings: int = 131072


@dataclass
class GARSpec:
    """Everything the engine needs to size its kernels and arenas."""

    name: str = "gar-1b"
    vision: VisionSpec = field(default_factory=VisionSpec)
    text: TextSpec = field(default_factory=TextSpec)
    projector_pooling_ratio: int = 2
    image_token_id: int = 128002
    video_token_id: int = 128003
    prompt_numbers: int = 5
    crop_tokens_ids: List[int] = field(
        default_factory=lambda: [128004, 128005, 128008, 128010, 128011]
    )
    eos_token_id: Optional[int] = 128009
    pad_token_id: Optional[int] = 128001

    @property
    def tokens_per_tile(self) -> int:
        g = self.vision.grid // self.projector_pooling_ratio
        return g * g

    @property
    def kernel_size(self):
        return [self.vision.patch_size, self.vision.patch_size]

    @property
    def mask_path_embedding_out_channels(self) -> int:
        return self.vision.embed_dim

    def to_dict(self):
        return dataclasses.asdict(self)

    @staticmethod
    def from_dict(d) -> "GARSpec":
        d = dict(d)
        d["vision"] = VisionSpec(**d["vision"])
        d["text"] = TextSpec(**d["text"])
        return GARSpec(**d)


def gar_1b() -> GARSpec:
    """GAR-1B = PE-Core-L/14@448 + Llama-3.2-1B."""
    return GARSpec()


def gar_8b() -> GARSpec:
    """GAR-8B = PE-Core-G/14@448 + Llama-3.1-8B."""
    return GARSpec(
        name="gar-8b",
        vision=VisionSpec(embed_dim=1536, depth=47, num_heads=16, mlp_dim=8960, use_cls_token=False),
        text=TextSpec(hidden_size=4096, num_layers=32, num_heads=32, num_kv_heads=8, head_dim=128,
                      intermediate_size=14336, rope_factor=8.0),
    )


def gar_tiny() -> GARSpec:
    """Small architecture with the same structure (head_dim 64, cls token, GQA, 448px tiles,
    real special-token ids) used for CPU-speed parity fixtures."""
    return GARSpec(
        name="gar-tiny",
        vision=VisionSpec(embed_dim=128, depth=2, num_heads=2, mlp_dim=256),
        text=TextSpec(hidden_size=128, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=64,
                      intermediate_size=256),
    )


def gar_small() -> GARSpec:
    """Mid-size architecture for GPU smoke/parity runs that finish in seconds."""
    return GARSpec(
        name="gar-small",
        vision=VisionSpec(embed_dim=256, depth=3, num_heads=4, mlp_dim=1024),
        text=TextSpec(hidden_size=512, num_layers=3, num_heads=8, num_kv_heads=2, head_dim=64,
                      intermediate_size=1536),
    )


def gar_8b_mini() -> GARSpec:
    """GAR-8B's kernel shapes at toy depth: ViT head_dim 96 without cls token, Llama head_dim 128 with GQA."""
    return GARSpec(
        name="gar-8b-mini",
        vision=VisionSpec(embed_dim=192, depth=2, num_heads=2, mlp_dim=512, use_cls_token=False),
        text=TextSpec(hidden_size=512, num_layers=2, num_heads=8, num_kv_heads=2, head_dim=128,
                      intermediate_size=1024, rope_factor=8.0),
    )


PRESETS = {"gar-1b": gar_1b, "gar-8b": gar_8b, "gar-tiny": gar_tiny, "gar-small": gar_small,
           "gar-8b-mini": gar_8b_mini}
