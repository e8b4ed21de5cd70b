"""``GARConfig`` with the reference's field names (reference:
projects/grasp_any_region/hf_models/configuration_gar.py:10-76) plus ``.spec`` -- the resolved
architecture constants the B200 engine is sized from."""
from __future__ import annotations

import copy
from typing import Optional

from transformers.configuration_utils import PretrainedConfig

from .config import GARSpec, TextSpec, VisionSpec

# timm architecture name -> PE-ViT hyper-parameters (SURVEY.md §8 model constants)
_TIMM_ARCH = {
    "vit_pe_core_large_patch14_336": dict(embed_dim=1024, depth=23, num_heads=16, mlp_dim=4096, use_cls_token=True),
    "vit_pe_core_gigantic_patch14_448": dict(embed_dim=1536, depth=47, num_heads=16, mlp_dim=8960, use_cls_token=False),
}


def _vision_spec(vc: dict, use_cls: Optional[bool]) -> VisionSpec:
    args = dict(vc.get("model_args") or {})
    # PE-Core towers inside Perception-LM run without timm's post-transformer norm (`norm` = Identity, SURVEY.md
    # Appendix A); a checkpoint that asks for it needs a kernel this path does not have.
    if args.get("use_post_transformer_norm", False):
        raise ValueError("vision_config.model_args.use_post_transformer_norm=True is not supported by the B200 path")
    base = dict(_TIMM_ARCH.get(vc.get("architecture", ""), {}))
    for k in ("embed_dim", "depth", "num_heads"):
        if k in args:
            base[k] = args[k]
    if "mlp_dim" in args:
        base["mlp_dim"] = args["mlp_dim"]
    elif "mlp_ratio" in args and "embed_dim" in base:
        base["mlp_dim"] = int(round(base["embed_dim"] * args["mlp_ratio"]))
    if use_cls is not None:
        base["use_cls_token"] = bool(use_cls)
    img = args.get("img_size", [448, 448])
    ref = args.get("ref_feat_shape", [32, 32])
    img = img[0] if isinstance(img, (list, tuple)) else img
    ref = ref[0] if isinstance(ref, (list, tuple)) else ref
    return VisionSpec(img_size=img, ref_feat_shape=ref, patch_size=img // ref, **base)


def _text_spec(tc: dict) -> TextSpec:
    rp = tc.get("rope_parameters") or tc.get("rope_scaling") or {}
    hidden, heads = tc.get("hidden_size", 2048), tc.get("num_attention_heads", 32)
    return TextSpec(
        hidden_size=hidden, num_layers=tc.get("num_hidden_layers", 16), num_heads=heads,
        num_kv_heads=tc.get("num_key_value_heads", heads), head_dim=tc.get("head_dim") or hidden // heads,
        intermediate_size=tc.get("intermediate_size", 8192), vocab_size=tc.get("vocab_size", 128262),
        rms_eps=tc.get("rms_norm_eps", 1e-5),
        rope_theta=rp.get("rope_theta", tc.get("rope_theta", 500000.0)), rope_factor=rp.get("factor", 32.0),
        rope_low_freq_factor=rp.get("low_freq_factor", 1.0), rope_high_freq_factor=rp.get("high_freq_factor", 4.0),
        rope_original_max_pos=rp.get("original_max_position_embeddings", 8192),
        max_position_embeddings=tc.get("max_position_embeddings", 131072))


class GARConfig(PretrainedConfig):
    model_type = "GAR"
    is_composition = True

    def __init__(self, mllm_config=None, prompt_numbers=5,
                 crop_tokens_ids=(128004, 128005, 128008, 128010, 128011), use_flash_attn=True, **kwargs):
        super().__init__(**kwargs)
        self.mllm_config = copy.deepcopy(mllm_config) if mllm_config is not None else {}
        self.prompt_numbers = prompt_numbers
        self.crop_tokens_ids = list(crop_tokens_ids)
        assert len(self.crop_tokens_ids) == self.prompt_numbers, (
            f"{self.crop_tokens_ids} crop_tokens_ids length should be {self.prompt_numbers}")
        self.use_flash_attn = use_flash_attn
        spec = self.spec
        self.patch_size_h = self.patch_size_w = spec.vision.patch_size
        self.kernel_size = [self.patch_size_h, self.patch_size_w]
        self.mask_path_embedding_out_channels = spec.vision.embed_dim

    @property
    def spec(self) -> GARSpec:
        mc = self.mllm_config if isinstance(self.mllm_config, dict) else self.mllm_config.to_dict()
        vision = _vision_spec(mc.get("vision_config") or {}, mc.get("vision_use_cls_token"))
        text = _text_spec(mc.get("text_config") or {})
        return GARSpec(name=mc.get("name", "gar"), vision=vision, text=text,
                       projector_pooling_ratio=mc.get("projector_pooling_ratio", 2),
                       image_token_id=mc.get("image_token_id", 128002),
                       video_token_id=mc.get("video_token_id", 128003),
                       prompt_numbers=self.prompt_numbers, crop_tokens_ids=list(self.crop_tokens_ids))

    @classmethod
    def from_spec(cls, spec: GARSpec) -> "GARConfig":
        v, t = spec.vision, spec.text
        mllm = dict(
            name=spec.name,
            vision_config=dict(architecture="pe_vit", model_args=dict(
                img_size=[v.img_size] * 2, ref_feat_shape=[v.ref_feat_shape] * 2, embed_dim=v.embed_dim,
                depth=v.depth, num_heads=v.num_heads, mlp_dim=v.mlp_dim)),
            text_config=dict(model_type="llama", hidden_size=t.hidden_size, num_hidden_layers=t.num_layers,
                             num_attention_heads=t.num_heads, num_key_value_heads=t.num_kv_heads,
                             head_dim=t.head_dim, intermediate_size=t.intermediate_size, vocab_size=t.vocab_size,
                             rms_norm_eps=t.rms_eps, max_position_embeddings=t.max_position_embeddings,
                             rope_parameters=dict(rope_type="llama3", rope_theta=t.rope_theta, factor=t.rope_factor,
                                                  low_freq_factor=t.rope_low_freq_factor,
                                                  high_freq_factor=t.rope_high_freq_factor,
                                                  original_max_position_embeddings=t.rope_original_max_pos)),
            vision_use_cls_token=v.use_cls_token, projector_pooling_ratio=spec.projector_pooling_ratio,
            image_token_id=spec.image_token_id, video_token_id=spec.video_token_id)
        return cls(mllm_config=mllm, prompt_numbers=spec.prompt_numbers, crop_tokens_ids=list(spec.crop_tokens_ids))

    def to_dict(self):
        output = copy.deepcopy(self.__dict__)
        output["mllm_config"] = copy.deepcopy(self.mllm_config)
        output["model_type"] = self.__class__.model_type
        return output
