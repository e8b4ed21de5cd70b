"""SURVEY §8(b) loader contract + next-row N2, CPU half: a checkpoint directory laid out the way the reference's
converter writes it (hf_models/convert_to_hf.py:112-135: `auto_map` naming configuration_gar.GARConfig /
modeling_gar.GARModel, a PerceptionLM-shaped `mllm_config`) with INTEGRATION.md route A's two re-export files loads
through `AutoConfig` / `AutoModel.from_pretrained(dir, trust_remote_code=True, torch_dtype=bf16, device_map=...)` --
the call every reference caller makes (demo/gar_with_mask.py:84-94) -- and yields the B200 `GARModel`."""
import json
import os
import shutil

import pytest
import torch

from grasp_any_region_b200 import GARConfig, GARModel, gar_tiny

AUTO_MAP = {"AutoConfig": "configuration_gar.GARConfig", "AutoModel": "modeling_gar.GARModel",
            "AutoModelForCausalLM": "modeling_gar.GARModel"}


def plm_shaped_mllm_config(spec, style="5.x", **model_args_extra):
    """`PerceptionLMConfig(...).to_dict()` built with transformers' own config classes: TimmWrapper vision_config
    (architecture + model_args) and a llama text_config with llama3 rope scaling; style="4.56" rewrites the rope keys
    the way the pinned transformers 4.56.2 serialises them (rope_theta + rope_scaling)."""
    from transformers import LlamaConfig, PerceptionLMConfig
    from transformers.models.timm_wrapper.configuration_timm_wrapper import TimmWrapperConfig

    v, t = spec.vision, spec.text
    vc = TimmWrapperConfig(architecture="vit_pe_core_large_patch14_336", model_args=dict(
        embed_dim=v.embed_dim, depth=v.depth, num_heads=v.num_heads, mlp_ratio=v.mlp_dim / v.embed_dim,
        img_size=[v.img_size] * 2, ref_feat_shape=[v.ref_feat_shape] * 2, init_values=0.1, global_pool="",
        use_post_transformer_norm=False, **model_args_extra))
    rope = dict(rope_type="llama3", rope_theta=t.rope_theta, factor=t.rope_factor, low_freq_factor=t.rope_low_freq_factor,
                high_freq_factor=t.rope_high_freq_factor, original_max_position_embeddings=t.rope_original_max_pos)
    tc = LlamaConfig(hidden_size=t.hidden_size, num_hidden_layers=t.num_layers, num_attention_heads=t.num_heads,
                     num_key_value_heads=t.num_kv_heads, head_dim=t.head_dim, intermediate_size=t.intermediate_size,
                     vocab_size=t.vocab_size, rms_norm_eps=t.rms_eps, max_position_embeddings=t.max_position_embeddings,
                     rope_parameters=rope)
    d = PerceptionLMConfig(vision_config=vc, text_config=tc, vision_use_cls_token=v.use_cls_token,
                           projector_pooling_ratio=spec.projector_pooling_ratio).to_dict()
    if style == "4.56":
        rp = d["text_config"].pop("rope_parameters")
        d["text_config"]["rope_theta"] = rp.pop("rope_theta")
        d["text_config"]["rope_scaling"] = rp
    return d


def write_checkpoint_dir(path, spec, seed=0, style="5.x"):
    cfg = GARConfig(mllm_config=plm_shaped_mllm_config(spec, style), prompt_numbers=spec.prompt_numbers,
                    crop_tokens_ids=list(spec.crop_tokens_ids), auto_map=dict(AUTO_MAP))
    model = GARModel(cfg).random_init_(seed)
    model.save_pretrained(path)
    # INTEGRATION.md route A: the two files the README tells users to copy next to the weights, as re-exports
    with open(os.path.join(path, "modeling_gar.py"), "w") as f:
        f.write("from grasp_any_region_b200.modeling_gar import GARModel  # noqa: F401\n")
    with open(os.path.join(path, "configuration_gar.py"), "w") as f:
        f.write("from grasp_any_region_b200.configuration_gar import GARConfig  # noqa: F401\n")
    return model


@pytest.mark.parametrize("style", ["5.x", "4.56"])
def test_auto_model_resolves_auto_map_to_the_b200_class(tmp_path, style):
    from transformers import AutoConfig, AutoModel

    spec = gar_tiny()
    d = str(tmp_path / f"ckpt_{style}")
    src = write_checkpoint_dir(d, spec, style=style)
    assert json.load(open(os.path.join(d, "config.json")))["auto_map"] == AUTO_MAP
    cfg = AutoConfig.from_pretrained(d, trust_remote_code=True)
    assert type(cfg).__name__ == "GARConfig" and cfg.prompt_numbers == 5
    assert cfg.spec.vision == spec.vision and cfg.spec.text == spec.text      # PLM-shaped dict parsed back exactly
    model = AutoModel.from_pretrained(d, trust_remote_code=True, torch_dtype=torch.bfloat16, device_map="cpu")
    assert type(model).__name__ == "GARModel" and type(model).__module__.endswith("modeling_gar")
    assert model.dtype == torch.bfloat16 and model.config.prompt_numbers == 5
    a, b = src.state_dict(), model.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert model.get_input_embeddings().weight.shape == (spec.text.vocab_size, spec.text.hidden_size)
    assert model.lm_head.weight.shape == (spec.text.vocab_size, spec.text.hidden_size)


def test_from_pretrained_refuses_a_checkpoint_with_missing_parameters(tmp_path):
    import safetensors.torch as st
    from transformers import AutoModel

    d = str(tmp_path / "ckpt")
    write_checkpoint_dir(d, gar_tiny())
    bad = str(tmp_path / "bad")
    os.makedirs(bad)
    for f in ("config.json", "modeling_gar.py", "configuration_gar.py"):
        shutil.copy(os.path.join(d, f), os.path.join(bad, f))
    sd = st.load_file(os.path.join(d, "model.safetensors"))
    del sd["mask_patch_embedding.weight"]
    st.save_file(sd, os.path.join(bad, "model.safetensors"), metadata={"format": "pt"})
    with pytest.raises(ValueError, match="missing 1 parameter"):
        AutoModel.from_pretrained(bad, trust_remote_code=True, torch_dtype=torch.bfloat16)


def test_config_rejects_post_transformer_norm_and_sharded_device_maps(tmp_path):
    spec = gar_tiny()
    mc = plm_shaped_mllm_config(spec)
    mc["vision_config"]["model_args"]["use_post_transformer_norm"] = True
    with pytest.raises(ValueError, match="use_post_transformer_norm"):
        GARConfig(mllm_config=mc)
    d = str(tmp_path / "ckpt")
    write_checkpoint_dir(d, spec)
    with pytest.raises(ValueError, match="ONE device"):
        GARModel.from_pretrained(d, device_map={"mllm": "cpu", "mask_patch_embedding": "cuda:0"})


def test_engine_is_dropped_when_parameters_change():
    """The engine keeps fused private copies of some weights; load_state_dict / .to() must invalidate it."""
    m = GARModel(GARConfig.from_spec(gar_tiny())).random_init_(0)

    class FakeEngine:
        closed = False

        def close(self):
            self.closed = True

    for mutate in (lambda: m.load_state_dict(m.state_dict()), lambda: m.to(torch.bfloat16), lambda: m.random_init_(1),
                   lambda: m.configure_engine(max_batch=2)):
        fake = FakeEngine()
        m._engine = fake
        mutate()
        assert m._engine is None and fake.closed
