"""Drop-in ``GARModel`` for the reference's region-captioning path, backed by libgar_b200.so.

Mirrors the reference interface (projects/grasp_any_region/hf_models/modeling_gar.py:17-72 ctor and
attributes, :74-95 ``forward(data, data_samples, mode)``, :295-308 ``generate(...)``) with the same
argument meaning and error behaviour, so ``demo/gar_with_mask.py`` / ``gar_relationship.py`` and the
evaluation drivers run unchanged.  Parameters live under the reference's HF state-dict names.

Host code here is plumbing only: dtype/device moves, the box arithmetic the reference does in Python
floats (:367-387), and argument checks.  All math runs in the CUDA library; nothing falls back to
torch ops or the CPU.
"""
from __future__ import annotations

import dataclasses
import weakref
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib
from .config import GARSpec
from .configuration_gar import GARConfig
from .engine import LM, PJ, REPLAY_FRAME, REPLAY_MERGED, REPLAY_MERGED_LAST, VT, Engine

try:  # same base class as the reference when transformers is importable
    from transformers.modeling_utils import PreTrainedModel as _Base

    _HAS_HF = True
except Exception:  # pragma: no cover
    _Base = nn.Module
    _HAS_HF = False


def parameter_shapes(spec: GARSpec) -> Dict[str, tuple]:
    """HF state-dict layout of a GAR checkpoint."""
    v, t = spec.vision, spec.text
    d, D = v.embed_dim, t.hidden_size
    out = {"mask_patch_embedding.weight": (d, 3, v.patch_size, v.patch_size),
           VT + "patch_embed.proj.weight": (d, 3, v.patch_size, v.patch_size),
           VT + "pos_embed": (1, v.tokens, d), VT + "norm_pre.weight": (d,), VT + "norm_pre.bias": (d,)}
    if v.use_cls_token:
        out[VT + "cls_token"] = (1, 1, d)
    for i in range(v.depth):
        b = f"{VT}blocks.{i}."
        out.update({b + "norm1.weight": (d,), b + "norm1.bias": (d,), b + "attn.qkv.weight": (3 * d, d),
                    b + "attn.qkv.bias": (3 * d,), b + "attn.proj.weight": (d, d), b + "attn.proj.bias": (d,),
                    b + "gamma_1": (d,), b + "norm2.weight": (d,), b + "norm2.bias": (d,),
                    b + "mlp.fc1.weight": (v.mlp_dim, d), b + "mlp.fc1.bias": (v.mlp_dim,),
                    b + "mlp.fc2.weight": (d, v.mlp_dim), b + "mlp.fc2.bias": (d,), b + "gamma_2": (d,)})
    out.update({PJ + "linear_1.weight": (D, d), PJ + "linear_1.bias": (D,), PJ + "linear_2.weight": (D, D),
                PJ + "linear_2.bias": (D,), LM + "embed_tokens.weight": (t.vocab_size, D), LM + "norm.weight": (D,),
                "mllm.lm_head.weight": (t.vocab_size, D)})
    q, kv = t.num_heads * t.head_dim, t.num_kv_heads * t.head_dim
    for i in range(t.num_layers):
        b = f"{LM}layers.{i}."
        out.update({b + "input_layernorm.weight": (D,), b + "post_attention_layernorm.weight": (D,),
                    b + "self_attn.q_proj.weight": (q, D), b + "self_attn.k_proj.weight": (kv, D),
                    b + "self_attn.v_proj.weight": (kv, D), b + "self_attn.o_proj.weight": (D, q),
                    b + "mlp.gate_proj.weight": (t.intermediate_size, D),
                    b + "mlp.up_proj.weight": (t.intermediate_size, D),
                    b + "mlp.down_proj.weight": (D, t.intermediate_size)})
    return out


class _Node(nn.Module):
    """Parameter container reproducing the reference's module paths (no forward of its own)."""


class _MLLMNode(_Node):
    """``model.mllm``: the parameters of the vendored PerceptionLMForConditionalGeneration plus its ``generate`` with the
    replay branches of ``prepare_inputs_for_generation`` (GARModel._mllm_generate)."""

    def generate(self, *args, **kwargs):
        return self._owner()._mllm_generate(*args, **kwargs)


def _attach(root: nn.Module, dotted: str, shape, dtype):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Node())
        mod = getattr(mod, p)
    mod.register_parameter(parts[-1], nn.Parameter(torch.empty(shape, dtype=dtype), requires_grad=False))


@dataclasses.dataclass
class GenerateOutput:
    """What callers read from HF's GenerateDecoderOnlyOutput: only the NEW tokens (quirk 9)."""
    sequences: torch.LongTensor


class GARModel(_Base):
    config_class = GARConfig
    main_input_name = "pixel_values"
    base_model_prefix = "language_model"
    _no_split_modules = ["LlamaDecoderLayer"]
    _supports_flash_attn_2 = True
    supports_gradient_checkpointing = True

    def __init__(self, config: GARConfig, mllm=None, mask_patch_embedding=None, use_flash_attn=True):
        if _HAS_HF:
            super().__init__(config)
        else:  # pragma: no cover
            super().__init__()
            self.config = config
        self.prompt_numbers = config.prompt_numbers
        self.crop_tokens_ids = config.crop_tokens_ids
        self.spec = config.spec
        node = _MLLMNode()
        object.__setattr__(node, "_owner", weakref.ref(self))  # not a submodule link: no cycle in .modules()
        self.add_module("mllm", node)
        for name, shape in parameter_shapes(self.spec).items():
            _attach(self, name, shape, torch.bfloat16)
        if mllm is not None or mask_patch_embedding is not None:
            # reference ctor lets callers inject modules; accept anything exposing a state_dict
            sd = {}
            if mllm is not None:
                sd.update({"mllm." + k: v for k, v in mllm.state_dict().items()})
            if mask_patch_embedding is not None:
                sd.update({"mask_patch_embedding." + k: v for k, v in mask_patch_embedding.state_dict().items()})
            self.load_state_dict(sd, strict=False)
        self._engine: Optional[Engine] = None
        self._engine_caps = dict(max_tiles=37, max_seq=12288, max_batch=8, max_new_tokens=1024, max_prefill_rows=0)
        if _HAS_HF and hasattr(self, "post_init"):
            self.post_init()

    # ------------------------------------------------------------------ reference surface
    @property
    def lm_head(self):
        return self.mllm.lm_head

    def get_input_embeddings(self):
        return self.mllm.model.language_model.embed_tokens

    def get_output_embeddings(self):
        return self.mllm.lm_head

    def _init_weights(self, module):  # parameters always come from a checkpoint / random_init_
        pass

    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        """HF loader (the path every reference caller takes, demo/gar_with_mask.py:84-94).  A checkpoint that lacks
        any parameter of the path is an error here: there is no initialiser to fall back to, and a silently
        uninitialised tensor would only show up as garbage captions."""
        want_info = kwargs.pop("output_loading_info", False)
        # Every reference caller passes device_map="cuda:0".  A single-device map needs no dispatch machinery (and
        # transformers would demand `accelerate` for it): load, then move the whole model -- the engine is one replica
        # per GPU by design.  Anything that really shards a model across devices is refused.
        device = None
        dm = kwargs.pop("device_map", None)
        if dm is not None:
            if isinstance(dm, dict):
                targets = set(dm.values())
                if len(targets) != 1:
                    raise ValueError(f"device_map must name ONE device (one full replica per GPU), got {dm}")
                dm = targets.pop()
            if dm in ("auto", "balanced", "balanced_low_0", "sequential"):
                dm = "cuda:0" if torch.cuda.is_available() else "cpu"
            device = torch.device(f"cuda:{dm}" if isinstance(dm, int) else dm)
        out = super().from_pretrained(*args, output_loading_info=True, **kwargs)
        model, info = out if isinstance(out, tuple) else (out, {})
        if device is not None:
            model = model.to(device)
        missing = sorted(info.get("missing_keys", []) or [])
        if missing:
            raise ValueError(f"checkpoint is missing {len(missing)} parameter(s) of the GAR path, e.g. {missing[:4]}")
        return (model, info) if want_info else model

    # The engine holds fused / padded private copies of some weights: anything that changes or moves the parameters
    # (load_state_dict, .to(), .cuda(), .half() ...) drops it, and the next call rebuilds it from the live tensors.
    def _drop_engine(self):
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.close()
            self._engine = None

    def _apply(self, fn, *args, **kwargs):
        self._drop_engine()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._drop_engine()
        return super().load_state_dict(*args, **kwargs)

    @property
    def dtype(self):
        return torch.bfloat16

    # ------------------------------------------------------------------ engine management
    def random_init_(self, seed: int = 0) -> "GARModel":
        """Deterministic random weights of the reference architecture (benchmarks have no checkpoint)."""
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name.endswith("gamma_1") or name.endswith("gamma_2"):
                    p.fill_(0.1)
                elif "norm" in name and name.endswith("weight"):
                    p.copy_((1.0 + 0.05 * torch.randn(p.shape, generator=g)).to(p.dtype))
                else:
                    p.copy_((0.02 * torch.randn(p.shape, generator=g)).to(p.dtype))
        self._drop_engine()
        return self

    def configure_engine(self, **caps) -> "GARModel":
        """Capacity of the device arenas: max_tiles, max_seq, max_batch, max_new_tokens, max_prefill_rows
        (max_tiles >= n*T and max_prefill_rows >= n*S let generate_batch encode n requests in one pass)."""
        self._engine_caps.update(caps)
        self._drop_engine()
        return self

    def engine(self) -> Engine:
        if self._engine is None:
            p = next(self.parameters())
            if p.device.type != "cuda":
                raise _lib.GarError("GARModel must live on a CUDA device (model.to('cuda:0')): no CPU path exists")
            if p.dtype != torch.bfloat16:
                raise _lib.GarError(f"the B200 engine computes in bf16; got parameters in {p.dtype}")
            self._engine = Engine(self.spec, self.state_dict(), device=str(p.device), **self._engine_caps)
        return self._engine

    # ------------------------------------------------------------------ request plumbing
    def _prepare_request(self, eng: Engine, pixel_values, global_mask_values, aspect_ratios, bboxes, input_ids,
                         batch_idx=0):
        """Device-side copies of one request in the dtypes the engine reads (bf16 / int64).  The (tiny) ids go
        first so the row plan can validate the prompt while the pixel copies are still in flight."""
        dev = eng.device
        if input_ids.shape[0] != 1:
            raise AssertionError("Currently only support batch_size=1")
        ids = input_ids[batch_idx].to(dev, non_blocking=True).to(torch.long).contiguous()
        px = pixel_values.to(dev, non_blocking=True).to(torch.bfloat16)
        if px.dim() == 5:
            px = px.flatten(0, 1)
        assert px.dim() == 4, (
            f"pixel_values should be of shape (batch_size * num_tiles, channels, height, width). But got {px.shape}.")
        mv = None
        if global_mask_values is not None:
            # the reference decodes ids in the tensor's own dtype (bf16 after the dataset's cast, quirk 7)
            mv = global_mask_values.to(dev, non_blocking=True).to(torch.bfloat16).contiguous()
        ar = aspect_ratios[batch_idx]
        return px.contiguous(), mv, ids, bboxes[batch_idx], (int(ar[0]), int(ar[1]))

    def _prepare_raw_request(self, eng: Engine, r: dict):
        """A request that still carries the uint8 image and uint8 region-id map (preprocess_device.build_raw_request):
        ~4 MB cross PCIe instead of ~41 MB of bf16 tiles; resize / tiling / normalisation run on the device, bit-identical
        to the host builder."""
        from .preprocess_device import DevicePreprocessor

        dev = eng.device
        if r["input_ids"].shape[0] != 1:
            raise AssertionError("Currently only support batch_size=1")
        pre = getattr(self, "_device_pre", None)
        if pre is None or pre.device != dev:
            pre = self._device_pre = DevicePreprocessor(self.spec, dev)
        ids = r["input_ids"][0].to(dev, non_blocking=True).to(torch.long).contiguous()
        img = r["image_u8"].to(dev, non_blocking=True)
        idm = r["idmap_u8"].to(dev, non_blocking=True) if r.get("idmap_u8") is not None else None
        px, mv, (ncw, nch) = pre(img, idm, int(r.get("max_num_tiles", 16)))
        return px, mv, ids, r["bboxes"][0], (ncw, nch)

    @staticmethod
    def _raise_on_status_word(st: int):
        if st & 1:
            raise ValueError("Image features and image tokens do not match")
        if st & 4:
            raise KeyError("bboxes has no entry for a crop token present in input_ids")
        if st & 2:
            raise AssertionError("each crop-token run must be 256 contiguous tokens")

    def _prepare(self, eng: Engine, requests: List[dict]):
        """Host -> device copies and (for raw uint8 requests) the device preprocessing of `requests` on the CURRENT stream;
        returns the per-request tuples `_stage` consumes."""
        return [self._prepare_raw_request(eng, r) if "image_u8" in r else
                self._prepare_request(eng, r["pixel_values"], r.get("global_mask_values"), r["aspect_ratios"],
                                      r["bboxes"], r["input_ids"]) for r in requests]

    def _stage(self, eng: Engine, slot0: int, requests: List[dict], prepared=None, ready=None) -> None:
        """ViT -> assemble -> prefill of `requests` into KV slots slot0.. on the engine stream.  Consecutive requests
        of equal shape (tile count, prompt length) share one ViT pass and one prefill pass as far as the engine's
        arenas allow; the rest go one by one.  Each pass first runs the row plan and raises the reference's
        exceptions (ValueError / KeyError / AssertionError) BEFORE any feature or logit of that pass is computed.
        prepared / ready: the output of `_prepare` made earlier on another stream and the event that marks it done."""
        if prepared is None:
            prepared = self._prepare(eng, requests)
            eng.stream.wait_stream(torch.cuda.current_stream(eng.device))
        else:
            eng.stream.wait_event(ready)
        i = 0
        with torch.cuda.stream(eng.stream):
            while i < len(prepared):
                T, S, masked = prepared[i][0].shape[0], prepared[i][2].numel(), prepared[i][1] is not None
                fit = max(1, min(eng.max_tiles // T, eng.max_prefill_rows // S))
                j = i + 1
                while (j < len(prepared) and j - i < fit and prepared[j][0].shape[0] == T
                       and prepared[j][2].numel() == S and (prepared[j][1] is not None) == masked):
                    j += 1
                run = prepared[i:j]
                for st in eng.plan_batch(slot0 + i, run):
                    self._raise_on_status_word(st)
                eng.request_batch(slot0 + i, run)
                i = j
        for p in prepared:
            for t in (p[0], p[1], p[2]):
                if t is not None:
                    t.record_stream(eng.stream)

    @staticmethod
    def _generation_args(generation_config, generate_kwargs):
        max_new, eos, do_sample = 20, None, False
        if generation_config is not None:
            max_new = getattr(generation_config, "max_new_tokens", None) or max_new
            eos = getattr(generation_config, "eos_token_id", None)
            do_sample = bool(getattr(generation_config, "do_sample", False))
        max_new = generate_kwargs.pop("max_new_tokens", max_new)
        eos = generate_kwargs.pop("eos_token_id", eos)
        if do_sample or generate_kwargs.pop("do_sample", False):
            raise NotImplementedError(
                "only greedy decoding (do_sample=False) is implemented, as every reference caller uses")
        if isinstance(eos, (list, tuple)):
            eos = [int(x) for x in eos]
        return int(max_new), eos

    @torch.no_grad()
    def generate(self, pixel_values: Optional[torch.FloatTensor] = None,
                 global_mask_values: Optional[torch.Tensor] = None,
                 aspect_ratios: Optional[torch.Tensor] = None, bboxes=None,
                 input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                 generation_config=None, output_hidden_states: Optional[bool] = None,
                 return_dict: Optional[bool] = None, **generate_kwargs) -> GenerateOutput:
        if pixel_values is None:
            raise NotImplementedError("text-only generate() is outside the region-captioning hot path")
        if attention_mask is not None and not bool((attention_mask != 0).all()):
            # every reference caller passes all ones (eval_dataset.py:143); padding would change the prefill
            raise NotImplementedError("attention_mask with padded positions is not supported (batch 1, no padding)")
        max_new, eos = self._generation_args(generation_config, generate_kwargs)
        eng = self.engine()
        if max_new > eng.max_new_tokens:
            self.configure_engine(max_new_tokens=max_new)
            eng = self.engine()
        self._stage(eng, 0, [dict(pixel_values=pixel_values, global_mask_values=global_mask_values,
                                  aspect_ratios=aspect_ratios, bboxes=bboxes, input_ids=input_ids)])
        toks = eng.generate(1, max_new, eos)
        seq = torch.tensor([toks[0]], dtype=torch.long, device=eng.device)
        return GenerateOutput(sequences=seq)

    @torch.no_grad()
    def _mllm_generate(self, input_ids=None, pixel_values=None, mask_embeds=None, attention_mask=None,
                       feature_replay=None, feature_replay_video=None, crop_tokens=(128004,), bboxes=None,
                       aspect_ratios=None, processor=None, generation_config=None, global_mask_values=None,
                       pixel_values_videos=None, **generate_kwargs) -> GenerateOutput:
        """``model.mllm.generate(...)`` with the replay branches of the vendored PLM's ``prepare_inputs_for_generation``
        (models/modeling/modeling_perception_lm.py:574-860), same keyword names:

        * ``feature_replay=True``: merged-map RoI replay over ``crop_tokens`` -- reproducing the reference's splice into
          the ORIGINAL embeddings (:736-742): only the last crop token present keeps its replay;
        * ``feature_replay_video=True``: the tiles of ``pixel_values`` are frames; frame f's crop token
          ``processor.tokenizer.convert_tokens_to_ids("<|reserved_special_token_{2+f}|>")`` is replayed from frame f's
          own 16x16 map (:765-852);
        * neither: image-token scatter only.
        Returns HF's decoder-only ``.sequences`` for a call that passed ``input_ids``: prompt ids followed by the new ones.
        ``global_mask_values`` (B200 extension) feeds the engine's mask conv; a precomputed ``mask_embeds`` tensor is not
        accepted (the engine fuses that conv into the patch embedding)."""
        assert not (feature_replay and feature_replay_video)
        if pixel_values is None or input_ids is None:
            raise NotImplementedError("mllm.generate needs input_ids and pixel_values on this path")
        if mask_embeds is not None:
            raise NotImplementedError("pass global_mask_values=...; a precomputed mask_embeds tensor is not supported")
        if pixel_values_videos is not None:
            raise NotImplementedError("pixel_values_videos: frames go in as pixel_values (what the replay branch reads, :772)")
        if attention_mask is not None and not bool((attention_mask != 0).all()):
            raise NotImplementedError("attention_mask with padded positions is not supported (batch 1, no padding)")
        assert input_ids.shape[0] == 1, "Currently only support batch_size=1 for feature replay"
        T = pixel_values.shape[0] if pixel_values.dim() == 4 else pixel_values.shape[0] * pixel_values.shape[1]
        if feature_replay_video:
            assert processor is not None, "Need processor"
            if T > 8:
                raise NotImplementedError(f"per-frame replay supports at most 8 frames per request, got {T}")
            ids = [int(processor.tokenizer.convert_tokens_to_ids(f"<|reserved_special_token_{2 + f}|>")) for f in range(T)]
            mode = REPLAY_FRAME
            if aspect_ratios is None or aspect_ratios is True:
                aspect_ratios = [[1, 1]]
        elif feature_replay:
            ids, mode = [int(c) for c in crop_tokens], REPLAY_MERGED_LAST
        else:
            ids, mode = [], REPLAY_MERGED
            if aspect_ratios is None or aspect_ratios is True:
                aspect_ratios = [[1, 1]]
            bboxes = bboxes if bboxes is not None else [{}]
        if len(ids) > 8:
            raise NotImplementedError(f"at most 8 crop tokens per request, got {len(ids)}")
        max_new, eos = self._generation_args(generation_config, generate_kwargs)
        eng = self.engine()
        if max_new > eng.max_new_tokens:
            self.configure_engine(max_new_tokens=max_new)
            eng = self.engine()
        eng.set_replay(mode, ids)
        try:
            self._stage(eng, 0, [dict(pixel_values=pixel_values, global_mask_values=global_mask_values,
                                      aspect_ratios=aspect_ratios, bboxes=bboxes, input_ids=input_ids)])
            toks = eng.generate(1, max_new, eos)
        finally:
            eng.set_replay()
        new = torch.tensor([toks[0]], dtype=torch.long, device=eng.device)
        return GenerateOutput(sequences=torch.cat([input_ids.to(eng.device).to(torch.long), new], dim=1))

    @torch.no_grad()
    def generate_batch(self, requests: List[dict], max_new_tokens: int = 128, eos_token_id=None):
        """B200 scheduler extension: independent requests (each the dict the reference's dataset emits)
        share one batched decode.  Per-request results equal the batch-1 ``generate`` results."""
        eng = self.engine()
        if len(requests) > eng.max_batch:
            raise ValueError(f"{len(requests)} requests > engine max_batch {eng.max_batch}")
        self._stage(eng, 0, requests)
        return eng.generate(len(requests), int(max_new_tokens), eos_token_id)

    @torch.no_grad()
    def generate_sharded(self, requests: List[dict], max_new_tokens: int = 128, eos_token_id=None,
                         already_sharded: bool = False) -> List[List[int]]:
        """Data-parallel serving over the process group (one process per GPU, a full replica each): every
        rank passes the SAME global request list and decodes its contiguous shard -- or, with already_sharded=True, each
        rank passes only its own shard (equal sizes) -- and the path's single collective, an all-gather of the fixed-shape
        token buffer (NCCL over NVLink), returns every caption's ids in global request order on every rank.  Without an
        initialised process group this is generate_batch."""
        import torch.distributed as dist

        from .parallel import gather_tokens, pack_tokens, shard_range

        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        n_global = len(requests) * world if already_sharded else len(requests)
        lo, hi = (0, len(requests)) if already_sharded else shard_range(n_global, rank, world)
        eng = self.engine()
        local: List[List[int]] = []
        for i in range(lo, hi, eng.max_batch):
            local += self.generate_batch(requests[i:min(hi, i + eng.max_batch)], max_new_tokens, eos_token_id)
        rows = (n_global + world - 1) // world
        return gather_tokens(pack_tokens(local, rows, int(max_new_tokens), eng.device), n_global)

    @torch.no_grad()
    def generate_batches(self, batches: List[List[dict]], max_new_tokens: int = 128,
                         eos_token_id=None, overlap: str = "encode", ways: int = 2) -> List[List[List[int]]]:
        """Throughput serving of a stream of request batches (B200 scheduler extension).

        overlap="decode" (the bench's mode): `ways` consecutive batches are encoded back to back (ViT + assemble +
        prefill, tensor-bound, each alone on the GPU) and then DECODED CONCURRENTLY -- every batch in its own engine
        twin (own KV cache, workspaces, stream and step graph; one shared copy of the weights), driven by its own host
        thread.  A decode step is a chain of ~80 dependent launches that reaches 0.58 of the HBM peak by itself; a
        second and third independent chain fill its launch / drain gaps (measured on B200, 8 x 4.7k-token contexts:
        0.98 ms per step alone, 0.78 with two groups in flight, 0.71 with three).  Nothing is shared between the
        groups but read-only weights, so every request's tokens equal its batch-1 result.  The next round's host ->
        device copies and device preprocessing run under the current round's decode as in overlap="prepare".

        overlap="prepare": only the NEXT batch's host -> device copies and device preprocessing (uint8 image / id map ->
        bf16 tiles) run under the current batch's decode, on their own stream from a helper thread; ViT, prefill and
        decode of the batches stay back to back on the engine stream (one KV-slot group).  The copies and the
        preprocessing launches are latency- and host-bound, the decode leaves the host idle: this hides them for free.

        overlap="encode":

        Two KV-slot groups ping-pong: while group g decodes on the engine's decode stream, the next batch's
        ViT + assemble + prefill is enqueued (from a helper thread) on the encode stream into group 1-g.  The decode
        is an HBM-bound weight stream and the encoder is tensor-bound, so they are complementary; with
        ``engine().enable_partition(k)`` the two streams own disjoint SM sets (k SMs decode, the rest encode) and
        really run side by side instead of interleaving at kernel boundaries.  Every request's result equals its
        batch-1 ``generate`` result."""
        import threading

        eng = self.engine()
        if not batches:
            return []
        if overlap == "prepare":
            return self._generate_batches_prepare(eng, batches, max_new_tokens, eos_token_id, threading)
        if overlap == "decode":
            return self._generate_batches_decode(eng, batches, max_new_tokens, eos_token_id, threading, int(ways))
        if overlap != "encode":
            raise ValueError(f"overlap must be 'encode', 'prepare' or 'decode', got {overlap!r}")
        G = max(len(b) for b in batches)
        if 2 * G > eng.max_batch:
            raise ValueError(f"pipelining needs engine max_batch >= 2 x batch ({2 * G}), got {eng.max_batch}")
        dev = eng.device
        plain = (eng.stream, eng.decode_stream)
        if eng.enc_stream is not None:
            # the partition streams take over for the duration of the call (they see everything the plain ones did)
            eng.enc_stream.wait_stream(eng.stream)
            eng.dec_stream.wait_stream(eng.decode_stream)
            eng.stream, eng.decode_stream = eng.enc_stream, eng.dec_stream
        try:
            return self._generate_batches(eng, batches, G, dev, max_new_tokens, eos_token_id, threading)
        finally:
            if eng.stream is not plain[0]:
                plain[0].wait_stream(eng.stream)
                plain[1].wait_stream(eng.decode_stream)
            eng.stream, eng.decode_stream = plain

    def _generate_batches_prepare(self, eng, batches, max_new_tokens, eos_token_id, threading):
        dev = eng.device
        if max(len(b) for b in batches) > eng.max_batch:
            raise ValueError(f"batch larger than the engine's max_batch {eng.max_batch}")
        prep = getattr(self, "_prep_stream", None)
        if prep is None or prep.device != dev:
            prep = self._prep_stream = torch.cuda.Stream(device=dev)

        def prepare(bi: int):
            torch.cuda.set_device(dev)
            with torch.cuda.stream(prep):
                prepared = self._prepare(eng, batches[bi])
                ev = torch.cuda.Event()
                ev.record(prep)
            return prepared, ev

        prep.wait_stream(torch.cuda.current_stream(dev))
        out: List[List[List[int]]] = []
        prepared, ev = prepare(0)
        for i, batch in enumerate(batches):
            self._stage(eng, 0, batch, prepared, ev)
            box, th = {}, None
            if i + 1 < len(batches):
                def work(bi=i + 1):
                    try:
                        box["p"] = prepare(bi)
                    except BaseException as ex:  # surfaced on the caller's thread
                        box["err"] = ex
                th = threading.Thread(target=work)
                th.start()
            out.append(eng.generate(len(batch), int(max_new_tokens), eos_token_id))
            if th is not None:
                th.join()
                if "err" in box:
                    raise box["err"]
                prepared, ev = box["p"]
        return out

    def _generate_batches_decode(self, eng, batches, max_new_tokens, eos_token_id, threading, ways):
        if ways < 1 or ways > 4:
            raise ValueError(f"ways must be 1..4, got {ways}")
        dev = eng.device
        if max(len(b) for b in batches) > eng.max_batch:
            raise ValueError(f"batch larger than the engine's max_batch {eng.max_batch}")
        engines = [eng] + eng.twins(min(ways, len(batches)) - 1)
        prep = getattr(self, "_prep_stream", None)
        if prep is None or prep.device != dev:
            prep = self._prep_stream = torch.cuda.Stream(device=dev)

        def prepare(first: int, count: int):
            torch.cuda.set_device(dev)
            # ONE event after the whole round's copies / preprocessing: letting the first encode start while the later
            # batches were still being prepared was measured to cost 18 % end to end -- the preprocessing is hundreds of
            # small launches on a second stream, and next to the encoder's persistent kernels each of them only gets
            # SMs at a kernel boundary.  From the second round on all of this runs under the previous round's decode.
            with torch.cuda.stream(prep):
                prepared = [self._prepare(eng, batches[first + k]) for k in range(count)]
                ev = torch.cuda.Event()
                ev.record(prep)
            return prepared, [ev] * count

        def decode(e, n, box, k):
            try:
                torch.cuda.set_device(dev)
                box[k] = e.generate(n, int(max_new_tokens), eos_token_id)
            except BaseException as ex:  # surfaced on the caller's thread
                box["err"] = ex

        prep.wait_stream(torch.cuda.current_stream(dev))
        out: List[List[List[int]]] = []
        i = 0
        count = min(ways, len(batches))
        prepared, ev = prepare(0, count)
        while i < len(batches):
            # encodes back to back: each ViT / prefill pass owns the GPU (persistent tensor-core kernels), the first
            # decode waits for the last encode so that all chains start together
            for k in range(count):
                if k > 0:
                    engines[k].stream.wait_stream(engines[k - 1].stream)
                self._stage(engines[k], 0, batches[i + k], prepared[k], ev[k])
            for k in range(count - 1):
                engines[k].stream.wait_stream(engines[count - 1].stream)
            nxt = min(ways, len(batches) - (i + count))
            pbox, pth = {}, None
            if nxt > 0:
                def work(first=i + count, c=nxt):
                    try:
                        pbox["p"] = prepare(first, c)
                    except BaseException as ex:
                        pbox["err"] = ex
                pth = threading.Thread(target=work)
                pth.start()
            box = {}
            # an engine's very first decode sets kernel attributes and captures its step graph: those run one by one
            cold = [k for k in range(count) if not engines[k].decoded_once]
            for k in cold:
                decode(engines[k], len(batches[i + k]), box, k)
            ths = [threading.Thread(target=decode, args=(engines[k], len(batches[i + k]), box, k))
                   for k in range(1, count) if k not in cold]
            for t in ths:
                t.start()
            if 0 not in cold:
                decode(engines[0], len(batches[i]), box, 0)
            for t in ths:
                t.join()
            if pth is not None:
                pth.join()
            for b in (box, pbox):
                if "err" in b:
                    raise b["err"]
            out.extend(box[k] for k in range(count))
            i += count
            if nxt > 0:
                prepared, ev = pbox["p"]
                count = nxt
        return out

    def _generate_batches(self, eng, batches, G, dev, max_new_tokens, eos_token_id, threading):
        timing = []  # (encode start, encode end, decode start, decode end) events per batch -> eng.pipeline_stats

        def stage(bi: int, group: int):
            torch.cuda.set_device(dev)
            a = torch.cuda.Event(enable_timing=True)
            a.record(eng.stream)
            self._stage(eng, group * G, batches[bi])
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(eng.stream)
            return a, ev

        out: List[List[List[int]]] = []
        e0, ev = stage(0, 0)
        for i, batch in enumerate(batches):
            g = i & 1
            eng.decode_stream.wait_event(ev)
            box, th = {}, None
            if i + 1 < len(batches):
                def work(bi=i + 1, grp=1 - g):
                    try:
                        box["ev"] = stage(bi, grp)
                    except BaseException as ex:  # surfaced on the caller's thread
                        box["err"] = ex
                th = threading.Thread(target=work)
                th.start()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record(eng.decode_stream)
            toks = eng.generate(len(batch), int(max_new_tokens), eos_token_id, slot0=g * G, stream=eng.decode_stream)
            d1.record(eng.decode_stream)
            timing.append((e0, ev, d0, d1))
            if th is not None:
                th.join()
                if "err" in box:
                    raise box["err"]
                e0, ev = box["ev"]
            out.append(toks)
        eng.decode_stream.synchronize()
        eng.pipeline_stats = {"encode_ms": [a.elapsed_time(b) for a, b, _, _ in timing],
                              "decode_ms": [c.elapsed_time(d) for _, _, c, d in timing]}
        return out

    def forward(self, data, data_samples=None, mode="loss"):
        """Teacher-forced LM loss, reference modeling_gar.py:74-246 (only mode="loss" exists there)."""
        if mode != "loss":
            if mode in ("predict", "tensor"):
                raise UnboundLocalError("outputs")  # reference falls through to an unbound name (:239-246)
            raise NotImplementedError
        input_ids = data["input_ids"]
        assert input_ids.shape[0] == 1, "Currently only support batch_size=1"
        eng = self.engine()
        # stage ViT + assemble only, then all-position logits
        dev = eng.device
        px = data["pixel_values"].to(dev).to(torch.bfloat16)
        if px.dim() == 5:
            px = px.flatten(0, 1)
        mv = data["global_mask_values"].to(dev).to(torch.bfloat16).contiguous()
        ids = input_ids[0].to(dev).to(torch.long).contiguous()
        ar = data["aspect_ratios"][0]
        eng.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(eng.stream):
            req = (px.contiguous(), mv, ids, data["bboxes"][0], (int(ar[0]), int(ar[1])))
            for st in eng.plan_batch(0, [req]):
                self._raise_on_status_word(st)
            eng.vit_encode(req[0], mv)
            eng.assemble(0, ids, px.shape[0], data["bboxes"][0], int(ar[0]), int(ar[1]))
            logits = eng.prefill_all_logits(0, ids.numel())
        torch.cuda.current_stream(dev).wait_stream(eng.stream)
        labels = data["labels"][0].to(dev).clone()
        for c in self.crop_tokens_ids:  # replayed rows carry no label (reference :180-191)
            labels[ids == c] = -100
        V = self.spec.text.vocab_size
        # HF ForCausalLMLoss: upcast, shift, mean over non-ignored targets
        lg = logits[:-1, :V].float()
        loss = torch.nn.functional.cross_entropy(lg, labels[1:], ignore_index=-100)
        return {"loss": loss}
