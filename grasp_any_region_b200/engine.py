"""Python owner of one ``gar_engine`` (include/gar_b200.h): lays the reference's HF state dict out
the way the kernels want it (fused qkv / gate-up, K-padded conv matrices, padded lm_head, RoPE
tables), keeps those tensors alive, and drives the stage entry points.

State-dict key layout follows the reference checkpoint (modeling_gar.py:47-60,
convert_to_hf.py:126-135): ``mllm.model.vision_tower.timm_model.*``,
``mllm.model.multi_modal_projector.linear_{1,2}.*``, ``mllm.model.language_model.*``,
``mllm.lm_head.weight``, ``mask_patch_embedding.weight``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import GarModelDesc, check
from .config import GARSpec
from .rope import llama_cos_sin, rope_table_chunk_major, vit_rope_table

VT = "mllm.model.vision_tower.timm_model."
PJ = "mllm.model.multi_modal_projector."
LM = "mllm.model.language_model."


REPLAY_MERGED, REPLAY_MERGED_LAST, REPLAY_FRAME = 0, 1, 2  # include/gar_b200.h GAR_REPLAY_*


def roi_boxes_f32(spec: GARSpec, bboxes: Dict[str, Sequence[float]], ncw: int, nch: int, crop_ids=None,
                  per_frame: bool = False):
    """Box arithmetic exactly as modeling_gar.py:366-387 does it on the host: python doubles, then one
    rounding to fp32 (torch.tensor(..., dtype=float32)).  Returns (rois flat list, has list, scale).
    per_frame: the map is one tile (modeling_perception_lm.py:787-801, feat_h = feat_w = 16)."""
    tpt = int(math.isqrt(spec.tokens_per_tile))
    if per_frame:
        ncw = nch = 1
    feat_h, feat_w = nch * tpt, ncw * tpt
    orig_h, orig_w = feat_h * 28, feat_w * 28
    spatial_scale = feat_w / orig_w
    rois, has = [], []
    for cid in (spec.crop_tokens_ids if crop_ids is None else crop_ids):
        box = bboxes.get(str(cid)) if bboxes is not None else None
        if box is None:
            rois.append([0.0] * 5)
            has.append(0)
            continue
        x1, y1, x2, y2 = [float(v) for v in box]
        vals = [0.0, x1 * orig_w * spatial_scale, y1 * orig_h * spatial_scale, x2 * orig_w * spatial_scale,
                y2 * orig_h * spatial_scale]
        rois.append(torch.tensor(vals, dtype=torch.float32).tolist())
        has.append(1)
    return rois, has, spatial_scale


class Engine:
    def __init__(self, spec: GARSpec, state_dict: Dict[str, torch.Tensor], device: str = "cuda:0",
                 max_tiles: int = 17, max_seq: int = 6144, max_batch: int = 8, max_new_tokens: int = 128,
                 max_prefill_rows: int = 0, share: Optional["Engine"] = None):
        """share: another Engine of the same spec whose (already fused / padded) weight tensors this one points at
        instead of building its own copies from `state_dict` -- see `twins`."""
        if not torch.cuda.is_available():
            raise _lib.GarError("grasp_any_region_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.spec = spec
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        v, t = spec.vision, spec.text
        self.vocab_padded = ((t.vocab_size + 255) // 256) * 256
        self.max_tiles, self.max_seq, self.max_batch, self.max_new_tokens = max_tiles, max_seq, max_batch, max_new_tokens
        d = GarModelDesc()
        d.vit_dim, d.vit_depth, d.vit_heads, d.vit_mlp = v.embed_dim, v.depth, v.num_heads, v.mlp_dim
        d.vit_use_cls, d.img_size, d.patch_size, d.ln_eps = int(v.use_cls_token), v.img_size, v.patch_size, v.ln_eps
        d.pool_ratio = spec.projector_pooling_ratio
        d.llm_dim, d.llm_layers, d.llm_heads, d.llm_kv_heads = t.hidden_size, t.num_layers, t.num_heads, t.num_kv_heads
        d.llm_head_dim, d.llm_inter, d.vocab, d.vocab_padded = t.head_dim, t.intermediate_size, t.vocab_size, self.vocab_padded
        d.rms_eps = t.rms_eps
        d.prompt_numbers = spec.prompt_numbers
        d.image_token_id = spec.image_token_id
        for i in range(8):
            d.crop_token_ids[i] = spec.crop_tokens_ids[i] if i < len(spec.crop_tokens_ids) else -1
        d.max_tiles, d.max_seq, d.max_batch, d.max_new_tokens = max_tiles, max_seq, max_batch, max_new_tokens
        d.max_prefill_rows = max(int(max_prefill_rows), max_seq)
        self.max_prefill_rows = d.max_prefill_rows
        self._desc = d
        h = C.c_void_p()
        check(self.lib.gar_engine_create(C.byref(d), C.byref(h)), "gar_engine_create")
        self._h = h
        self._keep: Dict[str, torch.Tensor] = {}
        self.stream = torch.cuda.Stream(device=self.device)
        # decode of one request group may overlap the next group's ViT+prefill (see GARModel.generate_batches)
        self.decode_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.enc_stream: Optional[torch.cuda.Stream] = None   # SM-partition streams (enable_partition)
        self.dec_stream: Optional[torch.cuda.Stream] = None
        self.partition_sms = None
        self.last_steps_run = 0
        self.replay_mode, self.crop_ids = REPLAY_MERGED, list(spec.crop_tokens_ids)
        self.trace = None  # debug: set to [] to collect (label, CUDA event) marks between the stage entry points
        self._caps = dict(max_tiles=max_tiles, max_seq=max_seq, max_batch=max_batch, max_new_tokens=max_new_tokens,
                          max_prefill_rows=max_prefill_rows)
        self._twins: List["Engine"] = []
        self._options: Dict[str, int] = {}
        self.decoded_once = False  # the first decode sets kernel attributes and captures the step graph
        if share is not None:
            for name, tensor in share._keep.items():  # same device tensors: one copy of the weights in HBM
                self._keep[name] = tensor
                check(self.lib.gar_engine_set_weight(self._h, name.encode(), C.c_void_p(tensor.data_ptr())),
                      f"set_weight {name}")
            for k, v in share._options.items():
                self.set_option(k, v)
            if share.replay_mode != REPLAY_MERGED or share.crop_ids != list(spec.crop_tokens_ids):
                self.set_replay(share.replay_mode, share.crop_ids)
        else:
            self._load_weights(state_dict)

    def twins(self, n: int) -> List["Engine"]:
        """n more engines over the SAME weight tensors, each with its own workspaces, KV cache, streams and decode graph
        (created on first use, then kept): independent request groups whose decode loops run concurrently
        (GARModel.generate_batches(overlap="decode")).  Options and replay settings set on this engine follow."""
        while len(self._twins) < n:
            self._twins.append(Engine(self.spec, None, device=str(self.device), share=self, **self._caps))
        return self._twins[:n]

    def enable_partition(self, decode_sms: int) -> bool:
        """Split the GPU's SMs into a decode partition (>= decode_sms, driver granularity) and an encode partition
        (the rest); `enc_stream` / `dec_stream` then confine what is enqueued on them to their SM sets, so
        GARModel.generate_batches overlaps decode(i) with ViT + prefill(i+1) on disjoint SMs.  decode_sms <= 0 turns it
        off.  Returns False (and stays unpartitioned) when the driver cannot create green contexts."""
        if decode_sms <= 0:
            self.enc_stream = self.dec_stream = self.partition_sms = None
            return False
        es, ds, en, dn = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        rc = self.lib.gar_engine_partition(self._h, int(decode_sms), C.byref(es), C.byref(en), C.byref(ds), C.byref(dn))
        if rc != 0:
            self.partition_error = self.lib.gar_last_error().decode(errors="replace")
            self.enc_stream = self.dec_stream = self.partition_sms = None
            return False
        self.enc_stream = torch.cuda.ExternalStream(es.value, device=self.device)
        self.dec_stream = torch.cuda.ExternalStream(ds.value, device=self.device)
        self.partition_sms = (en.value, dn.value)
        return True

    def set_replay(self, mode: int = REPLAY_MERGED, crop_ids=None):
        """Which reference replay loop the following plan / assemble calls reproduce (gar_engine_set_replay) and which
        crop-token ids index the boxes; the defaults restore GARModel.generate's loop over spec.crop_tokens_ids."""
        ids = list(self.spec.crop_tokens_ids) if crop_ids is None else [int(c) for c in crop_ids]
        if not ids:
            ids = [-2]  # no replay at all: an id no token can have
        arr = (C.c_longlong * len(ids))(*ids)
        check(self.lib.gar_engine_set_replay(self._h, int(mode), None if crop_ids is None else arr,
                                             0 if crop_ids is None else len(ids)), "gar_engine_set_replay")
        self.replay_mode, self.crop_ids = int(mode), ids
        for t in self._twins:
            t.set_replay(mode, crop_ids)

    # ------------------------------------------------------------------ weights
    def _set(self, name: str, tensor: torch.Tensor):
        tensor = tensor.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        self._keep[name] = tensor
        check(self.lib.gar_engine_set_weight(self._h, name.encode(), C.c_void_p(tensor.data_ptr())),
              f"set_weight {name}")

    def _load_weights(self, sd: Dict[str, torch.Tensor]):
        spec, v, t = self.spec, self.spec.vision, self.spec.text
        kpad = ((3 * v.patch_size ** 2 + 63) // 64) * 64

        def conv_matrix(w):
            m = w.reshape(w.shape[0], -1)
            out = torch.zeros((m.shape[0], kpad), dtype=m.dtype, device=m.device)
            out[:, : m.shape[1]] = m
            return out

        self._set("vit.patch_w", conv_matrix(sd[VT + "patch_embed.proj.weight"]))
        self._set("vit.mask_w", conv_matrix(sd["mask_patch_embedding.weight"]))
        if v.use_cls_token:
            self._set("vit.cls", sd[VT + "cls_token"].reshape(-1))
        self._set("vit.pos", sd[VT + "pos_embed"].reshape(v.tokens, v.embed_dim))
        self._set("vit.norm_pre.w", sd[VT + "norm_pre.weight"])
        self._set("vit.norm_pre.b", sd[VT + "norm_pre.bias"])
        # Heads narrower than the attention kernel's tile (PE-G: 96) are zero-padded to 64/128 per head: padded q/k
        # dims add 0 to every score, padded v dims give zero outputs that hit zero columns of proj.
        hd, H, dv = v.head_dim, v.num_heads, v.embed_dim
        hdp = 64 if hd <= 64 else 128

        def pad_heads_rows(w):  # (3*H*hd, ...) -> (3*H*hdp, ...)
            if hdp == hd:
                return w
            w = w.reshape(3, H, hd, *w.shape[1:])
            out = torch.zeros((3, H, hdp, *w.shape[3:]), dtype=w.dtype, device=w.device)
            out[:, :, :hd] = w
            return out.reshape(3 * H * hdp, *w.shape[3:])

        def pad_heads_cols(w):  # (dv, H*hd) -> (dv, H*hdp)
            if hdp == hd:
                return w
            out = torch.zeros((w.shape[0], H, hdp), dtype=w.dtype, device=w.device)
            out[:, :, :hd] = w.reshape(w.shape[0], H, hd)
            return out.reshape(w.shape[0], H * hdp)

        rope = vit_rope_table(v)  # (g*g, 2*hd) = cat(sin, cos)
        if hdp != hd:
            sin, cos = rope[:, :hd], rope[:, hd:]
            rp = torch.zeros((rope.shape[0], 2 * hdp), dtype=rope.dtype)
            rp[:, :hd], rp[:, hdp:hdp + hd] = sin, cos
            rope = rp
        self._set("vit.rope", rope)
        self._set("vit.rope_cm", rope_table_chunk_major(rope, hdp))  # layout of the fused qkv-GEMM epilogue
        for i in range(v.depth):
            b = f"{VT}blocks.{i}."
            self._set(f"vit.{i}.ln1.w", sd[b + "norm1.weight"])
            self._set(f"vit.{i}.ln1.b", sd[b + "norm1.bias"])
            self._set(f"vit.{i}.qkv.w", pad_heads_rows(sd[b + "attn.qkv.weight"]))
            self._set(f"vit.{i}.qkv.b", pad_heads_rows(sd[b + "attn.qkv.bias"]))
            self._set(f"vit.{i}.proj.w", pad_heads_cols(sd[b + "attn.proj.weight"]))
            self._set(f"vit.{i}.proj.b", sd[b + "attn.proj.bias"])
            self._set(f"vit.{i}.g1", sd[b + "gamma_1"])
            self._set(f"vit.{i}.ln2.w", sd[b + "norm2.weight"])
            self._set(f"vit.{i}.ln2.b", sd[b + "norm2.bias"])
            self._set(f"vit.{i}.fc1.w", sd[b + "mlp.fc1.weight"])
            self._set(f"vit.{i}.fc1.b", sd[b + "mlp.fc1.bias"])
            self._set(f"vit.{i}.fc2.w", sd[b + "mlp.fc2.weight"])
            self._set(f"vit.{i}.fc2.b", sd[b + "mlp.fc2.bias"])
            self._set(f"vit.{i}.g2", sd[b + "gamma_2"])
        self._set("proj.l1.w", sd[PJ + "linear_1.weight"])
        self._set("proj.l1.b", sd[PJ + "linear_1.bias"])
        self._set("proj.l2.w", sd[PJ + "linear_2.weight"])
        self._set("proj.l2.b", sd[PJ + "linear_2.bias"])
        self._set("llm.embed", sd[LM + "embed_tokens.weight"])
        for i in range(t.num_layers):
            b = f"{LM}layers.{i}."
            self._set(f"llm.{i}.ln1", sd[b + "input_layernorm.weight"])
            self._set(f"llm.{i}.qkv.w", torch.cat([sd[b + "self_attn.q_proj.weight"], sd[b + "self_attn.k_proj.weight"],
                                                   sd[b + "self_attn.v_proj.weight"]], 0))
            self._set(f"llm.{i}.o.w", sd[b + "self_attn.o_proj.weight"])
            self._set(f"llm.{i}.ln2", sd[b + "post_attention_layernorm.weight"])
            gw, uw = sd[b + "mlp.gate_proj.weight"], sd[b + "mlp.up_proj.weight"]
            inter, k = gw.shape  # rows interleaved in 32-row groups [gate | up] (GAR_EPI_SWIGLU layout)
            self._set(f"llm.{i}.gu.w", torch.stack([gw.reshape(inter // 32, 32, k), uw.reshape(inter // 32, 32, k)],
                                                   1).reshape(2 * inter, k))
            self._set(f"llm.{i}.down.w", sd[b + "mlp.down_proj.weight"])
        self._set("llm.norm", sd[LM + "norm.weight"])
        head = sd["mllm.lm_head.weight"]
        padded = torch.zeros((self.vocab_padded, head.shape[1]), dtype=head.dtype, device=head.device)
        padded[: head.shape[0]] = head
        self._set("llm.lm_head", padded)
        cos, sin = llama_cos_sin(t, self.max_seq)
        self._set("llm.cos", cos)
        self._set("llm.sin", sin)

    # ------------------------------------------------------------------ stages
    def _s(self):
        return C.c_void_p(self.stream.cuda_stream)

    def buffer(self, name: str, shape, dtype=torch.bfloat16) -> torch.Tensor:
        """Copy of an engine-internal buffer (parity taps)."""
        ptr, nbytes = C.c_void_p(), C.c_longlong()
        check(self.lib.gar_engine_get_buffer(self._h, name.encode(), C.byref(ptr), C.byref(nbytes)), "get_buffer")
        out = torch.empty(shape, dtype=dtype, device=self.device)
        assert out.numel() * out.element_size() <= nbytes.value, (name, shape, nbytes.value)
        check(self.lib.gar_copy_d2d(C.c_void_p(out.data_ptr()), ptr, out.numel() * out.element_size(), self._s()),
              "gar_copy_d2d")
        self.stream.synchronize()
        return out

    def vit_encode(self, pixel_values: torch.Tensor, mask_values: Optional[torch.Tensor]):
        T = pixel_values.shape[0]
        check(self.lib.gar_engine_vit_encode(self._h, C.c_void_p(pixel_values.data_ptr()),
                                             None if mask_values is None else C.c_void_p(mask_values.data_ptr()),
                                             T, self._s()), "gar_engine_vit_encode")

    def assemble(self, slot: int, input_ids: torch.Tensor, T: int, bboxes, ncw: int, nch: int):
        rois, has, scale = roi_boxes_f32(self.spec, bboxes, ncw, nch, self.crop_ids, self.replay_mode == REPLAY_FRAME)
        n = len(has)
        c_rois = (C.c_float * (5 * n))(*[x for r in rois for x in r])
        c_has = (C.c_int * n)(*has)
        S = input_ids.numel()
        check(self.lib.gar_engine_assemble(self._h, slot, C.c_void_p(input_ids.data_ptr()), S, T, c_rois, c_has,
                                           float(scale), ncw, nch, self._s()), "gar_engine_assemble")

    def prefill(self, slot: int, S: int):
        check(self.lib.gar_engine_prefill(self._h, slot, S, self._s()), "gar_engine_prefill")

    def prefill_all_logits(self, slot: int, S: int) -> torch.Tensor:
        out = torch.empty((S, self.vocab_padded), dtype=torch.bfloat16, device=self.device)
        check(self.lib.gar_engine_prefill_all_logits(self._h, slot, S, C.c_void_p(out.data_ptr()), self._s()),
              "gar_engine_prefill_all_logits")
        return out

    def request(self, slot: int, pixel_values, global_mask_values, input_ids, bboxes, aspect_ratio):
        """ViT -> assemble -> prefill for one (image, mask-set, prompt) request into KV slot `slot`.
        Inputs must already be on the device in bf16 / int64; runs on the engine stream."""
        self.request_batch(slot, [(pixel_values, global_mask_values, input_ids, bboxes, aspect_ratio)])

    def plan_batch(self, slot0: int, requests) -> List[int]:
        """R6 row plan + validation of `requests` (same tuples as request_batch) BEFORE anything else runs; returns
        one GAR_STATUS_* word per request.  Synchronises the engine stream (idle at this point)."""
        n = len(requests)
        T, S = requests[0][0].shape[0], requests[0][2].numel()
        has = []
        for (_, _, _, bboxes, _) in requests:
            has += [1 if (bboxes is not None and bboxes.get(str(c)) is not None) else 0 for c in self.crop_ids]
        c_ids = (C.c_void_p * n)(*[r[2].data_ptr() for r in requests])
        st = (C.c_int * n)()
        check(self.lib.gar_engine_plan_batch(self._h, slot0, n, c_ids, S, T, (C.c_int * len(has))(*has), st, self._s()),
              "gar_engine_plan_batch")
        return list(st)

    def request_batch(self, slot0: int, requests, planned: bool = False):
        """ViT -> assemble -> prefill for several requests in ONE pass (one GEMM / attention launch covers all of
        them): requires equal tile counts and prompt lengths.  `requests` = list of
        (pixel_values, global_mask_values, input_ids(S,), bboxes, (ncw, nch)) already on the device.
        Per-request results are identical to a one-request pass; only the launch count and tile quantisation change.
        Every request is patchified straight from its own tensors (no concatenated copy)."""
        n = len(requests)
        T = requests[0][0].shape[0]
        S = requests[0][2].numel()
        assert all(r[0].shape[0] == T and r[2].numel() == S for r in requests)
        assert n * T <= self.max_tiles and n * S <= self.max_prefill_rows and slot0 + n <= self.max_batch
        with_mask = requests[0][1] is not None
        assert all((r[1] is not None) == with_mask for r in requests)
        self._mark("stage_begin")
        for i, r in enumerate(requests):
            check(self.lib.gar_engine_patchify_at(self._h, C.c_void_p(r[0].data_ptr()),
                                                  C.c_void_p(r[1].data_ptr()) if with_mask else None, T, i * T,
                                                  self._s()), "gar_engine_patchify_at")
        check(self.lib.gar_engine_vit_encode_patches(self._h, n * T, int(with_mask), self._s()),
              "gar_engine_vit_encode_patches")
        self._mark("vit_done")
        flat_rois, flat_has, ncws, nchs, scale = [], [], [], [], 1.0
        for (_, _, ids, bboxes, (ncw, nch)) in requests:
            rois, has, scale = roi_boxes_f32(self.spec, bboxes, ncw, nch, self.crop_ids,
                                             self.replay_mode == REPLAY_FRAME)
            flat_rois += [x for r in rois for x in r]
            flat_has += list(has)
            ncws.append(int(ncw))
            nchs.append(int(nch))
        c_ids = (C.c_void_p * n)(*[r[2].data_ptr() for r in requests])
        check(self.lib.gar_engine_assemble_batch(self._h, slot0, n, c_ids, S, T, (C.c_float * len(flat_rois))(*flat_rois),
                                                 (C.c_int * len(flat_has))(*flat_has), float(scale),
                                                 (C.c_int * n)(*ncws), (C.c_int * n)(*nchs), self._s()),
              "gar_engine_assemble_batch")
        self._mark("assemble_done")
        check(self.lib.gar_engine_prefill_batch(self._h, slot0, n, S, self._s()), "gar_engine_prefill_batch")
        self._mark("prefill_done")

    def _mark(self, label: str):
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(self.stream)
            self.trace.append((label, ev))

    def trace_ms(self):
        """[(label, ms since the previous mark)] of the marks collected so far (synchronises)."""
        torch.cuda.synchronize(self.device)
        out, prev = [], None
        for label, ev in self.trace or []:
            out.append((label, 0.0 if prev is None or label == "stage_begin" else prev.elapsed_time(ev)))
            prev = ev
        return out

    def generate(self, batch: int, max_new_tokens: int, eos_token_id=None, slot0: int = 0,
                 stream: Optional[torch.cuda.Stream] = None) -> List[List[int]]:
        """Greedy decode of KV slots [slot0, slot0+batch); blocks until the tokens are on the host.
        eos_token_id: None, an int, or a list of up to four stop ids (HF GenerationConfig.eos_token_id).
        Decoding stops as soon as every sequence has emitted a stop id; `last_steps_run` = decode steps launched."""
        if eos_token_id is None:
            eos = []
        elif isinstance(eos_token_id, (list, tuple)):
            eos = [int(x) for x in eos_token_id]
        else:
            eos = [int(eos_token_id)]
        if len(eos) > 4:
            raise _lib.GarError(f"at most 4 eos_token_id values are supported, got {len(eos)}")
        buf = (C.c_int * (batch * max_new_tokens))()
        steps = C.c_int(0)
        s = C.c_void_p((stream or self.stream).cuda_stream)
        check(self.lib.gar_engine_generate_eos(self._h, slot0, batch, max_new_tokens, (C.c_int * max(1, len(eos)))(*eos),
                                               len(eos), buf, C.byref(steps), s), "gar_engine_generate")
        self.last_steps_run = steps.value
        self.decoded_once = True
        out = []
        for b in range(batch):
            row = [buf[b * max_new_tokens + i] for i in range(max_new_tokens)]
            out.append([t for t in row if t >= 0])
        return out

    def status(self, slot: int) -> int:
        st = C.c_int()
        check(self.lib.gar_engine_status(self._h, slot, C.byref(st)), "gar_engine_status")
        return st.value

    def launch_count(self) -> int:
        return int(self.lib.gar_engine_launch_count(self._h)) + sum(t.launch_count() for t in self._twins)

    def set_option(self, name: str, value: int):
        check(self.lib.gar_engine_set_option(self._h, name.encode(), int(value)), "gar_engine_set_option")
        self._options[name] = int(value)
        self.decoded_once = False  # the C side dropped its decode graphs
        for t in self._twins:
            t.set_option(name, value)

    def profile(self, enable: bool):
        check(self.lib.gar_engine_profile(self._h, int(enable)), "gar_engine_profile")

    def profile_read(self):
        """{class: (ms, work, launches)} summed over the launches recorded since profile(True)."""
        n = 6
        ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
        check(self.lib.gar_engine_profile_read(self._h, ms, work, cnt), "gar_engine_profile_read")
        names = ["gemm", "attn", "assemble", "decode_attn", "gemv", "plan"]
        return {names[i]: (ms[i], work[i], cnt[i]) for i in range(n)}

    def close(self):
        for t in getattr(self, "_twins", []):
            t.close()
        self._twins = []
        if getattr(self, "_h", None):
            self.lib.gar_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
