"""Thin torch-tensor wrappers over the per-kernel C-ABI entry points (include/gar_b200.h).
torch supplies device memory and the stream; all math happens inside libgar_b200.so."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from ._lib import check

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_LS_RES, EPI_RES = range(5)
EPI_SWIGLU = 6


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype=torch.bfloat16):
    if not t.is_cuda:
        raise _lib.GarError("gar_b200 ops need CUDA tensors (no CPU path)")
    assert t.dtype == dtype, (t.dtype, dtype)


def gemm(a, b, epi=EPI_NONE, bias=None, gamma=None, res=None, out=None, block_n=0):
    """out[M,N] = epi(a[M,K] @ b[N,K]^T)"""
    _req(a), _req(b)
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty((M, N // 2 if epi == EPI_SWIGLU else N), dtype=torch.bfloat16, device=a.device)
    check(_lib.load().gar_gemm_bf16(_p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0), M, N, K, epi,
                                    _p(bias), _p(gamma), _p(res), 0 if res is None else res.stride(0), block_n,
                                    _stream()), "gar_gemm_bf16")
    return out


def gemm_rope2d(a, b, bias, rope, ntok, npt, hd, rope_cols, chunk_major=False):
    """rope: the (ntok - npt, 2 * hd) cat(sin, cos) table (re-laid out chunk-major here), or, with chunk_major=True, the
    output of rope.rope_table_chunk_major (what the engine uploads once)."""
    if not chunk_major:
        from .rope import rope_table_chunk_major

        rope = rope_table_chunk_major(rope, hd)
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    check(_lib.load().gar_gemm_bf16_rope2d(_p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0), M, N, K,
                                           _p(bias), _p(rope), ntok, npt, hd, rope_cols, _stream()),
          "gar_gemm_bf16_rope2d")
    return out


def gemm_skinny(x, w, epi=EPI_NONE, res=None, out=None):
    """out[M<=16,N] = epi(x[M,K] @ w[N,K]^T) through the decode weight-streaming kernel."""
    _req(x), _req(w)
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().gar_gemm_skinny_bf16(_p(x), x.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K,
                                           epi, _p(res), 0 if res is None else res.stride(0), _stream()),
          "gar_gemm_skinny_bf16")
    return out


def gemv_decode(x, w, epi=0, res=None, norm_w=None, sumsq_in=None, eps=1e-5, want_sumsq=False, out=None):
    """Decode linear layer (gar_gemv_decode).  Returns (out, sumsq_out or None)."""
    _req(x), _req(w)
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N // 2 if epi == 2 else N), dtype=torch.bfloat16, device=x.device)
    sq = torch.zeros((N // 16, 8), dtype=torch.float32, device=x.device) if want_sumsq else None
    rows = C.c_int(0)
    n_part = 0 if sumsq_in is None else sumsq_in.shape[0]
    check(_lib.load().gar_gemv_decode(_p(x), x.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K, epi,
                                      _p(res), 0 if res is None else res.stride(0), _p(norm_w), _p(sumsq_in), n_part,
                                      eps, _p(sq), C.byref(rows), _stream()), "gar_gemv_decode")
    if sq is not None:
        sq = sq[: N // rows.value]
    return out, sq


def layernorm(x, w, b, eps):
    _req(x)
    y = torch.empty_like(x)
    check(_lib.load().gar_layernorm(_p(x), x.stride(0), _p(w), _p(b), _p(y), y.stride(0), x.shape[0], x.shape[1], eps,
                                    _stream()), "gar_layernorm")
    return y


def rmsnorm(x, w, eps):
    _req(x)
    y = torch.empty_like(x)
    check(_lib.load().gar_rmsnorm(_p(x), x.stride(0), _p(w), _p(y), y.stride(0), x.shape[0], x.shape[1], eps,
                                  _stream()), "gar_rmsnorm")
    return y


def patchify(pixel_values, mask_values, patch, ldk, prompt_numbers):
    _req(pixel_values)
    T, _, img, _ = pixel_values.shape
    g = img // patch
    a_img = torch.zeros((T * g * g, ldk), dtype=torch.bfloat16, device=pixel_values.device)
    a_mask = torch.zeros_like(a_img) if mask_values is not None else None
    check(_lib.load().gar_patchify(_p(pixel_values), _p(mask_values), _p(a_img), _p(a_mask), T, img, patch, ldk,
                                   prompt_numbers, _stream()), "gar_patchify")
    return a_img, a_mask


def vit_embed_ln(patch, cls, pos, w, b, T, ntok, npt, eps):
    d = patch.shape[-1]
    x = torch.empty((T * ntok, d), dtype=torch.bfloat16, device=patch.device)
    check(_lib.load().gar_vit_embed_ln(_p(patch), _p(cls), _p(pos), _p(w), _p(b), _p(x), T, ntok, npt, d, eps,
                                       _stream()), "gar_vit_embed_ln")
    return x


def vit_rope_(qkv, rope, ntok, npt, heads, hd):
    check(_lib.load().gar_vit_rope(_p(qkv), _p(rope), qkv.shape[0], ntok, npt, heads, hd, _stream()), "gar_vit_rope")
    return qkv


def llama_rope_kv_(qkv, hq, hkv, hd, cos_t, sin_t, kcache, vcache, max_seq, pos0=0, pos_dev=None, slot0=0):
    check(_lib.load().gar_llama_rope_kv(_p(qkv), qkv.shape[0], hq, hkv, hd, _p(cos_t), _p(sin_t), pos0, _p(pos_dev),
                                        None, slot0, _p(kcache), _p(vcache), max_seq, 0, _stream()), "gar_llama_rope_kv")
    return qkv


def swiglu(gu, interleaved=False):
    rows, two_i = gu.shape
    act = torch.empty((rows, two_i // 2), dtype=torch.bfloat16, device=gu.device)
    check(_lib.load().gar_swiglu(_p(gu), _p(act), rows, two_i // 2, int(interleaved), _stream()), "gar_swiglu")
    return act


def interleave_gate_up(gate_w, up_w, block=32):
    """[gate(I,K); up(I,K)] -> rows interleaved in `block`-row groups [gate | up] (GAR_EPI_SWIGLU layout)."""
    I, K = gate_w.shape
    g = gate_w.reshape(I // block, block, K)
    u = up_w.reshape(I // block, block, K)
    return torch.stack([g, u], 1).reshape(2 * I, K).contiguous()


def pool2x2(x, T, ntok, npt, grid):
    D = x.shape[-1]
    y = torch.empty((T, (grid // 2) ** 2, D), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().gar_pool2x2(_p(x), _p(y), T, ntok, npt, grid, D, _stream()), "gar_pool2x2")
    return y


def attn_fwd_packed_qkv(qkv, T, ntok, heads, hd, causal=False):
    """ViT layout: qkv [T*ntok, 3*heads*hd] -> out [T*ntok, heads*hd]."""
    d = heads * hd
    out = torch.empty((T * ntok, d), dtype=torch.bfloat16, device=qkv.device)
    check(_lib.load().gar_attn_fwd(_p(qkv), 3 * d, ntok * 3 * d, 0, _p(qkv), _p(qkv), 3 * d, ntok * 3 * d, T, d, 2 * d,
                                   hd, 1, 0, _p(out), d, ntok * d, 0, T, heads, heads, hd, ntok, ntok, ntok,
                                   1 if causal else 0, 1.0 / math.sqrt(hd), _stream()), "gar_attn_fwd")
    return out


def attn_fwd_cache(q, q_ld, kcache_slot, vcache_slot, hq, hkv, hd, S, max_seq):
    """Llama prefill layout: q rows [S, q_ld] (head h at col h*hd), cache slot [hkv, max_seq, hd]."""
    out = torch.empty((S, hq * hd), dtype=torch.bfloat16, device=q.device)
    check(_lib.load().gar_attn_fwd(_p(q), q_ld, S * q_ld, 0, _p(kcache_slot), _p(vcache_slot), hd, max_seq * hd, hkv,
                                   0, 0, 0, 0, 1, _p(out), hq * hd, S * hq * hd, 0, 1, hq, hkv, hd, S, S, max_seq, 1,
                                   1.0 / math.sqrt(hd), _stream()), "gar_attn_fwd")
    return out


def decode_attn(qkv, kcache, vcache, pos_dev, hq, hkv, hd, max_seq, splits=8, fused_combine=True):
    B = qkv.shape[0]
    tickets = torch.zeros((B * hkv,), dtype=torch.int32, device=qkv.device) if fused_combine else None
    o_part = torch.empty((B, hq, splits, hd), dtype=torch.float32, device=qkv.device)
    ml = torch.empty((B, hq, splits, 2), dtype=torch.float32, device=qkv.device)
    out = torch.empty((B, hq * hd), dtype=torch.bfloat16, device=qkv.device)
    check(_lib.load().gar_decode_attn(_p(qkv), qkv.stride(0), _p(kcache), _p(vcache), _p(pos_dev), None, B, hq, hkv, hd,
                                      max_seq, splits, 1.0 / math.sqrt(hd), _p(o_part), _p(ml), _p(tickets), _p(out),
                                      hq * hd, _stream()), "gar_decode_attn")
    return out


def decode_attn_rope(qkv, kcache, vcache, pos_dev, hq, hkv, hd, max_seq, cos_t, sin_t, splits=8):
    """Fused RoPE + KV append + decode attention (qkv rows un-rotated; caches updated in place)."""
    B = qkv.shape[0]
    tickets = torch.zeros((B * hkv,), dtype=torch.int32, device=qkv.device)
    o_part = torch.empty((B, hq, splits, hd), dtype=torch.float32, device=qkv.device)
    ml = torch.empty((B, hq, splits, 2), dtype=torch.float32, device=qkv.device)
    out = torch.empty((B, hq * hd), dtype=torch.bfloat16, device=qkv.device)
    check(_lib.load().gar_decode_attn_rope(_p(qkv), qkv.stride(0), _p(kcache), _p(vcache), _p(pos_dev), None, B, hq, hkv,
                                           hd, max_seq, splits, 1.0 / math.sqrt(hd), _p(cos_t), _p(sin_t), _p(o_part),
                                           _p(ml), _p(tickets), _p(out), hq * hd, _stream()), "gar_decode_attn_rope")
    return out


REPLAY_MERGED, REPLAY_MERGED_LAST, REPLAY_FRAME = 0, 1, 2  # include/gar_b200.h GAR_REPLAY_*


def assemble_embeds(input_ids, table, feats, crop_ids, rois, has_roi, spatial_scale, ncw, nch, tpt, image_token_id,
                    replay_mode=REPLAY_MERGED):
    """input_ids (S,) int64 cuda; feats (T, tpt*tpt, D); rois list of 5-float lists (one per crop id)."""
    S = input_ids.numel()
    D = table.shape[1]
    n = len(crop_ids)
    out = torch.empty((S, D), dtype=torch.bfloat16, device=table.device)
    plan = torch.empty((S + 8,), dtype=torch.int32, device=table.device)
    status = torch.zeros((2,), dtype=torch.int32, device=table.device)
    c_ids = (C.c_longlong * n)(*crop_ids)
    flat = [float(v) for r in rois for v in r]
    c_rois = (C.c_float * (5 * n))(*flat)
    c_has = (C.c_int * n)(*[int(h) for h in has_roi])
    check(_lib.load().gar_assemble_embeds(_p(input_ids), S, _p(table), table.shape[0], _p(feats),
                                          feats.shape[0] * feats.shape[1], D, c_ids, c_rois, c_has, n,
                                          float(spatial_scale), ncw, nch, tpt, image_token_id, int(replay_mode), _p(out),
                                          _p(plan), _p(status), _stream()), "gar_assemble_embeds")
    return out, status


def assemble_embeds_batch(input_ids_list, table, feats, crop_ids, rois, has_roi, spatial_scale, grids, tpt,
                          image_token_id, replay_mode=REPLAY_MERGED):
    """Batched form: input_ids_list = nreq (S,) int64 cuda tensors; feats (nreq, T, tpt*tpt, D); rois[r][k] 5 floats;
    has_roi[r][k]; grids[r] = (ncw, nch).  Returns out (nreq, S, D), status (nreq, 2)."""
    n = len(input_ids_list)
    S = input_ids_list[0].numel()
    D = table.shape[1]
    k = len(crop_ids)
    out = torch.empty((n, S, D), dtype=torch.bfloat16, device=table.device)
    plan = torch.empty((n * S + 8 * n,), dtype=torch.int32, device=table.device)
    status = torch.zeros((n, 2), dtype=torch.int32, device=table.device)
    c_ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in input_ids_list])
    c_ids = (C.c_longlong * k)(*crop_ids)
    c_rois = (C.c_float * (5 * k * n))(*[float(v) for rr in rois for r in rr for v in r])
    c_has = (C.c_int * (k * n))(*[int(h) for hh in has_roi for h in hh])
    c_w = (C.c_int * n)(*[int(g[0]) for g in grids])
    c_h = (C.c_int * n)(*[int(g[1]) for g in grids])
    check(_lib.load().gar_assemble_embeds_batch(n, c_ptrs, S, _p(table), table.shape[0], _p(feats),
                                                feats.shape[1] * feats.shape[2], D, c_ids, c_rois, c_has, k,
                                                float(spatial_scale), c_w, c_h, tpt, image_token_id, int(replay_mode),
                                                _p(out), _p(plan), _p(status), _stream()), "gar_assemble_embeds_batch")
    return out, status


def argmax_step(logits, V, two_stage=True):
    B = logits.shape[0]
    nxt = torch.empty((B,), dtype=torch.int32, device=logits.device)
    ws = torch.empty((B * ((V + 4095) // 4096) * 2,), dtype=torch.float32, device=logits.device) if two_stage else None
    check(_lib.load().gar_argmax_step(_p(logits), logits.stride(0), V, B, _p(nxt), None, 0, None, None, None, -1, 0, 0,
                                      _p(ws), _stream()), "gar_argmax_step")
    return nxt
