// HBM-bound vectorised kernels of the region-captioning path (no tensor cores on these):
// mask-id decode + patchify (im2col of the stride==kernel conv), cls/pos-embed + LayerNorm,
// LayerNorm, RMSNorm, 2-D interleaved RoPE (PE ViT), rotate-half RoPE + KV-cache write (Llama),
// SwiGLU, 2x2 token average pool, embedding gather, argmax.
// Every kernel reproduces the reference's bf16 rounding points (one rounding per torch op).
#include "common.cuh"

namespace {

// -------------------------------------------------------------------------------------
// R2/R3 + K1 operand prep: pixel_values / global_mask_values (T,3,H,W) -> patch rows.
//   reference: modeling_gar.py:315-328 (mask-id decode, binary mask), timm PatchEmbed.proj and
//   GARModel.mask_patch_embedding (Conv2d k=s=14, bias-free) -> here an im2col so both convs
//   run as tcgen05 GEMMs.  Row = (tile, py, px); column = c*P*P + ky*P + kx (conv weight order),
//   zero-padded to ldk.
// -------------------------------------------------------------------------------------
__device__ __forceinline__ float mask_value_to_binary(float v, int prompt_numbers) {
  // torch: round((v + 1.0) / 2.0 * 255.0) evaluated in bf16, .long(), clamp(0, P), != P
  float a = bf16r(v + 1.0f);
  a = bf16r(a / 2.0f);
  a = bf16r(a * 255.0f);
  a = bf16r(rintf(a));
  long long id = static_cast<long long>(a);
  id = id < 0 ? 0 : (id > prompt_numbers ? prompt_numbers : id);
  return id != prompt_numbers ? 1.0f : 0.0f;
}

__global__ void patchify_kernel(const bf16* __restrict__ px, const bf16* __restrict__ mv, bf16* __restrict__ a_img,
                                bf16* __restrict__ a_mask, int T, int img, int patch, int ldk, int prompt_numbers) {
  const int grid = img / patch;
  const long long total = static_cast<long long>(T) * 3 * img * grid;  // one thread per (t, c, y, px)
  const int half = patch / 2;                                            // patch is even (14): 7 bf16x2 per segment
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pxi = static_cast<int>(i % grid);
    long long r = i / grid;
    const int y = static_cast<int>(r % img);
    r /= img;
    const int c = static_cast<int>(r % 3);
    const int t = static_cast<int>(r / 3);
    const int py = y / patch, ky = y % patch;
    const size_t src = ((static_cast<size_t>(t) * 3 + c) * img + y) * img + static_cast<size_t>(pxi) * patch;
    const size_t dst = (static_cast<size_t>(t) * grid * grid + static_cast<size_t>(py) * grid + pxi) * ldk +
                       static_cast<size_t>(c) * patch * patch + ky * patch;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(px + src);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(a_img + dst);
    for (int j = 0; j < half; ++j) d32[j] = __ldg(s32 + j);
    if (mv != nullptr) {
      const uint32_t* m32 = reinterpret_cast<const uint32_t*>(mv + src);
      uint32_t* dm = reinterpret_cast<uint32_t*>(a_mask + dst);
      for (int j = 0; j < half; ++j) {
        const float2 f = unpack_bf16x2(__ldg(m32 + j));
        dm[j] = pack_bf16x2(mask_value_to_binary(f.x, prompt_numbers), mask_value_to_binary(f.y, prompt_numbers));
      }
    }
  }
}

// -------------------------------------------------------------------------------------
// warp-per-row normalisation helpers
// -------------------------------------------------------------------------------------
constexpr int NORM_MAXV = 16;  // up to 16 x 16B per lane kept in registers (d <= 4096)

template <bool RMS>
__device__ __forceinline__ void norm_row(const uint4* __restrict__ xin, const bf16* __restrict__ w,
                                         const bf16* __restrict__ b, uint4* __restrict__ yout, int nvec, float eps,
                                         int lane, const uint4* __restrict__ addin) {
  // optional addin: y = norm(bf16(x + addin))   (pos-embed add fused in front of norm_pre)
  uint4 cache[NORM_MAXV];
  float sum = 0.f;
  const int d = nvec * 8;
#pragma unroll
  for (int j = 0; j < NORM_MAXV; ++j) {
    const int v = lane + j * 32;
    if (v < nvec) {
      uint4 x = xin[v];
      if (addin) {
        const uint4 a = __ldg(addin + v);
        uint32_t* xw = reinterpret_cast<uint32_t*>(&x);
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(xw[k]), g = unpack_bf16x2(aw[k]);
          xw[k] = pack_bf16x2(f.x + g.x, f.y + g.y);
        }
      }
      cache[j] = x;
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(xw[k]);
        if (RMS) sum += f.x * f.x + f.y * f.y;
        else sum += f.x + f.y;
      }
    }
  }
  sum = warp_sum(sum);
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(sum / d + eps);
  } else {
    mean = sum / d;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NORM_MAXV; ++j) {
      const int v = lane + j * 32;
      if (v < nvec) {
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(&cache[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(xw[k]);
          var += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
      }
    }
    var = warp_sum(var);
    rstd = rsqrtf(var / d + eps);
  }
#pragma unroll
  for (int j = 0; j < NORM_MAXV; ++j) {
    const int v = lane + j * 32;
    if (v < nvec) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w) + v);
      uint4 bv = make_uint4(0, 0, 0, 0);
      if (!RMS) bv = __ldg(reinterpret_cast<const uint4*>(b) + v);
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(&cache[j]);
      const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bv);
      uint4 o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(xw[k]), g = unpack_bf16x2(ww[k]);
        if (RMS) {
          // LlamaRMSNorm: weight * (x32 * rsqrt(var + eps)).to(bf16)
          ow[k] = pack_bf16x2(g.x * bf16r(f.x * rstd), g.y * bf16r(f.y * rstd));
        } else {
          const float2 h = unpack_bf16x2(bw[k]);
          ow[k] = pack_bf16x2((f.x - mean) * rstd * g.x + h.x, (f.y - mean) * rstd * g.y + h.y);
        }
      }
      yout[v] = o;
    }
  }
}

// Fixed-width variant (d = 256 * NV): the row lives in NV registers-vectors per lane with no predication, and the next
// row of the warp is already in flight while the current one is reduced and written (two rows = 2 * d * 2 bytes in
// flight per warp).  Same operation order per row as norm_row.
template <bool RMS, int NV>
__device__ __forceinline__ void norm_row_regs(const uint4 (&cache)[NV], const bf16* __restrict__ w,
                                              const bf16* __restrict__ b, uint4* __restrict__ yout, float eps, int lane) {
  constexpr int d = NV * 256;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&cache[j]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(xw[k]);
      if (RMS) sum += f.x * f.x + f.y * f.y;
      else sum += f.x + f.y;
    }
  }
  sum = warp_sum(sum);
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(sum / d + eps);
  } else {
    mean = sum / d;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(&cache[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(xw[k]);
        var += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
      }
    }
    var = warp_sum(var);
    rstd = rsqrtf(var / d + eps);
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int v = lane + j * 32;
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w) + v);
    uint4 bv = make_uint4(0, 0, 0, 0);
    if (!RMS) bv = __ldg(reinterpret_cast<const uint4*>(b) + v);
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&cache[j]);
    const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wv);
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bv);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(xw[k]), g = unpack_bf16x2(ww[k]);
      if (RMS) {
        ow[k] = pack_bf16x2(g.x * bf16r(f.x * rstd), g.y * bf16r(f.y * rstd));
      } else {
        const float2 h = unpack_bf16x2(bw[k]);
        ow[k] = pack_bf16x2((f.x - mean) * rstd * g.x + h.x, (f.y - mean) * rstd * g.y + h.y);
      }
    }
    yout[v] = o;
  }
}

template <bool RMS, int NV>
__global__ void __launch_bounds__(128) norm_kernel_fixed(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w,
                                                         const bf16* __restrict__ b, bf16* __restrict__ y, int ldy,
                                                         int rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * (blockDim.x >> 5);
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint4 cur[NV], nxt[NV];
  if (row < rows) {
#pragma unroll
    for (int j = 0; j < NV; ++j) cur[j] = ld_nc_v4(reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx) + lane + j * 32);
  }
  while (row < rows) {
    const int nrow = row + stride;
    if (nrow < rows) {
#pragma unroll
      for (int j = 0; j < NV; ++j) nxt[j] = ld_nc_v4(reinterpret_cast<const uint4*>(x + static_cast<size_t>(nrow) * ldx) + lane + j * 32);
    }
    norm_row_regs<RMS, NV>(cur, w, b, reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * ldy), eps, lane);
#pragma unroll
    for (int j = 0; j < NV; ++j) cur[j] = nxt[j];
    row = nrow;
  }
}

template <bool RMS>
__global__ void norm_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w,
                            const bf16* __restrict__ b, bf16* __restrict__ y, int ldy, int rows, int d, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    norm_row<RMS>(reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx), w, b,
                  reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * ldy), d / 8, eps, lane, nullptr);
  }
}

// K4: x[t, n] = LN_pre( bf16( (n < npt ? cls : patch[t, n - npt]) + pos_embed[n] ) )
//   reference: timm Eva._pos_embed + norm_pre, called at modeling_perception_lm.py:197-198
__global__ void vit_embed_ln_kernel(const bf16* __restrict__ patch, const bf16* __restrict__ cls,
                                    const bf16* __restrict__ pos, const bf16* __restrict__ w,
                                    const bf16* __restrict__ b, bf16* __restrict__ x, int T, int ntok, int npt, int d,
                                    float eps) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rows = T * ntok;
  const int npatch = ntok - npt;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const int t = row / ntok, n = row % ntok;
    const bf16* src = n < npt ? cls : patch + (static_cast<size_t>(t) * npatch + (n - npt)) * d;
    norm_row<false>(reinterpret_cast<const uint4*>(src), w, b, reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * d),
                    d / 8, eps, lane, reinterpret_cast<const uint4*>(pos + static_cast<size_t>(n) * d));
  }
}

// -------------------------------------------------------------------------------------
// PE-ViT 2-D RoPE on q,k in place inside the fused qkv buffer [rows, 3*d] (interleaved pairs,
// prefix tokens untouched).  reference: timm AttentionRope / apply_rot_embed_cat (Appendix A).
// rope table row = cat(sin[hd], cos[hd]) in bf16.
// -------------------------------------------------------------------------------------
__global__ void vit_rope_kernel(bf16* __restrict__ qkv, const bf16* __restrict__ rope, int rows, int ntok, int npt,
                                int heads, int hd) {
  const int d = heads * hd;
  const int vec_per_row = 2 * d / 8;  // q and k parts
  const long long total = static_cast<long long>(rows) * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / vec_per_row);
    const int v = static_cast<int>(i % vec_per_row);
    const int n = row % ntok;
    if (n < npt) continue;
    const int col = v * 8;              // column inside [q | k]
    const int e = col % hd;             // element inside the head
    uint4* p = reinterpret_cast<uint4*>(qkv + static_cast<size_t>(row) * 3 * d + col);
    const bf16* tr = rope + static_cast<size_t>(n - npt) * 2 * hd;
    const uint4 sv = __ldg(reinterpret_cast<const uint4*>(tr + e));
    const uint4 cv = __ldg(reinterpret_cast<const uint4*>(tr + hd + e));
    uint4 x = *p;
    uint32_t* xw = reinterpret_cast<uint32_t*>(&x);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(&sv);
    const uint32_t* cw = reinterpret_cast<const uint32_t*>(&cv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(xw[k]), s = unpack_bf16x2(sw[k]), c = unpack_bf16x2(cw[k]);
      const float o0 = bf16r(f.x * c.x) + bf16r(-f.y * s.x);
      const float o1 = bf16r(f.y * c.y) + bf16r(f.x * s.y);
      xw[k] = pack_bf16x2(o0, o1);
    }
    *p = x;
  }
}

// -------------------------------------------------------------------------------------
// Llama RoPE (rotate-half) on q in place + roped K / V written into the KV cache.
//   reference: HF modeling_llama.py apply_rotary_pos_emb (:138-168) + DynamicCache.update.
//   qkv row layout: [Hq*hd | Hkv*hd | Hkv*hd]; cos/sin tables [max_pos, hd/2] bf16.
//   positions: pos0 + row (prefill) or pos_dev[row] (decode, one row per sequence slot).
//   cache layout: [slot][kv_head][max_seq][hd]
// -------------------------------------------------------------------------------------
__global__ void llama_rope_kv_kernel(bf16* __restrict__ qkv, int rows, int hq, int hkv, int hd,
                                     const bf16* __restrict__ cos_t, const bf16* __restrict__ sin_t, int pos0,
                                     const int* __restrict__ pos_dev, const int* __restrict__ slot_dev, int slot0,
                                     bf16* __restrict__ kcache, bf16* __restrict__ vcache, int max_seq, int seq_len) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = hd / 2;
  const int vec_per_head = half / 8;  // each work item: 8 low elements + their 8 high partners
  const int heads_total = hq + 2 * hkv;
  const int ld = heads_total * hd;
  const long long total = static_cast<long long>(rows) * heads_total * vec_per_head;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec_per_head);
    long long r = i / vec_per_head;
    const int head = static_cast<int>(r % heads_total);
    const int row = static_cast<int>(r / heads_total);
    // seq_len > 0: rows are nreq back-to-back sequences of seq_len positions in consecutive slots (batched prefill)
    const int pos = pos_dev ? pos_dev[row] : pos0 + (seq_len > 0 ? row % seq_len : row);
    const int slot = slot_dev ? slot_dev[row]
                              : (slot0 >= 0 ? slot0 + (seq_len > 0 ? row / seq_len : 0) : row);  // slot0 < 0: row b -> slot b
    bf16* base = qkv + static_cast<size_t>(row) * ld + static_cast<size_t>(head) * hd + v * 8;
    uint4 lo = *reinterpret_cast<const uint4*>(base);
    uint4 hi = *reinterpret_cast<const uint4*>(base + half);
    if (head < hq + hkv) {
      const uint4 cv = __ldg(reinterpret_cast<const uint4*>(cos_t + static_cast<size_t>(pos) * half + v * 8));
      const uint4 sv = __ldg(reinterpret_cast<const uint4*>(sin_t + static_cast<size_t>(pos) * half + v * 8));
      uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
      uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
      const uint32_t* cw = reinterpret_cast<const uint32_t*>(&cv);
      const uint32_t* sw = reinterpret_cast<const uint32_t*>(&sv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = unpack_bf16x2(lw[k]), b = unpack_bf16x2(hw[k]);
        const float2 c = unpack_bf16x2(cw[k]), s = unpack_bf16x2(sw[k]);
        // out_lo = x_lo*cos + (-x_hi)*sin ; out_hi = x_hi*cos + x_lo*sin   (each product rounded to bf16)
        lw[k] = pack_bf16x2(bf16r(a.x * c.x) + bf16r(-b.x * s.x), bf16r(a.y * c.y) + bf16r(-b.y * s.y));
        hw[k] = pack_bf16x2(bf16r(b.x * c.x) + bf16r(a.x * s.x), bf16r(b.y * c.y) + bf16r(a.y * s.y));
      }
    }
    if (head < hq) {
      *reinterpret_cast<uint4*>(base) = lo;
      *reinterpret_cast<uint4*>(base + half) = hi;
    } else {
      const int kvh = head < hq + hkv ? head - hq : head - hq - hkv;
      bf16* cache = head < hq + hkv ? kcache : vcache;
      bf16* dst = cache + ((static_cast<size_t>(slot) * hkv + kvh) * max_seq + pos) * hd + v * 8;
      *reinterpret_cast<uint4*>(dst) = lo;
      *reinterpret_cast<uint4*>(dst + half) = hi;
    }
  }
}

// SwiGLU: act = bf16( bf16(silu(g)) * u ) with gu rows = [gate(I) | up(I)]   (HF LlamaMLP :182-184)
__global__ void swiglu_kernel(const bf16* __restrict__ gu, bf16* __restrict__ act, int rows, int inter,
                              int interleaved) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = inter / 8;
  const long long total = static_cast<long long>(rows) * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / vec_per_row);
    const int v = static_cast<int>(i % vec_per_row);
    // feature f = v*8: plain layout gate at f, up at inter+f; interleaved: 64*(f/32) + f%32 and +32
    const int f = v * 8;
    const size_t goff = interleaved ? static_cast<size_t>(64 * (f / 32) + (f % 32)) : static_cast<size_t>(f);
    const size_t uoff = interleaved ? goff + 32 : static_cast<size_t>(inter + f);
    const uint4 g = *reinterpret_cast<const uint4*>(gu + static_cast<size_t>(row) * 2 * inter + goff);
    const uint4 u = *reinterpret_cast<const uint4*>(gu + static_cast<size_t>(row) * 2 * inter + uoff);
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g);
    const uint32_t* uw = reinterpret_cast<const uint32_t*>(&u);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = unpack_bf16x2(gw[k]), b = unpack_bf16x2(uw[k]);
      ow[k] = pack_bf16x2(bf16r(silu_f(a.x)) * b.x, bf16r(silu_f(a.y)) * b.y);
    }
    *reinterpret_cast<uint4*>(act + static_cast<size_t>(row) * inter + v * 8) = o;
  }
}

// 2x2 adaptive average pool over the token grid, dropping prefix (cls) tokens.
//   reference: modeling_perception_lm.py:42-60 (+ :263-264 cls drop).  x [T, ntok, D] -> y [T, (g/2)^2, D]
__global__ void pool2x2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int T, int ntok, int npt, int grid,
                               int D) {
  const int og = grid / 2;
  const int vec = D / 8;
  const long long total = static_cast<long long>(T) * og * og * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    long long r = i / vec;
    const int pw = static_cast<int>(r % og);
    r /= og;
    const int ph = static_cast<int>(r % og);
    const int t = static_cast<int>(r / og);
    const bf16* base = x + (static_cast<size_t>(t) * ntok + npt + static_cast<size_t>(2 * ph) * grid + 2 * pw) * D + v * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(base);
    const uint4 b = *reinterpret_cast<const uint4*>(base + D);
    const uint4 c = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(grid) * D);
    const uint4 e = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(grid + 1) * D);
    const uint32_t *aw = reinterpret_cast<const uint32_t*>(&a), *bw = reinterpret_cast<const uint32_t*>(&b),
                   *cw = reinterpret_cast<const uint32_t*>(&c), *ew = reinterpret_cast<const uint32_t*>(&e);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_bf16x2(aw[k]), fb = unpack_bf16x2(bw[k]), fc = unpack_bf16x2(cw[k]),
                   fe = unpack_bf16x2(ew[k]);
      ow[k] = pack_bf16x2((((fa.x + fb.x) + fc.x) + fe.x) / 4.0f, (((fa.y + fb.y) + fc.y) + fe.y) / 4.0f);
    }
    *reinterpret_cast<uint4*>(y + ((static_cast<size_t>(t) * og + ph) * og + pw) * D + v * 8) = o;
  }
}

// rows of the embedding table gathered by (device) token ids: h[b] = table[tok[b]]
__global__ void embed_gather_kernel(const bf16* __restrict__ table, const long long* __restrict__ ids64,
                                    const int* __restrict__ ids32, bf16* __restrict__ out, int rows, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec = D / 8;
  const long long total = static_cast<long long>(rows) * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / vec), v = static_cast<int>(i % vec);
    const long long id = ids64 ? ids64[row] : ids32[row];
    reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * D)[v] =
        __ldg(reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * D) + v);
  }
}

// greedy step: tok[b] = argmax_v logits[b, v] (lowest index on ties, like torch.argmax), then
// bookkeeping for the device-side decode loop: append to out_tokens[b, step], bump position,
// latch EOS.  Stage 1 (optional): grid (chunks, batch) reduces 4096-wide chunks into (max, idx) pairs;
// stage 2: one block per sequence reduces the pairs (or scans the row itself when no workspace is given).
constexpr int ARGMAX_CHUNK = 4096;

__device__ __forceinline__ void argmax_merge(float& best, int& bidx, float ob, int oi) {
  if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
}

__device__ __forceinline__ void argmax_block_reduce(float& best, int& bidx, float* sb, int* si) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    argmax_merge(best, bidx, ob, oi);
  }
  if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bidx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? sb[threadIdx.x] : -INFINITY;
    bidx = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      argmax_merge(best, bidx, ob, oi);
    }
  }
}

__global__ void argmax_partial_kernel(const bf16* __restrict__ logits, int ld, int V, float* __restrict__ pmax,
                                      int* __restrict__ pidx, int chunks) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sb[32];
  __shared__ int si[32];
  const int c = blockIdx.x, b = blockIdx.y;
  const bf16* row = logits + static_cast<size_t>(b) * ld;
  const int lo = c * ARGMAX_CHUNK, hi = min(V, lo + ARGMAX_CHUNK);
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int v0 = lo + threadIdx.x * 8; v0 < hi; v0 += blockDim.x * 8) {
    if (v0 + 8 <= hi) {
      const uint4 raw = *reinterpret_cast<const uint4*>(row + v0);  // ld % 8 == 0 and lo % 8 == 0
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(rw[j]);
        argmax_merge(best, bidx, f.x, v0 + 2 * j);
        argmax_merge(best, bidx, f.y, v0 + 2 * j + 1);
      }
    } else {
      for (int v = v0; v < hi; ++v) argmax_merge(best, bidx, bf2f(row[v]), v);
    }
  }
  argmax_block_reduce(best, bidx, sb, si);
  if (threadIdx.x == 0) {
    pmax[static_cast<size_t>(b) * chunks + c] = best;
    pidx[static_cast<size_t>(b) * chunks + c] = bidx;
  }
}

__global__ void argmax_step_kernel(const bf16* __restrict__ logits, int ld, int V, const float* __restrict__ pmax,
                                   const int* __restrict__ pidx, int chunks, int* __restrict__ next_tok,
                                   int* __restrict__ out_tokens, int out_ld, int* __restrict__ step_ptr,
                                   int* __restrict__ pos, int* __restrict__ finished, int4 eos4, int advance_pos,
                                   int bump_step) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sb[32];
  __shared__ int si[32];
  const int b = blockIdx.x;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  if (pmax != nullptr) {
    for (int c = threadIdx.x; c < chunks; c += blockDim.x)
      argmax_merge(best, bidx, pmax[static_cast<size_t>(b) * chunks + c], pidx[static_cast<size_t>(b) * chunks + c]);
  } else {
    const bf16* row = logits + static_cast<size_t>(b) * ld;
    for (int v = threadIdx.x; v < V; v += blockDim.x) argmax_merge(best, bidx, bf2f(row[v]), v);
  }
  argmax_block_reduce(best, bidx, sb, si);
  if (threadIdx.x == 0) {
    const int step = step_ptr ? *step_ptr : 0;
    const int fin = finished ? finished[b] : 0;
    // eos4: up to four stop ids (-1 = unused); a finished sequence keeps feeding its first stop id (HF pads likewise)
    const int eos = eos4.x;
    const int tok = fin ? (eos >= 0 ? eos : bidx) : bidx;
    next_tok[b] = tok;
    if (out_tokens) out_tokens[static_cast<size_t>(b) * out_ld + step] = fin ? -1 : tok;
    const bool stop = (eos4.x >= 0 && tok == eos4.x) || (eos4.y >= 0 && tok == eos4.y) || (eos4.z >= 0 && tok == eos4.z) ||
                      (eos4.w >= 0 && tok == eos4.w);
    if (finished && stop) finished[b] = 1;
    if (pos && advance_pos) pos[b] += 1;
  }
  // the last block to finish bumps the step counter (replaces a separate 1-thread launch)
  if (bump_step && step_ptr) {
    __shared__ int last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      // ticket lives in step_ptr[1]
      const int t = atomicAdd(step_ptr + 1, 1);
      last = (t == static_cast<int>(gridDim.x) - 1);
      if (last) { step_ptr[1] = 0; step_ptr[0] += 1; }
    }
  }
}

__global__ void bump_step_kernel(int* step) { *step += 1; }

inline int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(gar_num_sms()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int gar_patchify(const void* pixel_values, const void* mask_values, void* a_img, void* a_mask, int T, int img,
                 int patch, int ldk, int prompt_numbers, void* stream) {
  if (img % patch || (patch & 1) || ldk < 3 * patch * patch || (ldk % 8)) {
    gar_set_error("gar_patchify: bad geometry img=%d patch=%d ldk=%d", img, patch, ldk);
    return GAR_ERR_SHAPE;
  }
  const long long total = static_cast<long long>(T) * 3 * img * (img / patch);
  patchify_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(pixel_values), static_cast<const bf16*>(mask_values), static_cast<bf16*>(a_img),
      static_cast<bf16*>(a_mask), T, img, patch, ldk, prompt_numbers);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

int gar_layernorm(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int rows, int d, float eps,
                  void* stream) {
  if (d % 8 || d > NORM_MAXV * 256 || ldx % 8 || ldy % 8) {
    gar_set_error("gar_layernorm: d=%d unsupported (need d%%8==0, d<=%d)", d, NORM_MAXV * 256);
    return GAR_ERR_SHAPE;
  }
  const bf16 *xp = static_cast<const bf16*>(x), *wp = static_cast<const bf16*>(w), *bp = static_cast<const bf16*>(b);
  bf16* yp = static_cast<bf16*>(y);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // x must not alias y for the streaming-load variant (the engine never normalises in place)
  const int grid = grid_for(static_cast<long long>(rows) * 32, 128);
  if (d == 1024 && x != y) norm_kernel_fixed<false, 4><<<grid, 128, 0, st>>>(xp, ldx, wp, bp, yp, ldy, rows, eps);
  else if (d == 1536 && x != y) norm_kernel_fixed<false, 6><<<grid, 128, 0, st>>>(xp, ldx, wp, bp, yp, ldy, rows, eps);
  else norm_kernel<false><<<grid, 128, 0, st>>>(xp, ldx, wp, bp, yp, ldy, rows, d, eps);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

int gar_rmsnorm(const void* x, int ldx, const void* w, void* y, int ldy, int rows, int d, float eps, void* stream) {
  if (d % 8 || d > NORM_MAXV * 256 || ldx % 8 || ldy % 8) {
    gar_set_error("gar_rmsnorm: d=%d unsupported (need d%%8==0, d<=%d)", d, NORM_MAXV * 256);
    return GAR_ERR_SHAPE;
  }
  const dim3 grid(grid_for(static_cast<long long>(rows) * 32, 128));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bf16 *xp = static_cast<const bf16*>(x), *wp = static_cast<const bf16*>(w), *nb = nullptr;
  bf16* yp = static_cast<bf16*>(y);
  if (d == 2048 && x != y && rows >= 64)
    GAR_CUDA_CHECK(gar_launch(norm_kernel_fixed<true, 8>, grid, dim3(128), 0, st, true, dim3(1, 1, 1), xp, ldx, wp, nb, yp, ldy,
                              rows, eps));
  else if (d == 4096 && x != y && rows >= 64)
    GAR_CUDA_CHECK(gar_launch(norm_kernel_fixed<true, 16>, grid, dim3(128), 0, st, true, dim3(1, 1, 1), xp, ldx, wp, nb, yp,
                              ldy, rows, eps));
  else
    GAR_CUDA_CHECK(gar_launch(norm_kernel<true>, grid, dim3(128), 0, st, true, dim3(1, 1, 1), xp, ldx, wp, nb, yp, ldy, rows,
                              d, eps));
  return GAR_OK;
}

int gar_vit_embed_ln(const void* patch, const void* cls, const void* pos, const void* w, const void* b, void* x, int T,
                     int ntok, int npt, int d, float eps, void* stream) {
  if (d % 8 || d > NORM_MAXV * 256 || (npt && !cls)) return GAR_ERR_SHAPE;
  vit_embed_ln_kernel<<<grid_for(static_cast<long long>(T) * ntok * 32, 128), 128, 0,
                        static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(patch), static_cast<const bf16*>(cls), static_cast<const bf16*>(pos),
      static_cast<const bf16*>(w), static_cast<const bf16*>(b), static_cast<bf16*>(x), T, ntok, npt, d, eps);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

int gar_vit_rope(void* qkv, const void* rope, int rows, int ntok, int npt, int heads, int hd, void* stream) {
  if (hd % 8) return GAR_ERR_SHAPE;
  const long long total = static_cast<long long>(rows) * (2 * heads * hd / 8);
  vit_rope_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(qkv), static_cast<const bf16*>(rope), rows, ntok, npt, heads, hd);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

int gar_llama_rope_kv(void* qkv, int rows, int hq, int hkv, int hd, const void* cos_t, const void* sin_t, int pos0,
                      const int* pos_dev, const int* slot_dev, int slot0, void* kcache, void* vcache, int max_seq,
                      int seq_len, void* stream) {
  if (hd % 16) return GAR_ERR_SHAPE;
  const long long total = static_cast<long long>(rows) * (hq + 2 * hkv) * (hd / 16);
  GAR_CUDA_CHECK(gar_launch(llama_rope_kv_kernel, dim3(grid_for(total, 256)), dim3(256), 0,
                            static_cast<cudaStream_t>(stream), true, dim3(1, 1, 1), static_cast<bf16*>(qkv), rows, hq,
                            hkv, hd, static_cast<const bf16*>(cos_t), static_cast<const bf16*>(sin_t), pos0, pos_dev,
                            slot_dev, slot0, static_cast<bf16*>(kcache), static_cast<bf16*>(vcache), max_seq, seq_len));
  return GAR_OK;
}

int gar_swiglu(const void* gu, void* act, int rows, int inter, int interleaved, void* stream) {
  if (inter % 8 || (interleaved && inter % 32)) return GAR_ERR_SHAPE;
  GAR_CUDA_CHECK(gar_launch(swiglu_kernel, dim3(grid_for(static_cast<long long>(rows) * (inter / 8), 256)), dim3(256),
                            0, static_cast<cudaStream_t>(stream), true, dim3(1, 1, 1), static_cast<const bf16*>(gu),
                            static_cast<bf16*>(act), rows, inter, interleaved));
  return GAR_OK;
}

int gar_pool2x2(const void* x, void* y, int T, int ntok, int npt, int grid, int D, void* stream) {
  if (D % 8 || grid % 2 || ntok != grid * grid + npt) return GAR_ERR_SHAPE;
  const long long total = static_cast<long long>(T) * (grid / 2) * (grid / 2) * (D / 8);
  pool2x2_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), T, ntok, npt, grid, D);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

int gar_embed_gather(const void* table, const long long* ids64, const int* ids32, void* out, int rows, int D,
                     void* stream) {
  if (D % 8 || (!ids64 && !ids32)) return GAR_ERR_SHAPE;
  GAR_CUDA_CHECK(gar_launch(embed_gather_kernel, dim3(grid_for(static_cast<long long>(rows) * (D / 8), 256)),
                            dim3(256), 0, static_cast<cudaStream_t>(stream), true, dim3(1, 1, 1),
                            static_cast<const bf16*>(table), ids64, ids32, static_cast<bf16*>(out), rows, D));
  return GAR_OK;
}

int gar_argmax_step_eos(const void* logits, int ld, int V, int batch, int* next_tok, int* out_tokens, int out_ld,
                        int* step_ptr, int* pos, int* finished, const int* eos_ids, int n_eos, int advance_pos,
                        int bump_step, void* partial_ws, void* stream);

int gar_argmax_step(const void* logits, int ld, int V, int batch, int* next_tok, int* out_tokens, int out_ld,
                    int* step_ptr, int* pos, int* finished, int eos, int advance_pos, int bump_step, void* partial_ws,
                    void* stream) {
  const int ids[1] = {eos};
  return gar_argmax_step_eos(logits, ld, V, batch, next_tok, out_tokens, out_ld, step_ptr, pos, finished, ids,
                             eos >= 0 ? 1 : 0, advance_pos, bump_step, partial_ws, stream);
}

int gar_argmax_step_eos(const void* logits, int ld, int V, int batch, int* next_tok, int* out_tokens, int out_ld,
                        int* step_ptr, int* pos, int* finished, const int* eos_ids, int n_eos, int advance_pos,
                        int bump_step, void* partial_ws, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_eos < 0 || n_eos > 4 || (n_eos > 0 && !eos_ids)) return GAR_ERR_ARG;
  int4 eos = make_int4(-1, -1, -1, -1);
  if (n_eos > 0) eos.x = eos_ids[0];
  if (n_eos > 1) eos.y = eos_ids[1];
  if (n_eos > 2) eos.z = eos_ids[2];
  if (n_eos > 3) eos.w = eos_ids[3];
  const int chunks = (V + ARGMAX_CHUNK - 1) / ARGMAX_CHUNK;
  float* pmax = nullptr;
  int* pidx = nullptr;
  if (partial_ws != nullptr && (ld % 8) == 0) {
    pmax = static_cast<float*>(partial_ws);
    pidx = reinterpret_cast<int*>(pmax + static_cast<size_t>(batch) * chunks);
    GAR_CUDA_CHECK(gar_launch(argmax_partial_kernel, dim3(chunks, batch), dim3(256), 0, s, true, dim3(1, 1, 1),
                              static_cast<const bf16*>(logits), ld, V, pmax, pidx, chunks));
  }
  GAR_CUDA_CHECK(gar_launch(argmax_step_kernel, dim3(batch), dim3(pmax ? 64 : 1024), 0, s, true, dim3(1, 1, 1),
                            static_cast<const bf16*>(logits), ld, V, static_cast<const float*>(pmax),
                            static_cast<const int*>(pidx), chunks, next_tok, out_tokens, out_ld, step_ptr, pos,
                            finished, eos, advance_pos, bump_step));
  return GAR_OK;
}

int gar_bump_step(int* step, void* stream) {
  bump_step_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(step);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

}  // extern "C"
