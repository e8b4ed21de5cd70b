// Single-query (decode) GQA attention over the pre-allocated KV cache, split along the KV length
// ("flash-decoding"), plus the split combine.  HBM-bound: every K/V byte of every live sequence is
// read exactly once per step; no tensor cores (q_len == 1).
//
// Replaces flash-attn 2 with q_len=1 + DynamicCache torch.cat per layer per step behind HF Llama
// (SURVEY.md §2b K10).  Numerics follow flash-attn: fp32 scores/softmax, P rounded to bf16 before
// the PV product, fp32 accumulation, one bf16 rounding at the end.
//
// cache layout [slot][kv_head][max_seq][hd]; q is the fused qkv row of the step (roped in place).
#include "common.cuh"

namespace {

constexpr int DEC_THREADS = 128;

// Each 8-lane group owns one KV row at a time (lane l%8 holds elements [8*(l%8), +8) of each 64-wide
// slice); a warp covers 4 consecutive rows per load instruction (512 contiguous bytes) and keeps
// ROWS_U such instructions in flight for K and for V before touching the data.
constexpr int ROWS_U = 4;

template <int HD, int GROUP>
__global__ void __launch_bounds__(DEC_THREADS)
decode_attn_split_kernel(const bf16* __restrict__ qkv, int ldq, const bf16* __restrict__ kcache,
                         const bf16* __restrict__ vcache, const int* __restrict__ pos_dev,
                         const int* __restrict__ slot_dev, int hq, int hkv, int max_seq, int splits, float scale_log2,
                         float* __restrict__ o_part, float* __restrict__ ml_part) {
  constexpr int VPL = HD / 64;  // 16-byte vectors per lane per row
  pdl_launch_dependents();
  pdl_wait();
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int len = pos_dev[b] + 1;  // keys 0..pos (the step's own K/V were just appended)
  const int slot = slot_dev ? slot_dev[b] : b;
  int chunk = (len + splits - 1) / splits;
  chunk = (chunk + 15) & ~15;
  const int k_begin = split * chunk;
  const int k_end = min(len, k_begin + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, grp = lane >> 3;

  float q[GROUP][VPL][8];
#pragma unroll
  for (int g = 0; g < GROUP; ++g)
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const uint4 raw = *reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(b) * ldq +
                                                        static_cast<size_t>(kvh * GROUP + g) * HD + v * 64 + sub * 8);
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(rw[j]);
        q[g][v][2 * j] = f.x * scale_log2;  // fold softmax scale * log2(e) into q once
        q[g][v][2 * j + 1] = f.y * scale_log2;
      }
    }
  float m[GROUP], l[GROUP], acc[GROUP][VPL][8];
#pragma unroll
  for (int g = 0; g < GROUP; ++g) {
    m[g] = -INFINITY;
    l[g] = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[g][v][j] = 0.f;
  }
  const bf16* kbase = kcache + (static_cast<size_t>(slot) * hkv + kvh) * max_seq * HD;
  const bf16* vbase = vcache + (static_cast<size_t>(slot) * hkv + kvh) * max_seq * HD;
  constexpr int ROWS_PER_WARP_ITER = 4 * ROWS_U;
  constexpr int ROWS_PER_ITER = (DEC_THREADS / 32) * ROWS_PER_WARP_ITER;
  // warp-uniform trip count: the shuffles need all 32 lanes, so tail rows are predicated, not skipped
  for (int rb = k_begin + warp * ROWS_PER_WARP_ITER; rb < k_end; rb += ROWS_PER_ITER) {
    uint4 kr[ROWS_U][VPL], vr[ROWS_U][VPL];
#pragma unroll
    for (int u = 0; u < ROWS_U; ++u) {
      const int r = rb + u * 4 + grp;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        kr[u][v] = make_uint4(0, 0, 0, 0);
        if (r < k_end) kr[u][v] = *reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(r) * HD + v * 64 + sub * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS_U; ++u) {
      const int r = rb + u * 4 + grp;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        vr[u][v] = make_uint4(0, 0, 0, 0);
        if (r < k_end) vr[u][v] = *reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(r) * HD + v * 64 + sub * 8);
      }
    }
    float sc[GROUP][ROWS_U];
#pragma unroll
    for (int u = 0; u < ROWS_U; ++u) {
      float kf[VPL][8];
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const uint32_t* kw = reinterpret_cast<const uint32_t*>(&kr[u][v]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = unpack_bf16x2(kw[j]);
          kf[v][2 * j] = a.x;
          kf[v][2 * j + 1] = a.y;
        }
      }
#pragma unroll
      for (int g = 0; g < GROUP; ++g) {
        float d = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int j = 0; j < 8; ++j) d += q[g][v][j] * kf[v][j];
        sc[g][u] = d;
      }
    }
#pragma unroll
    for (int g = 0; g < GROUP; ++g)
#pragma unroll
      for (int u = 0; u < ROWS_U; ++u) {
        float d = sc[g][u];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        sc[g][u] = (rb + u * 4 + grp < k_end) ? d : -INFINITY;
      }
    float pb[GROUP][ROWS_U];
#pragma unroll
    for (int g = 0; g < GROUP; ++g) {
      float mx = m[g];
#pragma unroll
      for (int u = 0; u < ROWS_U; ++u) mx = fmaxf(mx, sc[g][u]);
      const float m_use = (mx == -INFINITY) ? 0.f : mx;
      const float alpha = exp2f(m[g] - m_use);
      float ps = 0.f;
#pragma unroll
      for (int u = 0; u < ROWS_U; ++u) {
        const float pe = exp2f(sc[g][u] - m_use);
        ps += pe;
        pb[g][u] = bf16r(pe);  // flash-attn rounds P to bf16 before the PV product
      }
      l[g] = l[g] * alpha + ps;
      m[g] = mx;
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[g][v][j] *= alpha;
    }
#pragma unroll
    for (int u = 0; u < ROWS_U; ++u) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const uint32_t* vw = reinterpret_cast<const uint32_t*>(&vr[u][v]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 c = unpack_bf16x2(vw[j]);
#pragma unroll
          for (int g = 0; g < GROUP; ++g) {
            acc[g][v][2 * j] += pb[g][u] * c.x;
            acc[g][v][2 * j + 1] += pb[g][u] * c.y;
          }
        }
      }
    }
  }
  // merge the 4 row-groups of a warp, then the warps, through shared memory
  __shared__ float s_m[DEC_THREADS / 8][GROUP], s_l[DEC_THREADS / 8][GROUP];
  __shared__ float s_acc[DEC_THREADS / 8][GROUP][HD];
  const int gid = warp * 4 + grp;  // 16 row-groups per CTA
#pragma unroll
  for (int g = 0; g < GROUP; ++g) {
    if (sub == 0) { s_m[gid][g] = m[g]; s_l[gid][g] = l[g]; }
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) s_acc[gid][g][v * 64 + sub * 8 + j] = acc[g][v][j];
  }
  __syncthreads();
  constexpr int NG = DEC_THREADS / 8;
  for (int idx = threadIdx.x; idx < GROUP * HD; idx += DEC_THREADS) {
    const int g = idx / HD, e = idx % HD;
    float mm = -INFINITY;
#pragma unroll
    for (int i = 0; i < NG; ++i) mm = fmaxf(mm, s_m[i][g]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const float w = (s_m[i][g] == -INFINITY) ? 0.f : exp2f(s_m[i][g] - mm);
      ll += s_l[i][g] * w;
      oo += s_acc[i][g][e] * w;
    }
    const int h = kvh * GROUP + g;
    const size_t pidx = (static_cast<size_t>(b) * hq + h) * splits + split;
    o_part[pidx * HD + e] = oo;
    if (e == 0) { ml_part[pidx * 2] = mm; ml_part[pidx * 2 + 1] = ll; }
  }
}

template <int HD>
__global__ void decode_attn_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part,
                                           bf16* __restrict__ out, int ldo, int hq, int splits) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y;
  const int e = threadIdx.x;  // HD threads
  const size_t base = (static_cast<size_t>(b) * hq + h) * splits;
  float mm = -INFINITY;
  for (int s = 0; s < splits; ++s) mm = fmaxf(mm, ml_part[(base + s) * 2]);
  float ll = 0.f, oo = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float ms = ml_part[(base + s) * 2];
    const float w = (ms == -INFINITY) ? 0.f : exp2f(ms - mm);
    ll += ml_part[(base + s) * 2 + 1] * w;
    oo += o_part[(base + s) * HD + e] * w;
  }
  out[static_cast<size_t>(b) * ldo + h * HD + e] = __float2bfloat16_rn(ll > 0.f ? oo / ll : 0.f);
}

}  // namespace

extern "C" int gar_decode_attn(const void* qkv, int ldq, const void* kcache, const void* vcache, const int* pos_dev,
                               const int* slot_dev, int batch, int hq, int hkv, int head_dim, int max_seq, int splits,
                               float scale, void* o_part_ws, void* ml_part_ws, void* out, int ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int group = hq / hkv;
  if (hq % hkv || splits < 1 || splits > 32) return GAR_ERR_ARG;
  const float sl2 = scale * 1.4426950408889634f;
  dim3 grid(splits, hkv, batch);
#define LAUNCH_DEC(HD_, G_)                                                                                        \
  GAR_CUDA_CHECK(gar_launch(decode_attn_split_kernel<HD_, G_>, grid, dim3(DEC_THREADS), 0, stream, true, dim3(1, 1, 1), \
                            static_cast<const bf16*>(qkv), ldq, static_cast<const bf16*>(kcache),                  \
                            static_cast<const bf16*>(vcache), pos_dev, slot_dev, hq, hkv, max_seq, splits, sl2,    \
                            static_cast<float*>(o_part_ws), static_cast<float*>(ml_part_ws)))
  if (head_dim == 64 && group == 4) LAUNCH_DEC(64, 4);
  else if (head_dim == 64 && group == 2) LAUNCH_DEC(64, 2);
  else if (head_dim == 64 && group == 1) LAUNCH_DEC(64, 1);
  else if (head_dim == 128 && group == 4) LAUNCH_DEC(128, 4);
  else {
    gar_set_error("gar_decode_attn: head_dim=%d group=%d not built", head_dim, group);
    return GAR_ERR_SHAPE;
  }
#undef LAUNCH_DEC
  GAR_CUDA_CHECK(cudaGetLastError());
  dim3 cgrid(hq, batch);
  if (head_dim == 64)
    GAR_CUDA_CHECK(gar_launch(decode_attn_combine_kernel<64>, cgrid, dim3(64), 0, stream, true, dim3(1, 1, 1),
                              static_cast<const float*>(o_part_ws), static_cast<const float*>(ml_part_ws),
                              static_cast<bf16*>(out), ldo, hq, splits));
  else
    GAR_CUDA_CHECK(gar_launch(decode_attn_combine_kernel<128>, cgrid, dim3(128), 0, stream, true, dim3(1, 1, 1),
                              static_cast<const float*>(o_part_ws), static_cast<const float*>(ml_part_ws),
                              static_cast<bf16*>(out), ldo, hq, splits));
  return GAR_OK;
}
