// Single-query (decode) GQA attention over the pre-allocated KV cache, split along the KV length
// ("flash-decoding"), plus the split combine.  HBM-bound: every K/V byte of every live sequence is
// read exactly once per step; no tensor cores (q_len == 1).
//
// Replaces flash-attn 2 with q_len=1 + DynamicCache torch.cat per layer per step behind HF Llama
// (SURVEY.md §2b K10).  Numerics follow flash-attn: fp32 scores/softmax, P rounded to bf16 before
// the PV product, fp32 accumulation, one bf16 rounding at the end.
//
// cache layout [slot][kv_head][max_seq][hd]; q is the fused qkv row of the step (roped in place).
#include "../../include/gar_b200.h"
#include "common.cuh"

namespace {

constexpr int DEC_THREADS = 128;

// Fused split combine: the last CTA of a (sequence, kv head) to finish -- decided by an atomic ticket --
// merges all splits of its GROUP heads and writes the bf16 attention output, saving a launch per layer.
template <int HD, int GROUP>
__device__ __forceinline__ void fused_combine_tail(int* __restrict__ tickets, const float* o_part, const float* ml_part,
                                                   bf16* __restrict__ out, int ldo, int b, int kvh, int hq, int hkv,
                                                   int splits) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(&tickets[b * hkv + kvh], 1);
    s_last = (t == splits - 1);
    if (s_last) tickets[b * hkv + kvh] = 0;  // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int idx = threadIdx.x; idx < GROUP * HD; idx += blockDim.x) {
    const int h = kvh * GROUP + idx / HD, e = idx % HD;
    const size_t base = (static_cast<size_t>(b) * hq + h) * splits;
    float mm = -INFINITY;
    for (int sidx = 0; sidx < splits; ++sidx) mm = fmaxf(mm, __ldcg(ml_part + (base + sidx) * 2));
    float ll = 0.f, oo = 0.f;
    for (int sidx = 0; sidx < splits; ++sidx) {
      const float ms = __ldcg(ml_part + (base + sidx) * 2);
      const float w = (ms == -INFINITY) ? 0.f : exp2f(ms - mm);
      ll += __ldcg(ml_part + (base + sidx) * 2 + 1) * w;
      oo += __ldcg(o_part + (base + sidx) * HD + e) * w;
    }
    out[static_cast<size_t>(b) * ldo + h * HD + e] = __float2bfloat16_rn(ll > 0.f ? oo / ll : 0.f);
  }
}

// Each 8-lane group owns one KV row at a time (lane l%8 holds elements [8*(l%8), +8) of each 64-wide
// slice); a warp covers 4 consecutive rows per load instruction (512 contiguous bytes) and keeps
// ROWS_U such instructions in flight for K and for V before touching the data.
constexpr int ROWS_U = 4;

template <int HD, int GROUP>
__global__ void __launch_bounds__(DEC_THREADS)
decode_attn_split_kernel(const bf16* __restrict__ qkv, int ldq, const bf16* __restrict__ kcache,
                         const bf16* __restrict__ vcache, const int* __restrict__ pos_dev,
                         const int* __restrict__ slot_dev, int hq, int hkv, int max_seq, int splits, float scale_log2,
                         float* __restrict__ o_part, float* __restrict__ ml_part, int* __restrict__ tickets,
                         bf16* __restrict__ out, int ldo) {
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
  if (tickets != nullptr) fused_combine_tail<HD, GROUP>(tickets, o_part, ml_part, out, ldo, b, kvh, hq, hkv, splits);
}

// ---------------------------------------------------------------------------------------------
// Tensor-core variant (mma.sync m16n8k16, flash-decoding layout): the CUDA-core kernel above is
// instruction-bound (~22 warp-instructions per key); here a warp handles 16 keys with 16 MMAs:
//   S[head, key]   = q (A, rows = the GROUP query heads of this kv head, zero padded to 16) x K^T (B, loaded
//                    straight from the cache with 16-byte loads and a permuted-d fragment mapping)
//   O[head, d]    += P (A, re-used from the S accumulator fragments) x V (B, via ldmatrix.trans from a
//                    swizzled per-warp shared-memory stage)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// rotate-half RoPE of one lane's slice: lo = dims [8tig, +8), hi = dims [32+8tig, +8) of a 64-wide head;
// identical arithmetic (and rounding points) to llama_rope_kv_kernel
__device__ __forceinline__ void rope_lo_hi(uint4& lo, uint4& hi, const bf16* cos_row, const bf16* sin_row, int tig) {
  const uint4 cv = __ldg(reinterpret_cast<const uint4*>(cos_row + 8 * tig));
  const uint4 sv = __ldg(reinterpret_cast<const uint4*>(sin_row + 8 * tig));
  uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
  uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
  const uint32_t* cw = reinterpret_cast<const uint32_t*>(&cv);
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(&sv);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = unpack_bf16x2(lw[k]), b2 = unpack_bf16x2(hw[k]);
    const float2 c = unpack_bf16x2(cw[k]), sn = unpack_bf16x2(sw[k]);
    lw[k] = pack_bf16x2(bf16r(a.x * c.x) + bf16r(-b2.x * sn.x), bf16r(a.y * c.y) + bf16r(-b2.y * sn.y));
    hw[k] = pack_bf16x2(bf16r(b2.x * c.x) + bf16r(a.x * sn.x), bf16r(b2.y * c.y) + bf16r(a.y * sn.y));
  }
}

// ROPE: the step's q/k arrive un-rotated in `qkv`; q is rotated in registers, and the one CTA whose key range holds
// the new position rotates that K row on the fly, appends it and the V row to the cache (DynamicCache.update) and
// attends over it -- the separate RoPE/KV-append launch of the decode step disappears.
template <int GROUP, bool ROPE>
__global__ void __launch_bounds__(DEC_THREADS)
decode_attn_mma_kernel(const bf16* __restrict__ qkv, int ldq, bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                       const int* __restrict__ pos_dev, const int* __restrict__ slot_dev, int hq, int hkv, int max_seq,
                       int splits, float scale_log2, float* __restrict__ o_part, float* __restrict__ ml_part,
                       int* __restrict__ tickets, bf16* __restrict__ out, int ldo, const bf16* __restrict__ cos_t,
                       const bf16* __restrict__ sin_t) {
  constexpr int HD = 64;
  pdl_launch_dependents();
  pdl_wait();
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int len = pos_dev[b] + 1;
  const int slot = slot_dev ? slot_dev[b] : b;
  int chunk = (len + splits - 1) / splits;
  chunk = (chunk + 15) & ~15;
  const int k_begin = split * chunk;
  const int k_end = min(len, k_begin + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  constexpr int NW = DEC_THREADS / 32;

  __shared__ __align__(16) bf16 vstage[NW][16 * HD];  // per-warp V block, 16-byte chunks XOR-swizzled by key
  __shared__ float s_m[NW][8], s_l[NW][8];
  __shared__ float s_o[NW][8][HD];

  // q fragments (A operand rows = heads; rows >= GROUP and rows 8..15 are zero)
  uint32_t qa[4] = {0, 0, 0, 0}, qb[4] = {0, 0, 0, 0};
  if (g < GROUP) {
    const bf16* qp = qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(kvh * GROUP + g) * HD;
    uint4 x = *reinterpret_cast<const uint4*>(qp + 8 * tig);
    uint4 y = *reinterpret_cast<const uint4*>(qp + 32 + 8 * tig);
    if (ROPE) rope_lo_hi(x, y, cos_t + static_cast<size_t>(len - 1) * (HD / 2), sin_t + static_cast<size_t>(len - 1) * (HD / 2), tig);
    qa[0] = x.x; qa[1] = x.y; qa[2] = x.z; qa[3] = x.w;
    qb[0] = y.x; qb[1] = y.y; qb[2] = y.z; qb[3] = y.w;
  }
  bf16* kbase = kcache + (static_cast<size_t>(slot) * hkv + kvh) * max_seq * HD;
  bf16* vbase = vcache + (static_cast<size_t>(slot) * hkv + kvh) * max_seq * HD;
  const int new_pos = len - 1;  // ROPE: this row is not in the cache yet
  const bf16* knew = qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(hq + kvh) * HD;
  const bf16* vnew = qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(hq + hkv + kvh) * HD;

  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[n][i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;  // per (lane row g): partial over this lane's key columns

  const uint32_t vs_base = smem_u32(&vstage[warp][0]);
  // K fragments (key kb+g and kb+8+g; d permuted as [8tig..8tig+7] and [32+8tig..]) and the V block
  // (128 chunks of 16 B, 4 per lane) of one 16-key step; loads are predicated on the key range.
  struct Step { uint4 k0a, k0b, k1a, k1b, vv[4]; };
  auto load_step = [&](Step& st, int kb) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    st.k0a = st.k0b = st.k1a = st.k1b = z;
    auto load_k = [&](int key, uint4& lo, uint4& hi) {
      if (key >= k_end) return;
      if (ROPE && key == new_pos) {
        lo = *reinterpret_cast<const uint4*>(knew + 8 * tig);
        hi = *reinterpret_cast<const uint4*>(knew + 32 + 8 * tig);
        rope_lo_hi(lo, hi, cos_t + static_cast<size_t>(new_pos) * (HD / 2), sin_t + static_cast<size_t>(new_pos) * (HD / 2), tig);
        *reinterpret_cast<uint4*>(kbase + static_cast<size_t>(key) * HD + 8 * tig) = lo;
        *reinterpret_cast<uint4*>(kbase + static_cast<size_t>(key) * HD + 32 + 8 * tig) = hi;
      } else {
        const bf16* kp = kbase + static_cast<size_t>(key) * HD;
        lo = *reinterpret_cast<const uint4*>(kp + 8 * tig);
        hi = *reinterpret_cast<const uint4*>(kp + 32 + 8 * tig);
      }
    };
    load_k(kb + g, st.k0a, st.k0b);
    load_k(kb + 8 + g, st.k1a, st.k1b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = lane + 32 * i, key = id >> 3, c = id & 7;
      st.vv[i] = z;
      if (kb + key < k_end) {
        if (ROPE && kb + key == new_pos) {
          st.vv[i] = *reinterpret_cast<const uint4*>(vnew + 8 * c);
          *reinterpret_cast<uint4*>(vbase + static_cast<size_t>(kb + key) * HD + 8 * c) = st.vv[i];
        } else {
          st.vv[i] = *reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(kb + key) * HD + 8 * c);
        }
      }
    }
  };
  // register double buffering: the loads of step i+1 are in flight while step i is computed (the kernel is
  // otherwise bound by one DRAM round trip per step)
  Step cur, nxt;
  const int kb_first = k_begin + warp * 16;
  if (kb_first < k_end) load_step(cur, kb_first);
  for (int kb = kb_first; kb < k_end; kb += NW * 16) {
    if (kb + NW * 16 < k_end) load_step(nxt, kb + NW * 16);
    const uint4 k0a = cur.k0a, k0b = cur.k0b, k1a = cur.k1a, k1b = cur.k1b;
    uint4 vv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) vv[i] = cur.vv[i];
    __syncwarp();  // previous iteration's ldmatrix reads are done
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = lane + 32 * i, key = id >> 3, c = id & 7;
      *reinterpret_cast<uint4*>(&vstage[warp][key * HD + ((c ^ (key & 7)) << 3)]) = vv[i];
    }
    // ---- S = q K^T for two 8-key groups
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    mma16816(s0, qa[0], 0u, qa[1], 0u, k0a.x, k0a.y);
    mma16816(s0, qa[2], 0u, qa[3], 0u, k0a.z, k0a.w);
    mma16816(s0, qb[0], 0u, qb[1], 0u, k0b.x, k0b.y);
    mma16816(s0, qb[2], 0u, qb[3], 0u, k0b.z, k0b.w);
    mma16816(s1, qa[0], 0u, qa[1], 0u, k1a.x, k1a.y);
    mma16816(s1, qa[2], 0u, qa[3], 0u, k1a.z, k1a.w);
    mma16816(s1, qb[0], 0u, qb[1], 0u, k1b.x, k1b.y);
    mma16816(s1, qb[2], 0u, qb[3], 0u, k1b.z, k1b.w);
    // lane holds row g (head), keys kb + 2tig, +1 (s0[0], s0[1]) and kb + 8 + 2tig, +1 (s1[0], s1[1])
    float v0 = s0[0] * scale_log2, v1 = s0[1] * scale_log2, v2 = s1[0] * scale_log2, v3 = s1[1] * scale_log2;
    v0 = (kb + 2 * tig < k_end) ? v0 : -INFINITY;
    v1 = (kb + 2 * tig + 1 < k_end) ? v1 : -INFINITY;
    v2 = (kb + 8 + 2 * tig < k_end) ? v2 : -INFINITY;
    v3 = (kb + 8 + 2 * tig + 1 < k_end) ? v3 : -INFINITY;
    float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = ex2_approx(m_run - m_use);
    const float p0 = ex2_approx(v0 - m_use), p1 = ex2_approx(v1 - m_use), p2 = ex2_approx(v2 - m_use),
                p3 = ex2_approx(v3 - m_use);
    l_run = l_run * alpha + ((p0 + p1) + (p2 + p3));
    m_run = m_new;
    const uint32_t pa0 = pack_bf16x2(p0, p1), pa2 = pack_bf16x2(p2, p3);  // P as A fragment (rows 8..15 zero)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= alpha;
      o[n][1] *= alpha;
    }
    __syncwarp();  // V stage fully written
    // ---- O += P V : B fragments through ldmatrix.trans (matrix rows = keys, cols = 8 d)
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // two 8-wide d chunks per ldmatrix.x4
      const int mat = lane >> 3, r = lane & 7;
      const int key = (mat & 1) * 8 + r;
      const int c = np * 2 + (mat >> 1);
      const uint32_t addr = vs_base + (key * HD + ((c ^ (key & 7)) << 3)) * 2;
      uint32_t b0, b1, b2, b3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                   : "r"(addr));
      mma16816(o[np * 2], pa0, 0u, pa2, 0u, b0, b1);
      mma16816(o[np * 2 + 1], pa0, 0u, pa2, 0u, b2, b3);
    }
    cur = nxt;
  }
  // ---- per-warp row statistics: combine the 4 key-column lanes of each head row
  float m_w = fmaxf(m_run, __shfl_xor_sync(0xffffffffu, m_run, 1));
  m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, 2));
  const float fix = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_w);
  float l_w = l_run * fix;
  l_w += __shfl_xor_sync(0xffffffffu, l_w, 1);
  l_w += __shfl_xor_sync(0xffffffffu, l_w, 2);
  // (all four lanes of a row always share m_run: the max was reduced across them every block, so fix == 1
  //  whenever the row saw a key; the rescale keeps the code robust to future per-lane maxima)
  if (g < GROUP) {
    if (tig == 0) { s_m[warp][g] = m_w; s_l[warp][g] = l_w; }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s_o[warp][g][n * 8 + 2 * tig] = o[n][0] * fix;
      s_o[warp][g][n * 8 + 2 * tig + 1] = o[n][1] * fix;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < GROUP * HD; idx += DEC_THREADS) {
    const int gg = idx / HD, e = idx % HD;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) mm = fmaxf(mm, s_m[w][gg]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float wgt = (s_m[w][gg] == -INFINITY) ? 0.f : ex2_approx(s_m[w][gg] - mm);
      ll += s_l[w][gg] * wgt;
      oo += s_o[w][gg][e] * wgt;
    }
    const int h = kvh * GROUP + gg;
    const size_t pidx = (static_cast<size_t>(b) * hq + h) * splits + split;
    o_part[pidx * HD + e] = oo;
    if (e == 0) { ml_part[pidx * 2] = mm; ml_part[pidx * 2 + 1] = ll; }
  }
  if (tickets != nullptr) fused_combine_tail<HD, GROUP>(tickets, o_part, ml_part, out, ldo, b, kvh, hq, hkv, splits);
}

// ---------------------------------------------------------------------------------------------
// The same flash-decoding step for any head_dim that is a multiple of 32 (instantiated for 128: Llama-3.1-8B), on a
// tighter register budget: K blocks stay double-buffered in registers, V blocks go HBM -> shared memory by cp.async
// into a two-stage per-warp ring (no registers), and the dead half of every accumulator tile (rows 8..15 of the
// 16-row MMA tile never hold a query head) is ONE shared pair of always-zero registers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816_top(float& d0, float& d1, float& z0, float& z1, uint32_t a0, uint32_t a2,
                                             uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d0), "+f"(d1), "+f"(z0), "+f"(z1)
      : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int HD, int GROUP, bool ROPE>
__global__ void __launch_bounds__(DEC_THREADS)
decode_attn_mma_wide_kernel(const bf16* __restrict__ qkv, int ldq, bf16* __restrict__ kcache, bf16* __restrict__ vcache,
                            const int* __restrict__ pos_dev, const int* __restrict__ slot_dev, int hq, int hkv,
                            int max_seq, int splits, float scale_log2, float* __restrict__ o_part,
                            float* __restrict__ ml_part, int* __restrict__ tickets, bf16* __restrict__ out, int ldo,
                            const bf16* __restrict__ cos_t, const bf16* __restrict__ sin_t) {
  constexpr int NC = HD / 32;    // 16-byte chunks per lane per K row: dims [32c + 8tig, +8), c < NC
  constexpr int CPR = HD / 8;    // 16-byte chunks per V row
  constexpr int VCH = 16 * CPR / 32;  // V chunks per lane per 16-key block
  static_assert(HD % 64 == 0 && GROUP <= 8, "rotate-half pairs chunk c with chunk c + NC/2");
  pdl_launch_dependents();
  pdl_wait();
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int len = pos_dev[b] + 1;
  const int slot = slot_dev ? slot_dev[b] : b;
  int chunk = (len + splits - 1) / splits;
  chunk = (chunk + 15) & ~15;
  const int k_begin = split * chunk;
  const int k_end = min(len, k_begin + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  constexpr int NW = DEC_THREADS / 32;

  // per-warp V ring: 2 stages x 16 keys x HD; 16-byte chunk c of row `key` at (c & ~7) | ((c ^ key) & 7) (conflict-free
  // ldmatrix.trans); after the key loop the warp's own stages hold its partial output rows for the cross-warp merge
  __shared__ __align__(16) bf16 vstage[NW][2][16 * HD];
  __shared__ float s_m[NW][8], s_l[NW][8];
  static_assert(sizeof(float) * 8 * HD <= sizeof(bf16) * 2 * 16 * HD, "partial output rows must fit the warp's ring");

  const bf16* cos_q = ROPE ? cos_t + static_cast<size_t>(len - 1) * (HD / 2) : nullptr;
  const bf16* sin_q = ROPE ? sin_t + static_cast<size_t>(len - 1) * (HD / 2) : nullptr;
  // q fragments (A operand rows = heads; rows >= GROUP and rows 8..15 are zero)
  uint4 qf[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) qf[c] = make_uint4(0, 0, 0, 0);
  if (g < GROUP) {
    const bf16* qp = qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(kvh * GROUP + g) * HD;
#pragma unroll
    for (int c = 0; c < NC; ++c) qf[c] = *reinterpret_cast<const uint4*>(qp + 32 * c + 8 * tig);
    if (ROPE) {
#pragma unroll
      for (int c = 0; c < NC / 2; ++c) rope_lo_hi(qf[c], qf[c + NC / 2], cos_q + 32 * c, sin_q + 32 * c, tig);
    }
  }
  bf16* kbase = kcache + (static_cast<size_t>(slot) * hkv + kvh) * max_seq * HD;
  bf16* vbase = vcache + (static_cast<size_t>(slot) * hkv + kvh) * max_seq * HD;
  const int new_pos = len - 1;  // ROPE: this row is not in the cache yet
  const bf16* knew = qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(hq + kvh) * HD;
  const bf16* vnew = qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(hq + hkv + kvh) * HD;

  float o[HD / 8][2];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o[n][0] = o[n][1] = 0.f;
  float z0 = 0.f, z1 = 0.f;              // rows 8..15 of every accumulator tile: always zero
  float m_run = -INFINITY, l_run = 0.f;  // per (lane row g): partial over this lane's key columns

  struct KStep { uint4 k0[NC], k1[NC]; };  // key kb+g and kb+8+g
  auto load_k = [&](KStep& st, int kb) {
    auto one = [&](int key, uint4 (&kr)[NC]) {
#pragma unroll
      for (int c = 0; c < NC; ++c) kr[c] = make_uint4(0, 0, 0, 0);
      if (key >= k_end) return;
      if (ROPE && key == new_pos) {
#pragma unroll
        for (int c = 0; c < NC; ++c) kr[c] = *reinterpret_cast<const uint4*>(knew + 32 * c + 8 * tig);
#pragma unroll
        for (int c = 0; c < NC / 2; ++c)
          rope_lo_hi(kr[c], kr[c + NC / 2], cos_t + static_cast<size_t>(new_pos) * (HD / 2) + 32 * c,
                     sin_t + static_cast<size_t>(new_pos) * (HD / 2) + 32 * c, tig);
#pragma unroll
        for (int c = 0; c < NC; ++c)
          *reinterpret_cast<uint4*>(kbase + static_cast<size_t>(key) * HD + 32 * c + 8 * tig) = kr[c];
      } else {
        const bf16* kp = kbase + static_cast<size_t>(key) * HD;
#pragma unroll
        for (int c = 0; c < NC; ++c) kr[c] = *reinterpret_cast<const uint4*>(kp + 32 * c + 8 * tig);
      }
    };
    one(kb + g, st.k0);
    one(kb + 8 + g, st.k1);
  };
  // V block of one step into ring stage `stg`: cp.async for cache rows, a direct store for the step's own row (taken
  // from `qkv`, also appended to the cache) and zeros past the key range
  auto issue_v = [&](int stg, int kb) {
#pragma unroll
    for (int i = 0; i < VCH; ++i) {
      const int id = lane + 32 * i, key = id / CPR, c = id % CPR;
      bf16* dst = &vstage[warp][stg][key * HD + (((c & ~7) | ((c ^ key) & 7)) << 3)];
      if (kb + key >= k_end) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      } else if (ROPE && kb + key == new_pos) {
        const uint4 v = *reinterpret_cast<const uint4*>(vnew + 8 * c);
        *reinterpret_cast<uint4*>(vbase + static_cast<size_t>(kb + key) * HD + 8 * c) = v;
        *reinterpret_cast<uint4*>(dst) = v;
      } else {
        cp_async16(smem_u32(dst), vbase + static_cast<size_t>(kb + key) * HD + 8 * c);
      }
    }
    cp_async_commit();
  };

  KStep cur, nxt;
  const int kb_first = k_begin + warp * 16;
  if (kb_first < k_end) {
    load_k(cur, kb_first);
    issue_v(0, kb_first);
  }
  int it = 0;
  for (int kb = kb_first; kb < k_end; kb += NW * 16, ++it) {
    __syncwarp();  // the ldmatrix reads of the previous step are done: its ring stage may be refilled
    if (kb + NW * 16 < k_end) {
      load_k(nxt, kb + NW * 16);
      issue_v((it + 1) & 1, kb + NW * 16);
    } else {
      cp_async_commit();  // keeps the group count uniform for the wait below
    }
    // ---- S = q K^T for two 8-key groups
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      mma16816_top(s00, s01, z0, z1, qf[c].x, qf[c].y, cur.k0[c].x, cur.k0[c].y);
      mma16816_top(s00, s01, z0, z1, qf[c].z, qf[c].w, cur.k0[c].z, cur.k0[c].w);
      mma16816_top(s10, s11, z0, z1, qf[c].x, qf[c].y, cur.k1[c].x, cur.k1[c].y);
      mma16816_top(s10, s11, z0, z1, qf[c].z, qf[c].w, cur.k1[c].z, cur.k1[c].w);
    }
    // lane holds row g (head), keys kb + 2tig, +1 (s00, s01) and kb + 8 + 2tig, +1 (s10, s11)
    float v0 = s00 * scale_log2, v1 = s01 * scale_log2, v2 = s10 * scale_log2, v3 = s11 * scale_log2;
    v0 = (kb + 2 * tig < k_end) ? v0 : -INFINITY;
    v1 = (kb + 2 * tig + 1 < k_end) ? v1 : -INFINITY;
    v2 = (kb + 8 + 2 * tig < k_end) ? v2 : -INFINITY;
    v3 = (kb + 8 + 2 * tig + 1 < k_end) ? v3 : -INFINITY;
    float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = ex2_approx(m_run - m_use);
    const float p0 = ex2_approx(v0 - m_use), p1 = ex2_approx(v1 - m_use), p2 = ex2_approx(v2 - m_use),
                p3 = ex2_approx(v3 - m_use);
    l_run = l_run * alpha + ((p0 + p1) + (p2 + p3));
    m_run = m_new;
    const uint32_t pa0 = pack_bf16x2(p0, p1), pa2 = pack_bf16x2(p2, p3);  // P as A fragment (rows 8..15 zero)
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      o[n][0] *= alpha;
      o[n][1] *= alpha;
    }
    cp_async_wait<1>();  // this lane's share of the current V block has landed ...
    __syncwarp();        // ... and so has every other lane's (plus the direct stores of issue_v)
    // ---- O += P V : B fragments through ldmatrix.trans (matrix rows = keys, cols = 8 d)
    const uint32_t vs_base = smem_u32(&vstage[warp][it & 1][0]);
#pragma unroll
    for (int np = 0; np < HD / 16; ++np) {  // two 8-wide d chunks per ldmatrix.x4
      const int mat = lane >> 3, r = lane & 7;
      const int key = (mat & 1) * 8 + r;
      const int c = np * 2 + (mat >> 1);
      const uint32_t addr = vs_base + (key * HD + (((c & ~7) | ((c ^ key) & 7)) << 3)) * 2;
      uint32_t b0, b1, b2, b3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                   : "r"(addr));
      mma16816_top(o[np * 2][0], o[np * 2][1], z0, z1, pa0, pa2, b0, b1);
      mma16816_top(o[np * 2 + 1][0], o[np * 2 + 1][1], z0, z1, pa0, pa2, b2, b3);
    }
    cur = nxt;
  }
  cp_async_wait<0>();
  __syncwarp();  // every lane is done with the warp's ring: it now holds the warp's partial output rows
  // ---- per-warp row statistics: combine the 4 key-column lanes of each head row
  float m_w = fmaxf(m_run, __shfl_xor_sync(0xffffffffu, m_run, 1));
  m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, 2));
  const float fix = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_w);
  float l_w = l_run * fix;
  l_w += __shfl_xor_sync(0xffffffffu, l_w, 1);
  l_w += __shfl_xor_sync(0xffffffffu, l_w, 2);
  float (*so_w)[HD] = reinterpret_cast<float (*)[HD]>(&vstage[warp][0][0]);  // this warp's [8][HD] floats
  if (g < GROUP) {
    if (tig == 0) { s_m[warp][g] = m_w; s_l[warp][g] = l_w; }
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      so_w[g][n * 8 + 2 * tig] = o[n][0] * fix;
      so_w[g][n * 8 + 2 * tig + 1] = o[n][1] * fix;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < GROUP * HD; idx += DEC_THREADS) {
    const int gg = idx / HD, e = idx % HD;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) mm = fmaxf(mm, s_m[w][gg]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float wgt = (s_m[w][gg] == -INFINITY) ? 0.f : ex2_approx(s_m[w][gg] - mm);
      ll += s_l[w][gg] * wgt;
      oo += reinterpret_cast<const float (*)[HD]>(&vstage[w][0][0])[gg][e] * wgt;
    }
    const int h = kvh * GROUP + gg;
    const size_t pidx = (static_cast<size_t>(b) * hq + h) * splits + split;
    o_part[pidx * HD + e] = oo;
    if (e == 0) { ml_part[pidx * 2] = mm; ml_part[pidx * 2 + 1] = ll; }
  }
  if (tickets != nullptr) fused_combine_tail<HD, GROUP>(tickets, o_part, ml_part, out, ldo, b, kvh, hq, hkv, splits);
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

static int decode_attn_impl(const void* qkv, int ldq, void* kcache, void* vcache, const int* pos_dev,
                            const int* slot_dev, int batch, int hq, int hkv, int head_dim, int max_seq, int splits,
                            float scale, void* o_part_ws, void* ml_part_ws, int* tickets, void* out, int ldo,
                            const void* cos_t, const void* sin_t, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool rope = cos_t != nullptr;
  bf16* kc_w = static_cast<bf16*>(kcache);
  bf16* vc_w = static_cast<bf16*>(vcache);
  const int group = hq / hkv;
  const bool wide_ok = head_dim == 128 && (group == 4 || group == 8 || group == 1);
  if (rope && !wide_ok && !(head_dim == 64 && (group == 4 || group == 2 || group == 1 || group == 8))) {
    gar_set_error("gar_decode_attn_rope: fused RoPE needs head_dim 64 or 128 (got %d, group %d)", head_dim, group);
    return GAR_ERR_SHAPE;
  }
  if (hq % hkv || splits < 1 || splits > 32) return GAR_ERR_ARG;
  const float sl2 = scale * 1.4426950408889634f;
  dim3 grid(splits, hkv, batch);
#define LAUNCH_DEC(HD_, G_)                                                                                        \
  GAR_CUDA_CHECK(gar_launch(decode_attn_split_kernel<HD_, G_>, grid, dim3(DEC_THREADS), 0, stream, true, dim3(1, 1, 1), \
                            static_cast<const bf16*>(qkv), ldq, static_cast<const bf16*>(kcache),                  \
                            static_cast<const bf16*>(vcache), pos_dev, slot_dev, hq, hkv, max_seq, splits, sl2,    \
                            static_cast<float*>(o_part_ws), static_cast<float*>(ml_part_ws), tickets,              \
                            static_cast<bf16*>(out), ldo))
#define LAUNCH_MMA(G_)                                                                                              \
  do {                                                                                                              \
    if (rope)                                                                                                       \
      GAR_CUDA_CHECK(gar_launch(decode_attn_mma_kernel<G_, true>, grid, dim3(DEC_THREADS), 0, stream, true,         \
                                dim3(1, 1, 1), static_cast<const bf16*>(qkv), ldq, kc_w, vc_w, pos_dev, slot_dev, hq,  \
                                hkv, max_seq, splits, sl2, static_cast<float*>(o_part_ws),                          \
                                static_cast<float*>(ml_part_ws), tickets, static_cast<bf16*>(out), ldo,             \
                                static_cast<const bf16*>(cos_t), static_cast<const bf16*>(sin_t)));                 \
    else                                                                                                            \
      GAR_CUDA_CHECK(gar_launch(decode_attn_mma_kernel<G_, false>, grid, dim3(DEC_THREADS), 0, stream, true,        \
                                dim3(1, 1, 1), static_cast<const bf16*>(qkv), ldq, kc_w, vc_w, pos_dev, slot_dev, hq,  \
                                hkv, max_seq, splits, sl2, static_cast<float*>(o_part_ws),                          \
                                static_cast<float*>(ml_part_ws), tickets, static_cast<bf16*>(out), ldo,             \
                                static_cast<const bf16*>(nullptr), static_cast<const bf16*>(nullptr)));             \
  } while (0)
#define LAUNCH_WIDE_1(G_, R_)                                                                                       \
  GAR_CUDA_CHECK(gar_launch(decode_attn_mma_wide_kernel<128, G_, R_>, grid, dim3(DEC_THREADS), 0, stream, true,     \
                            dim3(1, 1, 1), static_cast<const bf16*>(qkv), ldq, kc_w, vc_w, pos_dev, slot_dev, hq,   \
                            hkv, max_seq, splits, sl2, static_cast<float*>(o_part_ws),                              \
                            static_cast<float*>(ml_part_ws), tickets, static_cast<bf16*>(out), ldo,                 \
                            static_cast<const bf16*>(cos_t), static_cast<const bf16*>(sin_t)))
#define LAUNCH_WIDE(G_)                                                                                             \
  do {                                                                                                              \
    if (rope) LAUNCH_WIDE_1(G_, true);                                                                              \
    else LAUNCH_WIDE_1(G_, false);                                                                                  \
  } while (0)
  if (wide_ok && (rope || gar_global_option("decode_attn_mma"))) {
    if (group == 4) LAUNCH_WIDE(4);
    else if (group == 8) LAUNCH_WIDE(8);
    else LAUNCH_WIDE(1);
  } else if (head_dim == 64 && (rope || gar_global_option("decode_attn_mma")) && (group == 4 || group == 2 || group == 1 || group == 8)) {
    if (group == 4) LAUNCH_MMA(4);
    else if (group == 2) LAUNCH_MMA(2);
    else if (group == 1) LAUNCH_MMA(1);
    else LAUNCH_MMA(8);
  } else if (head_dim == 64 && group == 4) LAUNCH_DEC(64, 4);
  else if (head_dim == 64 && group == 2) LAUNCH_DEC(64, 2);
  else if (head_dim == 64 && group == 1) LAUNCH_DEC(64, 1);
  else if (head_dim == 128 && group == 4) LAUNCH_DEC(128, 4);
  else {
    gar_set_error("gar_decode_attn: head_dim=%d group=%d not built", head_dim, group);
    return GAR_ERR_SHAPE;
  }
#undef LAUNCH_DEC
#undef LAUNCH_MMA
#undef LAUNCH_WIDE
#undef LAUNCH_WIDE_1
  GAR_CUDA_CHECK(cudaGetLastError());
  if (tickets != nullptr) return GAR_OK;  // the split kernel's last CTA per (sequence, kv head) already combined
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

extern "C" int gar_decode_attn(const void* qkv, int ldq, const void* kcache, const void* vcache, const int* pos_dev,
                               const int* slot_dev, int batch, int hq, int hkv, int head_dim, int max_seq, int splits,
                               float scale, void* o_part_ws, void* ml_part_ws, int* tickets, void* out, int ldo,
                               void* stream) {
  return decode_attn_impl(qkv, ldq, const_cast<void*>(kcache), const_cast<void*>(vcache), pos_dev, slot_dev, batch, hq,
                          hkv, head_dim, max_seq, splits, scale, o_part_ws, ml_part_ws, tickets, out, ldo, nullptr,
                          nullptr, stream);
}

// Same, with HF apply_rotary_pos_emb + DynamicCache.update fused in: qkv holds the step's UN-rotated q/k/v rows,
// cos/sin tables are [max_pos, head_dim/2]; the new K/V rows are appended to the cache at pos_dev[b].  head_dim 64.
extern "C" int gar_decode_attn_rope(const void* qkv, int ldq, void* kcache, void* vcache, const int* pos_dev,
                                    const int* slot_dev, int batch, int hq, int hkv, int head_dim, int max_seq,
                                    int splits, float scale, const void* cos_t, const void* sin_t, void* o_part_ws,
                                    void* ml_part_ws, int* tickets, void* out, int ldo, void* stream) {
  if (!cos_t || !sin_t) return GAR_ERR_ARG;
  return decode_attn_impl(qkv, ldq, kcache, vcache, pos_dev, slot_dev, batch, hq, hkv, head_dim, max_seq, splits, scale,
                          o_part_ws, ml_part_ws, tickets, out, ldo, cos_t, sin_t, stream);
}
