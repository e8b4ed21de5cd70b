// One persistent kernel per greedy decode step (R9): embedding gather, every Llama layer (RMSNorm -> qkv, RoPE + KV
// append + GQA attention, o-proj + residual, RMSNorm -> gate/up + SwiGLU, down + residual) and the lm_head.
//
// Replaces, per generated token, the ~80 dependent launches of HF Llama's decode step behind
// `mllm.generate` (modeling_gar.py:418-426).  A decode step at batch <= 8 is a pure HBM
// stream (2.47 GB of weights + the live KV cache for GAR-1B); as separate kernels every launch pays its own
// DRAM round trip, ramp and tail, and HBM idles in between.  Here:
//   * one CTA per SM stays resident for the whole step; phases are separated by a grid barrier (generation
//     counter in global memory) instead of kernel boundaries;
//   * a producer warp per CTA streams that CTA's weight tiles (16 output rows x <=2048 k, whole rows as 1-D bulk
//     TMA copies) into a 2-stage shared-memory ring and runs AHEAD across phase and layer boundaries -- weights
//     never depend on activations -- so the HBM stream does not stop while the consumers sit in a barrier or
//     in the attention phase (~19 MB of weights in flight chip-wide; tools/micro/ring_bw.cu: 7.1 TB/s);
//   * 8 consumer warps turn a stage into mma.sync fragments (conflict-free LDS.128: row pitch = 64 mod 128 B)
//     with exactly the arithmetic of gemv_decode.cu: fp32 accumulation, fixed-order reduction over the 8
//     k-slices, the reference's bf16 rounding points in the fused RMSNorm / SwiGLU / residual epilogues;
//   * attention runs as two independent 4-warp units per CTA, flash-decoding over split-KV items with the math of
//     decode_attn.cu (RoPE of q and of the new K row, KV append, split partials, ticketed combine); K/V blocks
//     travel global -> shared with cp.async, three 16-key blocks deep per warp, so no score/value tile ever sits
//     in registers and two blocks per warp are always in flight.
// With ~225 KB of shared memory per CTA the L1 is a few KB: a register spill costs an L2 round trip.  Every phase
// body therefore exists exactly once in the code (one loop over a phase table) and keeps its state small; the
// producer lives in its own warpgroup so setmaxnreg can hand its registers to the consumers.
// Activations written by other CTAs are always read with ld.global.cg (L1 is not coherent inside one kernel).
#include "../../include/gar_b200.h"
#include "common.cuh"

namespace {

constexpr int MG_CW = 8;                 // consumer warps
constexpr int MG_CT = MG_CW * 32;        // consumer threads
constexpr int MG_THREADS = MG_CT + 128;  // + the producer warpgroup (warp 8 streams, warps 9-11 only donate registers)
constexpr int MG_REG_PROD = 40;          // setmaxnreg: 384 x 168 launch registers -> producers 40, consumers 232
constexpr int MG_REG_CONS = 232;
constexpr int MG_STAGES = 2;
constexpr int MG_KS_MAX = 2048;          // k-width of one stage
constexpr int MG_PAD = 64;               // row pitch = KS*2 + 64 bytes
constexpr int MG_STAGE_BYTES = 16 * (MG_KS_MAX * 2 + MG_PAD);
constexpr int MG_PART_FLOATS = MG_CW * 16 * 9;
constexpr int MG_UW_MAX = 4;             // warps per attention unit: 2 (default) or 4 (args.unit_warps)
constexpr int MG_HD = 64;
constexpr int MG_KVB = 3;                // 16-key K/V blocks in flight per warp
constexpr int MG_BLK_BYTES = 2 * 16 * MG_HD * 2;           // K block + V block
constexpr int MG_WARP_KV = MG_KVB * MG_BLK_BYTES;          // 12 KB per warp
constexpr int MG_ATT_BYTES = MG_CW * MG_WARP_KV;           // 96 KB; also holds the GEMV reduction buffers
constexpr int MG_SMEM = MG_STAGES * MG_STAGE_BYTES + MG_ATT_BYTES;
static_assert(2 * MG_PART_FLOATS * 4 <= MG_ATT_BYTES, "GEMV reduction buffers alias the attention region");

struct MegaLayer {
  const bf16 *wqkv, *wo, *wgu, *wdown, *ln1, *ln2;
  bf16 *kc, *vc;  // this layer's cache, already offset to the first slot of the decode group
};

constexpr int MG_MAX_LAYERS = 32;  // the layer table travels in the kernel parameters (constant bank, no load latency)

struct MegaArgs {
  MegaLayer layers[MG_MAX_LAYERS];
  int n_layers;
  const bf16 *table, *lm_head, *final_norm, *cos_t, *sin_t;
  const int *next_tok, *pos;
  bf16 *h, *qkv, *att, *act, *logits;
  float *sumsq_a, *sumsq_b, *o_part, *ml_part;
  int* tickets;
  unsigned long long* gbar;  // arrival counter of the grid barrier (never reset)
  int batch, D, I, hq, hkv, qkv_w, vocab_padded, max_seq, splits, unit_warps;
  float eps, scale_log2;
  unsigned long long* timeline;  // optional debug: globaltimer stamps of CTA 0 (work end / barrier end per phase)
};

enum { MG_EPI_NONE = 0, MG_EPI_RES = 1, MG_EPI_SWIGLU = 2 };

// one GEMV phase: out[M, N] = epi( norm?(X)[M, K] * W[N, K]^T )
struct GemvRun {
  const bf16* W;
  int ldw, tiles, K, KS, epi;
  const bf16* X;
  int ldx;
  const bf16* norm_w;
  const float* sumsq_in;
  int n_part;
  bf16* out;
  int ldo;
  const bf16* res;
  float* sumsq_out;
};

__device__ __forceinline__ int ks_of(int K) { return K < MG_KS_MAX ? K : MG_KS_MAX; }

// Phase p of the step: p = 5 * layer + {0 qkv, 1 attention, 2 o-proj, 3 gate/up, 4 down}; p = 5 * n_layers: lm_head.
// Returns false for the attention phase (no weight stream).
__device__ __forceinline__ bool phase_gemv(const MegaArgs& a, int p, GemvRun& r) {
  const int D = a.D, I = a.I, hqd = a.hq * MG_HD;
  const int n_part_h = p < 5 ? 1 : D / 16;  // the first RMSNorm statistics come from the embedding gather
  if (p == 5 * a.n_layers) {
    r = {a.lm_head, D, a.vocab_padded / 16, D, ks_of(D), MG_EPI_NONE, a.h, D, a.final_norm, a.sumsq_a, n_part_h,
         a.logits, a.vocab_padded, nullptr, nullptr};
    return true;
  }
  const MegaLayer& ly = a.layers[p / 5];
  switch (p % 5) {
    case 0:
      r = {ly.wqkv, D, a.qkv_w / 16, D, ks_of(D), MG_EPI_NONE, a.h, D, ly.ln1, a.sumsq_a, n_part_h, a.qkv, a.qkv_w,
           nullptr, nullptr};
      return true;
    case 2:
      r = {ly.wo, hqd, D / 16, hqd, ks_of(hqd), MG_EPI_RES, a.att, hqd, nullptr, nullptr, 0, a.h, D, a.h, a.sumsq_b};
      return true;
    case 3:
      r = {ly.wgu, D, 2 * I / 16, D, ks_of(D), MG_EPI_SWIGLU, a.h, D, ly.ln2, a.sumsq_b, D / 16, a.act, I, nullptr,
           nullptr};
      return true;
    case 4:
      r = {ly.wdown, I, D / 16, I, ks_of(I), MG_EPI_RES, a.act, I, nullptr, nullptr, 0, a.h, D, a.h, a.sumsq_a};
      return true;
    default:
      return false;
  }
}

__device__ __forceinline__ void tile_rows(int swiglu, int t, int& lo, int& hi) {
  if (swiglu) {  // gu rows are interleaved [32 gate | 32 up]: a tile = 8 gate rows + their 8 up rows
    lo = (t >> 2) * 64 + (t & 3) * 8;
    hi = lo + 32;
  } else {
    lo = t * 16;
    hi = lo + 8;
  }
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 16-byte global -> shared copy without registers; src_bytes = 0 zero-fills (rows past the key range)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 ldcg128(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Grid barrier over the consumer threads of all CTAs (the producer warps never take part: they only wait for ring
// space): ONE monotonically increasing 64-bit arrival counter -- every barrier adds exactly gridDim.x to it, so the
// k-th barrier of a launch completes at (value at launch, a multiple of gridDim.x) + k * gridDim.x.  One release
// reduction + an acquire poll per CTA.  Bounded spin: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned long long* cnt, unsigned long long& target, unsigned long long* tl,
                                             int& tl_n) {
  named_bar_sync(1, MG_CT);
  if (tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tl[tl_n++] = globaltimer_ns();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(cnt) : "memory");
    unsigned long long spins = 0;
    while (ld_acquire_u64(cnt) < target) {
      if (++spins > (1ull << 28)) {
        printf("gar_b200: decode grid barrier timeout block=%d target=%llu\n", blockIdx.x, target);
        __trap();
      }
    }
  }
  named_bar_sync(1, MG_CT);
  if (tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tl[tl_n++] = globaltimer_ns();
}

// ---------------------------------------------------------------------------------------------------------------
// producer: this CTA's tiles of one GEMV phase -> ring
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void produce_phase(const GemvRun& d, uint32_t ring, uint32_t full0, uint32_t empty0,
                                              uint32_t& it, int lane) {
  const int nch = d.K / d.KS;
  const uint32_t row_bytes = static_cast<uint32_t>(d.KS) * 2u, pitch = row_bytes + MG_PAD;
  for (int t = blockIdx.x; t < d.tiles; t += gridDim.x) {
    int lo, hi;
    tile_rows(d.epi == MG_EPI_SWIGLU, t, lo, hi);
    for (int c = 0; c < nch; ++c, ++it) {
      const uint32_t st = it % MG_STAGES, ph = (it / MG_STAGES) & 1u;
      mbar_wait(empty0 + 8 * st, ph ^ 1u);
      if (lane == 0) mbar_expect_tx(full0 + 8 * st, 16u * row_bytes);
      __syncwarp();
      if (lane < 16) {
        const int row = lane < 8 ? lo + lane : hi + (lane - 8);
        bulk_g2s(ring + st * MG_STAGE_BYTES + lane * pitch,
                 d.W + static_cast<size_t>(row) * d.ldw + static_cast<size_t>(c) * d.KS, row_bytes, full0 + 8 * st);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// consumers: this CTA's tiles of one GEMV phase (same arithmetic as gemv_decode.cu)
// ---------------------------------------------------------------------------------------------------------------
// activation fragments of one warp k-slice (B operand: batch row g, 8 consecutive k per step): raw loads first, the
// RMSNorm x' = bf16(w * bf16(x * rstd)) (the reference's rounding points) applied once the statistics are known
template <int STEPS>
__device__ __forceinline__ void load_x_raw(uint4 (&xv)[STEPS], uint4 (&nw)[STEPS], const bf16* X, int ldx,
                                           const bf16* norm_w, bool xvalid, int g, int k0) {
#pragma unroll
  for (int u = 0; u < STEPS; ++u) {
    xv[u] = make_uint4(0, 0, 0, 0);
    nw[u] = make_uint4(0, 0, 0, 0);
    if (xvalid) xv[u] = ldcg128(X + static_cast<size_t>(g) * ldx + k0 + u * 32);
    if (norm_w != nullptr) nw[u] = __ldg(reinterpret_cast<const uint4*>(norm_w + k0 + u * 32));
  }
}
template <int STEPS>
__device__ __forceinline__ void norm_x(uint4 (&xv)[STEPS], const uint4 (&nw)[STEPS], float rstd) {
#pragma unroll
  for (int u = 0; u < STEPS; ++u) {
    uint32_t* xw = reinterpret_cast<uint32_t*>(&xv[u]);
    const uint32_t* ww = reinterpret_cast<const uint32_t*>(&nw[u]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(xw[j]), w2 = unpack_bf16x2(ww[j]);
      xw[j] = pack_bf16x2(w2.x * bf16r(f.x * rstd), w2.y * bf16r(f.y * rstd));
    }
  }
}

template <int STEPS>  // k-steps of 32 per warp and stage: KS = 256 * STEPS
__device__ __forceinline__ void consume_tiles(const GemvRun& d, int M, float eps, uint32_t ring, float* part_base,
                                              uint32_t full0, uint32_t empty0, uint32_t& it) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int nch = d.K / d.KS;
  const int slice = d.KS / MG_CW;
  const uint32_t pitch = static_cast<uint32_t>(d.KS) * 2u + MG_PAD;
  const bool xvalid = g < M;

  // every load this phase start depends on goes out first (activations, norm weights, statistics): one L2 round trip
  uint4 xv[STEPS], xn[STEPS], nw[STEPS];
  const int kw = warp * slice + 8 * tig;
  load_x_raw<STEPS>(xn, nw, d.X, d.ldx, d.norm_w, xvalid, g, kw);
  float rstd = 0.f;
  if (d.norm_w != nullptr) {
    // fixed-order sum of the producer phase's per-tile sums of squares for batch row g: lane tig owns partials
    // tig + 4j (+64, +128, ...); two 16-wide batches of loads are in flight together
    float acc16[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc16[j] = 0.f;
    if (xvalid) {
      for (int p0 = tig; p0 < d.n_part; p0 += 128) {
        float v0[16], v1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int p = p0 + 4 * j;
          v0[j] = p < d.n_part ? __ldcg(d.sumsq_in + static_cast<size_t>(p) * 8 + g) : 0.f;
          v1[j] = p + 64 < d.n_part ? __ldcg(d.sumsq_in + static_cast<size_t>(p + 64) * 8 + g) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          acc16[j] += v0[j];
          acc16[j] += v1[j];
        }
      }
    }
#pragma unroll
    for (int w = 8; w > 0; w >>= 1)
#pragma unroll
      for (int j = 0; j < w; ++j) acc16[j] += acc16[j + w];
    float s = acc16[0];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    rstd = rsqrtf(s / static_cast<float>(d.K) + eps);
  }

  int tcount = 0;
  for (int t = blockIdx.x; t < d.tiles; t += gridDim.x, ++tcount) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < nch; ++c, ++it) {
      if (c == 0 ? tcount == 0 || nch > 1 : true) {
        // adopt the fragments loaded ahead (xn) and, in a multi-chunk phase, put the next chunk's loads in flight
#pragma unroll
        for (int u = 0; u < STEPS; ++u) xv[u] = xn[u];
        if (d.norm_w != nullptr) {
          if (nch > 1 && (tcount > 0 || c > 0)) {  // (the norm weights are only prefetched at the phase start)
#pragma unroll
            for (int u = 0; u < STEPS; ++u) nw[u] = __ldg(reinterpret_cast<const uint4*>(d.norm_w + c * d.KS + kw + u * 32));
          }
          norm_x<STEPS>(xv, nw, rstd);
        }
        if (nch > 1) {
          const int cn = c + 1 < nch ? c + 1 : 0;  // the next tile starts over at chunk 0
#pragma unroll
          for (int u = 0; u < STEPS; ++u) {
            xn[u] = make_uint4(0, 0, 0, 0);
            if (xvalid) xn[u] = ldcg128(d.X + static_cast<size_t>(g) * d.ldx + cn * d.KS + kw + u * 32);
          }
        }
      }
      const uint32_t st = it % MG_STAGES, ph = (it / MG_STAGES) & 1u;
      mbar_wait(full0 + 8 * st, ph);
      const uint32_t base = ring + st * MG_STAGE_BYTES + g * pitch + (warp * slice + 8 * tig) * 2;
      constexpr int H0 = (STEPS + 1) / 2;  // two batches of fragments: <= 32 weight registers live at a time
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        constexpr int HN = H0;
        uint4 wl[HN], wh[HN];
#pragma unroll
        for (int u = 0; u < HN; ++u) {
          if (hb * H0 + u < STEPS) {
            wl[u] = lds128(base + (hb * H0 + u) * 64);
            wh[u] = lds128(base + 8 * pitch + (hb * H0 + u) * 64);
          } else {
            wl[u] = wh[u] = make_uint4(0, 0, 0, 0);
          }
        }
        if (hb == (STEPS > 1 ? 1 : 0)) {  // the last fragments of the stage are in registers: the producer may refill it
          __syncwarp();
          if (lane == 0) mbar_arrive(empty0 + 8 * st);
        }
#pragma unroll
        for (int u = 0; u < HN; ++u) {
          if (hb * H0 + u < STEPS) {
            mma16816(acc, wl[u].x, wh[u].x, wl[u].y, wh[u].y, xv[hb * H0 + u].x, xv[hb * H0 + u].y);
            mma16816(acc, wl[u].z, wh[u].z, wl[u].w, wh[u].w, xv[hb * H0 + u].z, xv[hb * H0 + u].w);
          }
        }
      }
    }
    // ---- fixed-order reduction over the 8 k-slices, then the fused epilogue (two part buffers: one barrier/tile)
    float* part = part_base + (tcount & 1) * MG_PART_FLOATS;
    part[(warp * 16 + g) * 9 + 2 * tig] = acc[0];
    part[(warp * 16 + g) * 9 + 2 * tig + 1] = acc[1];
    part[(warp * 16 + g + 8) * 9 + 2 * tig] = acc[2];
    part[(warp * 16 + g + 8) * 9 + 2 * tig + 1] = acc[3];
    named_bar_sync(1, MG_CT);
    const int idx = threadIdx.x;
    if (d.epi == MG_EPI_SWIGLU) {
      if (idx < 64) {  // 8 features x 8 batch rows: rows 0-7 of the tile are gate rows, rows 8-15 their up rows
        const int f = idx & 7, b = idx >> 3;
        float gsum = 0.f, usum = 0.f;
#pragma unroll
        for (int w = 0; w < MG_CW; ++w) {
          gsum += part[(w * 16 + f) * 9 + b];
          usum += part[(w * 16 + 8 + f) * 9 + b];
        }
        const float y = bf16r(silu_f(bf16r(gsum))) * bf16r(usum);
        if (b < M) d.out[static_cast<size_t>(b) * d.ldo + (t >> 2) * 32 + (t & 3) * 8 + f] = __float2bfloat16_rn(y);
      }
    } else if (idx < 128) {
      const int f = idx & 15, b = idx >> 4;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < MG_CW; ++w) s += part[(w * 16 + f) * 9 + b];
      float y = bf16r(s);
      const int gf = t * 16 + f;
      float o = 0.f;
      if (b < M) {
        if (d.epi == MG_EPI_RES) y += bf2f(__ldcg(d.res + static_cast<size_t>(b) * d.ldo + gf));
        const bf16 ob = __float2bfloat16_rn(y);
        d.out[static_cast<size_t>(b) * d.ldo + gf] = ob;
        o = bf2f(ob);
      }
      if (d.sumsq_out != nullptr) {
        float q = o * o;  // the 16 features of one batch row sit in one half-warp
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
        if (f == 0) d.sumsq_out[static_cast<size_t>(t) * 8 + b] = q;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// attention unit (128 threads): one (split, kv head, sequence) item of flash-decoding, GROUP = 4 query heads per
// kv head, head_dim 64 -- the math of decode_attn_mma_kernel<4, true>
// ---------------------------------------------------------------------------------------------------------------
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

// shared-memory layout of one 16-key block: K [16][64] with 16-byte chunk c of key r at c ^ ((r & 1) << 2) (the two
// rows a quarter-warp reads land in different bank halves), then V [16][64] with chunk c at c ^ (r & 7) (ldmatrix)
__device__ __forceinline__ uint32_t k_off(int key, int c) { return key * 128 + ((c ^ ((key & 1) << 2)) << 4); }
__device__ __forceinline__ uint32_t v_off(int key, int c) { return 16 * 128 + key * 128 + ((c ^ (key & 7)) << 4); }

template <int GROUP>
__device__ __forceinline__ void attn_unit_item(const MegaArgs& a, const MegaLayer& ly, int split, int kvh, int b,
                                               uint8_t* usmem, int ubar) {
  constexpr int HD = MG_HD;
  const int NW = a.unit_warps, UT = NW * 32;
  const int utid = threadIdx.x & (UT - 1);
  const int lane = utid & 31, warp = utid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int len = __ldcg(a.pos + b) + 1;
  int chunk = (len + a.splits - 1) / a.splits;
  chunk = (chunk + 15) & ~15;
  const int k_begin = split * chunk;
  const int k_end = min(len, k_begin + chunk);
  const int ldq = a.qkv_w;

  const uint32_t wbuf = smem_u32(usmem) + warp * MG_WARP_KV;   // this warp's MG_KVB blocks
  float* s_o = reinterpret_cast<float*>(usmem);                // [NW][8][HD], aliases the blocks after the loop
  float* s_m = s_o + MG_UW_MAX * 8 * HD;                       // [NW][8]
  float* s_l = s_m + MG_UW_MAX * 8;

  bf16* kbase = ly.kc + (static_cast<size_t>(b) * a.hkv + kvh) * a.max_seq * HD;
  bf16* vbase = ly.vc + (static_cast<size_t>(b) * a.hkv + kvh) * a.max_seq * HD;
  const int new_pos = len - 1;  // this step's K/V row: not in the cache yet

  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[n][i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  // block `kb` -> buffer `buf`: 8 lanes per key, 16-byte chunks, zero-fill outside [.., k_end) and for the new row
  auto issue_block = [&](int kb, int buf) {
    if (kb < k_end) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int id = lane + 32 * i, key = id >> 3, c = id & 7;
        const int kk = kb + key;
        const bool live = kk < k_end && kk != new_pos;
        const size_t goff = static_cast<size_t>(live ? kk : 0) * HD + 8 * c;
        cp_async16(wbuf + buf * MG_BLK_BYTES + k_off(key, c), kbase + goff, live ? 16u : 0u);
        cp_async16(wbuf + buf * MG_BLK_BYTES + v_off(key, c), vbase + goff, live ? 16u : 0u);
      }
    }
    cp_async_commit();
  };
  const int kb_first = k_begin + warp * 16;
  issue_block(kb_first, 0);
  issue_block(kb_first + NW * 16, 1);
  // (the K/V stream is in flight before the query is fetched and rotated)
  uint32_t qa[4] = {0, 0, 0, 0}, qb[4] = {0, 0, 0, 0};
  if (g < GROUP) {
    const bf16* qp = a.qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(kvh * GROUP + g) * HD;
    uint4 x = ldcg128(qp + 8 * tig);
    uint4 y = ldcg128(qp + 32 + 8 * tig);
    rope_lo_hi(x, y, a.cos_t + static_cast<size_t>(len - 1) * (HD / 2), a.sin_t + static_cast<size_t>(len - 1) * (HD / 2), tig);
    qa[0] = x.x; qa[1] = x.y; qa[2] = x.z; qa[3] = x.w;
    qb[0] = y.x; qb[1] = y.y; qb[2] = y.z; qb[3] = y.w;
  }
  // DynamicCache.update of the step's row, prepared while the first blocks are in flight: the unit whose key range
  // holds new_pos rotates K (lanes 0-3: chunks c and c+4), copies V (lanes 8-15) and appends both to the cache; the
  // warp that later meets new_pos in a block only drops the registers into shared memory
  uint4 nk_lo = make_uint4(0, 0, 0, 0), nk_hi = nk_lo;
  const bool owns_new = new_pos >= k_begin && new_pos < k_end && ((new_pos - k_begin) >> 4) % NW == warp;
  if (owns_new) {
    if (lane < 4) {
      const bf16* knew = a.qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(a.hq + kvh) * HD;
      nk_lo = ldcg128(knew + 8 * lane);
      nk_hi = ldcg128(knew + 32 + 8 * lane);
      rope_lo_hi(nk_lo, nk_hi, a.cos_t + static_cast<size_t>(new_pos) * (HD / 2),
                 a.sin_t + static_cast<size_t>(new_pos) * (HD / 2), lane);
      *reinterpret_cast<uint4*>(kbase + static_cast<size_t>(new_pos) * HD + 8 * lane) = nk_lo;
      *reinterpret_cast<uint4*>(kbase + static_cast<size_t>(new_pos) * HD + 32 + 8 * lane) = nk_hi;
    } else if (lane >= 8 && lane < 16) {
      const bf16* vnew = a.qkv + static_cast<size_t>(b) * ldq + static_cast<size_t>(a.hq + a.hkv + kvh) * HD;
      nk_lo = ldcg128(vnew + 8 * (lane - 8));
      *reinterpret_cast<uint4*>(vbase + static_cast<size_t>(new_pos) * HD + 8 * (lane - 8)) = nk_lo;
    }
  }
  int buf = 0;
  for (int kb = kb_first; kb < k_end; kb += NW * 16) {
    issue_block(kb + 2 * NW * 16, buf == 0 ? 2 : buf - 1);  // two blocks ahead, into the buffer consumed last
    cp_async_wait<2>();
    __syncwarp();  // every lane's copies of this block have landed (the fix-up below overwrites other lanes' rows)
    const uint32_t bb = wbuf + buf * MG_BLK_BYTES;
    if (owns_new && new_pos >= kb && new_pos < kb + 16) {
      const int key = new_pos - kb;
      if (lane < 4) {
        sts128(bb + k_off(key, lane), nk_lo);
        sts128(bb + k_off(key, lane + 4), nk_hi);
      } else if (lane >= 8 && lane < 16) {
        sts128(bb + v_off(key, lane - 8), nk_lo);
      }
    }
    __syncwarp();
    // ---- S = q K^T for two 8-key groups (K fragments: keys g and g+8, d permuted as [8tig,+8) and [32+8tig,+8))
    const uint4 k0a = lds128(bb + k_off(g, tig)), k0b = lds128(bb + k_off(g, tig + 4));
    const uint4 k1a = lds128(bb + k_off(g + 8, tig)), k1b = lds128(bb + k_off(g + 8, tig + 4));
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    mma16816(s0, qa[0], 0u, qa[1], 0u, k0a.x, k0a.y);
    mma16816(s0, qa[2], 0u, qa[3], 0u, k0a.z, k0a.w);
    mma16816(s0, qb[0], 0u, qb[1], 0u, k0b.x, k0b.y);
    mma16816(s0, qb[2], 0u, qb[3], 0u, k0b.z, k0b.w);
    mma16816(s1, qa[0], 0u, qa[1], 0u, k1a.x, k1a.y);
    mma16816(s1, qa[2], 0u, qa[3], 0u, k1a.z, k1a.w);
    mma16816(s1, qb[0], 0u, qb[1], 0u, k1b.x, k1b.y);
    mma16816(s1, qb[2], 0u, qb[3], 0u, k1b.z, k1b.w);
    float v0 = s0[0] * a.scale_log2, v1 = s0[1] * a.scale_log2, v2 = s1[0] * a.scale_log2, v3 = s1[1] * a.scale_log2;
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
    // ---- O += P V : B fragments through ldmatrix.trans (matrix rows = keys, cols = 8 d)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int mat = lane >> 3, r = lane & 7;
      const int key = (mat & 1) * 8 + r;
      const int c = np * 2 + (mat >> 1);
      uint32_t b0, b1, b2, b3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                   : "r"(bb + v_off(key, c)));
      mma16816(o[np * 2], pa0, 0u, pa2, 0u, b0, b1);
      mma16816(o[np * 2 + 1], pa0, 0u, pa2, 0u, b2, b3);
    }
    __syncwarp();  // every lane is done with this block before it is refilled
    buf = buf == MG_KVB - 1 ? 0 : buf + 1;
  }
  cp_async_wait<0>();
  float m_w = fmaxf(m_run, __shfl_xor_sync(0xffffffffu, m_run, 1));
  m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, 2));
  const float fix = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_w);
  float l_w = l_run * fix;
  l_w += __shfl_xor_sync(0xffffffffu, l_w, 1);
  l_w += __shfl_xor_sync(0xffffffffu, l_w, 2);
  named_bar_sync(ubar, UT);  // every warp of the unit is done with its blocks: the partial buffers alias them
  if (g < GROUP) {
    if (tig == 0) { s_m[warp * 8 + g] = m_w; s_l[warp * 8 + g] = l_w; }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      s_o[(warp * 8 + g) * HD + n * 8 + 2 * tig] = o[n][0] * fix;
      s_o[(warp * 8 + g) * HD + n * 8 + 2 * tig + 1] = o[n][1] * fix;
    }
  }
  named_bar_sync(ubar, UT);
  for (int idx = utid; idx < GROUP * HD; idx += UT) {
    const int gg = idx / HD, e = idx % HD;
    float mm = -INFINITY;
    for (int w = 0; w < NW; ++w) mm = fmaxf(mm, s_m[w * 8 + gg]);
    float ll = 0.f, oo = 0.f;
    for (int w = 0; w < NW; ++w) {
      const float wgt = (s_m[w * 8 + gg] == -INFINITY) ? 0.f : ex2_approx(s_m[w * 8 + gg] - mm);
      ll += s_l[w * 8 + gg] * wgt;
      oo += s_o[(w * 8 + gg) * HD + e] * wgt;
    }
    const int h = kvh * GROUP + gg;
    const size_t pidx = (static_cast<size_t>(b) * a.hq + h) * a.splits + split;
    a.o_part[pidx * HD + e] = oo;
    if (e == 0) { a.ml_part[pidx * 2] = mm; a.ml_part[pidx * 2 + 1] = ll; }
  }
  // ---- ticketed combine: the last unit of this (sequence, kv head) merges the splits of its GROUP heads
  __threadfence();
  named_bar_sync(ubar, UT);
  int* flag = reinterpret_cast<int*>(s_l + MG_UW_MAX * 8);
  if (utid == 0) {
    const int t = atomicAdd(&a.tickets[b * a.hkv + kvh], 1);
    const int last = (t == a.splits - 1);
    if (last) a.tickets[b * a.hkv + kvh] = 0;
    *flag = last;
  }
  named_bar_sync(ubar, UT);
  if (*flag) {  // unit-uniform
    __threadfence();
    // all statistics in one round trip, then every thread's partial O values in one more (the sums keep the
    // per-kernel combine's order: s = 0, 1, ...)
    float* c_m = s_o;                // [GROUP][32]
    float* c_l = s_o + GROUP * 32;   // [GROUP][32]
    // each thread owns GROUP*HD/UT outputs (2 or 4) and walks the splits in batches of 8 with every load of a batch in
    // flight together (the first batch goes out with the statistics); per output the sums still run s = 0, 1, ...
    // (the per-kernel combine's order)
    constexpr int MAXO = GROUP * HD / 64;
    const int nout = GROUP * HD / UT;
    float ov[MAXO][8];
    auto load_batch = [&](int s0) {
#pragma unroll
      for (int j = 0; j < MAXO; ++j) {
        const int idx = utid + j * UT, gg = idx / HD, e = idx % HD;
        const size_t base = (static_cast<size_t>(b) * a.hq + kvh * GROUP + gg) * a.splits;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          ov[j][q] = (j < nout && s0 + q < a.splits) ? __ldcg(a.o_part + (base + s0 + q) * HD + e) : 0.f;
      }
    };
    load_batch(0);
    for (int i = utid; i < GROUP * a.splits; i += UT) {
      const int gg = i / a.splits, sp = i % a.splits;
      const size_t base = (static_cast<size_t>(b) * a.hq + kvh * GROUP + gg) * a.splits;
      const float2 ml = __ldcg(reinterpret_cast<const float2*>(a.ml_part + (base + sp) * 2));
      c_m[gg * 32 + sp] = ml.x;
      c_l[gg * 32 + sp] = ml.y;
    }
    named_bar_sync(ubar, UT);
    float mm[MAXO], ll[MAXO], oo[MAXO];
#pragma unroll
    for (int j = 0; j < MAXO; ++j) {
      mm[j] = -INFINITY;
      ll[j] = 0.f;
      oo[j] = 0.f;
      if (j < nout) {
        const int gg = (utid + j * UT) / HD;
        for (int sp = 0; sp < a.splits; ++sp) mm[j] = fmaxf(mm[j], c_m[gg * 32 + sp]);
      }
    }
    for (int s0 = 0; s0 < a.splits; s0 += 8) {
      if (s0 > 0) load_batch(s0);
#pragma unroll
      for (int j = 0; j < MAXO; ++j) {
        const int gg = (utid + j * UT) / HD;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (j < nout && s0 + q < a.splits) {
            const float ms = c_m[gg * 32 + s0 + q];
            const float w = (ms == -INFINITY) ? 0.f : exp2f(ms - mm[j]);
            ll[j] += c_l[gg * 32 + s0 + q] * w;
            oo[j] += ov[j][q] * w;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < MAXO; ++j) {
      if (j < nout) {
        const int idx = utid + j * UT, gg = idx / HD, e = idx % HD;
        a.att[static_cast<size_t>(b) * (a.hq * HD) + (kvh * GROUP + gg) * HD + e] =
            __float2bfloat16_rn(ll[j] > 0.f ? oo[j] / ll[j] : 0.f);
      }
    }
  }
  named_bar_sync(ubar, UT);  // the unit's shared memory is free for its next item
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * MG_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_u32(smem);
  uint8_t* att_smem = smem + MG_STAGES * MG_STAGE_BYTES;
  float* part_base = reinterpret_cast<float*>(att_smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * MG_STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MG_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, MG_CW);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int n_phases = 5 * a.n_layers + 1;
  uint32_t it = 0;

  if (warp >= MG_CW) {
    // ================= producer warpgroup: warp 8 streams this CTA's weight tiles in consumption order ==========
    setmaxnreg_dec<MG_REG_PROD>();
    if (warp != MG_CW) return;
    for (int p = 0; p < n_phases; ++p) {
      GemvRun r;
      if (phase_gemv(a, p, r)) produce_phase(r, ring, full0, empty0, it, lane);
    }
    return;
  }

  // ================= consumers =================
  setmaxnreg_inc<MG_REG_CONS>();
  unsigned long long target = 0;
  int tl_n = 0;
  unsigned long long* tl = a.timeline;
  if (tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tl[tl_n++] = globaltimer_ns();
  if (threadIdx.x == 0) {  // only thread 0 ever uses it; no barrier of this launch can complete before this read
    const unsigned long long c0 = ld_acquire_u64(a.gbar);
    target = c0 - c0 % gridDim.x;
  }

  // ---- token embedding of the step's input ids + the first RMSNorm statistics (n_part = 1)
  if (blockIdx.x < a.batch) {
    const int b = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(a.table + static_cast<size_t>(__ldcg(a.next_tok + b)) * a.D);
    uint4* dst = reinterpret_cast<uint4*>(a.h + static_cast<size_t>(b) * a.D);
    float q = 0.f;
    for (int v = threadIdx.x; v < a.D / 8; v += MG_CT) {
      const uint4 x = __ldg(src + v);
      dst[v] = x;
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(xw[j]);
        q += f.x * f.x + f.y * f.y;
      }
    }
    q = warp_sum(q);
    if (lane == 0) part_base[warp] = q;
    named_bar_sync(1, MG_CT);
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < MG_CW; ++w) t += part_base[w];
      a.sumsq_a[b] = t;
    }
  }
  grid_barrier(a.gbar, target, tl, tl_n);

  const int units = MG_CW / a.unit_warps;       // attention units per CTA
  const int unit = warp / a.unit_warps;         // attention unit of this warp
  for (int p = 0; p < n_phases; ++p) {
    GemvRun r;
    if (phase_gemv(a, p, r)) {
      switch (r.KS / 256) {
        case 1: consume_tiles<1>(r, a.batch, a.eps, ring, part_base, full0, empty0, it); break;
        case 2: consume_tiles<2>(r, a.batch, a.eps, ring, part_base, full0, empty0, it); break;
        case 4: consume_tiles<4>(r, a.batch, a.eps, ring, part_base, full0, empty0, it); break;
        case 6: consume_tiles<6>(r, a.batch, a.eps, ring, part_base, full0, empty0, it); break;
        default: consume_tiles<8>(r, a.batch, a.eps, ring, part_base, full0, empty0, it); break;
      }
    } else {
      // RoPE + KV append + attention, two 4-warp units per CTA
      const MegaLayer& ly = a.layers[p / 5];
      const int items = a.splits * a.hkv * a.batch;
      for (int item = units * blockIdx.x + unit; item < items; item += units * gridDim.x) {
        const int split = item % a.splits, kvh = (item / a.splits) % a.hkv, b = item / (a.splits * a.hkv);
        attn_unit_item<4>(a, ly, split, kvh, b, att_smem + unit * (a.unit_warps * MG_WARP_KV), 2 + unit);
      }
    }
    // debug: every CTA's arrival time at the barrier after phase 26 (layer 5 attention) / 28 (layer 5 gate-up)
    if (tl != nullptr && threadIdx.x == 0 && (p == 26 || p == 28)) tl[512 + (p == 28 ? 256 : 0) + blockIdx.x] = globaltimer_ns();
    if (p + 1 < n_phases) grid_barrier(a.gbar, target, tl, tl_n);
  }
  if (tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tl[tl_n++] = globaltimer_ns();
}

}  // namespace

static unsigned long long* g_mega_timeline = nullptr;
// debug hook: device buffer of >= 16 + 10 * n_layers uint64 receiving CTA 0's globaltimer stamps (NULL disables):
// [start, (work end, barrier end) x (embed, then per layer qkv, attention, o, gate/up, down), lm_head end]
extern "C" int gar_decode_mega_debug_timeline(void* dev_buf) {
  g_mega_timeline = static_cast<unsigned long long*>(dev_buf);
  return GAR_OK;
}

extern "C" int gar_decode_mega_supported(int batch, int D, int I, int hq, int hkv, int head_dim, int vocab_padded) {
  auto ok_k = [](int K) {  // one stage = min(K, 2048) columns, 256 * {1, 2, 4, 6, 8}
    const int ks = K < MG_KS_MAX ? K : MG_KS_MAX, s = ks / 256;
    return K % 256 == 0 && K % ks == 0 && (s == 1 || s == 2 || s == 4 || s == 6 || s == 8);
  };
  return batch >= 1 && batch <= 8 && head_dim == MG_HD && hq == 4 * hkv && ok_k(D) && ok_k(I) && ok_k(hq * head_dim) &&
         D % 16 == 0 && (2 * I) % 64 == 0 && ((hq + 2 * hkv) * head_dim) % 16 == 0 && vocab_padded % 16 == 0;
}

// layers: n_layers (<= 32) x {wqkv, wo, wgu, wdown, ln1, ln2, kcache, vcache} device pointers in a HOST array
// (gar_mega_layer; copied into the kernel parameters); gbar: 8 zero-initialised bytes, 8-byte aligned; sumsq_a/b: (D/16 + 1) * 8 floats; the remaining buffers as in the per-kernel decode step.
extern "C" int gar_decode_mega_step(const gar_mega_layer* layers, int n_layers, const void* table, const void* lm_head,
                                    const void* final_norm, const void* cos_t, const void* sin_t, const int* next_tok,
                                    const int* pos, void* h, void* qkv, void* att, void* act, void* logits,
                                    float* sumsq_a, float* sumsq_b, float* o_part, float* ml_part, int* tickets,
                                    unsigned* gbar, int batch, int D, int I, int hq, int hkv, int head_dim,
                                    int vocab_padded, int max_seq, int splits, int unit_warps, float eps, float scale,
                                    void* stream_) {
  if (!gar_decode_mega_supported(batch, D, I, hq, hkv, head_dim, vocab_padded) || splits < 1 || splits > 32 ||
      (unit_warps != 2 && unit_warps != 4) || n_layers < 1 || n_layers > MG_MAX_LAYERS || !layers) {
    gar_set_error("gar_decode_mega_step: shape not supported (batch=%d D=%d I=%d hq=%d hkv=%d hd=%d)", batch, D, I, hq,
                  hkv, head_dim);
    return GAR_ERR_SHAPE;
  }
  MegaArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < n_layers; ++i)
    a.layers[i] = {static_cast<const bf16*>(layers[i].wqkv), static_cast<const bf16*>(layers[i].wo),
                   static_cast<const bf16*>(layers[i].wgu), static_cast<const bf16*>(layers[i].wdown),
                   static_cast<const bf16*>(layers[i].ln1), static_cast<const bf16*>(layers[i].ln2),
                   static_cast<bf16*>(layers[i].kcache), static_cast<bf16*>(layers[i].vcache)};
  a.n_layers = n_layers;
  a.table = static_cast<const bf16*>(table);
  a.lm_head = static_cast<const bf16*>(lm_head);
  a.final_norm = static_cast<const bf16*>(final_norm);
  a.cos_t = static_cast<const bf16*>(cos_t);
  a.sin_t = static_cast<const bf16*>(sin_t);
  a.next_tok = next_tok;
  a.pos = pos;
  a.h = static_cast<bf16*>(h);
  a.qkv = static_cast<bf16*>(qkv);
  a.att = static_cast<bf16*>(att);
  a.act = static_cast<bf16*>(act);
  a.logits = static_cast<bf16*>(logits);
  a.sumsq_a = sumsq_a;
  a.sumsq_b = sumsq_b;
  a.o_part = o_part;
  a.ml_part = ml_part;
  a.tickets = tickets;
  a.gbar = reinterpret_cast<unsigned long long*>(gbar);
  a.batch = batch; a.D = D; a.I = I; a.hq = hq; a.hkv = hkv;
  a.qkv_w = (hq + 2 * hkv) * head_dim;
  a.vocab_padded = vocab_padded;
  a.max_seq = max_seq;
  a.splits = splits;
  a.unit_warps = unit_warps;
  a.eps = eps;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.timeline = g_mega_timeline;
  static bool attr_set = false;
  if (!attr_set) {
    GAR_CUDA_CHECK(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
    attr_set = true;
  }
  // cooperative launch: the grid barrier needs every CTA resident (one per SM by shared-memory footprint)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gar_num_sms());
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = MG_SMEM;
  cfg.stream = static_cast<cudaStream_t>(stream_);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, decode_mega_kernel, a));
  return GAR_OK;
}
