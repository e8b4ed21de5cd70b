// Fused softmax(Q K^T * scale) V for head_dim 64 / 128 on tcgen05 tensor cores with TMEM accumulators.
//
// Replaces F.scaled_dot_product_attention in timm's AttentionRope (PE ViT, non-causal, N=1025) and
// flash-attn 2 `flash_attn_func` behind HF Llama (causal, GQA) on the prefill -- SURVEY.md §2b K5/K9.
// Same numerics contract as those kernels: S and the online softmax in fp32, P rounded to bf16 for
// the PV product, fp32 O accumulation, one bf16 rounding of O / l at the end.
//
// A tile = one (sequence, head, 128-query block); two CTAs co-reside per SM (head_dim 64) so one CTA's softmax
// overlaps the other's MMAs.  The default kernel is persistent (attn_fwd_persist_kernel below: resident CTAs walk tiles
// handed out by a device-side counter); attn_fwd_kernel is the same tile code as one CTA per tile.  Roles (256 threads
// = two warpgroups):
//   warp 0     TMA producer: Q, K and V double-buffered (3-D tensor maps so rows past a sequence end are zero-filled
//              on load and clipped on store)
//   warp 1     MMA issuer : S = Q K^T  (128x128x64, both operands K-major, 128B swizzle)
//                           O += P V   (128x64x128, P from TENSOR MEMORY (A operand), V MN-major from smem)
//   warps 2-3  idle (they complete the control warpgroup so setmaxnreg can hand its registers over)
//   warps 4-7  softmax    : tcgen05.ld S row-per-thread -> online softmax -> P (bf16) into tensor memory by tcgen05.st
//                           (swizzled smem only in the BKV 64 configuration), rescale O in TMEM, final O/l -> smem ->
//                           TMA store
// Registers: the kernel launches at MAXREG per thread; the control warpgroup drops to 48 (56 in the persistent kernel)
// and the softmax warpgroup (a whole 128-wide score row per thread lives in registers) takes the difference, so the
// row-per-thread softmax compiles without local-memory spills.
#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int ATT_THREADS = 256;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows of 128 B (one 64-wide, 128B-swizzled sub-tile)
constexpr uint32_t TM_S = 0;

// Tile configurations (BKV = keys per block):
//   HD 64,  BKV 64 : 64 KB smem, 128 TMEM columns, three CTAs per SM
//   HD 64,  BKV 128: 112 KB smem, 256 TMEM columns, two CTAs per SM
//   HD 128, BKV 128: 192 KB smem, one CTA per SM (GAR-8B Llama; PE-G's 96-wide heads are zero-padded to 128)
template <int HD, int BKV>
struct AttCfg {
  static constexpr int NSUB = HD / 64;
  static constexpr int KV_TILE = BKV * 64 * 2;         // one 64-wide sub-tile of a K/V block
  static constexpr int Q_BYTES = NSUB * TILE_BYTES;
  static constexpr int KV_BYTES = NSUB * KV_TILE;
  static constexpr int SMEM_Q = 0;
  static constexpr int SMEM_K = Q_BYTES;                      // 2 stages
  static constexpr int SMEM_V = SMEM_K + 2 * KV_BYTES;        // 2 stages: the load of V(j+1) no longer waits for PV(j)
  // P (the bf16 probabilities, A operand of the PV MMA) lives in TENSOR MEMORY next to S and O whenever the column
  // budget allows (BKV/2 packed columns): the softmax warps tcgen05.st it and the MMA reads it from there, so the
  // 32 KB smem write + 32 KB smem read per KV block (more than half of the kernel's shared-memory traffic) and the
  // proxy fence disappear.  HD 64 / BKV 64 keeps P in smem (three CTAs x 256 columns would not fit 512).
  static constexpr bool P_TMEM = !(HD == 64 && BKV == 64);
  static constexpr int SMEM_P = SMEM_V + 2 * KV_BYTES;        // BKV/64 sub-tiles [128 x 64] (smem-P configurations)
  static constexpr int SMEM_TOTAL = P_TMEM ? SMEM_P : SMEM_P + (BKV / 64) * TILE_BYTES;
  // no alignment slack: the dynamic segment is declared 1024-byte aligned (checked at run time), so two CTAs of the
  // HD 64 / BKV 128 configuration (112 KB each) still fit the SM's 228 KB
  static constexpr int SMEM_BYTES = SMEM_TOTAL;
  static constexpr int CTAS_PER_SM = HD == 64 ? (BKV == 64 ? 3 : 2) : 1;
  // 64K regs / (256 thr * CTAs), multiple of 8; control warpgroup keeps REG_CTRL, the softmax warpgroup gets the rest
  static constexpr int MAXREG = CTAS_PER_SM == 3 ? 80 : (CTAS_PER_SM == 2 ? 128 : 255);
  static constexpr bool REALLOC = CTAS_PER_SM > 1;
  static constexpr int REG_CTRL = 48;
  static constexpr int REG_SOFTMAX = 2 * MAXREG - REG_CTRL;  // 112 / 208
  static constexpr int TM_NEED = BKV + HD + (P_TMEM ? BKV / 2 : 0);  // S: BKV columns, O: HD columns, P: BKV/2
  static constexpr int TMEM_COLS = TM_NEED <= 128 ? 128 : (TM_NEED <= 256 ? 256 : 512);
  static constexpr uint32_t TM_O = BKV;
  static constexpr uint32_t TM_P = BKV + HD;
};

struct AttnParams {
  int q_len, kv_len;       // per sequence
  int causal;              // query i attends keys <= i + (kv_len - q_len)
  int q_col0, k_col0, v_col0, o_col0;   // column (element) offsets of head 0 inside the row
  int q_head_cols, kv_head_cols;        // column step per (kv) head
  int group;               // q heads per kv head
  int kv_batch_mul, kv_batch_head_mul;  // kv batch coordinate = b*kv_batch_mul + kvh*kv_batch_head_mul
  float scale_log2;
  long long* timeline;  // optional debug: per-phase clock64 stamps of CTA (0,0,0), softmax warp 2 lane 0 and MMA lane
};

// Row-per-thread softmax pieces, specialised on whether the block needs masking so the common
// (unmasked) path carries no compares and the score registers are never written back.
template <bool MASKED, int NCOLS>
__device__ __forceinline__ float row_max_n(const uint32_t (&s)[NCOLS], int kv0, int lim) {
  // 8 independent FMNMX3 chains (depth NCOLS/16) so the dependent-issue latency is covered by one warp alone
  float mx[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) mx[q] = -INFINITY;
#pragma unroll
  for (int i = 0; i < NCOLS; i += 16) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float a = __uint_as_float(s[i + 2 * q]), b = __uint_as_float(s[i + 2 * q + 1]);
      if (MASKED) {
        a = (kv0 + i + 2 * q <= lim) ? a : -INFINITY;
        b = (kv0 + i + 2 * q + 1 <= lim) ? b : -INFINITY;
      }
      mx[q] = fmax3(mx[q], a, b);
    }
  }
  return fmaxf(fmax3(mx[0], mx[1], mx[2]), fmax3(fmax3(mx[3], mx[4], mx[5]), mx[6], mx[7]));
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2: two elements per FMA-pipe issue slot) and exp2 on the FMA pipe ----
// The timeline of the row-per-thread softmax (tools/prof_kernels.py attn_timeline) puts ~60 % of a KV block into the
// exp phase, and that phase is MUFU.EX2-bound: 16 lanes/clk/SM against 16 384 exponentials per 128 x 128 block.  A share
// of the exponentials (POLY pairs out of every 8 column pairs) is therefore evaluated on the FMA pipe instead:
// 2^x = 2^n * p(f), n = round(x) by the magic-number add, f = x - n in [-0.5, 0.5], p a degree-3 minimax polynomial
// (max relative error 7.5e-5, far below the bf16 rounding P gets anyway), 2^n applied as an integer add on the exponent.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 exp2_poly2(f32x2 x2) {
  float x0, x1;
  upk2(x2, x0, x1);
  x2 = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));  // keep 2^n a normal number
  const f32x2 j = add2(x2, pk2(12582912.0f, 12582912.0f));         // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const f32x2 n = add2(j, pk2(-12582912.0f, -12582912.0f));
  const f32x2 f = fma2(n, pk2(-1.0f, -1.0f), x2);
  f32x2 pl = fma2(pk2(0.0551716685f, 0.0551716685f), f, pk2(0.2426111251f, 0.2426111251f));
  pl = fma2(pl, f, pk2(0.6932609677f, 0.6932609677f));
  pl = fma2(pl, f, pk2(0.9999280572f, 0.9999280572f));
  float j0, j1, p0, p1;
  upk2(j, j0, j1);
  upk2(pl, p0, p1);
  // 2^n applied as ONE integer multiply-add on the exponent field (IMAD runs on the FMA pipe; a shift + add pair would
  // cost two ALU-pipe slots, and the ALU pipe already carries the row max, the clamps and the bf16 packing)
  p0 = __int_as_float(__float_as_int(j0) * (1 << 23) + __float_as_int(p0));
  p1 = __int_as_float(__float_as_int(j1) * (1 << 23) + __float_as_int(p1));
  return pk2(p0, p1);
}

// which of every 8 column pairs take the polynomial: POLY = 0 none (all MUFU), 1 -> 2/8, 2 -> 3/8, 3 -> 4/8
__host__ __device__ constexpr uint32_t poly_mask(int poly) { return poly == 0 ? 0x00u : (poly == 1 ? 0x11u : (poly == 2 ? 0x49u : 0x55u)); }

// unmasked block: P = exp2(S * scale - m) -> bf16 -> swizzled smem; returns the row sum of the fp32 P values
// destination of 8 packed columns (4 words) of P: swizzled smem (TM = false: rbase / sw as documented at the call) or,
// TM = true, the caller's register array `out` -- the tensor-memory copy of P is written by the caller AFTER it has
// waited for PV(j-1), so that wait hides behind the exponentials instead of in front of them
template <bool TM, int NW>
__device__ __forceinline__ void p_emit(int c, const uint32_t (&w)[4], uint32_t (&out)[NW], uint32_t rbase, uint32_t sw) {
  if (TM) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out[c * 4 + i] = w[i];
  } else {
    const uint32_t addr = rbase + (c >> 3) * TILE_BYTES + (((c & 7) ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
}

template <int NCOLS, int POLY, bool TM>
__device__ __forceinline__ float exp_pack_store_fast(const uint32_t (&s)[NCOLS], float scale_log2, float neg_m,
                                                     uint32_t rbase, uint32_t sw, uint32_t (&out)[TM ? NCOLS / 2 : 1]) {
  const f32x2 sc2 = pk2(scale_log2, scale_log2), nm2 = pk2(neg_m, neg_m);
  f32x2 rs2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int c = 0; c < NCOLS / 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const f32x2 x2 = fma2(pk2(__uint_as_float(s[c * 8 + 2 * i]), __uint_as_float(s[c * 8 + 2 * i + 1])), sc2, nm2);
      f32x2 p2;
      if ((poly_mask(POLY) >> ((c * 4 + i) & 7)) & 1u) {
        p2 = exp2_poly2(x2);
      } else {
        float x0, x1;
        upk2(x2, x0, x1);
        p2 = pk2(ex2_approx(x0), ex2_approx(x1));
      }
      rs2[i] = add2(rs2[i], p2);
      float p0, p1;
      upk2(p2, p0, p1);
      w[i] = pack_bf16x2(p0, p1);
    }
    p_emit<TM>(c, w, out, rbase, sw);
  }
  float a0, a1, b0, b1, c0, c1, d0, d1;
  upk2(rs2[0], a0, a1);
  upk2(rs2[1], b0, b1);
  upk2(rs2[2], c0, c1);
  upk2(rs2[3], d0, d1);
  return ((a0 + a1) + (b0 + b1)) + ((c0 + c1) + (d0 + d1));
}

template <bool MASKED, int NCOLS, bool TM>
__device__ __forceinline__ float exp_pack_store_n(const uint32_t (&s)[NCOLS], int kv0, int lim, float scale_log2,
                                                  float neg_m, uint32_t rbase, uint32_t sw,
                                                  uint32_t (&out)[TM ? NCOLS / 2 : 1]) {
  float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < NCOLS / 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float p0 = ex2_approx(fmaf(__uint_as_float(s[c * 8 + 2 * i]), scale_log2, neg_m));
      float p1 = ex2_approx(fmaf(__uint_as_float(s[c * 8 + 2 * i + 1]), scale_log2, neg_m));
      if (MASKED) {
        p0 = (kv0 + c * 8 + 2 * i <= lim) ? p0 : 0.f;
        p1 = (kv0 + c * 8 + 2 * i + 1 <= lim) ? p1 : 0.f;
      }
      rs4[i] += p0 + p1;
      w[i] = pack_bf16x2(p0, p1);
    }
    p_emit<TM>(c, w, out, rbase, sw);
  }
  return (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
}

template <int HD, int BKV, int POLY>
__global__ void __launch_bounds__(ATT_THREADS) __maxnreg__((AttCfg<HD, BKV>::MAXREG))
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                     AttnParams p) {
  using Cfg = AttCfg<HD, BKV>;
  constexpr int NSUB = Cfg::NSUB, SMEM_Q = Cfg::SMEM_Q, SMEM_K = Cfg::SMEM_K, SMEM_V = Cfg::SMEM_V, SMEM_P = Cfg::SMEM_P;
  constexpr int Q_BYTES = Cfg::Q_BYTES, KV_BYTES = Cfg::KV_BYTES, KV_TILE = Cfg::KV_TILE;
  constexpr uint32_t TM_O = Cfg::TM_O;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[14];
  __shared__ uint32_t tmem_ptr_sh;
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();  // the 128B-swizzled tiles need the 1024-byte alignment the declaration asks for
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t B_QFULL = bar0, B_KFULL0 = bar0 + 8, B_KEMPTY0 = bar0 + 24, B_VFULL0 = bar0 + 40,
                 B_VEMPTY0 = bar0 + 56, B_SFULL = bar0 + 72, B_PFULL = bar0 + 80, B_PVDONE = bar0 + 88,
                 B_SCONS = bar0 + 96;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_qb = (p.q_len + BQ - 1) / BQ;
  const int qb = num_qb - 1 - blockIdx.x;  // heaviest (causal) blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / p.group;
  const int q0 = qb * BQ;
  const int shift = p.kv_len - p.q_len;
  int kv_end = p.kv_len;
  if (p.causal) kv_end = min(p.kv_len, q0 + BQ + shift);
  const int num_kvb = (kv_end + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    mbar_init(B_QFULL, 1);
    mbar_init(B_KFULL0, 1);
    mbar_init(B_KFULL0 + 8, 1);
    mbar_init(B_KEMPTY0, 1);
    mbar_init(B_KEMPTY0 + 8, 1);
    mbar_init(B_VFULL0, 1);
    mbar_init(B_VFULL0 + 8, 1);
    mbar_init(B_VEMPTY0, 1);
    mbar_init(B_VEMPTY0 + 8, 1);
    mbar_init(B_SFULL, 1);
    mbar_init(B_PFULL, 128);
    mbar_init(B_PVDONE, 1);
    mbar_init(B_SCONS, 128);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_ptr_sh), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_sh);

  if (warp < 4) {
  if constexpr (Cfg::REALLOC) setmaxnreg_dec<Cfg::REG_CTRL>();
  if (warp == 0) {
    if (lane == 0) {
      const int kb_coord = b * p.kv_batch_mul + kvh * p.kv_batch_head_mul;
      mbar_expect_tx(B_QFULL, Q_BYTES);
#pragma unroll
      for (int u = 0; u < NSUB; ++u)
        tma_load_3d(sbase + SMEM_Q + u * TILE_BYTES, &tmap_q, p.q_col0 + h * p.q_head_cols + 64 * u, q0, b, B_QFULL);
      for (int j = 0; j < num_kvb; ++j) {
        const int st = j & 1;
        mbar_wait(B_KEMPTY0 + 8 * st, ((j >> 1) & 1) ^ 1u);
        mbar_expect_tx(B_KFULL0 + 8 * st, KV_BYTES);
#pragma unroll
        for (int u = 0; u < NSUB; ++u)
          tma_load_3d(sbase + SMEM_K + st * KV_BYTES + u * KV_TILE, &tmap_k, p.k_col0 + kvh * p.kv_head_cols + 64 * u,
                      j * BKV, kb_coord, B_KFULL0 + 8 * st);
        mbar_wait(B_VEMPTY0 + 8 * st, ((j >> 1) & 1) ^ 1u);
        mbar_expect_tx(B_VFULL0 + 8 * st, KV_BYTES);
#pragma unroll
        for (int u = 0; u < NSUB; ++u)
          tma_load_3d(sbase + SMEM_V + st * KV_BYTES + u * KV_TILE, &tmap_v, p.v_col0 + kvh * p.kv_head_cols + 64 * u,
                      j * BKV, kb_coord, B_VFULL0 + 8 * st);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);   // P K-major, V MN-major
      mbar_wait(B_QFULL, 0);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(B_KFULL0 + 8 * st, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          // 64-wide sub-tile (k >> 2), 16-element step (k & 3) inside its 128B swizzle row
          const uint64_t qdesc = umma_desc_sw128(sbase + SMEM_Q + (k >> 2) * TILE_BYTES, 0, 1024) + 2u * (k & 3);
          const uint64_t kdesc = umma_desc_sw128(sbase + SMEM_K + st * KV_BYTES + (k >> 2) * KV_TILE, 0, 1024) + 2u * (k & 3);
          umma_bf16_ss(tmem + TM_S, qdesc, kdesc, idesc_s, k != 0);
        }
        umma_commit(B_KEMPTY0 + 8 * st);
        umma_commit(B_SFULL);  // also covers every earlier MMA of this thread
      };
      // Pipeline per KV block j:   S(j) -> [softmax loads S(j) into registers] -> S(j+1) issued at once
      // (overlaps the softmax math of block j) -> P(j) ready -> O += P(j) V(j).
      const bool tlm = p.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == 2;
#define TLM_STAMP(i) if (tlm) p.timeline[512 + j * 8 + (i)] = clock64()
      issue_s(0);
      for (int j = 0; j < num_kvb; ++j) {
        TLM_STAMP(0);
        if (j + 1 < num_kvb) {
          mbar_wait(B_SCONS, j & 1);  // every softmax thread holds S(j) in registers: TMEM S is free
          TLM_STAMP(1);
          issue_s(j + 1);
        }
        TLM_STAMP(2);
        mbar_wait(B_PFULL, j & 1);  // P(j) in smem, O rescaled
        TLM_STAMP(3);
        mbar_wait(B_VFULL0 + 8 * (j & 1), (j >> 1) & 1);
        TLM_STAMP(4);
        tc_fence_after();
        // V tile: NSUB sub-tiles [kv rows][64 hd] of 128B rows => MN-major B operand (N = HD): LBO = stride
        // between 64-wide N blocks (one sub-tile), SBO = 8 kv rows; 16 kv rows per MMA = 2048 B
        const uint64_t vdesc = umma_desc_sw128(sbase + SMEM_V + (j & 1) * KV_BYTES, KV_TILE, 1024);
        const int nk = (BKV > 32 && kv_end - j * BKV <= 32) ? 2 : BKV / 16;  // short tail: 32 keys of P/V only
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          if (k < nk) {
            if constexpr (Cfg::P_TMEM) {
              umma_bf16_ts(tmem + TM_O, tmem + Cfg::TM_P + 8u * k, vdesc + 128u * k, idesc_o, (j | k) != 0);
            } else {
              const uint64_t pdesc = umma_desc_sw128(sbase + SMEM_P + (k >> 2) * TILE_BYTES, 0, 1024) + 2u * (k & 3);
              umma_bf16_ss(tmem + TM_O, pdesc, vdesc + 128u * k, idesc_o, (j | k) != 0);
            }
          }
        }
        umma_commit(B_VEMPTY0 + 8 * (j & 1));
        umma_commit(B_PVDONE);
        TLM_STAMP(5);
      }
    }
  }
  } else {
    if constexpr (Cfg::REALLOC) setmaxnreg_inc<Cfg::REG_SOFTMAX>();
    // ---------------- softmax warps: thread <-> query row ----------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int q_pos = q0 + row + shift;  // last key this row may see (causal)
    float m_run = -INFINITY, l_run = 0.f;
    // debug: per-phase clock64 stamps of one mid-grid CTA's first softmax thread (p.timeline is NULL in production)
    const bool tl = p.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == 2 && threadIdx.x == 128;
#define TL_STAMP(i) if (tl) p.timeline[j * 8 + (i)] = clock64()
    // Tails cost what they hold, not a full tile: a softmax warp whose 32 query rows all lie past q_len only keeps the
    // barrier protocol in step (the ViT's 1025-token tiles leave a 1-row last query block), and a last KV block with
    // <= 32 live keys is handled on 32 score columns (and 2 of the BKV/16 PV MMAs) instead of BKV.
    const bool warp_live = q0 + quarter * 32 < p.q_len;
    const uint32_t p_rbase = Cfg::P_TMEM ? tmem + lane_addr + Cfg::TM_P : sbase + SMEM_P + (row >> 3) * 1024 + (row & 7) * 128;
    for (int j = 0; j < num_kvb; ++j) {
      TL_STAMP(0);
      mbar_wait(B_SFULL, j & 1);
      tc_fence_after();
      TL_STAMP(1);
      if (!warp_live) {
        tc_fence_before();
        mbar_arrive(B_SCONS);
        if (j > 0) mbar_wait(B_PVDONE, (j - 1) & 1);  // keeps this arrival in P's phase j (see the MMA loop)
        mbar_arrive(B_PFULL);
        continue;
      }
      const int kv0 = j * BKV;
      const int lim = p.causal ? min(p.kv_len - 1, q_pos) : p.kv_len - 1;
      auto block = [&](auto nc_tag, bool need_mask) {
        constexpr int NC = decltype(nc_tag)::value;
        uint32_t s[NC];
#pragma unroll
        for (int c = 0; c < NC / 32; ++c) {
          uint32_t(&sc)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]);
          tmem_ld_32x32b_x32(tmem + lane_addr + TM_S + c * 32, sc);
        }
        tmem_ld_wait();
        TL_STAMP(2);
        tc_fence_before();
        mbar_arrive(B_SCONS);  // S(j) is in registers: the MMA warp may overwrite TMEM S with S(j+1)
        const float mx = need_mask ? row_max_n<true, NC>(s, kv0, lim) : row_max_n<false, NC>(s, kv0, lim);
        // Lazy rescale: keep exponentiating against the old reference max until the true max has grown by
        // more than 2^8; P then stays <= 256 (harmless in bf16/fp32) and most O rescales disappear.
        const float m_cand = fmaxf(m_run, mx * p.scale_log2);
        const bool bump = (m_cand - m_run) > 8.0f || m_run == -INFINITY;
        const float m_new = bump ? m_cand : m_run;
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = bump ? ex2_approx(m_run - m_use) : 1.0f;
        TL_STAMP(3);
        // O and the P buffer are quiescent only once PV(j-1) has retired
        auto wait_pv_rescale = [&]() {
          if (j > 0) {
            mbar_wait(B_PVDONE, (j - 1) & 1);
            tc_fence_after();
            TL_STAMP(4);
            const bool any = __any_sync(0xffffffffu, alpha != 1.0f);
            if (any) {
#pragma unroll
              for (int c = 0; c < HD / 32; ++c) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(tmem + lane_addr + TM_O + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x32(tmem + lane_addr + TM_O + c * 32, o);
              }
              tmem_st_wait();
            }
          }
        };
        uint32_t pw[Cfg::P_TMEM ? NC / 2 : 1];
        if constexpr (!Cfg::P_TMEM) wait_pv_rescale();  // smem P is written while it is computed
        // P = exp2(S*scale - m) -> bf16.  Tensor-memory P: row r on lane r, two keys per 32-bit column (the A operand
        // layout of the PV MMA); the packed words stay in registers until PV(j-1) has retired, so that wait hides
        // behind the exponentials.  smem P (BKV 64): canonical K-major 128B-swizzled layout, row r, 16B chunk c at
        // r*128 + ((c ^ (r&7))<<4)
        const float rs = need_mask ? exp_pack_store_n<true, NC, Cfg::P_TMEM>(s, kv0, lim, p.scale_log2, -m_use, p_rbase, row & 7, pw)
                                   : exp_pack_store_fast<NC, POLY, Cfg::P_TMEM>(s, p.scale_log2, -m_use, p_rbase, row & 7, pw);
        if constexpr (Cfg::P_TMEM) {
          wait_pv_rescale();
#pragma unroll
          for (int c = 0; c < NC / 32; ++c)
            tmem_st_32x32b_x16(p_rbase + 16 * c, *reinterpret_cast<uint32_t(*)[16]>(&pw[16 * c]));
        }
        l_run = l_run * alpha + rs;
        m_run = m_new;
      };
      if (BKV > 32 && kv_end - kv0 <= 32) {
        block(std::integral_constant<int, 32>{}, true);
      } else {
        const bool need_mask = (kv0 + BKV > p.kv_len) || (p.causal && kv0 + BKV - 1 > q0 + shift);
        block(std::integral_constant<int, BKV>{}, need_mask);
      }
      TL_STAMP(5);
      if constexpr (Cfg::P_TMEM) tmem_st_wait();   // P (and a rescaled O) are in tensor memory
      else fence_proxy_async_smem();               // P in smem must be visible to the async proxy
      tc_fence_before();
      mbar_arrive(B_PFULL);
      TL_STAMP(6);
    }
    // ---------------- finalize: O / l -> bf16 -> smem (reuse Q tile) -> TMA store ----------------
    mbar_wait(B_PVDONE, (num_kvb - 1) & 1);
    tc_fence_after();
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    const uint32_t rbase = sbase + SMEM_Q + (row >> 3) * 1024 + (row & 7) * 128;
    const uint32_t sw = row & 7;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      if (!warp_live) break;  // rows past q_len are clipped by the TMA store
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem + lane_addr + TM_O + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = pack_bf16x2(__uint_as_float(o[q * 8 + 2 * i]) * inv_l, __uint_as_float(o[q * 8 + 2 * i + 1]) * inv_l);
        const uint32_t addr = rbase + (c >> 1) * TILE_BYTES + ((static_cast<uint32_t>((c & 1) * 4 + q) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (warp == 4 && lane == 0) {
#pragma unroll
      for (int u = 0; u < NSUB; ++u)
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&tmap_o)),
                     "r"(sbase + SMEM_Q + u * TILE_BYTES), "r"(p.o_col0 + h * p.q_head_cols + 64 * u), "r"(q0), "r"(b)
                     : "memory");
      tma_store_commit();
      tma_store_wait_read<0>();  // the staging tile has been read: the CTA may retire while the write drains
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Persistent variant (default for the tensor-memory-P configurations).  The per-tile CTA above pays, for every
// (sequence, head, query block): barrier init + TMEM allocation + a CTA-wide sync, a cold Q/K load before the first
// S MMA (~2 300 cycles in which the softmax warpgroup only waits), and a finalize during which its MMA and TMA warps
// idle -- together about a quarter of a ViT tile's 32 k cycles (profiles/r02_attn_timeline_ptmem.txt: 24 k of them
// in the KV loop).  Here a CTA stays resident and walks tiles handed out by a device-side counter in the SAME order
// the hardware would have launched them (query blocks of one (head, sequence) heaviest-first, heads next):
//   * the TMA thread fetches the next tile index one tile ahead (atomicAdd), and as soon as the K/V ring has room
//     loads the next tile's Q (second Q buffer) and first K/V blocks while the current tile is still in its KV loop;
//   * the MMA thread issues S(0) of the next tile the moment the last S of the current one sits in the softmax
//     warps' registers, so it runs under the current tile's last exponentials and its finalize;
//   * the softmax warpgroup goes from one tile's finalize straight into the next tile's first block.
// The K/V ring, the S / P / PV hand-offs and their parities simply keep counting KV blocks across tiles (`g`); the two
// Q buffers double as the O staging tiles and are handed back to the TMA thread (QEMPTY) once the TMA store has
// read them.  The tile index travels through shared memory next to the QFULL arrive (release) / wait (acquire).
// The last CTA to retire resets the counter pair, so a launch needs no memset.
template <int HD, int BKV>
struct AttCfgP : AttCfg<HD, BKV> {
  using Base = AttCfg<HD, BKV>;
  static_assert(Base::P_TMEM, "persistent attention keeps P in tensor memory");
  static constexpr int SMEM_Q = 0;                                   // 2 buffers
  static constexpr int SMEM_K = 2 * Base::Q_BYTES;                   // 2 stages
  static constexpr int SMEM_V = SMEM_K + 2 * Base::KV_BYTES;         // 2 stages
  static constexpr int SMEM_BYTES = SMEM_V + 2 * Base::KV_BYTES;     // HD 64: 96 KB (2 CTAs/SM), HD 128: 192 KB
  // the control threads carry the tile walk (tile index, coordinates, running block counter) on top of the per-tile
  // kernel's state: 8 more registers for them keep the MMA issue loop free of local-memory traffic
  static constexpr int REG_CTRL = 56;
  static constexpr int REG_SOFTMAX = 2 * Base::MAXREG - REG_CTRL;   // 200
};

template <int HD, int BKV, int POLY>
__global__ void __launch_bounds__(ATT_THREADS) __maxnreg__((AttCfg<HD, BKV>::MAXREG))
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                        AttnParams p, int heads, int total_tiles, unsigned int* sched) {
  using Cfg = AttCfgP<HD, BKV>;
  constexpr int NSUB = Cfg::NSUB, SMEM_Q = Cfg::SMEM_Q, SMEM_K = Cfg::SMEM_K, SMEM_V = Cfg::SMEM_V;
  constexpr int Q_BYTES = Cfg::Q_BYTES, KV_BYTES = Cfg::KV_BYTES, KV_TILE = Cfg::KV_TILE;
  constexpr uint32_t TM_O = Cfg::TM_O;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_ptr_sh;
  __shared__ int tile_sh[2];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t B_QFULL0 = bar0, B_QEMPTY0 = bar0 + 16, B_KFULL0 = bar0 + 32, B_KEMPTY0 = bar0 + 48,
                 B_VFULL0 = bar0 + 64, B_VEMPTY0 = bar0 + 80, B_SFULL = bar0 + 96, B_PFULL = bar0 + 104,
                 B_PVDONE = bar0 + 112, B_SCONS = bar0 + 120;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_qb = (p.q_len + BQ - 1) / BQ;
  const int shift = p.kv_len - p.q_len;
  // tile t -> (query block, head, sequence): query blocks fastest and heaviest (causal) first, as the per-tile grid
  auto tile_coords = [&](int t, int& q0, int& h, int& b, int& kv_end) {
    const int r = t / num_qb;
    q0 = (num_qb - 1 - (t - r * num_qb)) * BQ;
    b = r / heads;
    h = r - b * heads;
    kv_end = p.causal ? min(p.kv_len, q0 + BQ + shift) : p.kv_len;
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(B_QFULL0 + 8 * i, 1);
      mbar_init(B_QEMPTY0 + 8 * i, 1);
      mbar_init(B_KFULL0 + 8 * i, 1);
      mbar_init(B_KEMPTY0 + 8 * i, 1);
      mbar_init(B_VFULL0 + 8 * i, 1);
      mbar_init(B_VEMPTY0 + 8 * i, 1);
    }
    mbar_init(B_SFULL, 1);
    mbar_init(B_PFULL, 128);
    mbar_init(B_PVDONE, 1);
    mbar_init(B_SCONS, 128);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_ptr_sh), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_sh);
  volatile int* tile_v = tile_sh;

  if (warp < 4) {
  if constexpr (Cfg::REALLOC) setmaxnreg_dec<Cfg::REG_CTRL>();
  if (warp == 0) {
    if (lane == 0) {
      int t = blockIdx.x;  // the first tile of every CTA is static; the counter hands out the rest
      uint32_t g = 0;      // KV blocks loaded so far: ring position and parity keep counting across tiles
      for (int n = 0;; ++n) {
        const int buf = n & 1;
        // Q buffer `buf` last held tile n-2: its O tile must have left through the TMA store
        if (n >= 2) mbar_wait(B_QEMPTY0 + 8 * buf, ((n >> 1) - 1) & 1);
        tile_v[buf] = t;
        if (t >= total_tiles) {
          mbar_arrive(B_QFULL0 + 8 * buf);  // sentinel: the other roles leave their loops
          break;
        }
        const int t_next = static_cast<int>(gridDim.x + atomicAdd(sched, 1u));  // needed only after this tile's loads
        int q0, h, b, kv_end;
        tile_coords(t, q0, h, b, kv_end);
        const int num_kvb = (kv_end + BKV - 1) / BKV;
        const int kvh = h / p.group;
        const int kb_coord = b * p.kv_batch_mul + kvh * p.kv_batch_head_mul;
        mbar_expect_tx(B_QFULL0 + 8 * buf, Q_BYTES);
#pragma unroll
        for (int u = 0; u < NSUB; ++u)
          tma_load_3d(sbase + SMEM_Q + buf * Q_BYTES + u * TILE_BYTES, &tmap_q, p.q_col0 + h * p.q_head_cols + 64 * u, q0,
                      b, B_QFULL0 + 8 * buf);
        for (int j = 0; j < num_kvb; ++j, ++g) {
          const uint32_t st = g & 1, par = ((g >> 1) & 1) ^ 1u;
          mbar_wait(B_KEMPTY0 + 8 * st, par);
          mbar_expect_tx(B_KFULL0 + 8 * st, KV_BYTES);
#pragma unroll
          for (int u = 0; u < NSUB; ++u)
            tma_load_3d(sbase + SMEM_K + st * KV_BYTES + u * KV_TILE, &tmap_k, p.k_col0 + kvh * p.kv_head_cols + 64 * u,
                        j * BKV, kb_coord, B_KFULL0 + 8 * st);
          mbar_wait(B_VEMPTY0 + 8 * st, par);
          mbar_expect_tx(B_VFULL0 + 8 * st, KV_BYTES);
#pragma unroll
          for (int u = 0; u < NSUB; ++u)
            tma_load_3d(sbase + SMEM_V + st * KV_BYTES + u * KV_TILE, &tmap_v, p.v_col0 + kvh * p.kv_head_cols + 64 * u,
                        j * BKV, kb_coord, B_VFULL0 + 8 * st);
        }
        t = t_next;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, HD, 0, 1);
      uint32_t g = 0;  // KV blocks issued so far (all tiles)
      for (int n = 0;; ++n) {
        const int buf = n & 1;
        mbar_wait(B_QFULL0 + 8 * buf, (n >> 1) & 1);
        const int t = tile_v[buf];
        if (t >= total_tiles) break;
        int q0, h, b, kv_end;
        tile_coords(t, q0, h, b, kv_end);
        const int num_kvb = (kv_end + BKV - 1) / BKV;
        const uint32_t qbase = sbase + SMEM_Q + buf * Q_BYTES;
        auto issue_s = [&](uint32_t gj) {
          const uint32_t st = gj & 1;
          mbar_wait(B_KFULL0 + 8 * st, (gj >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint64_t qdesc = umma_desc_sw128(qbase + (k >> 2) * TILE_BYTES, 0, 1024) + 2u * (k & 3);
            const uint64_t kdesc = umma_desc_sw128(sbase + SMEM_K + st * KV_BYTES + (k >> 2) * KV_TILE, 0, 1024) + 2u * (k & 3);
            umma_bf16_ss(tmem + TM_S, qdesc, kdesc, idesc_s, k != 0);
          }
          umma_commit(B_KEMPTY0 + 8 * st);
          umma_commit(B_SFULL);
        };
        // the previous tile's last S must sit in the softmax warps' registers before TMEM S is overwritten
        if (g > 0) mbar_wait(B_SCONS, (g - 1) & 1);
        issue_s(g);
        for (int j = 0; j < num_kvb; ++j) {
          const uint32_t gj = g + j;
          if (j + 1 < num_kvb) {
            mbar_wait(B_SCONS, gj & 1);
            issue_s(gj + 1);
          }
          mbar_wait(B_PFULL, gj & 1);  // P(j) in tensor memory, O rescaled (j == 0: the previous tile's O has been read out)
          mbar_wait(B_VFULL0 + 8 * (gj & 1), (gj >> 1) & 1);
          tc_fence_after();
          const uint64_t vdesc = umma_desc_sw128(sbase + SMEM_V + (gj & 1) * KV_BYTES, KV_TILE, 1024);
          const int nk = (BKV > 32 && kv_end - j * BKV <= 32) ? 2 : BKV / 16;
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            if (k < nk) umma_bf16_ts(tmem + TM_O, tmem + Cfg::TM_P + 8u * k, vdesc + 128u * k, idesc_o, (j | k) != 0);
          }
          umma_commit(B_VEMPTY0 + 8 * (gj & 1));
          umma_commit(B_PVDONE);
        }
        g += num_kvb;
      }
    }
  }
  } else {
    if constexpr (Cfg::REALLOC) setmaxnreg_inc<Cfg::REG_SOFTMAX>();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t p_rbase = tmem + lane_addr + Cfg::TM_P;
    uint32_t g = 0;
    for (int n = 0;; ++n) {
      const int buf = n & 1;
      mbar_wait(B_QFULL0 + 8 * buf, (n >> 1) & 1);
      const int t = tile_v[buf];
      if (t >= total_tiles) break;
      int q0, h, b, kv_end;
      tile_coords(t, q0, h, b, kv_end);
      const int num_kvb = (kv_end + BKV - 1) / BKV;
      const int q_pos = q0 + row + shift;
      const bool warp_live = q0 + quarter * 32 < p.q_len;
      const int lim = p.causal ? min(p.kv_len - 1, q_pos) : p.kv_len - 1;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < num_kvb; ++j) {
        const uint32_t gj = g + j;
        mbar_wait(B_SFULL, gj & 1);
        tc_fence_after();
        if (!warp_live) {
          tc_fence_before();
          mbar_arrive(B_SCONS);
          if (j > 0) mbar_wait(B_PVDONE, (gj - 1) & 1);
          mbar_arrive(B_PFULL);
          continue;
        }
        const int kv0 = j * BKV;
        auto block = [&](auto nc_tag, bool need_mask) {
          constexpr int NC = decltype(nc_tag)::value;
          uint32_t s[NC];
#pragma unroll
          for (int c = 0; c < NC / 32; ++c) {
            uint32_t(&sc)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]);
            tmem_ld_32x32b_x32(tmem + lane_addr + TM_S + c * 32, sc);
          }
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(B_SCONS);
          const float mx = need_mask ? row_max_n<true, NC>(s, kv0, lim) : row_max_n<false, NC>(s, kv0, lim);
          const float m_cand = fmaxf(m_run, mx * p.scale_log2);
          const bool bump = (m_cand - m_run) > 8.0f || m_run == -INFINITY;
          const float m_new = bump ? m_cand : m_run;
          const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
          const float alpha = bump ? ex2_approx(m_run - m_use) : 1.0f;
          uint32_t pw[NC / 2];
          const float rs = need_mask ? exp_pack_store_n<true, NC, true>(s, kv0, lim, p.scale_log2, -m_use, p_rbase, 0, pw)
                                     : exp_pack_store_fast<NC, POLY, true>(s, p.scale_log2, -m_use, p_rbase, 0, pw);
          if (j > 0) {  // O and the P columns are quiescent only once PV(j-1) has retired (j == 0: finalize waited)
            mbar_wait(B_PVDONE, (gj - 1) & 1);
            tc_fence_after();
            const bool any = __any_sync(0xffffffffu, alpha != 1.0f);
            if (any) {
#pragma unroll
              for (int c = 0; c < HD / 32; ++c) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(tmem + lane_addr + TM_O + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x32(tmem + lane_addr + TM_O + c * 32, o);
              }
              tmem_st_wait();
            }
          }
#pragma unroll
          for (int c = 0; c < NC / 32; ++c)
            tmem_st_32x32b_x16(p_rbase + 16 * c, *reinterpret_cast<uint32_t(*)[16]>(&pw[16 * c]));
          l_run = l_run * alpha + rs;
          m_run = m_new;
        };
        if (BKV > 32 && kv_end - kv0 <= 32) {
          block(std::integral_constant<int, 32>{}, true);
        } else {
          const bool need_mask = (kv0 + BKV > p.kv_len) || (p.causal && kv0 + BKV - 1 > q0 + shift);
          block(std::integral_constant<int, BKV>{}, need_mask);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(B_PFULL);
      }
      g += num_kvb;
      // ---- finalize: O / l -> bf16 -> this tile's Q buffer -> TMA store; S(0) of the next tile is already running ----
      mbar_wait(B_PVDONE, (g - 1) & 1);
      tc_fence_after();
      const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
      const uint32_t rbase = sbase + SMEM_Q + buf * Q_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
      const uint32_t sw = row & 7;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        if (!warp_live) break;
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem + lane_addr + TM_O + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack_bf16x2(__uint_as_float(o[q * 8 + 2 * i]) * inv_l, __uint_as_float(o[q * 8 + 2 * i + 1]) * inv_l);
          const uint32_t addr = rbase + (c >> 1) * TILE_BYTES + ((static_cast<uint32_t>((c & 1) * 4 + q) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      }
      tc_fence_before();  // this thread's O reads are ordered before its next PFULL arrive lets PV(0) overwrite O
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (warp == 4 && lane == 0) {
#pragma unroll
        for (int u = 0; u < NSUB; ++u)
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmap_o)),
                       "r"(sbase + SMEM_Q + buf * Q_BYTES + u * TILE_BYTES), "r"(p.o_col0 + h * p.q_head_cols + 64 * u),
                       "r"(q0), "r"(b)
                       : "memory");
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(B_QEMPTY0 + 8 * buf);  // the TMA thread may load tile n+2's Q into this buffer
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, Cfg::TMEM_COLS);
  }
  if (threadIdx.x == 0) {  // the TMA thread: its counter fetches precede this arrival
    __threadfence();
    if (atomicAdd(sched + 1, 1u) == gridDim.x - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace

static long long* g_attn_timeline = nullptr;
// debug hook: device buffer of >= 8*num_kv_blocks int64 receiving clock64 stamps (NULL disables)
extern "C" int gar_attn_debug_timeline(void* dev_buf) {
  g_attn_timeline = static_cast<long long*>(dev_buf);
  return GAR_OK;
}

// Tile counters of the persistent kernel: {next tile, retired CTAs} pairs, zero at rest (the last CTA of a launch resets
// its pair).  Launches rotate through 64 pairs so that two launches in flight on different streams never share one.
static unsigned int* attn_sched_slot() {
  static std::mutex mu;
  static unsigned int* base = nullptr;
  static unsigned int next = 0;
  std::lock_guard<std::mutex> lk(mu);
  if (base == nullptr) {
    if (cudaMalloc(&base, 64 * 2 * sizeof(unsigned int)) != cudaSuccess) {
      gar_set_error("attention: cudaMalloc of the tile counters failed: %s", cudaGetErrorString(cudaGetLastError()));
      base = nullptr;
      return nullptr;
    }
    if (cudaMemset(base, 0, 64 * 2 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
  }
  return base + 2 * (next++ & 63u);
}

// Generic strided description (all in bf16 elements):
//   Q/O: [batch][q_len rows][row width], K/V: [kv_batches][kv rows][row width]
template <int HD, int BKV, int POLY>
static int launch_attn(const void* q, long long q_row_stride, long long q_batch_stride, const void* k, const void* v,
                       long long kv_row_stride, long long kv_batch_stride, int kv_batches, void* o,
                       long long o_row_stride, long long o_batch_stride, int batch, int heads, int q_len,
                       int kv_rows_alloc, const AttnParams& p, cudaStream_t stream) {
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = gar_make_tmap_3d(&tq, q, q_row_stride, q_len, batch, q_row_stride, q_batch_stride, 64, BQ, 1, GAR_SW_128))) return rc;
  if ((rc = gar_make_tmap_3d(&tk, k, kv_row_stride, kv_rows_alloc, kv_batches, kv_row_stride, kv_batch_stride, 64, BKV, 1,
                             GAR_SW_128)))
    return rc;
  if ((rc = gar_make_tmap_3d(&tv, v, kv_row_stride, kv_rows_alloc, kv_batches, kv_row_stride, kv_batch_stride, 64, BKV, 1,
                             GAR_SW_128)))
    return rc;
  if ((rc = gar_make_tmap_3d(&to, o, o_row_stride, q_len, batch, o_row_stride, o_batch_stride, 64, BQ, 1, GAR_SW_128))) return rc;
  using Cfg = AttCfg<HD, BKV>;
  static bool attr_set = false;
  if (!attr_set) {
    GAR_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<HD, BKV, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set = true;
  }
  if constexpr (Cfg::P_TMEM) {
    if (gar_global_option("attn_persist") && p.timeline == nullptr) {
      using CfgP = AttCfgP<HD, BKV>;
      static bool attr_set_p = false;
      if (!attr_set_p) {
        GAR_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_persist_kernel<HD, BKV, POLY>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, CfgP::SMEM_BYTES));
        attr_set_p = true;
      }
      unsigned int* sched = attn_sched_slot();
      if (sched == nullptr) return GAR_ERR_CUDA;
      const int total = ((q_len + BQ - 1) / BQ) * heads * batch;
      const int resident = Cfg::CTAS_PER_SM * gar_grid_sms();
      attn_fwd_persist_kernel<HD, BKV, POLY><<<dim3(total < resident ? total : resident), ATT_THREADS, CfgP::SMEM_BYTES,
                                               stream>>>(tq, tk, tv, to, p, heads, total, sched);
      GAR_CUDA_CHECK(cudaGetLastError());
      return GAR_OK;
    }
  }
  dim3 grid((q_len + BQ - 1) / BQ, heads, batch);
  attn_fwd_kernel<HD, BKV, POLY><<<grid, ATT_THREADS, Cfg::SMEM_BYTES, stream>>>(tq, tk, tv, to, p);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

extern "C" int gar_attn_fwd(const void* q, long long q_row_stride, long long q_batch_stride, int q_col0,
                            const void* k, const void* v, long long kv_row_stride, long long kv_batch_stride,
                            int kv_batches, int k_col0, int v_col0, int kv_head_cols, int kv_batch_mul,
                            int kv_batch_head_mul, void* o, long long o_row_stride, long long o_batch_stride,
                            int o_col0, int batch, int heads, int kv_heads, int head_dim, int q_len, int kv_len,
                            int kv_rows_alloc, int causal, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (head_dim != 64 && head_dim != 128) {
    gar_set_error("gar_attn_fwd: head_dim %d not built (64 and 128; pad 96 to 128)", head_dim);
    return GAR_ERR_SHAPE;
  }
  if (heads % kv_heads || q_len <= 0 || kv_len <= 0 || q_len > kv_len) return GAR_ERR_ARG;
  AttnParams p;
  p.q_len = q_len;
  p.kv_len = kv_len;
  p.causal = causal;
  p.q_col0 = q_col0;
  p.k_col0 = k_col0;
  p.v_col0 = v_col0;
  p.o_col0 = o_col0;
  p.q_head_cols = head_dim;
  p.kv_head_cols = kv_head_cols;
  p.group = heads / kv_heads;
  p.kv_batch_mul = kv_batch_mul;
  p.kv_batch_head_mul = kv_batch_head_mul;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.timeline = g_attn_timeline;
#define GAR_ATTN_ARGS q, q_row_stride, q_batch_stride, k, v, kv_row_stride, kv_batch_stride, kv_batches, o, o_row_stride, \
                      o_batch_stride, batch, heads, q_len, kv_rows_alloc, p, stream
  const int poly = gar_global_option("attn_poly");  // exponentials on the FMA pipe: 0 none, 1 2/8 (default), 2 3/8, 3 4/8
  // head_dim 128 has half the exponentials per MMA flop: the polynomial share only from attn_poly >= 2
  if (head_dim == 128) return poly < 2 ? launch_attn<128, 128, 0>(GAR_ATTN_ARGS) : launch_attn<128, 128, 2>(GAR_ATTN_ARGS);
  if (gar_global_option("attn_bkv") == 128) {
    switch (poly) {
      case 0: return launch_attn<64, 128, 0>(GAR_ATTN_ARGS);
      case 1: return launch_attn<64, 128, 1>(GAR_ATTN_ARGS);
      case 3: return launch_attn<64, 128, 3>(GAR_ATTN_ARGS);
      default: return launch_attn<64, 128, 2>(GAR_ATTN_ARGS);
    }
  }
  return poly == 0 ? launch_attn<64, 64, 0>(GAR_ATTN_ARGS) : launch_attn<64, 64, 2>(GAR_ATTN_ARGS);
#undef GAR_ATTN_ARGS
  return GAR_OK;
}
