// C[M,N] = epilogue(A[M,K] * B[N,K]^T)   bf16 in, fp32 accumulate in TMEM, bf16 out.
//
// Replaces the cuBLAS(Lt) GEMMs + ATen bias/GELU/mul/add launches behind the reference's
// nn.Linear / nn.Conv2d(k=s) calls (timm EvaBlock qkv/proj/fc1/fc2, PerceptionLM projector
// modeling_perception_lm.py:85-92, HF Llama q/k/v/o/gate/up/down -- SURVEY.md §2b K1,K2,K5,K6,K9).
//
// B200-native structure (no CUTLASS): persistent CTAs, one 128 x BLOCK_N output tile at a time.
//   warp 0      TMA producer   : cp.async.bulk.tensor 128B-swizzled A/B k-slabs into a smem ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16)
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       : tcgen05.ld accumulator -> registers -> fused epilogue with the
//                                reference's bf16 rounding points -> swizzled smem -> TMA store
// Two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of tile i+1.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int C_CHUNK = 32;                               // columns per TMA store box
constexpr int C_STAGE_BYTES = BLOCK_M * C_CHUNK * 2;      // 8 KB, 64B-swizzled rows

enum Epi { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_GELU = 2, EPI_BIAS_LS_RES = 3, EPI_RES = 4, EPI_BIAS_ROPE2D = 5, EPI_SWIGLU = 6 };

// CTA2: cta_group::2 -- a cluster of two CTAs computes a 256 x 256 tile; each CTA stages its 128 rows of A and
// 128 of the 256 rows of B, so every SM's shared-memory traffic per MMA halves for B (the 1-CTA 128x256 tile
// needs ~188 B/clk of smem traffic against a 128 B/clk port) and the ring gets two more stages.
template <int BLOCK_N, bool CTA2 = false>
struct Cfg {
  static constexpr int STAGES = CTA2 ? 6 : (BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8));
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages (power of two >= 32)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * C_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// Raster of the persistent tile walk: tile index -> (row block, column panel).  Column panels in groups of n_group
// (the last group may be narrower), n fastest inside a group, then m, then the group.  Shared by the kernel and the
// host-side probe (gar_gemm_raster_probe) the CPU tests use.
__host__ __device__ __forceinline__ void raster_tile_mn(int tile, int num_m, int num_n, int n_group, int& m_blk, int& n_blk) {
  const int per_group = num_m * n_group;
  const int grp = tile / per_group, r = tile - grp * per_group;
  const int rest = num_n - grp * n_group;
  const int ng = n_group < rest ? n_group : rest;
  m_blk = r / ng;
  n_blk = grp * n_group + (r - m_blk * ng);
}

// Column panels per raster group.  With `fly` tiles in flight (one per CTA / CTA pair) and ng panels per group, the
// operands being read at any moment are fly/ng row blocks of A and ng panels of B; the walk stays L2-friendly while
// that working set fits a budget well inside the 126 MB L2 (default 40 MB: the output stream passes through the same
// cache and each L2 half keeps its own copy of shared lines).  The widest group that fits is taken (A is streamed
// once per group); when not even the narrowest working set fits (long K: ViT fc2, down_proj) the plain n-fastest
// walk is kept.  Measured with ncu on GAR-1B's SwiGLU GEMM (67 MB of weights): 2.14 -> 0.68 GB of DRAM reads per
// launch, 1.73 -> 1.60 ms; GAR-8B's (235 MB): 7.1 -> 1.5 GB, 1.73 -> 1.42 ms (profiles/r02_gemm_group_ncu.txt).
static int raster_group_panels(int num_n, int fly, int tile_m, int block_n, int K) {
  const double a_block = static_cast<double>(tile_m) * K * 2, b_panel = static_cast<double>(block_n) * K * 2;
  const double budget = static_cast<double>(gar_global_option("gemm_group_kb")) * 1024.0;
  const int forced = gar_global_option("gemm_group_force");  // test hook: panels per group
  if (forced > 0) return std::min(forced, num_n);
  if (budget > 0) {
    for (int groups = 1; groups <= num_n; ++groups) {
      const int ng = (num_n + groups - 1) / groups;
      if ((static_cast<double>(fly) / ng) * a_block + ng * b_panel <= budget) return ng;
    }
  }
  return num_n;
}

struct EpiArgs {
  const bf16* bias;
  const bf16* gamma;
  const bf16* res;
  int ldres;
  // EPI_BIAS_ROPE2D: PE-ViT interleaved-pair RoPE on columns < rope_cols of rows whose token index
  // (row % ntok) is >= npt; table in chunk-major layout [sin | cos][hd / 8][ntok - npt][8]
  const bf16* rope;
  int ntok, npt, hd, rope_cols;
};

template <int BLOCK_N, int EPI, bool CTA2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, int M, int N, int K, int n_group, EpiArgs ea) {
  using C = Cfg<BLOCK_N, CTA2>;
  static_assert(!CTA2 || BLOCK_N == 256, "the CTA-pair kernel uses 256 x 256 tiles");
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  constexpr int TILE_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
  const int tile0 = CTA2 ? blockIdx.x / 2 : blockIdx.x;
  const int tile_step = CTA2 ? gridDim.x / 2 : gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_ab = smem_base;
  const uint32_t smem_c = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = smem_c + 4 * C_STAGE_BYTES;
  // barrier layout (8 B each): full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2] ; then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::STAGES + 4);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_ptr_smem - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + TILE_M - 1) / TILE_M;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / BLOCK_K;
  // Raster: column panels are walked in groups of `n_group`; inside a group n runs fastest, then m, so the CTAs in flight
  // share a few row blocks of A and the group's panels of B (n_group x BLOCK_N x K x 2 bytes, sized by the host to stay
  // L2-resident for all num_m row blocks); A is streamed once per group.  n_group >= num_n is the plain n-fastest order.
  // (The SwiGLU gate/up weights are 67 MB: walked as one group they fell out of L2 between row blocks and were
  // re-fetched from HBM ~25 times per launch -- 2.4 GB of DRAM traffic for 0.83 GB of operands.)
  auto tile_mn = [&](int tile, int& m_blk, int& n_blk) { raster_tile_mn(tile, num_m, num_n, n_group, m_blk, n_blk); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), (CTA2 ? 2 : 1) * (EPI_THREADS / 32));  // pair: both CTAs' epilogue warps release it
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTA2) {
      tmem_alloc_2cta(tmem_ptr_smem, C::TMEM_COLS);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_ptr_smem, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // Register hand-over (per warpgroup): the control warpgroup (TMA, MMA, allocator, one idle warp) keeps 56 registers,
  // the two epilogue warpgroups get 224 -- room for double-buffered TMEM loads and a whole tile's residual rows in
  // flight.  The instruction sits at the top of each role branch so ptxas allocates the branch against its own budget.
  if (warp < 4) setmaxnreg_dec<56>();
  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int m_blk, n_blk;
        tile_mn(tile, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_ab + stage * C::STAGE_BYTES;
          if (CTA2) {
            // both CTAs' bytes are accounted on the leader's barrier; the leader arms it for the pair
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            tma_load_2d_2cta(sa, &tmap_a, kb * BLOCK_K, m_blk * TILE_M + cta_rank * BLOCK_M, full_bar(stage));
            tma_load_2d_2cta(sa + C::A_BYTES, &tmap_b, kb * BLOCK_K, n_blk * BLOCK_N + cta_rank * (BLOCK_N / 2),
                             full_bar(stage));
          } else {
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, kb * BLOCK_K, m_blk * BLOCK_M, full_bar(stage));
            tma_load_2d(sa + C::A_BYTES, &tmap_b, kb * BLOCK_K, n_blk * BLOCK_N, full_bar(stage));
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_ab + stage * C::STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa, 0, 1024);
          const uint64_t bdesc = umma_desc_sw128(sa + C::A_BYTES, 0, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle row: +2 in the >>4 address field
            if (CTA2) umma_bf16_ss_2cta(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (CTA2) umma_commit_2cta(empty_bar(stage), 3);
          else umma_commit(empty_bar(stage));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CTA2) umma_commit_2cta(tfull_bar(acc), 3);
        else umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ----------------------------------
    setmaxnreg_inc<224>();
    const int ew = warp - 4;           // 0..7
    const int quarter = warp & 3;      // TMEM lane quarter this warp may touch (warp id % 4)
    const int half = ew >> 2;          // which half of the tile's columns
    const int row = quarter * 32 + lane;
    constexpr int COLS_PER_HALF = BLOCK_N / 2;
    constexpr int CHUNKS = COLS_PER_HALF / C_CHUNK;
    // staging: [half][2 buffers][128 rows x 64 B]; every warp owns the 32 rows of its TMEM lane quarter and stores them
    // with its OWN TMA store (box 32 columns x 32 rows), so no warp ever waits for another one inside the epilogue
    const uint32_t stage_c0 = smem_c + half * 2 * C_STAGE_BYTES + quarter * (C_STAGE_BYTES / 4);
    const bool issuer = lane == 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    uint32_t gate[EPI == EPI_SWIGLU ? 16 : 1];  // EPI_SWIGLU: bf16-rounded gate chunk (packed pairs) waiting for its up chunk
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      int m_blk, n_blk;
      tile_mn(tile, m_blk, n_blk);
      const int m_base = m_blk * TILE_M + static_cast<int>(cta_rank) * BLOCK_M;  // this CTA's 128 rows of the tile
      const int gm = m_base + row;
      // Residual rows: requested for ALL chunks of the tile before waiting for the accumulator, with a coalesced
      // mapping -- lane l fetches the 16-byte piece l%4 of rows 8j + l/4 (j < 4) of the warp's 32 x 64 B block, i.e.
      // 8 rows x 64 contiguous bytes per load instruction instead of 32 rows x 16 bytes.  The pieces reach their
      // row-per-thread owners through the warp's staging buffer just before it receives the C chunk.
      constexpr bool HAS_RES = EPI == EPI_BIAS_LS_RES || EPI == EPI_RES;
      uint4 rres[HAS_RES ? CHUNKS : 1][4];
      if (HAS_RES) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int gnp = n_blk * BLOCK_N + half * COLS_PER_HALF + c * C_CHUNK;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rrow = m_base + quarter * 32 + 8 * j + (lane >> 2);
            rres[c][j] = make_uint4(0, 0, 0, 0);
            if (rrow < M && gnp < N)
              rres[c][j] = *reinterpret_cast<const uint4*>(ea.res + static_cast<size_t>(rrow) * ea.ldres + gnp + 8 * (lane & 3));
          }
        }
      }
      // Column vectors (bias, LayerScale gamma) and RoPE table rows of chunk c+1 are requested while chunk c is
      // processed (chunk 0: here, before waiting for the accumulator): with 230 KB of shared memory the L1 is a few KB,
      // so each of these loads is an L2 round trip that must not sit between the TMEM load and the math.
      constexpr bool HAS_BIAS = EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_LS_RES || EPI == EPI_BIAS_ROPE2D;
      uint4 bnext[4], gnext[4], snext[4], cnext[4];
      bool rope_next = false;
      auto prefetch_side = [&](int c) {
        const int gnp = n_blk * BLOCK_N + half * COLS_PER_HALF + c * C_CHUNK;
        if (gnp >= N) return;
        if (HAS_BIAS) {
          const uint4* bp = reinterpret_cast<const uint4*>(ea.bias + gnp);
#pragma unroll
          for (int q = 0; q < 4; ++q) bnext[q] = __ldg(bp + q);
        }
        if (EPI == EPI_BIAS_LS_RES) {
          const uint4* gp = reinterpret_cast<const uint4*>(ea.gamma + gnp);
#pragma unroll
          for (int q = 0; q < 4; ++q) gnext[q] = __ldg(gp + q);
        }
        if (EPI == EPI_BIAS_ROPE2D) {
          const int tok = gm % ea.ntok;
          rope_next = gnp < ea.rope_cols && gm < M && tok >= ea.npt;
          if (rope_next) {
            // chunk-major table: [sin | cos][hd / 8][token][8] -- the 32 rows (consecutive tokens) of a warp read
            // 512 contiguous bytes per load instead of 32 separate 16-byte pieces
            const int nrt = ea.ntok - ea.npt, hc = ea.hd >> 3;
            const uint4* tr = reinterpret_cast<const uint4*>(ea.rope) + static_cast<size_t>((gnp % ea.hd) >> 3) * nrt + (tok - ea.npt);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              snext[q] = __ldg(tr + static_cast<size_t>(q) * nrt);
              cnext[q] = __ldg(tr + static_cast<size_t>(hc + q) * nrt);
            }
          }
        }
      };
      prefetch_side(0);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      // TMEM loads are double-buffered: chunk c+1 is requested before chunk c is processed
      uint32_t rbuf[2][32];
      const uint32_t tm_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N + half * COLS_PER_HALF;
      tmem_ld_32x32b_x32(tm_row, rbuf[0]);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = half * COLS_PER_HALF + c * C_CHUNK;
        const int gn = n_blk * BLOCK_N + col0;
        uint32_t (&r)[32] = rbuf[c & 1];
        const bool col_ok = gn < N;  // N % 32 == 0 is enforced on the host
        uint4 bvec[4], gvec[4], rsin[4], rcos[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { bvec[q] = bnext[q]; gvec[q] = gnext[q]; rsin[q] = snext[q]; rcos[q] = cnext[q]; }
        const bool do_rope = rope_next;
        if (c + 1 < CHUNKS) prefetch_side(c + 1);
        tmem_ld_wait();
        if (c + 1 < CHUNKS) tmem_ld_32x32b_x32(tm_row + (c + 1) * C_CHUNK, rbuf[(c + 1) & 1]);
        if (c == CHUNKS - 1) {
          // all of this warp's TMEM reads for the tile are done -> release the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTA2) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));  // the leader's MMA warp waits on it
            else mbar_arrive(tempty_bar(acc));
          }
        }
        // Epilogue arithmetic on packed words.  The reference's rounding points are one bf16 rounding per torch op; a
        // product or a sum of two bf16 values rounded once IS mul.rn.bf16x2 / add.rn.bf16x2 (the fp32 product of two
        // bf16 values is exact, and an fp32 sum of two bf16 values rounds to the same bf16 as the exact sum), so the
        // LayerScale multiply, the residual add and the RoPE products / sum run two elements per instruction with no
        // unpacking; the GELU polynomial runs on fp32x2.
        uint32_t packed[16];
        if (col_ok) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (HAS_BIAS) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 b = bvec[q];
              const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(bw[j]);
                v[q * 8 + j * 2] += f.x;
                v[q * 8 + j * 2 + 1] += f.y;
              }
            }
          }
          // first rounding point: the Linear's bf16 output, as packed pairs
#pragma unroll
          for (int i = 0; i < 16; ++i) packed[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          if (EPI == EPI_SWIGLU) {
            // weight rows are interleaved in 32-row blocks [gate | up]: even chunks carry gate features,
            // the following odd chunk the matching up features: act = bf16(bf16(silu(g)) * u)
            if ((c & 1) == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) gate[i] = packed[i];
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 g = unpack_bf16x2(gate[i]);
                packed[i] = mul_bf16x2(pack_bf16x2(silu_f(g.x), silu_f(g.y)), packed[i]);
              }
            }
          }
          if (EPI == EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 x = unpack_bf16x2(packed[i]);
              packed[i] = gelu_erf_fast_pack2(x.x, x.y);
            }
          }
          if (EPI == EPI_BIAS_ROPE2D) {
            if (do_rope) {
              // interleaved pairs (x0, x1): o0 = bf16(x0 c0) + bf16(-x1 s0), o1 = bf16(x1 c1) + bf16(x0 s1)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t sw4[4] = {rsin[q].x, rsin[q].y, rsin[q].z, rsin[q].w};
                const uint32_t cw4[4] = {rcos[q].x, rcos[q].y, rcos[q].z, rcos[q].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t x = packed[q * 4 + j];
                  const uint32_t xs = __byte_perm(x, x, 0x1032) ^ 0x00008000u;  // (-x1, x0)
                  packed[q * 4 + j] = add_bf16x2(mul_bf16x2(x, cw4[j]), mul_bf16x2(xs, sw4[j]));
                }
              }
            }
          }
          if (EPI == EPI_BIAS_LS_RES) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t gw[4] = {gvec[q].x, gvec[q].y, gvec[q].z, gvec[q].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) packed[q * 4 + j] = mul_bf16x2(packed[q * 4 + j], gw[j]);
            }
          }
          if (EPI == EPI_BIAS_LS_RES || EPI == EPI_RES) {
            // exchange: coalesced pieces -> the free staging buffer (same swizzled row layout as the C chunk) -> rows
            if (issuer) tma_store_wait_read<1>();
            __syncwarp();
            const uint32_t xbase = stage_c0 + cbuf * C_STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int xr = 8 * j + (lane >> 2);
              const uint4 x = rres[HAS_RES ? c : 0][j];
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                               xbase + xr * 64 + ((static_cast<uint32_t>(lane & 3) ^ ((xr >> 1) & 3)) << 4)),
                           "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w)
                           : "memory");
            }
            __syncwarp();
            uint4 xrow[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(xrow[q].x), "=r"(xrow[q].y), "=r"(xrow[q].z), "=r"(xrow[q].w)
                           : "r"(xbase + lane * 64 + ((static_cast<uint32_t>(q) ^ ((lane >> 1) & 3)) << 4)));
            __syncwarp();  // every lane has its row: the buffer may now receive the C chunk
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t xw[4] = {xrow[q].x, xrow[q].y, xrow[q].z, xrow[q].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) packed[q * 4 + j] = add_bf16x2(xw[j], packed[q * 4 + j]);
            }
          }
        }
        if (EPI == EPI_SWIGLU && (c & 1) == 0) continue;  // nothing to store for a gate chunk
        // staging buffer reuse: this warp's TMA store issued two chunks ago must have finished reading smem
        if (issuer) tma_store_wait_read<1>();
        __syncwarp();
        if (col_ok) {
          const uint32_t rbase = stage_c0 + cbuf * C_STAGE_BYTES + lane * 64;
          const uint32_t sw = (row >> 1) & 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t addr = rbase + ((static_cast<uint32_t>(q) ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[q * 4]),
                         "r"(packed[q * 4 + 1]), "r"(packed[q * 4 + 2]), "r"(packed[q * 4 + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (issuer && col_ok) {
          const int gn_out = EPI == EPI_SWIGLU ? (gn - C_CHUNK) / 2 : gn;
          tma_store_2d(&tmap_c, stage_c0 + cbuf * C_STAGE_BYTES, gn_out, m_base + quarter * 32);
        }
        if (issuer) tma_store_commit();
        cbuf ^= 1;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (issuer) tma_store_wait<0>();
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all();  // the peer may still signal this CTA's barriers / read its smem until here
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BLOCK_N, int EPI, bool CTA2>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, int K,
                const EpiArgs& ea, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, CTA2>;
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, EPI, CTA2>;
  static bool attr_set = false;
  if (!attr_set) {
    GAR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int tile_m = CTA2 ? 2 * BLOCK_M : BLOCK_M;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = ((M + tile_m - 1) / tile_m) * num_n;
  const int fly = CTA2 ? std::min(num_tiles, gar_grid_sms() / 2) : std::min(num_tiles, gar_grid_sms());
  const int n_group = raster_group_panels(num_n, fly, tile_m, BLOCK_N, K);
  if (CTA2) {
    const int clusters = std::min(num_tiles, gar_grid_sms() / 2);
    GAR_CUDA_CHECK(gar_launch(kern, dim3(2 * clusters), dim3(NUM_THREADS), C::SMEM_BYTES, stream, false, dim3(2, 1, 1), ta,
                              tb, tc, M, N, K, n_group, ea));
    return GAR_OK;
  }
  const int grid = num_tiles < gar_grid_sms() ? num_tiles : gar_grid_sms();
  kern<<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, tc, M, N, K, n_group, ea);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

template <int BLOCK_N, bool CTA2 = false>
int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, int K,
                 const EpiArgs& ea, cudaStream_t stream) {
  switch (epi) {
    case EPI_NONE: return launch_gemm<BLOCK_N, EPI_NONE, CTA2>(ta, tb, tc, M, N, K, ea, stream);
    case EPI_BIAS: return launch_gemm<BLOCK_N, EPI_BIAS, CTA2>(ta, tb, tc, M, N, K, ea, stream);
    case EPI_BIAS_GELU: return launch_gemm<BLOCK_N, EPI_BIAS_GELU, CTA2>(ta, tb, tc, M, N, K, ea, stream);
    case EPI_BIAS_LS_RES: return launch_gemm<BLOCK_N, EPI_BIAS_LS_RES, CTA2>(ta, tb, tc, M, N, K, ea, stream);
    case EPI_RES: return launch_gemm<BLOCK_N, EPI_RES, CTA2>(ta, tb, tc, M, N, K, ea, stream);
    case EPI_BIAS_ROPE2D: return launch_gemm<BLOCK_N, EPI_BIAS_ROPE2D, CTA2>(ta, tb, tc, M, N, K, ea, stream);
    case EPI_SWIGLU: return launch_gemm<BLOCK_N, EPI_SWIGLU, CTA2>(ta, tb, tc, M, N, K, ea, stream);
  }
  gar_set_error("gar_gemm_bf16: unknown epilogue %d", epi);
  return GAR_ERR_ARG;
}

}  // namespace

static int gemm_impl(const void* A, int lda, const void* B, int ldb, void* Cout, int ldc, int M, int N, int K, int epi,
                     const void* bias, const void* gamma, const void* res, int ldres, const void* rope, int ntok,
                     int npt, int hd, int rope_cols, int block_n_hint, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0 || (K % BLOCK_K) != 0 || (N % 32) != 0 || (lda % 8) || (ldb % 8) || (ldc % 8)) {
    gar_set_error("gar_gemm_bf16: unsupported shape M=%d N=%d K=%d lda=%d ldb=%d ldc=%d (need K%%64==0, N%%32==0, ld%%8==0)",
                  M, N, K, lda, ldb, ldc);
    return GAR_ERR_SHAPE;
  }
  if ((epi == EPI_BIAS || epi == EPI_BIAS_GELU || epi == EPI_BIAS_LS_RES || epi == EPI_BIAS_ROPE2D) && !bias)
    return GAR_ERR_ARG;
  if (epi == EPI_BIAS_ROPE2D && (!rope || hd % 32 || ntok <= 0 || rope_cols % 32)) return GAR_ERR_ARG;
  if (epi == EPI_BIAS_LS_RES && !gamma) return GAR_ERR_ARG;
  if ((epi == EPI_BIAS_LS_RES || epi == EPI_RES) && (!res || (ldres % 8))) return GAR_ERR_ARG;
  int bn = block_n_hint;
  if (bn != 64 && bn != 128 && bn != 256) {
    // widest tile that still gives every SM a tile; wide tiles halve A re-reads from L2
    const int num_m = (M + BLOCK_M - 1) / BLOCK_M;
    const int sms = gar_num_sms();
    bn = 64;
    const int cand[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      if (num_m * ((N + cand[i] - 1) / cand[i]) >= sms) { bn = cand[i]; break; }
    }
  }
  if (epi == EPI_SWIGLU) {
    if (N % 64) return GAR_ERR_SHAPE;
    if (bn < 128) bn = 128;  // a column group must hold whole [gate | up] chunk pairs
  }
  // CTA-pair kernel (256 x 256 tiles) whenever N allows it and there are enough tiles for every pair
  const bool cta2 = gar_global_option("gemm_2cta") != 0 && block_n_hint == 0 && N % 256 == 0 &&
                    ((M + 255) / 256) * (N / 256) >= gar_num_sms() / 2;
  if (cta2) bn = 256;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = gar_make_tmap_2d(&ta, A, K, M, lda, BLOCK_K, BLOCK_M, GAR_SW_128))) return rc;
  if ((rc = gar_make_tmap_2d(&tb, B, K, N, ldb, BLOCK_K, cta2 ? 128 : bn, GAR_SW_128))) return rc;
  // C is stored warp by warp: box = 32 columns x the 32 rows of one TMEM lane quarter
  if ((rc = gar_make_tmap_2d(&tc, Cout, epi == EPI_SWIGLU ? N / 2 : N, M, ldc, C_CHUNK, 32, GAR_SW_64))) return rc;
  EpiArgs ea{static_cast<const bf16*>(bias), static_cast<const bf16*>(gamma), static_cast<const bf16*>(res), ldres,
             static_cast<const bf16*>(rope), ntok, npt, hd, rope_cols};
  if (cta2) return dispatch_epi<256, true>(epi, ta, tb, tc, M, N, K, ea, stream);
  if (bn == 256) return dispatch_epi<256>(epi, ta, tb, tc, M, N, K, ea, stream);
  if (bn == 128) return dispatch_epi<128>(epi, ta, tb, tc, M, N, K, ea, stream);
  return dispatch_epi<64>(epi, ta, tb, tc, M, N, K, ea, stream);
}

// Host-only view of the raster (no CUDA call): panels per group for a launch of this geometry and the (row block,
// column panel) of `tile`; fly = tiles in flight (CTAs, or CTA pairs for 256-row tiles).
extern "C" int gar_gemm_raster_probe(int M, int N, int K, int tile_m, int block_n, int fly, int tile, int* n_group,
                                     int* m_blk, int* n_blk) {
  if (M <= 0 || N <= 0 || K <= 0 || tile_m <= 0 || block_n <= 0 || fly <= 0 || !n_group) return GAR_ERR_ARG;
  const int num_m = (M + tile_m - 1) / tile_m, num_n = (N + block_n - 1) / block_n;
  *n_group = raster_group_panels(num_n, std::min(fly, num_m * num_n), tile_m, block_n, K);
  if (m_blk && n_blk) {
    if (tile < 0 || tile >= num_m * num_n) return GAR_ERR_ARG;
    raster_tile_mn(tile, num_m, num_n, *n_group, *m_blk, *n_blk);
  }
  return GAR_OK;
}

extern "C" int gar_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* Cout, int ldc, int M, int N, int K,
                             int epi, const void* bias, const void* gamma, const void* res, int ldres,
                             int block_n_hint, void* stream) {
  if (epi == EPI_BIAS_ROPE2D) {
    gar_set_error("gar_gemm_bf16: GAR_EPI_BIAS_ROPE2D needs gar_gemm_bf16_rope2d");
    return GAR_ERR_ARG;
  }
  return gemm_impl(A, lda, B, ldb, Cout, ldc, M, N, K, epi, bias, gamma, res, ldres, nullptr, 0, 0, 0, 0, block_n_hint,
                   stream);
}

// qkv projection of the PE ViT with the 2-D RoPE of q,k fused into the epilogue:
// C = rope2d(bf16(A B^T + bias)) on columns < rope_cols for tokens (row % ntok) >= npt.
extern "C" int gar_gemm_bf16_rope2d(const void* A, int lda, const void* B, int ldb, void* Cout, int ldc, int M, int N,
                                    int K, const void* bias, const void* rope, int ntok, int npt, int hd,
                                    int rope_cols, void* stream) {
  return gemm_impl(A, lda, B, ldb, Cout, ldc, M, N, K, EPI_BIAS_ROPE2D, bias, nullptr, nullptr, 0, rope, ntok, npt, hd,
                   rope_cols, 0, stream);
}
