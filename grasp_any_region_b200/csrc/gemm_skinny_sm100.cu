// out[M<=16, N] = epi(X[M,K] * W[N,K]^T) for the decode step: a weight-streaming GEMM.
//
// Replaces the cuBLAS GEMV-like calls HF Llama issues per generated token (q/k/v/o/gate/up/down/lm_head at
// M = batch).  HBM-bound: every weight byte is read once per step, so the design
// goal is bytes-in-flight, not tensor throughput:
//   * swap-AB: the 128 weight rows of a tile are the MMA M dimension, the (padded to 16) batch rows
//     are N -> one tcgen05.mma (M128,N16,K16) per 16 k; the accumulator is 16 TMEM columns.
//   * split-K across a thread-block cluster (up to 8 CTAs): every CTA streams its K-slice of the
//     weight tile through a deep TMA ring, then the partial [128 x 16] fp32 tiles are reduced through
//     distributed shared memory in fixed rank order (deterministic), each CTA finishing 128/cluster rows.
//   * epilogue keeps the reference's rounding points: bf16(acc) then optional bf16(res + .)
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int SK_BLOCK_W = 128;  // weight rows per tile (MMA M)
constexpr int SK_BLOCK_K = 64;
constexpr int SK_NB = 16;        // padded batch (MMA N)
constexpr int SK_STAGES = 5;
constexpr int SK_W_BYTES = SK_BLOCK_W * SK_BLOCK_K * 2;  // 16 KB
constexpr int SK_X_BYTES = SK_NB * SK_BLOCK_K * 2;       // 2 KB
constexpr int SK_STAGE_BYTES = SK_W_BYTES + SK_X_BYTES;
constexpr int SK_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES + 1024;

__global__ void __launch_bounds__(SK_THREADS, 2)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, int M, int N,
                   int K, int kb_per_split, bf16* __restrict__ out, int ldo, const bf16* __restrict__ res, int ldres) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * SK_STAGES + 1];
  __shared__ uint32_t tmem_ptr_sh;
  __shared__ float partial[SK_BLOCK_W][SK_NB + 1];  // +1: conflict-free column reads
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t csize = cluster.num_blocks();
  const uint32_t crank = cluster.block_rank();
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (SK_STAGES + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * SK_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x;
  const int split = blockIdx.y;  // == cluster rank (cluster spans the whole y extent)
  const int num_kb_total = K / SK_BLOCK_K;
  const int kb0 = split * kb_per_split;
  const int num_kb = max(0, min(kb_per_split, num_kb_total - kb0));

  pdl_launch_dependents();  // let the next kernel's prologue (and its weight prefetch) start early
  if (threadIdx.x == 0) {
    for (int s = 0; s < SK_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_ptr_sh), 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_sh);

  if (warp == 0) {
    if (lane == 0) {
      // Weights never depend on the previous kernel: stream the first ring of W tiles before the
      // programmatic-dependency wait, so the weight fetch overlaps the predecessor's tail.
      const int pre = min(num_kb, SK_STAGES);
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(full_bar(kb), SK_STAGE_BYTES);
        tma_load_2d(sbase + kb * SK_STAGE_BYTES, &tmap_w, (kb0 + kb) * SK_BLOCK_K, n_blk * SK_BLOCK_W, full_bar(kb));
      }
      pdl_wait();
      for (int kb = 0; kb < pre; ++kb)
        tma_load_2d(sbase + kb * SK_STAGE_BYTES + SK_W_BYTES, &tmap_x, (kb0 + kb) * SK_BLOCK_K, 0, full_bar(kb));
      int stage = 0;
      uint32_t phase = 1;  // the ring has been filled once
      for (int kb = pre; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), SK_STAGE_BYTES);
        const uint32_t sa = sbase + stage * SK_STAGE_BYTES;
        tma_load_2d(sa, &tmap_w, (kb0 + kb) * SK_BLOCK_K, n_blk * SK_BLOCK_W, full_bar(stage));
        tma_load_2d(sa + SK_W_BYTES, &tmap_x, (kb0 + kb) * SK_BLOCK_K, 0, full_bar(stage));
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(SK_BLOCK_W, SK_NB, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * SK_STAGE_BYTES;
        const uint64_t wdesc = umma_desc_sw128(sa, 0, 1024);
        const uint64_t xdesc = umma_desc_sw128(sa + SK_W_BYTES, 0, 1024);
#pragma unroll
        for (int k = 0; k < SK_BLOCK_K / 16; ++k)
          umma_bf16_ss(tmem, wdesc + 2u * k, xdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(done_bar);
    }
  } else {
    // epilogue warps: TMEM lane quarter = warp % 4 ; thread <-> weight row (output feature)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    float v[SK_NB];
    pdl_wait();  // `res` / `out` belong to the dependency chain
    if (num_kb > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem + (static_cast<uint32_t>(quarter * 32) << 16))
          : "memory");
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < SK_NB; ++i) v[i] = __uint_as_float(r[i]);
    } else {
#pragma unroll
      for (int i = 0; i < SK_NB; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < SK_NB; ++i) partial[row][i] = v[i];
  }
  // ---- deterministic split-K reduction through distributed shared memory ----
  cluster.sync();
  if (warp >= 2) {
    const int t = threadIdx.x - 64;              // 0..127
    const int rows_per_rank = SK_BLOCK_W / csize;  // csize in {1,2,4,8}
    const int items = rows_per_rank * SK_NB;       // (row, batch) pairs this CTA finishes
    for (int it = t; it < items; it += 128) {
      const int b = it / rows_per_rank;          // batch index: consecutive threads -> consecutive features
      const int rr = it % rows_per_rank;
      const int row = crank * rows_per_rank + rr;
      float acc = 0.f;
      for (uint32_t r = 0; r < csize; ++r) {
        const float* peer = cluster.map_shared_rank(&partial[0][0], r);
        acc += peer[row * (SK_NB + 1) + b];
      }
      const int gn = n_blk * SK_BLOCK_W + row;
      if (b < M && gn < N) {
        float y = bf16r(acc);
        if (res != nullptr) y += bf2f(res[static_cast<size_t>(b) * ldres + gn]);
        out[static_cast<size_t>(b) * ldo + gn] = __float2bfloat16_rn(y);
      }
    }
  }
  cluster.sync();  // peers' shared memory must outlive every remote read
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

}  // namespace

// X [M<=16, K] (ldx), W [N, K] (ldw) -> out [M, N] (ldo); epi: 0 none, 4 = + residual (GAR_EPI_RES)
extern "C" int gar_gemm_skinny_bf16(const void* X, int ldx, const void* Wt, int ldw, void* out, int ldo, int M, int N,
                                    int K, int epi, const void* res, int ldres, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || M > SK_NB || N <= 0 || K <= 0 || (K % SK_BLOCK_K) || (ldx % 8) || (ldw % 8)) {
    gar_set_error("gar_gemm_skinny_bf16: unsupported shape M=%d N=%d K=%d", M, N, K);
    return GAR_ERR_SHAPE;
  }
  if (epi != 0 && epi != 4) return GAR_ERR_ARG;
  if (epi == 4 && !res) return GAR_ERR_ARG;
  CUtensorMap tw, tx;
  int rc;
  if ((rc = gar_make_tmap_2d(&tw, Wt, K, N, ldw, SK_BLOCK_K, SK_BLOCK_W, GAR_SW_128))) return rc;
  if ((rc = gar_make_tmap_2d(&tx, X, K, M, ldx, SK_BLOCK_K, SK_NB, GAR_SW_128))) return rc;
  const int n_tiles = (N + SK_BLOCK_W - 1) / SK_BLOCK_W;
  const int num_kb = K / SK_BLOCK_K;
  // enough CTAs to keep ~2 per SM streaming, cluster size a power of two <= 8 and <= k-blocks/4
  int split = 1;
  const int sms = gar_num_sms();
  while (split < 8 && n_tiles * split < sms && num_kb / (split * 2) >= 4) split *= 2;
  const int kb_per_split = (num_kb + split - 1) / split;
  static bool attr_set = false;
  if (!attr_set) {
    GAR_CUDA_CHECK(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
    attr_set = true;
  }
  GAR_CUDA_CHECK(gar_launch(gemm_skinny_kernel, dim3(n_tiles, split, 1), dim3(SK_THREADS), SK_SMEM, stream, true,
                            dim3(1, split, 1), tw, tx, M, N, K, kb_per_split, static_cast<bf16*>(out), ldo,
                            epi == 4 ? static_cast<const bf16*>(res) : static_cast<const bf16*>(nullptr), ldres));
  return GAR_OK;
}
