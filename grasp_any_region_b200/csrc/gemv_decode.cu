// Decode-step linear layers: out[M<=8, N] = epi( norm?(X)[M,K] * W[N,K]^T ), latency-first.
//
// Replaces, per generated token, HF Llama's q/k/v/o/gate/up/down/lm_head GEMV-like cuBLAS calls AND the
// RMSNorm / SiLU*up / residual elementwise launches around them.  At batch <= 8 these
// are pure weight streams (2.47 GB per step for GAR-1B); what decides the step time is the fixed latency
// of every launch, so this kernel has no TMEM allocation, no cluster, no TMA descriptor and no second
// pass: each warp streams its K-slice of 16..64 weight rows straight from HBM into mma.sync A-fragments
// (16-byte loads, a permuted-k mapping keeps A and B fragments consistent), the <=8 activation rows are
// the B operand read from L2, and one shared-memory reduction over the 8 K-slices finishes the tile.
// (Tensor throughput is irrelevant here -- HMMA is used only to keep the instruction count per weight
// byte low enough that the SMs never limit the stream.)
//
// Fusions, all with the reference's rounding points:
//   NORM  : LlamaRMSNorm of X applied on the fly: x' = bf16(w_k * bf16(x * rstd)); rstd from the
//           per-row partial sums of squares the producing kernel emitted (fixed-order sum).
//   EPI_RES    : out = bf16(res + bf16(acc)); also emits this tile's sum of squares of `out`.
//   EPI_SWIGLU : rows interleaved [32 gate | 32 up]: out = bf16(bf16(silu(bf16 g)) * bf16 u).
// Weight loads are issued BEFORE the programmatic-dependency wait, so they overlap the previous kernel.
#include "../../include/gar_b200.h"
#include "common.cuh"

namespace {

enum { GV_EPI_NONE = 0, GV_EPI_RES = 1, GV_EPI_SWIGLU = 2 };
constexpr int GV_NW_MAX = 4096;   // RMSNorm weights staged in shared memory up to this K (8 KB); longer rows read them from L2
constexpr int GV_PART_MAX = 256;  // producer partial sums of squares staged in shared memory up to this many CTAs (8 KB)

struct GvArgs {
  const bf16* W; int ldw;
  const bf16* X; int ldx;
  bf16* out; int ldo;
  const bf16* res; int ldres;
  int M, N, K;
  const float* sumsq_in; int n_part; const bf16* norm_w; float eps;  // NORM prologue
  float* sumsq_out;                                                   // EPI_RES: [gridDim.x][8]
  int stage_norm;                                                     // NORM operands through shared memory
  unsigned long long* timeline;  // optional debug: per-CTA globaltimer stamps [blockIdx.x][8] (NULL in production)
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GV_STAMP(i) if (a.timeline != nullptr && threadIdx.x == 0) a.timeline[static_cast<size_t>(blockIdx.x) * 8 + (i)] = gtime_ns()

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_stream(const bf16* p) {  // weights: read once, keep out of L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// WARPS K-slices per CTA: 8 (two CTAs per SM) or, for long K (down_proj), 16 so that a 128-CTA grid
// still keeps enough weight bytes in flight per SM.
template <int R, int EPI, bool NORM, int GV_WARPS>
__global__ void __launch_bounds__(GV_WARPS * 32, 16 / GV_WARPS) gemv_decode_kernel(GvArgs a) {
  constexpr int GV_THREADS = GV_WARPS * 32;
  constexpr int U = 8 / R;  // k-steps (32 k each) whose weight loads are in flight together
  pdl_launch_dependents();
  GV_STAMP(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int row0 = blockIdx.x * 16 * R;
  const int k_per_warp = a.K / GV_WARPS;
  const int kbeg = warp * k_per_warp;
  const int steps = k_per_warp / 32;

  float acc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;

  const bf16* wlo[R];
  const bf16* whi[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    wlo[r] = a.W + static_cast<size_t>(row0 + r * 16 + g) * a.ldw + kbeg + 8 * tig;
    whi[r] = wlo[r] + static_cast<size_t>(8) * a.ldw;
  }
  const bf16* xp = a.X + static_cast<size_t>(g) * a.ldx + kbeg + 8 * tig;
  const bool xvalid = g < a.M;

  uint4 wl[U][R], wh[U][R];
  // first chunk of weights before the dependency wait (they never depend on the previous kernel)
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (u < steps) {
        wl[u][r] = ldg_stream(wlo[r] + u * 32);
        wh[u][r] = ldg_stream(whi[r] + u * 32);
      }
    }
  // The RMSNorm weight vector never depends on the previous kernel: it is staged into shared memory by cp.async
  // BEFORE the dependency wait (no registers held, no round trip left on the critical path after the wait).
  __shared__ __align__(16) bf16 s_nw[NORM ? GV_NW_MAX : 8];
  __shared__ __align__(16) float s_ss[NORM ? GV_PART_MAX * 8 : 4];
  const bool nw_smem = NORM && a.stage_norm && a.K <= GV_NW_MAX;
  const bool ss_smem = NORM && a.stage_norm && a.n_part <= GV_PART_MAX;
  if (nw_smem) {
    for (int i = threadIdx.x; i < a.K / 8; i += GV_THREADS)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_nw + 8 * i)), "l"(a.norm_w + 8 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  GV_STAMP(1);
  pdl_wait();
  GV_STAMP(2);

  // ONE round trip after the wait: the producer's partial sums of squares (every thread fetches one 16-byte piece
  // of the [n_part][8] table) leave together with the activations of the first chunk
  float4 ssv[GV_PART_MAX * 2 / GV_THREADS > 0 ? GV_PART_MAX * 2 / GV_THREADS : 1];
  if (ss_smem) {
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(ssv) / sizeof(ssv[0])); ++i) {
      const int q = threadIdx.x + i * GV_THREADS;
      if (q < a.n_part * 2) ssv[i] = *reinterpret_cast<const float4*>(a.sumsq_in + 4 * q);
    }
  }
  uint4 xv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    xv[u] = make_uint4(0, 0, 0, 0);
    if (u < steps && xvalid) xv[u] = *reinterpret_cast<const uint4*>(xp + u * 32);
  }

  float rstd = 0.f;
  if (NORM) {
    if (ss_smem) {
#pragma unroll
      for (int i = 0; i < static_cast<int>(sizeof(ssv) / sizeof(ssv[0])); ++i) {
        const int q = threadIdx.x + i * GV_THREADS;
        if (q < a.n_part * 2) *reinterpret_cast<float4*>(s_ss + 4 * q) = ssv[i];
      }
    }
    if (nw_smem) asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (ss_smem || nw_smem) __syncthreads();
    // fixed-order sum of the partial sums of squares for batch row g (same order on every path and in every CTA)
    float acc16[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc16[j] = 0.f;
    if (xvalid) {
      for (int p0 = tig; p0 < a.n_part; p0 += 64) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int p = p0 + 4 * j;
          if (p < a.n_part) acc16[j] += ss_smem ? s_ss[p * 8 + g] : a.sumsq_in[static_cast<size_t>(p) * 8 + g];
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
    rstd = rsqrtf(s / static_cast<float>(a.K) + a.eps);
  }
  GV_STAMP(3);

  for (int s0 = 0; s0 < steps; s0 += U) {
    if (NORM) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (s0 + u < steps) {
          const int nwo = kbeg + 8 * tig + (s0 + u) * 32;
          const uint4 nw = nw_smem ? *reinterpret_cast<const uint4*>(s_nw + nwo) : __ldg(reinterpret_cast<const uint4*>(a.norm_w + nwo));
          uint32_t* xw = reinterpret_cast<uint32_t*>(&xv[u]);
          const uint32_t* ww = reinterpret_cast<const uint32_t*>(&nw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(xw[j]), w2 = unpack_bf16x2(ww[j]);
            xw[j] = pack_bf16x2(w2.x * bf16r(f.x * rstd), w2.y * bf16r(f.y * rstd));
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (s0 + u < steps) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          mma_bf16_16816(acc[r], wl[u][r].x, wh[u][r].x, wl[u][r].y, wh[u][r].y, xv[u].x, xv[u].y);
          mma_bf16_16816(acc[r], wl[u][r].z, wh[u][r].z, wl[u][r].w, wh[u][r].w, xv[u].z, xv[u].w);
        }
      }
    }
    // next chunk of weights and activations
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (s0 + U + u < steps) {
          wl[u][r] = ldg_stream(wlo[r] + (s0 + U + u) * 32);
          wh[u][r] = ldg_stream(whi[r] + (s0 + U + u) * 32);
        }
      }
      xv[u] = make_uint4(0, 0, 0, 0);
      if (s0 + U + u < steps && xvalid) xv[u] = *reinterpret_cast<const uint4*>(xp + (s0 + U + u) * 32);
    }
  }

  // ---- reduce the 8 K-slices (fixed order), then the fused epilogue ----
  GV_STAMP(4);
  __shared__ float part[GV_WARPS][16 * R][9];  // [warp][feature][batch], +1 pad
#pragma unroll
  for (int r = 0; r < R; ++r) {
    // c0,c1 -> (feature g, batch 2tig, 2tig+1); c2,c3 -> (feature g+8, ...)
    part[warp][r * 16 + g][2 * tig] = acc[r][0];
    part[warp][r * 16 + g][2 * tig + 1] = acc[r][1];
    part[warp][r * 16 + g + 8][2 * tig] = acc[r][2];
    part[warp][r * 16 + g + 8][2 * tig + 1] = acc[r][3];
  }
  __syncthreads();
  GV_STAMP(5);
  constexpr int FEATS = 16 * R;
  constexpr int OUT_FEATS = EPI == GV_EPI_SWIGLU ? FEATS / 2 : FEATS;
  for (int idx = threadIdx.x; idx < OUT_FEATS * 8; idx += GV_THREADS) {
    const int f = idx % OUT_FEATS, b = idx / OUT_FEATS;
    float y;
    if (EPI == GV_EPI_SWIGLU) {
      // FEATS = 64 rows = [32 gate | 32 up]
      float gsum = 0.f, usum = 0.f;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) {
        gsum += part[w][f][b];
        usum += part[w][32 + f][b];
      }
      y = bf16r(silu_f(bf16r(gsum))) * bf16r(usum);
    } else {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) s += part[w][f][b];
      y = bf16r(s);
    }
    const int gf = blockIdx.x * OUT_FEATS + f;
    float o = 0.f;
    if (b < a.M) {
      if (EPI == GV_EPI_RES) y += bf2f(a.res[static_cast<size_t>(b) * a.ldres + gf]);
      const bf16 ob = __float2bfloat16_rn(y);
      a.out[static_cast<size_t>(b) * a.ldo + gf] = ob;
      o = bf2f(ob);
    }
    if (EPI == GV_EPI_RES && a.sumsq_out != nullptr) {
      // OUT_FEATS in {16, 32}: the threads of one batch row sit in one (half-)warp
      float q = o * o;
#pragma unroll
      for (int off = OUT_FEATS / 2; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
      if (f == 0) a.sumsq_out[static_cast<size_t>(blockIdx.x) * 8 + b] = q;
    }
  }
  GV_STAMP(6);
}

// h[b] = table[tok[b]] and the row's sum of squares (the first layer's RMSNorm statistics)
__global__ void embed_gather_sumsq_kernel(const bf16* __restrict__ table, const int* __restrict__ ids, bf16* __restrict__ out,
                                          float* __restrict__ sumsq, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(ids[b]) * D);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(b) * D);
  float q = 0.f;
  for (int v = threadIdx.x; v < D / 8; v += blockDim.x) {
    const uint4 x = __ldg(src + v);
    dst[v] = x;
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&x);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(xw[j]);
      q += f.x * f.x + f.y * f.y;
    }
  }
  __shared__ float sq[32];
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) sq[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sq[w];
    sumsq[b] = t;  // n_part == 1 layout: [1][8]
  }
}

template <int R, int EPI>
int launch_gv(const GvArgs& a, bool norm, cudaStream_t s) {
  const int out_feats = (EPI == GV_EPI_SWIGLU ? 8 : 16) * R;
  const int n_out = EPI == GV_EPI_SWIGLU ? a.N / 2 : a.N;
  dim3 grid(n_out / out_feats);
  // 16 K-slices when K is long and the grid alone cannot fill the machine (R == 1 only: register budget)
  const bool wide = R == 1 && a.K % 512 == 0 && a.K >= 4096 && static_cast<int>(grid.x) < 2 * gar_num_sms() &&
                    gar_global_option("gemv_wide");
#define GV_LAUNCH(NORM_, W_) \
  GAR_CUDA_CHECK(gar_launch(gemv_decode_kernel<R, EPI, NORM_, W_>, grid, dim3(W_ * 32), 0, s, true, dim3(1, 1, 1), a))
  if constexpr (R == 1) if (wide) {
    if (norm) GV_LAUNCH(true, 16);
    else GV_LAUNCH(false, 16);
    return GAR_OK;
  }
  if (norm) GV_LAUNCH(true, 8);
  else GV_LAUNCH(false, 8);
#undef GV_LAUNCH
  return GAR_OK;
}

}  // namespace

// epi: 0 none, 1 residual (+ optional sumsq_out[(N/rows_per_cta)][8]), 2 swiglu (W rows interleaved, out is [M, N/2]).
// norm_w != NULL enables the RMSNorm prologue with sumsq_in[n_part][8].  Returns GAR_ERR_SHAPE when the shape does
// not fit this kernel (callers then use gar_gemm_skinny_bf16 / gar_rmsnorm / gar_swiglu).
static unsigned long long* g_gemv_timeline = nullptr;
// debug hook: device buffer of >= 8 * grid int64 receiving per-CTA globaltimer stamps of the NEXT gemv launches
// (0 entry, 1 first weights issued, 2 dependency wait done, 3 norm statistics ready, 4 weight stream consumed,
//  5 K-slice partials in shared memory, 6 epilogue stored); NULL disables
extern "C" int gar_gemv_debug_timeline(void* dev_buf) {
  g_gemv_timeline = static_cast<unsigned long long*>(dev_buf);
  return GAR_OK;
}

extern "C" int gar_gemv_decode(const void* X, int ldx, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                               int epi, const void* res, int ldres, const void* norm_w, const float* sumsq_in,
                               int n_part, float eps, float* sumsq_out, int* rows_per_cta, void* stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (M <= 0 || M > 8 || K % 256 || (ldx % 8) || (ldw % 8) || epi < 0 || epi > 2) return GAR_ERR_SHAPE;
  if (norm_w != nullptr && (sumsq_in == nullptr || n_part <= 0)) return GAR_ERR_ARG;
  GvArgs a;
  a.W = static_cast<const bf16*>(W); a.ldw = ldw;
  a.X = static_cast<const bf16*>(X); a.ldx = ldx;
  a.out = static_cast<bf16*>(out); a.ldo = ldo;
  a.res = static_cast<const bf16*>(res); a.ldres = ldres;
  a.M = M; a.N = N; a.K = K;
  a.sumsq_in = sumsq_in; a.n_part = n_part; a.norm_w = static_cast<const bf16*>(norm_w); a.eps = eps;
  a.sumsq_out = sumsq_out;
  a.timeline = g_gemv_timeline;
  a.stage_norm = gar_global_option("gemv_norm_smem");
  const bool norm = norm_w != nullptr;
  if (epi == GV_EPI_SWIGLU) {
    if (N % 64) return GAR_ERR_SHAPE;
    if (rows_per_cta) *rows_per_cta = 32;
    return launch_gv<4, GV_EPI_SWIGLU>(a, norm, s);
  }
  if (epi == GV_EPI_RES) {
    if (!res || N % 16) return GAR_ERR_SHAPE;
    if (N % 32 == 0 && N / 32 >= 2 * gar_num_sms()) {
      if (rows_per_cta) *rows_per_cta = 32;
      return launch_gv<2, GV_EPI_RES>(a, norm, s);
    }
    if (rows_per_cta) *rows_per_cta = 16;
    return launch_gv<1, GV_EPI_RES>(a, norm, s);
  }
  if (N % 16) return GAR_ERR_SHAPE;
  // wide outputs (lm_head): 64 rows per CTA; otherwise 16 rows per CTA for the most CTAs in flight
  if (N % 64 == 0 && N / 64 >= 4 * gar_num_sms()) {
    if (rows_per_cta) *rows_per_cta = 64;
    return launch_gv<4, GV_EPI_NONE>(a, norm, s);
  }
  if (N % 32 == 0 && N / 32 >= 2 * gar_num_sms()) {
    if (rows_per_cta) *rows_per_cta = 32;
    return launch_gv<2, GV_EPI_NONE>(a, norm, s);
  }
  if (rows_per_cta) *rows_per_cta = 16;
  return launch_gv<1, GV_EPI_NONE>(a, norm, s);
}

extern "C" int gar_embed_gather_sumsq(const void* table, const int* ids32, void* out, float* sumsq, int rows, int D,
                                      void* stream) {
  if (D % 8 || rows <= 0 || rows > 8) return GAR_ERR_SHAPE;
  GAR_CUDA_CHECK(gar_launch(embed_gather_sumsq_kernel, dim3(rows), dim3(256), 0, static_cast<cudaStream_t>(stream), true,
                            dim3(1, 1, 1), static_cast<const bf16*>(table), ids32, static_cast<bf16*>(out), sumsq, D));
  return GAR_OK;
}
