// Microbenchmark (round 2): what decides the HBM rate of a shared-memory weight ring fed by 1-D bulk async copies?
//   bytes in flight per SM (stages x stage bytes, CTAs per SM), the row-segment pattern (a [rows x K] matrix walked in
//   16-row x KS-column tiles vs contiguous tiles), and an L2 prefetch (cp.async.bulk.prefetch.L2) running ahead of the ring.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ring_bw2 ring_bw2.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// The matrix is [rows][K] bf16 (row pitch K*2 bytes).  Tile t = 16 rows; chunk c = KS columns.  CTA b handles tiles b, b+grid, ...
// pf_tiles: the producer issues L2 prefetches `pf_tiles` tiles ahead of the ring (0 = none).
template <int NSTAGE>
__global__ void __launch_bounds__(288) ring_kernel(const uint8_t* src, int tiles, int K, int KS, int pf_tiles, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_u32(smem), full0 = smem_u32(bars), empty0 = full0 + 8 * NSTAGE;
  const uint32_t row_bytes = KS * 2, pitch = row_bytes + 64, stage = 16 * pitch;
  const int nch = K / KS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t it = 0;
  if (warp == 8) {
    int pf_t = blockIdx.x;  // next tile to prefetch into L2
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      if (pf_tiles > 0) {
        while (pf_t < tiles && pf_t <= t + pf_tiles * (int)gridDim.x) {
          if (pf_t > t && lane < 16) {   // the whole 16 x K tile, row by row (K*2 bytes each, <= 16 KB ok in 4 KB pieces)
            const uint8_t* r = src + ((size_t)pf_t * 16 + lane) * (size_t)K * 2;
            for (uint32_t o = 0; o < (uint32_t)K * 2; o += 4096) l2_prefetch(r + o, (uint32_t)K * 2 - o < 4096 ? (uint32_t)K * 2 - o : 4096);
          }
          pf_t += gridDim.x;
        }
      }
      for (int c = 0; c < nch; ++c, ++it) {
        const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
        mbar_wait(empty0 + 8 * st, ph ^ 1);
        if (lane == 0) mbar_expect_tx(full0 + 8 * st, 16 * row_bytes);
        __syncwarp();
        if (lane < 16) bulk_g2s(ring + st * stage + lane * pitch, src + ((size_t)t * 16 + lane) * (size_t)K * 2 + (size_t)c * row_bytes, row_bytes, full0 + 8 * st);
      }
    }
    return;
  }
  unsigned acc = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    for (int c = 0; c < nch; ++c, ++it) {
      const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
      mbar_wait(full0 + 8 * st, ph);
      acc += *reinterpret_cast<volatile unsigned*>(smem + st * stage + (threadIdx.x * 16) % stage);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * st);
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int NSTAGE>
static float run(const uint8_t* src, int rows, int K, int KS, int ctas_per_sm, int pf_tiles, int grid_sms, unsigned* sink, uint8_t* flush, size_t flush_bytes) {
  const int smem = NSTAGE * 16 * (KS * 2 + 64);
  cudaFuncSetAttribute(ring_kernel<NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemsetAsync(flush, rep, flush_bytes);
    cudaEventRecord(a);
    ring_kernel<NSTAGE><<<grid_sms * ctas_per_sm, 288, smem>>>(src, rows / 16, K, KS, pf_tiles, sink);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const int K = 2048;
  const int rows = 65536;                       // 256 MiB matrix (a decode layer is ~120 MB of weights)
  const size_t bytes = (size_t)rows * K * 2, flush_bytes = (size_t)512 << 20;
  uint8_t *src, *flush; unsigned* sink;
  CHECK(cudaMalloc(&src, bytes)); CHECK(cudaMalloc(&flush, flush_bytes)); CHECK(cudaMalloc(&sink, 64));
  CHECK(cudaMemset(src, 1, bytes));
  cudaError_t e;
  printf("matrix %d x %d bf16 = %.0f MiB, 148 SMs unless noted\n", rows, K, bytes / 1048576.0);
  printf("%-58s %8s %8s\n", "config", "ms", "GB/s");
#define RUN(NS, KS, CPS, PF, SMS, label) { float ms = run<NS>(src, rows, K, KS, CPS, PF, SMS, sink, flush, flush_bytes); \
    e = cudaGetLastError(); printf("%-58s %8.3f %8.0f %s\n", label, ms, bytes / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e)); }
  RUN(3, 1024, 1, 0, 148, "3 x 33 KB, 1 CTA/SM              ( 99 KB/SM in flight)");
  RUN(3, 1024, 2, 0, 148, "3 x 33 KB, 2 CTA/SM              (198 KB/SM)");
  RUN(6, 1024, 1, 0, 148, "6 x 33 KB, 1 CTA/SM              (198 KB/SM)");
  RUN(3, 2048, 1, 0, 148, "3 x 66 KB, 1 CTA/SM              (198 KB/SM, whole rows)");
  RUN(2, 2048, 1, 0, 148, "2 x 66 KB, 1 CTA/SM              (132 KB/SM, the mega kernel's ring)");
  RUN(3, 1024, 1, 1, 148, "3 x 33 KB, 1 CTA/SM + L2 prefetch 1 tile ahead");
  RUN(3, 1024, 1, 2, 148, "3 x 33 KB, 1 CTA/SM + L2 prefetch 2 tiles ahead");
  RUN(3, 1024, 1, 4, 148, "3 x 33 KB, 1 CTA/SM + L2 prefetch 4 tiles ahead");
  RUN(3, 1024, 1, 8, 148, "3 x 33 KB, 1 CTA/SM + L2 prefetch 8 tiles ahead");
  RUN(2, 512, 1, 4, 148, "2 x 17 KB, 1 CTA/SM + L2 prefetch 4 tiles ahead");
  RUN(3, 1024, 1, 0, 32, " 32 SMs: 3 x 33 KB, 1 CTA/SM");
  RUN(3, 1024, 2, 0, 32, " 32 SMs: 3 x 33 KB, 2 CTA/SM");
  RUN(3, 1024, 2, 4, 32, " 32 SMs: 3 x 33 KB, 2 CTA/SM + L2 prefetch 4 tiles ahead");
  RUN(3, 2048, 1, 4, 32, " 32 SMs: 3 x 66 KB, 1 CTA/SM + L2 prefetch 4 tiles ahead");
  printf("OK\n");
  return 0;
}
