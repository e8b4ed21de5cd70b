// Microbenchmark: how fast can one CTA per SM stream HBM through a 3-stage shared-memory ring with 1-D bulk TMA copies?
// Variants: rows of 4 KB into a padded pitch (decode_mega.cu's layout), rows into an unpadded pitch, one 64 KB copy,
// and plain ld.global.v4 by 256 threads (no TMA).  Prints GB/s for each.
#include <cstdio>
#include <cstdint>
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

constexpr int STAGES = 3, ROWS = 16, ROWB = 4096;

// mode 0: 16 rows x 4 KB, pitch 4160; mode 1: pitch 4096; mode 2: one 64 KB copy; each "tile" = 64 KB contiguous
template <int MODE, int NSTAGE>
__global__ void __launch_bounds__(288, 1) ring_kernel(const uint8_t* src, size_t tiles, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_u32(smem), full0 = smem_u32(bars), empty0 = full0 + 8 * NSTAGE;
  constexpr uint32_t PITCH = MODE == 0 ? ROWB + 64 : ROWB;
  constexpr uint32_t STAGE = ROWS * (ROWB + 64);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t it = 0;
  if (warp == 8) {
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
      mbar_wait(empty0 + 8 * st, ph ^ 1);
      if (lane == 0) mbar_expect_tx(full0 + 8 * st, ROWS * ROWB);
      __syncwarp();
      const uint8_t* base = src + t * (size_t)(ROWS * ROWB);
      if (MODE == 2) {
        if (lane == 0) bulk_g2s(ring + st * STAGE, base, ROWS * ROWB, full0 + 8 * st);
      } else if (lane < ROWS) {
        bulk_g2s(ring + st * STAGE + lane * PITCH, base + lane * ROWB, ROWB, full0 + 8 * st);
      }
    }
    return;
  }
  unsigned acc = 0;
  for (size_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
    const uint32_t st = it % NSTAGE, ph = (it / NSTAGE) & 1;
    mbar_wait(full0 + 8 * st, ph);
    acc += *reinterpret_cast<volatile unsigned*>(smem + st * STAGE + threadIdx.x * 16);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * st);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void __launch_bounds__(256, 2) ldg_kernel(const uint4* src, size_t nvec, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < nvec; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
  const size_t bytes = (size_t)2 << 30;  // 2 GiB > L2
  uint8_t* src; unsigned* sink;
  CHECK(cudaMalloc(&src, bytes)); CHECK(cudaMemset(src, 1, bytes)); CHECK(cudaMalloc(&sink, 4));
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t tiles = bytes / (ROWS * ROWB);
  const int smem = 3 * ROWS * (ROWB + 64);
  CHECK(cudaFuncSetAttribute(ring_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CHECK(cudaFuncSetAttribute(ring_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CHECK(cudaFuncSetAttribute(ring_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CHECK(cudaFuncSetAttribute(ring_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float ms;
  ms = time_ms([&] { ring_kernel<0, 3><<<sms, 288, smem>>>(src, tiles, sink); }, 3);
  printf("ring rows 4KB pitch 4160, 3 stages : %7.1f GB/s\n", bytes / ms / 1e6);
  ms = time_ms([&] { ring_kernel<1, 3><<<sms, 288, smem>>>(src, tiles, sink); }, 3);
  printf("ring rows 4KB pitch 4096, 3 stages : %7.1f GB/s\n", bytes / ms / 1e6);
  ms = time_ms([&] { ring_kernel<2, 3><<<sms, 288, smem>>>(src, tiles, sink); }, 3);
  printf("ring one 64KB copy,       3 stages : %7.1f GB/s\n", bytes / ms / 1e6);
  ms = time_ms([&] { ring_kernel<0, 2><<<sms, 288, smem>>>(src, tiles, sink); }, 3);
  printf("ring rows 4KB pitch 4160, 2 stages : %7.1f GB/s\n", bytes / ms / 1e6);
  ms = time_ms([&] { ldg_kernel<<<sms * 2, 256>>>(reinterpret_cast<const uint4*>(src), bytes / 16, sink); }, 3);
  printf("ld.global.v4 x8 per thread, 2 CTA/SM: %7.1f GB/s\n", bytes / ms / 1e6);
  CHECK(cudaGetLastError());
  return 0;
}
