// Probe: can this driver partition the B200's SMs with green contexts, do runtime-API launches and CUDA graphs honour
// the partition, and do a bandwidth-bound kernel (small partition) and a compute-bound kernel (large partition) really
// overlap?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o green_ctx_probe green_ctx_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <set>
#include <vector>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, s_); return 1; } } while (0)
#define CR(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(r_)); return 1; } } while (0)

__global__ void smid_kernel(int* out, int spin) {
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  if (threadIdx.x == 0) out[blockIdx.x] = id;
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
}
__global__ void stream_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = src[i];
    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678u) dst[0] = acc;
}
__global__ void fma_kernel(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.0001f, c = 0.5f, d = 0.25f;
  for (int i = 0; i < iters; ++i) { a = fmaf(a, b, c); d = fmaf(d, b, a); c = fmaf(c, b, d); b = fmaf(b, 0.9999f, 1e-6f); }
  if (a + d + c == 123.f) out[0] = a;
}

static int distinct_sms(cudaStream_t s, int blocks, int* dbuf) {
  std::vector<int> h(blocks);
  smid_kernel<<<blocks, 64, 0, s>>>(dbuf, 200000);
  if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
  cudaMemcpy(h.data(), dbuf, blocks * sizeof(int), cudaMemcpyDeviceToHost);
  return (int)std::set<int>(h.begin(), h.end()).size();
}

int main(int argc, char** argv) {
  int want = argc > 1 ? atoi(argv[1]) : 32;
  CR(cudaSetDevice(0));
  CR(cudaFree(0));
  CUdevice dev;
  CK(cuDeviceGet(&dev, 0));
  CUdevResource all;
  CK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", all.sm.smCount);
  CUdevResource small, rest;
  unsigned n = 1;
  CK(cuDevSmResourceSplitByCount(&small, &n, &all, &rest, 0, want));
  printf("split(min %d): groups=%u small=%u rest=%u\n", want, n, small.sm.smCount, rest.sm.smCount);
  CUdevResourceDesc d_small, d_rest;
  CK(cuDevResourceGenerateDesc(&d_small, &small, 1));
  CK(cuDevResourceGenerateDesc(&d_rest, &rest, 1));
  CUgreenCtx g_small, g_rest;
  CK(cuGreenCtxCreate(&g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CK(cuGreenCtxCreate(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream s_small, s_rest;
  CK(cuGreenCtxStreamCreate(&s_small, g_small, CU_STREAM_NON_BLOCKING, 0));
  CK(cuGreenCtxStreamCreate(&s_rest, g_rest, CU_STREAM_NON_BLOCKING, 0));
  int* dbuf;
  CR(cudaMalloc(&dbuf, 4096 * sizeof(int)));
  cudaStream_t plain;
  CR(cudaStreamCreateWithFlags(&plain, cudaStreamNonBlocking));
  printf("distinct SMs seen: plain=%d small=%d rest=%d\n", distinct_sms(plain, 1024, dbuf), distinct_sms(s_small, 1024, dbuf),
         distinct_sms(s_rest, 1024, dbuf));
  // cooperative-style occupancy query inside the partition
  int dev_sms = 0;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  printf("cudaDevAttrMultiProcessorCount (unchanged by partitions): %d\n", dev_sms);

  // graph captured on the small-partition stream, replayed there
  cudaGraph_t graph; cudaGraphExec_t exec;
  CR(cudaStreamBeginCapture(s_small, cudaStreamCaptureModeThreadLocal));
  smid_kernel<<<1024, 64, 0, s_small>>>(dbuf, 200000);
  CR(cudaStreamEndCapture(s_small, &graph));
  CR(cudaGraphInstantiate(&exec, graph, 0));
  CR(cudaGraphLaunch(exec, s_small));
  CR(cudaStreamSynchronize(s_small));
  { std::vector<int> h(1024); cudaMemcpy(h.data(), dbuf, 4096, cudaMemcpyDeviceToHost);
    printf("graph replay on small stream: distinct SMs=%d\n", (int)std::set<int>(h.begin(), h.end()).size()); }
  CR(cudaGraphLaunch(exec, plain));
  CR(cudaStreamSynchronize(plain));
  { std::vector<int> h(1024); cudaMemcpy(h.data(), dbuf, 4096, cudaMemcpyDeviceToHost);
    printf("same graph replayed on a plain stream: distinct SMs=%d\n", (int)std::set<int>(h.begin(), h.end()).size()); }

  // overlap: HBM stream on the small partition vs FMA burn on the rest
  size_t bytes = (size_t)4 << 30;
  uint4 *src, *dst;
  CR(cudaMalloc(&src, bytes)); CR(cudaMalloc(&dst, 64));
  CR(cudaMemset(src, 1, bytes));
  float* fo; CR(cudaMalloc(&fo, 64));
  cudaEvent_t e0, e1, f0, f1;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
  auto run = [&](cudaStream_t sa, int blocks_a, cudaStream_t sb, int blocks_b, bool both, float* ta, float* tb) {
    cudaDeviceSynchronize();
    cudaEventRecord(e0, sa);
    stream_kernel<<<blocks_a, 512, 0, sa>>>(src, dst, bytes / 16);
    cudaEventRecord(e1, sa);
    if (both) {
      cudaEventRecord(f0, sb);
      fma_kernel<<<blocks_b, 1024, 0, sb>>>(fo, 400000);
      cudaEventRecord(f1, sb);
    }
    cudaDeviceSynchronize();
    cudaEventElapsedTime(ta, e0, e1);
    if (both) cudaEventElapsedTime(tb, f0, f1); else *tb = 0;
  };
  float ta, tb;
  run(plain, 148 * 4, plain, 0, false, &ta, &tb);
  printf("stream 4 GiB, all SMs alone            : %.3f ms = %.0f GB/s\n", ta, bytes / ta / 1e6);
  for (int blocks_per_sm = 2; blocks_per_sm <= 4; blocks_per_sm += 2) {
    run(s_small, small.sm.smCount * blocks_per_sm, s_rest, 0, false, &ta, &tb);
    printf("stream 4 GiB, %u-SM partition alone (%d blk/SM): %.3f ms = %.0f GB/s\n", small.sm.smCount, blocks_per_sm, ta, bytes / ta / 1e6);
  }
  cudaDeviceSynchronize();
  cudaEventRecord(f0, s_rest); fma_kernel<<<rest.sm.smCount * 2, 1024, 0, s_rest>>>(fo, 400000); cudaEventRecord(f1, s_rest);
  cudaDeviceSynchronize(); cudaEventElapsedTime(&tb, f0, f1);
  printf("fma burn on the %u-SM partition alone    : %.3f ms\n", rest.sm.smCount, tb);
  run(s_small, small.sm.smCount * 4, s_rest, rest.sm.smCount * 2, true, &ta, &tb);
  printf("both concurrently                       : stream %.3f ms = %.0f GB/s, fma %.3f ms\n", ta, bytes / ta / 1e6, tb);
  printf("OK\n");
  return 0;
}
