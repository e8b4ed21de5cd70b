// Host-side runtime shared by every kernel file: error string, SM count, TMA descriptor encode.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void gar_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* gar_last_error(void) { return g_err; }
extern "C" int gar_abi_version(void) { return 1; }

// plain device-to-device copy on a stream (parity taps read engine-internal buffers with it)
extern "C" int gar_copy_d2d(void* dst, const void* src, long long bytes, void* stream) {
  GAR_CUDA_CHECK(cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
  return GAR_OK;
}

static int g_pdl = 1;
static int g_decode_attn_mma = 1;
static int g_gemm_2cta = 1;
static int g_attn_poly = 1;  // share of the softmax exponentials evaluated on the FMA pipe: 0 none, 1 2/8 (fastest measured), 2 3/8, 3 4/8
static int g_gemv_wide = 1;          // long-K decode GEMVs (down_proj) run 16 K-slices per CTA (A/B switch)
static int g_gemv_norm_smem = 1;     // decode GEMV stages RMSNorm weights / statistics through shared memory (A/B switch)
static int g_attn_persist = 1;  // prefill / ViT attention as a persistent kernel with a device-side tile counter (A/B switch)
static int g_gemm_group_kb = 40960;  // tcgen05 GEMM raster: L2 working-set budget (KB) that sizes the column-panel groups; 0 = one group
static int g_gemm_group_force = 0;   // test hook: panels per raster group (0 = sized from gemm_group_kb)
static int g_attn_bkv = 128;  // measured: 128 (2 CTAs/SM) beats 64 (3 CTAs/SM) on both ViT and prefill shapes
int gar_pdl_enabled() { return g_pdl; }
void gar_set_pdl(int on) { g_pdl = on; }
int gar_global_option(const char* name) {
  if (!strcmp(name, "decode_attn_mma")) return g_decode_attn_mma;
  if (!strcmp(name, "pdl")) return g_pdl;
  if (!strcmp(name, "attn_bkv")) return g_attn_bkv;
  if (!strcmp(name, "attn_poly")) return g_attn_poly;
  if (!strcmp(name, "attn_persist")) return g_attn_persist;
  if (!strcmp(name, "gemm_group_kb")) return g_gemm_group_kb;
  if (!strcmp(name, "gemm_group_force")) return g_gemm_group_force;
  if (!strcmp(name, "gemm_2cta")) return g_gemm_2cta;
  if (!strcmp(name, "gemv_norm_smem")) return g_gemv_norm_smem;
  if (!strcmp(name, "gemv_wide")) return g_gemv_wide;
  return 0;
}
// process-wide kernel-selection switches (A/B comparisons in tests and bench): "pdl", "decode_attn_mma"
extern "C" int gar_set_global_option(const char* name, int value) {
  if (!name) return GAR_ERR_ARG;
  if (!strcmp(name, "decode_attn_mma")) g_decode_attn_mma = value;
  else if (!strcmp(name, "pdl")) g_pdl = value;
  else if (!strcmp(name, "gemm_2cta")) g_gemm_2cta = value;
  else if (!strcmp(name, "gemv_norm_smem")) g_gemv_norm_smem = value;
  else if (!strcmp(name, "gemv_wide")) g_gemv_wide = value;
  else if (!strcmp(name, "attn_bkv") && (value == 64 || value == 128)) g_attn_bkv = value;
  else if (!strcmp(name, "attn_poly") && value >= 0 && value <= 3) g_attn_poly = value;
  else if (!strcmp(name, "attn_persist")) g_attn_persist = value != 0;
  else if (!strcmp(name, "gemm_group_kb") && value >= 0) g_gemm_group_kb = value;
  else if (!strcmp(name, "gemm_group_force") && value >= 0) g_gemm_group_force = value;
  else {
    gar_set_error("gar_set_global_option: unknown option '%s'", name);
    return GAR_ERR_ARG;
  }
  return GAR_OK;
}

int gar_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

// SM budget of the calling thread's launches: persistent kernels size their grids with gar_grid_sms() so that a launch
// into an SM partition (green context, gar_engine_partition) does not queue CTAs the partition cannot hold.  Kernel
// VARIANT choices keep using the physical count (gar_num_sms), so results never depend on the partitioning.
static thread_local int g_sm_limit = 0;
void gar_set_sm_limit(int sms) { g_sm_limit = sms; }
int gar_grid_sms() {
  const int phys = gar_num_sms();
  return (g_sm_limit > 0 && g_sm_limit < phys) ? g_sm_limit : phys;
}

// ---- SM partitions (CUDA green contexts) through runtime-resolved driver entry points ----
namespace {
template <typename Fn>
Fn driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<Fn>(p);
}
}  // namespace

// Splits the device's SMs into a small partition of >= small_sms SMs and the rest, and creates one non-blocking stream
// in each.  Returns GAR_ERR_CUDA (with the driver's message) when the driver cannot do it; callers then run unpartitioned.
int gar_make_sm_partition(int small_sms, cudaStream_t* small_stream, int* small_count, cudaStream_t* rest_stream,
                          int* rest_count) {
  typedef CUresult (*GetDev)(CUdevice*, int);
  typedef CUresult (*GetRes)(CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*Split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*GenDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*GreenCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*GreenStream)(CUstream*, CUgreenCtx, unsigned int, int);
  auto get_dev = driver_fn<GetDev>("cuDeviceGet");
  auto get_res = driver_fn<GetRes>("cuDeviceGetDevResource");
  auto split = driver_fn<Split>("cuDevSmResourceSplitByCount");
  auto gen = driver_fn<GenDesc>("cuDevResourceGenerateDesc");
  auto create = driver_fn<GreenCreate>("cuGreenCtxCreate");
  auto mkstream = driver_fn<GreenStream>("cuGreenCtxStreamCreate");
  if (!get_dev || !get_res || !split || !gen || !create || !mkstream) {
    gar_set_error("gar_make_sm_partition: this driver has no green-context entry points");
    return GAR_ERR_CUDA;
  }
  int ordinal = 0;
  GAR_CUDA_CHECK(cudaGetDevice(&ordinal));
  GAR_CUDA_CHECK(cudaFree(nullptr));  // make sure the primary context exists
  CUdevice dev;
  CUdevResource all, small, rest;
  unsigned int groups = 1;
  CUdevResourceDesc d_small, d_rest;
  CUgreenCtx g_small, g_rest;
  CUstream s_small, s_rest;
  CUresult r;
#define GAR_DRV(call)                                                                 \
  if ((r = (call)) != CUDA_SUCCESS) {                                                 \
    gar_set_error("gar_make_sm_partition: %s failed with CUresult %d", #call, (int)r); \
    return GAR_ERR_CUDA;                                                              \
  }
  GAR_DRV(get_dev(&dev, ordinal));
  GAR_DRV(get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  GAR_DRV(split(&small, &groups, &all, &rest, 0, static_cast<unsigned int>(small_sms)));
  if (groups != 1 || rest.sm.smCount == 0) {
    gar_set_error("gar_make_sm_partition: cannot split %u SMs into %d + rest", all.sm.smCount, small_sms);
    return GAR_ERR_CUDA;
  }
  GAR_DRV(gen(&d_small, &small, 1));
  GAR_DRV(gen(&d_rest, &rest, 1));
  GAR_DRV(create(&g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  GAR_DRV(create(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  GAR_DRV(mkstream(&s_small, g_small, CU_STREAM_NON_BLOCKING, 0));
  GAR_DRV(mkstream(&s_rest, g_rest, CU_STREAM_NON_BLOCKING, 0));
#undef GAR_DRV
  *small_stream = s_small;
  *rest_stream = s_rest;
  *small_count = static_cast<int>(small.sm.smCount);
  *rest_count = static_cast<int>(rest.sm.smCount);
  return GAR_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static CUtensorMapSwizzle to_cu_swizzle(GarSwizzle sw) {
  switch (sw) {
    case GAR_SW_32: return CU_TENSOR_MAP_SWIZZLE_32B;
    case GAR_SW_64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case GAR_SW_128: return CU_TENSOR_MAP_SWIZZLE_128B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int gar_make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                     uint32_t box_inner, uint32_t box_outer, GarSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    gar_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return GAR_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu_swizzle(sw), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gar_set_error("cuTensorMapEncodeTiled(2d) failed: %d (base=%p dims=%llu,%llu pitch=%llu box=%u,%u)", (int)r, base,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner,
                  box_outer);
    return GAR_ERR_CUDA;
  }
  return GAR_OK;
}

int gar_make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t p1,
                     uint64_t p2, uint32_t b0, uint32_t b1, uint32_t b2, GarSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    gar_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return GAR_ERR_CUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {p1 * 2, p2 * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu_swizzle(sw), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gar_set_error("cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
    return GAR_ERR_CUDA;
  }
  return GAR_OK;
}
