// Request preprocessing on the device: uint8 image + uint8 region-id map in, the bf16
// `pixel_values` / `global_mask_values` tile tensors of GARModel.generate() out -- so a request crosses PCIe as
// ~4 MB of uint8 instead of ~41 MB of bf16 tiles.
//
// Replaces, bit for bit, what the reference's fast image processor does on the host
// (models/modeling/image_processing_perception_lm_fast.py:254-372: torchvision v2 `resize(..., BICUBIC)` of the uint8
// image onto the thumbnail and onto the tile canvas, `_split`, rescale 1/255, normalise mean = std = 0.5) and what its
// request builders do to the id map (evaluation/eval_dataset.py:56-99 + the hub processor's NEAREST resize,
// L -> RGB replication, 2 v / 255 - 1):
//   * torchvision resizes uint8 CPU tensors with ATen's native uint8 antialiased bicubic kernel: separable, horizontal
//     pass first, per-axis int16 fixed-point weights (cubic convolution a = -0.5, support 2 * max(scale, 1), weights
//     normalised then scaled by 2^precision), out = clamp((2^(precision-1) + sum w_j src_j) >> precision, 0, 255) with a
//     uint8 intermediate between the passes.  The weight tables are built on the host in double precision
//     (preprocess_device.py restates ATen's _compute_index_ranges_int16_weights); these kernels do the integer passes.
//   * the float steps reproduce torch's op order with explicit round-to-nearest intrinsics (no FMA contraction):
//     pixel: ((v * fl(1/255)) - 0.5) / 0.5 ; id map: ((2 * v) / 255) - 1 ; one bf16 rounding each.
// HBM-bound byte work, no tensor cores: one thread per output element, coalesced along x.
#include "../../include/gar_b200.h"
#include "common.cuh"

namespace {

// one separable pass over `planes` planes of (in_h, in_w) uint8: horizontal resizes x (out (in_h, out_size)),
// vertical resizes y (out (out_size, in_w))
__global__ void resize_aa_u8_pass_kernel(const uint8_t* __restrict__ src, int planes, int in_h, int in_w, int horizontal,
                                         int out_size, const int* __restrict__ idx_min, const int* __restrict__ idx_size,
                                         const short* __restrict__ weights, int max_interp, int precision,
                                         uint8_t* __restrict__ dst) {
  const int out_h = horizontal ? in_h : out_size, out_w = horizontal ? out_size : in_w;
  const long long total = static_cast<long long>(planes) * out_h * out_w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % out_w);
    const long long r = i / out_w;
    const int y = static_cast<int>(r % out_h), p = static_cast<int>(r / out_h);
    const int o = horizontal ? x : y;
    const int lo = __ldg(idx_min + o), n = __ldg(idx_size + o);
    const short* w = weights + static_cast<size_t>(o) * max_interp;
    const uint8_t* s = src + static_cast<size_t>(p) * in_h * in_w + (horizontal ? static_cast<size_t>(y) * in_w + lo
                                                                                : static_cast<size_t>(lo) * in_w + x);
    const int step = horizontal ? 1 : in_w;
    int acc = 1 << (precision - 1);
    for (int j = 0; j < n; ++j) acc += static_cast<int>(s[static_cast<size_t>(j) * step]) * static_cast<int>(__ldg(w + j));
    acc >>= precision;
    dst[i] = static_cast<uint8_t>(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

// canvas (3, nch*tile, ncw*tile) uint8 -> tiles (nch*ncw, 3, tile, tile) bf16, row-major over (tile row, tile column)
__global__ void u8_to_tiles_kernel(const uint8_t* __restrict__ canvas, int ncw, int nch, int tile, bf16* __restrict__ out) {
  const long long total = static_cast<long long>(nch) * ncw * 3 * tile * tile;
  const int W = ncw * tile, H = nch * tile;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % tile);
    long long r = i / tile;
    const int y = static_cast<int>(r % tile);
    r /= tile;
    const int c = static_cast<int>(r % 3), t = static_cast<int>(r / 3);
    const int ty = t / ncw, tx = t % ncw;
    const float v = static_cast<float>(canvas[(static_cast<size_t>(c) * H + ty * tile + y) * W + tx * tile + x]);
    // torch: (tiles.float() * (1.0 / 255.0) - 0.5) / 0.5 with the scalar rounded to fp32
    const float n = __fdiv_rn(__fsub_rn(__fmul_rn(v, 0.003921568859368563f), 0.5f), 0.5f);
    out[i] = __float2bfloat16_rn(n);
  }
}

// id map (H, W) uint8 -> (1 + ncw*nch, 3, tile, tile) bf16 = 2 v / 255 - 1, NEAREST through host-built index tables
// (thumbnail: ys_t/xs_t of length tile; canvas: ys_c of length nch*tile, xs_c of length ncw*tile), 3 identical channels
__global__ void idmap_to_tiles_kernel(const uint8_t* __restrict__ idmap, int W, const int* __restrict__ ys_t,
                                      const int* __restrict__ xs_t, const int* __restrict__ ys_c,
                                      const int* __restrict__ xs_c, int ncw, int nch, int tile, bf16* __restrict__ out) {
  const int T = 1 + ncw * nch;
  const long long total = static_cast<long long>(T) * tile * tile;
  const size_t plane = static_cast<size_t>(tile) * tile;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % tile);
    const long long r = i / tile;
    const int y = static_cast<int>(r % tile), t = static_cast<int>(r / tile);
    int sy, sx;
    if (t == 0) {
      sy = __ldg(ys_t + y);
      sx = __ldg(xs_t + x);
    } else {
      const int ty = (t - 1) / ncw, tx = (t - 1) % ncw;
      sy = __ldg(ys_c + ty * tile + y);
      sx = __ldg(xs_c + tx * tile + x);
    }
    const float v = static_cast<float>(idmap[static_cast<size_t>(sy) * W + sx]);
    const bf16 b = __float2bfloat16_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, v), 255.0f), 1.0f));
    bf16* o = out + static_cast<size_t>(t) * 3 * plane + static_cast<size_t>(y) * tile + x;
    o[0] = b;
    o[plane] = b;
    o[2 * plane] = b;
  }
}

inline int blocks_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = static_cast<long long>(gar_num_sms()) * 32;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int gar_resize_aa_u8_pass(const void* src, int planes, int in_h, int in_w, int horizontal, int out_size,
                                     const int* idx_min, const int* idx_size, const short* weights, int max_interp,
                                     int precision, void* dst, void* stream) {
  if (!src || !dst || planes <= 0 || in_h <= 0 || in_w <= 0 || out_size <= 0 || max_interp <= 0 || precision < 1 ||
      precision > 22 || !idx_min || !idx_size || !weights) {
    gar_set_error("gar_resize_aa_u8_pass: bad arguments (planes=%d in=%dx%d out=%d taps=%d precision=%d)", planes, in_h,
                  in_w, out_size, max_interp, precision);
    return GAR_ERR_ARG;
  }
  const long long total = static_cast<long long>(planes) * (horizontal ? in_h : out_size) * (horizontal ? out_size : in_w);
  resize_aa_u8_pass_kernel<<<blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), planes, in_h, in_w, horizontal, out_size, idx_min, idx_size, weights, max_interp,
      precision, static_cast<uint8_t*>(dst));
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

extern "C" int gar_u8_to_tiles(const void* canvas, int ncw, int nch, int tile, void* tiles_out, void* stream) {
  if (!canvas || !tiles_out || ncw <= 0 || nch <= 0 || tile <= 0) return GAR_ERR_ARG;
  const long long total = static_cast<long long>(nch) * ncw * 3 * tile * tile;
  u8_to_tiles_kernel<<<blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(canvas), ncw, nch, tile, static_cast<bf16*>(tiles_out));
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

extern "C" int gar_idmap_to_tiles(const void* idmap, int H, int W, const int* ys_thumb, const int* xs_thumb,
                                  const int* ys_canvas, const int* xs_canvas, int ncw, int nch, int tile,
                                  void* mask_values_out, void* stream) {
  if (!idmap || !mask_values_out || H <= 0 || W <= 0 || ncw <= 0 || nch <= 0 || tile <= 0 || !ys_thumb || !xs_thumb ||
      !ys_canvas || !xs_canvas)
    return GAR_ERR_ARG;
  const long long total = static_cast<long long>(1 + ncw * nch) * tile * tile;
  idmap_to_tiles_kernel<<<blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(idmap), W, ys_thumb, xs_thumb, ys_canvas, xs_canvas, ncw, nch, tile,
      static_cast<bf16*>(mask_values_out));
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}
