// Fused "assemble inputs_embeds + RoI-aligned feature replay" (SURVEY.md §2b K7+K8, rows R6+R7).
//
// Replaces, in the reference (hf_models/modeling_gar.py):
//   :332      embed_tokens(input_ids)                                  (index_select)
//   :341-346  get_placeholder_mask + masked_scatter of image_features  (bool mask S*D + scatter)
//   :348-414  per crop token: `in` scan + 2 .item() syncs, _merge permute+contiguous, .float() of the
//             whole map, torchvision.ops.roi_align, permute/cast, 3-way torch.cat
// with three launches for a whole batch of requests and no host synchronisation on the data path:
//   plan_rows    : one block per request scans input_ids -> per-row source plan (text / image ordinal / crop bin),
//                  the start of every region's crop run, and a status word (counts and run shapes validated).
//                  Depends only on input_ids, so the engine runs it BEFORE the ViT and the host reads the status early.
//   replay_rows  : one warp per crop row (grid = bins x regions x requests): roi_align for its (ph,pw) bin evaluated
//                  directly on the channel-contiguous (T,256,D) feature tensor (no merged map, no fp32 copy) in the
//                  operator's exact fp32 operation order (oracle/roi_align_ref.py), one bf16 rounding; latency-bound.
//   (replay_mode selects which reference loop is reproduced: GAR_REPLAY_MERGED = GARModel.generate; the two replay
//    branches of the vendored PLM's prepare_inputs_for_generation, modeling_perception_lm.py:648-852, are
//    GAR_REPLAY_MERGED_LAST -- same map, but only the last crop token present keeps its replay, the reference splices
//    every replay into the ORIGINAL embeddings -- and GAR_REPLAY_FRAME -- region k replays from tile k's own map.)
//   gather_rows  : one warp per text / image row of every request: 16-byte vector copies with all of a lane's loads in
//                  flight before its first store -- the HBM stream.  It starts together with replay_rows (programmatic
//                  dependent launch, dependency wait at its END): the two write disjoint rows.
// HBM-bound: ~2*S*D*2 bytes per request; no tensor cores on this path.
#include "../../include/gar_b200.h"
#include "common.cuh"

namespace {

constexpr int MAX_REGIONS = 8;
constexpr int KIND_TEXT = 0, KIND_IMAGE = 1, KIND_CROP = 2;

constexpr int REPLAY_PARTS = 1;  // warps per RoI replay row (4 measured slower: the per-warp sample set-up dominates)
constexpr int MAX_REQS = 8;  // requests per launch (kernel-parameter space); larger batches are chunked

struct RequestParams {
  const long long* ids;          // (S,) int64 device
  float rois[MAX_REGIONS][5];    // [0, x1, y1, x2, y2] already in "feature coords" (reference quirk 1)
  int has_roi[MAX_REGIONS];      // bbox supplied for this crop id
  int ncw, nch;                  // tile grid (w,h) of the map a region is replayed from
  int tile0[MAX_REGIONS];        // first tile of region k's map: 1 (thumbnail dropped) or, per-frame replay, k
};

struct AssembleParams {
  int n_regions;                 // number of crop token ids configured (prompt_numbers)
  int nreq;
  long long crop_ids[MAX_REGIONS];
  float spatial_scale;           // applied AGAIN inside roi_align, as the reference does
  int tpt;                       // tokens per tile side (16)
  long long image_token_id;
  int mode;                      // GAR_REPLAY_*
  RequestParams r[MAX_REQS];
};

// plan word: kind in bits [30,32), payload in [0,30)
__global__ void __launch_bounds__(1024) plan_rows_kernel(int S, int n_image_feats, const __grid_constant__ AssembleParams p,
                                 int* __restrict__ plan_all, int* __restrict__ status_all,
                                 int* __restrict__ region_start_all) {
  pdl_launch_dependents();
  const RequestParams& rq = p.r[blockIdx.x];
  const long long* __restrict__ ids = rq.ids;
  int* __restrict__ plan = plan_all + static_cast<size_t>(blockIdx.x) * S;
  int* __restrict__ status = status_all + 2 * blockIdx.x;
  __shared__ int s_warp[32];
  __shared__ int s_min[MAX_REGIONS], s_max[MAX_REGIONS], s_cnt[MAX_REGIONS];
  __shared__ int s_total;
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (tid < MAX_REGIONS) { s_min[tid] = 0x7fffffff; s_max[tid] = -1; s_cnt[tid] = 0; }
  __syncthreads();
  const int per = (S + nthr - 1) / nthr;
  const int lo = tid * per, hi = min(S, lo + per);
  // both passes read the thread's ids eight at a time (independent loads in flight before the first use)
  constexpr int PU = 8;
  int cnt = 0;
  // crop-token positions: per-thread min / max / count per region, reduced by warp shuffles, then ONE shared-memory
  // atomic per warp and region (a 256-token run hitting one shared address with 768 atomics serialised for ~8 us)
  int lmin[MAX_REGIONS], lmax[MAX_REGIONS], lcnt[MAX_REGIONS];
#pragma unroll
  for (int k = 0; k < MAX_REGIONS; ++k) { lmin[k] = 0x7fffffff; lmax[k] = -1; lcnt[k] = 0; }
  for (int s0 = lo; s0 < hi; s0 += PU) {
    long long idr[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) idr[u] = s0 + u < hi ? ids[s0 + u] : -1;
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      const int s = s0 + u;
      if (s >= hi) break;
      const long long id = idr[u];
      if (id == p.image_token_id) ++cnt;
#pragma unroll
      for (int k = 0; k < MAX_REGIONS; ++k) {
        if (k < p.n_regions && id == p.crop_ids[k]) {
          lmin[k] = min(lmin[k], s);
          lmax[k] = max(lmax[k], s);
          ++lcnt[k];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MAX_REGIONS; ++k) {
    if (k >= p.n_regions) break;
    const int wc = __reduce_add_sync(0xffffffffu, lcnt[k]);
    if (wc > 0) {  // warp-uniform
      const int wmin = __reduce_min_sync(0xffffffffu, lmin[k]);
      const int wmax = __reduce_max_sync(0xffffffffu, lmax[k]);
      if ((tid & 31) == 0) {
        atomicMin(&s_min[k], wmin);
        atomicMax(&s_max[k], wmax);
        atomicAdd(&s_cnt[k], wc);
      }
    }
  }
  // block exclusive scan of cnt
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = tid < (nthr >> 5) ? s_warp[tid] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (tid >= o) wi += v;
    }
    s_warp[tid] = wi - w;  // exclusive prefix of warp totals
    if (tid == 31) s_total = wi;
  }
  __syncthreads();
  int ord = s_warp[tid >> 5] + incl - cnt;
  // GAR_REPLAY_MERGED_LAST: only the last crop token (in crop-id order) present in the prompt is replayed; the runs of
  // the others stay what embed_tokens made of them (KIND_TEXT rows)
  int only = -1;
  if (p.mode == GAR_REPLAY_MERGED_LAST)
    for (int k = 0; k < p.n_regions; ++k)
      if (s_cnt[k] > 0) only = k;
  for (int s0 = lo; s0 < hi; s0 += PU) {
    long long idr[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) idr[u] = s0 + u < hi ? ids[s0 + u] : -1;
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      const int s = s0 + u;
      if (s >= hi) break;
      const long long id = idr[u];
      int word = KIND_TEXT << 30;
      if (id == p.image_token_id) {
        word = (KIND_IMAGE << 30) | ord;
        ++ord;
      } else {
        for (int k = 0; k < p.n_regions; ++k) {
          if (id == p.crop_ids[k] && rq.has_roi[k] && (p.mode != GAR_REPLAY_MERGED_LAST || k == only)) {
            word = (KIND_CROP << 30) | (k << 16) | ((s - s_min[k]) & 0xffff);
          }
        }
      }
      plan[s] = word;
    }
  }
  if (tid == 0) {
    int st = 0;
    if (s_total != n_image_feats) st |= 1;  // reference: ValueError "Image features and image tokens do not match"
    const int run = p.tpt * p.tpt;
    for (int k = 0; k < p.n_regions; ++k) {
      if (s_cnt[k] > 0) {
        if (!rq.has_roi[k]) st |= 4;                                         // reference: KeyError on bboxes[str(id)]
        if (s_cnt[k] != run || s_max[k] - s_min[k] + 1 != run) st |= 2;      // replay would change the length
      }
    }
    status[0] = st;
    status[1] = s_total;
  }
  // first position of every region's crop-token run (-1: absent, no box, or not a clean 256-run that fits the prompt):
  // the replay kernel's row index
  if (tid < MAX_REGIONS) {
    const int run = p.tpt * p.tpt;
    const bool ok = tid < p.n_regions && s_cnt[tid] > 0 && rq.has_roi[tid] && s_min[tid] + run <= S &&
                    (p.mode != GAR_REPLAY_MERGED_LAST || tid == only);
    region_start_all[blockIdx.x * MAX_REGIONS + tid] = ok ? s_min[tid] : -1;
  }
}

struct Sample {
  int pos[4];
  float w[4];
  int valid;
};

// torchvision bilinear pre-calc for one sample point; explicit _rn intrinsics forbid FMA contraction
__device__ __forceinline__ Sample make_sample(float y, float x, int H, int W) {
  Sample sm;
  sm.valid = !(y < -1.0f || y > static_cast<float>(H) || x < -1.0f || x > static_cast<float>(W));
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = static_cast<int>(y), x_low = static_cast<int>(x);
  int y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else { y_high = y_low + 1; }
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else { x_high = x_low + 1; }
  const float ly = __fsub_rn(y, static_cast<float>(y_low));
  const float lx = __fsub_rn(x, static_cast<float>(x_low));
  const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
  sm.w[0] = __fmul_rn(hy, hx);
  sm.w[1] = __fmul_rn(hy, lx);
  sm.w[2] = __fmul_rn(ly, hx);
  sm.w[3] = __fmul_rn(ly, lx);
  sm.pos[0] = y_low * W + x_low;
  sm.pos[1] = y_low * W + x_high;
  sm.pos[2] = y_high * W + x_low;
  sm.pos[3] = y_high * W + x_high;
  return sm;
}

// Text and image rows of every request: out row <- one table / feature row.  CH 16-byte vectors per lane are loaded
// before the first store, so a warp keeps CH*512 B in flight (12 resident blocks: ~100 KB per SM).  Launched as a
// programmatic dependent of replay_rows_kernel and waits for it only at its END: the two grids write disjoint rows, so
// the HBM stream runs next to the latency-bound replay instead of behind it, and whatever follows in the stream still
// sees both complete.  Every source index is clamped, so a malformed prompt (status bits set by plan_rows_kernel) can
// never make this kernel touch memory outside its request.
template <int CH>
__global__ void __launch_bounds__(128, 12)
gather_rows_kernel(const int* __restrict__ plan, const bf16* __restrict__ table, const bf16* __restrict__ feats,
                   bf16* __restrict__ out, int S, int D, int vocab, int rows_per_req_feats,
                   const __grid_constant__ AssembleParams p) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = D / 8;
  const int run = p.tpt * p.tpt;
  const long long total = static_cast<long long>(p.nreq) * S;
  for (long long g = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5); g < total;
       g += static_cast<long long>(gridDim.x) * wpb) {
    const int word = plan[g];
    int kind = (word >> 30) & 3;
    const int req = static_cast<int>(g / S), s = static_cast<int>(g - static_cast<long long>(req) * S);
    if (kind == KIND_CROP) {
      // replay_rows_kernel owns the rows of a clean run; anything else marked CROP (run too long / cut off: status
      // bit 2) is filled with its token embedding so no row of the output is ever left unwritten
      const int off = word & 0xffff, start = s - off;
      if (off < run && start + run <= S) continue;
      kind = KIND_TEXT;
    }
    const uint4* src;
    if (kind == KIND_IMAGE) {
      int ord = word & 0x3fffffff;
      ord = ord < rows_per_req_feats ? ord : rows_per_req_feats - 1;  // more <image> tokens than features (status bit 1)
      src = reinterpret_cast<const uint4*>(feats + (static_cast<size_t>(req) * rows_per_req_feats + ord) * D);
    } else {
      long long id = p.r[req].ids[s];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
      src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * D);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(g) * D);
    for (int v0 = 0; v0 < nvec; v0 += 32 * CH) {
      uint4 buf[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int v = v0 + c * 32 + lane;
        if (v < nvec) buf[c] = ld_nc_v4(src + v);
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int v = v0 + c * 32 + lane;
        if (v < nvec) dst[v] = buf[c];
      }
    }
  }
  pdl_wait();  // replay_rows_kernel (the programmatic predecessor) is complete before this grid is
}

// RoI replay rows: bin (ph, pw) of region k of request req.  One warp per crop row (the plan says which rows).
__global__ void __launch_bounds__(256)
replay_rows_kernel(const int* __restrict__ region_start, const bf16* __restrict__ feats_all, bf16* __restrict__ out, int S,
                   int D, int rows_per_req_feats, const __grid_constant__ AssembleParams p) {
  // grid = (ceil(tpt^2 / warps per block), regions, requests): warp <-> (request, region k, bin offset)
  pdl_launch_dependents();  // the gather kernel starts at once and streams next to this (latency-bound) grid
  const int req = blockIdx.z, k = blockIdx.y;
  const RequestParams& rq = p.r[req];
  if (!rq.has_roi[k]) return;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = D / 8;
  {
    // a crop row is shared by REPLAY_PARTS warps (each takes every REPLAY_PARTS-th 16-byte vector column): four times
    // the warps, a quarter of the dependent L2 round trips per warp -- the grid is latency-bound, not bandwidth-bound
    const int wid = blockIdx.x * wpb + (threadIdx.x >> 5);
    const int off = wid / REPLAY_PARTS, part = wid % REPLAY_PARTS;
    const int start = region_start[req * MAX_REGIONS + k];
    if (off >= p.tpt * p.tpt || start < 0 || start + off >= S) return;  // never outside this request's rows
    const long long g = static_cast<long long>(req) * S + start + off;
    const bf16* feats = feats_all + static_cast<size_t>(req) * rows_per_req_feats * D;
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(g) * D);
    const int tpt = p.tpt;
    const int ph = off / tpt, pw = off % tpt;
    const int H = rq.nch * tpt, W = rq.ncw * tpt;
    const float scale = p.spatial_scale;
    const float start_w = __fsub_rn(__fmul_rn(rq.rois[k][1], scale), 0.5f);
    const float start_h = __fsub_rn(__fmul_rn(rq.rois[k][2], scale), 0.5f);
    const float end_w = __fsub_rn(__fmul_rn(rq.rois[k][3], scale), 0.5f);
    const float end_h = __fsub_rn(__fmul_rn(rq.rois[k][4], scale), 0.5f);
    const float bin_h = __fdiv_rn(__fsub_rn(end_h, start_h), static_cast<float>(tpt));
    const float bin_w = __fdiv_rn(__fsub_rn(end_w, start_w), static_cast<float>(tpt));
    Sample sm[4];
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      const float yy = __fadd_rn(__fadd_rn(start_h, __fmul_rn(static_cast<float>(ph), bin_h)),
                                 __fdiv_rn(__fmul_rn(iy + 0.5f, bin_h), 2.0f));
#pragma unroll
      for (int ix = 0; ix < 2; ++ix) {
        const float xx = __fadd_rn(__fadd_rn(start_w, __fmul_rn(static_cast<float>(pw), bin_w)),
                                   __fdiv_rn(__fmul_rn(ix + 0.5f, bin_w), 2.0f));
        sm[iy * 2 + ix] = make_sample(yy, xx, H, W);
      }
    }
    // feature cell (y,x) of the region's map lives at tile tile0 + (y/tpt)*ncw + x/tpt, token (y%tpt)*tpt + x%tpt
    const uint4* cell[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int lin = sm[q].pos[c];
        const int y = lin / W, x = lin % W;
        const size_t row = (static_cast<size_t>(rq.tile0[k] + (y / tpt) * rq.ncw + x / tpt) * tpt + (y % tpt)) * tpt + (x % tpt);
        cell[q * 4 + c] = reinterpret_cast<const uint4*>(feats + row * D);
      }
    }
    for (int v = lane + 32 * part; v < nvec; v += 32 * REPLAY_PARTS) {
      // all 16 cell vectors of this output vector go out together (addresses of invalid samples are clamped, hence
      // loadable; their values are never used) -- one L2 round trip per output vector instead of four
      uint4 raws[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) raws[i] = __ldg(cell[i] + v);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!sm[q].valid) continue;
        float val[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 raw = raws[q * 4 + c];
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(rw[j]);
            const float t0 = __fmul_rn(sm[q].w[c], f.x), t1 = __fmul_rn(sm[q].w[c], f.y);
            val[2 * j] = c == 0 ? t0 : __fadd_rn(val[2 * j], t0);
            val[2 * j + 1] = c == 0 ? t1 : __fadd_rn(val[2 * j + 1], t1);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], val[j]);
      }
      uint4 o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        // output / count with count = 4: a division by a power of two is the same correctly rounded value as * 0.25
        ow[j] = pack_bf16x2(__fmul_rn(acc[2 * j], 0.25f), __fmul_rn(acc[2 * j + 1], 0.25f));
      dst[v] = o;
    }
  }
}

}  // namespace

namespace {
int fill_params(AssembleParams& p, int r0, int nr, const long long* const* input_ids, const long long* crop_ids,
                const float* rois, const int* has_roi, int n_regions, float spatial_scale, const int* ncw, const int* nch,
                int tpt, long long image_token_id, int mode) {
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.n_regions = n_regions;
  p.nreq = nr;
  for (int k = 0; k < MAX_REGIONS; ++k) p.crop_ids[k] = k < n_regions ? crop_ids[k] : -1;
  p.spatial_scale = spatial_scale;
  p.tpt = tpt;
  p.image_token_id = image_token_id;
  for (int r = 0; r < nr; ++r) {
    RequestParams& rq = p.r[r];
    rq.ids = input_ids[r0 + r];
    // per-frame replay: every region's map is ONE tile (its frame)
    rq.ncw = (ncw && mode != GAR_REPLAY_FRAME) ? ncw[r0 + r] : 1;
    rq.nch = (nch && mode != GAR_REPLAY_FRAME) ? nch[r0 + r] : 1;
    for (int k = 0; k < MAX_REGIONS; ++k) rq.tile0[k] = mode == GAR_REPLAY_FRAME ? k : 1;
    for (int k = 0; k < n_regions; ++k) {
      rq.has_roi[k] = has_roi[(r0 + r) * n_regions + k];
      for (int j = 0; j < 5; ++j) rq.rois[k][j] = rois ? rois[((r0 + r) * n_regions + k) * 5 + j] : 0.f;
    }
  }
  return GAR_OK;
}
}  // namespace

// Row plan + validation only (depends on nothing but input_ids: callers run it before the ViT and read the status
// word early, so a malformed prompt raises before any feature is computed, like the reference's ValueError).
extern "C" int gar_assemble_plan_batch(int nreq, const long long* const* input_ids, int S, int n_image_rows,
                                       const long long* crop_ids, const int* has_roi, int n_regions, int tpt,
                                       long long image_token_id, int replay_mode, int* plan_ws, int* status_dev,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_regions > MAX_REGIONS || S <= 0 || nreq <= 0 || tpt * tpt > 0xffff || replay_mode < 0 || replay_mode > 2) {
    gar_set_error("gar_assemble_plan_batch: bad args nreq=%d S=%d n_regions=%d", nreq, S, n_regions);
    return GAR_ERR_ARG;
  }
  for (int r0 = 0; r0 < nreq; r0 += MAX_REQS) {
    const int nr = nreq - r0 < MAX_REQS ? nreq - r0 : MAX_REQS;
    AssembleParams p;
    fill_params(p, r0, nr, input_ids, crop_ids, nullptr, has_roi, n_regions, 1.f, nullptr, nullptr, tpt, image_token_id,
                replay_mode);
    plan_rows_kernel<<<nr, 1024, 0, stream>>>(S, n_image_rows, p, plan_ws + static_cast<size_t>(r0) * S, status_dev + 2 * r0,
                                              plan_ws + static_cast<size_t>(nreq) * S + r0 * MAX_REGIONS);
    GAR_CUDA_CHECK(cudaGetLastError());
  }
  return GAR_OK;
}

// The HBM stream: every output row of every request in ONE launch (text / image gather rows and RoI replay rows by
// warp role), reading the plan written by gar_assemble_plan_batch.
extern "C" int gar_assemble_rows_batch(int nreq, const long long* const* input_ids, int S, const void* embed_table,
                                       int vocab, const void* image_features, int n_image_rows, int D,
                                       const long long* crop_ids, const float* rois, const int* has_roi, int n_regions,
                                       float spatial_scale, const int* ncw, const int* nch, int tpt,
                                       long long image_token_id, int replay_mode, void* out, const int* plan_ws,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_regions > MAX_REGIONS || D % 8 || S <= 0 || nreq <= 0 || tpt * tpt > 0xffff || n_image_rows <= 0 ||
      replay_mode < 0 || replay_mode > 2) {
    gar_set_error("gar_assemble_rows_batch: bad args nreq=%d S=%d D=%d n_regions=%d", nreq, S, D, n_regions);
    return GAR_ERR_ARG;
  }
  for (int r0 = 0; r0 < nreq; r0 += MAX_REQS) {
    const int nr = nreq - r0 < MAX_REQS ? nreq - r0 : MAX_REQS;
    AssembleParams p;
    fill_params(p, r0, nr, input_ids, crop_ids, rois, has_roi, n_regions, spatial_scale, ncw, nch, tpt, image_token_id,
                replay_mode);
    for (int r = 0; r < nr; ++r) {
      // every map cell a replay can address must exist: tiles tile0 .. tile0 + ncw*nch - 1 of this request's features
      for (int k = 0; k < n_regions; ++k) {
        const int need = (p.r[r].tile0[k] + p.r[r].ncw * p.r[r].nch) * tpt * tpt;
        if (p.r[r].has_roi[k] && need > n_image_rows) {
          gar_set_error("gar_assemble_rows_batch: region %d (grid %dx%d from tile %d) needs %d feature rows, got %d", k,
                        p.r[r].ncw, p.r[r].nch, p.r[r].tile0[k], need, n_image_rows);
          return GAR_ERR_SHAPE;
        }
      }
    }
    const bf16* feats = static_cast<const bf16*>(image_features) + static_cast<size_t>(r0) * n_image_rows * D;
    bf16* o = static_cast<bf16*>(out) + static_cast<size_t>(r0) * S * D;
    const int* region_start = plan_ws + static_cast<size_t>(nreq) * S + r0 * MAX_REGIONS;
    const bool replay = n_regions > 0;
    if (replay) {
      replay_rows_kernel<<<dim3((tpt * tpt * REPLAY_PARTS + 7) / 8, n_regions, nr), 256, 0, stream>>>(region_start, feats, o, S, D,
                                                                                       n_image_rows, p);
      GAR_CUDA_CHECK(cudaGetLastError());
    }
    const long long rows = static_cast<long long>(nr) * S;
    const int wpb = 4;
    // one row per warp; 4 vectors per lane in flight (2 passes over a 2048-wide row), 12 blocks/SM resident
    const long long gblocks = (rows + wpb - 1) / wpb;
    GAR_CUDA_CHECK(gar_launch(gather_rows_kernel<4>, dim3(static_cast<unsigned>(gblocks)), dim3(wpb * 32), 0, stream,
                              replay, dim3(1, 1, 1), plan_ws + static_cast<size_t>(r0) * S,
                              static_cast<const bf16*>(embed_table), feats, o, S, D, vocab, n_image_rows, p));
  }
  return GAR_OK;
}

// Batched form: request r reads input_ids[r] (device pointers in a HOST array), features
// image_features + r*n_image_rows*D, boxes rois[r][n_regions][5] / has_roi[r][n_regions] (host), tile grid
// ncw[r] x nch[r]; writes rows [r*S, (r+1)*S) of out / plan_ws and status_dev[2r..2r+1].  plan_ws: nreq*S + 8*nreq ints.
extern "C" int gar_assemble_embeds_batch(int nreq, const long long* const* input_ids, int S, const void* embed_table,
                                         int vocab, const void* image_features, int n_image_rows, int D,
                                         const long long* crop_ids, const float* rois, const int* has_roi,
                                         int n_regions, float spatial_scale, const int* ncw, const int* nch, int tpt,
                                         long long image_token_id, int replay_mode, void* out, int* plan_ws,
                                         int* status_dev, void* stream_) {
  int rc = gar_assemble_plan_batch(nreq, input_ids, S, n_image_rows, crop_ids, has_roi, n_regions, tpt, image_token_id,
                                   replay_mode, plan_ws, status_dev, stream_);
  if (rc) return rc;
  return gar_assemble_rows_batch(nreq, input_ids, S, embed_table, vocab, image_features, n_image_rows, D, crop_ids, rois,
                                 has_roi, n_regions, spatial_scale, ncw, nch, tpt, image_token_id, replay_mode, out,
                                 plan_ws, stream_);
}

// rois: n_regions x 5 fp32 host array; has_roi: n_regions ints; crop_ids: n_regions int64 (host)
extern "C" int gar_assemble_embeds(const long long* input_ids, int S, const void* embed_table, int vocab,
                                   const void* image_features, int n_image_rows, int D, const long long* crop_ids,
                                   const float* rois, const int* has_roi, int n_regions, float spatial_scale, int ncw,
                                   int nch, int tpt, long long image_token_id, int replay_mode, void* out, int* plan_ws,
                                   int* status_dev, void* stream_) {
  return gar_assemble_embeds_batch(1, &input_ids, S, embed_table, vocab, image_features, n_image_rows, D, crop_ids, rois,
                                   has_roi, n_regions, spatial_scale, &ncw, &nch, tpt, image_token_id, replay_mode, out,
                                   plan_ws, status_dev, stream_);
}
