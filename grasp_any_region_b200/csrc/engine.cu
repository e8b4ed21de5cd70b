// Device-resident pipeline for GARModel.generate() (reference hf_models/modeling_gar.py:295-428):
//   vit_encode  : R2-R5  mask-id decode, image+mask patch-embed GEMMs, PE-ViT blocks, projector, 2x2 pool
//   assemble    : R6-R7  embed + image scatter + RoI feature replay (one fused kernel pair)
//   prefill     : R8     causal Llama prefill into a pre-allocated KV-cache slot, last-position logits
//   generate    : R9     batched greedy decode, one CUDA graph per step, no host sync in the loop
// Requests are independent (the reference is batch-1): ViT + prefill run per request, decode is batched
// across the slots so the weight stream is shared.  All math is in the per-kernel entry points of
// gar_b200.h; this file only sequences them and owns the workspaces.
#include <map>
#include <string>
#include <vector>

#include "../../include/gar_b200.h"
#include "common.cuh"

extern "C" int gar_set_global_option(const char* name, int value);

#define TRY(x)                \
  do {                        \
    int _rc = (x);            \
    if (_rc) return _rc;      \
  } while (0)

struct gar_engine {
  gar_model_desc d;
  std::map<std::string, const void*> w;
  std::vector<void*> allocs;
  // geometry
  int grid, ntok, npt, hd_v, hdp_v, dq_v, kpad, tpt, tokens_per_tile;  // hdp_v: ViT head dim padded to 64/128
  // ViT workspaces
  bf16 *a_img, *a_mask, *pe, *x, *ln, *qkv_v, *att_v, *mlp_v, *p1, *p2, *feats;
  // LLM workspaces
  bf16 *emb, *h, *lnl, *qkv_l, *att_l, *gu, *act, *logits, *kcache, *vcache, *hdec, *lndec;
  // decode-private scratch: a decode group may run on its own stream while another group's ViT/prefill is in flight
  bf16 *d_qkv, *d_att, *d_gu, *d_act;
  int slot_base = 0;  // first KV slot of the group being decoded
  float *o_part, *ml_part, *argmax_ws;
  int *plan, *status, *next_tok, *out_tokens, *pos, *finished, *step, *tickets;
  long long* ids_scratch;
  int dec_splits;
  int use_skinny = 1;
  int fuse_rope = 1;
  int use_gemv = 1;
  int fuse_decode_rope = 1;
  float *sumsq_a, *sumsq_b;
  // replay loop reproduced by plan / assemble (gar_engine_set_replay): mode + the crop-token ids the boxes are indexed by
  int replay_mode = GAR_REPLAY_MERGED;
  long long replay_ids[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  int replay_n = 0;  // 0: the model's own crop_tokens_ids / prompt_numbers
  const long long* crop_ids() const { return replay_n > 0 ? replay_ids : d.crop_token_ids; }
  int n_regions() const { return replay_n > 0 ? replay_n : d.prompt_numbers; }
  // decode graph
  std::map<long long, cudaGraphExec_t> graphs;  // decode-step graphs keyed by (slot0, batch)
  std::map<long long, std::vector<int>> graph_eos;
  int eos_ids[4] = {-1, -1, -1, -1};  // stop ids of the decode in flight (GenerationConfig.eos_token_id may be a list)
  int n_eos = 0;
  int eos_poll_steps = 8;   // decode steps per early-stop poll (0 = never poll: always run max_new_tokens steps)
  int steps_run = 0;        // decode steps launched by the last generate call
  int* fin_host = nullptr;  // pinned: 2 x max_batch finished flags
  // Host copies go through pinned memory and event waits only: a pageable cudaMemcpyAsync is host-synchronous and was
  // measured to block the staging thread until the OTHER stream's decode had drained (pipelined serving).
  int* status_host = nullptr;  // pinned: 2 x max_batch status words
  int* out_host = nullptr;     // pinned: max_batch x max_new_tokens decoded ids
  cudaEvent_t plan_ev = nullptr, out_ev = nullptr;
  cudaEvent_t fin_ev[2] = {nullptr, nullptr};
  // SM partition (green contexts): launches into part_stream[i] may occupy part_sms[i] SMs (gar_engine_partition)
  cudaStream_t part_stream[2] = {nullptr, nullptr};
  int part_sms[2] = {0, 0};
  // plan_rows already ran for this (slot0, nreq, S, T) (gar_engine_plan_batch): assemble_batch skips it
  int plan_slot0 = -1, plan_nreq = 0, plan_S = 0, plan_T = 0;
  long long launches = 0;
  int launches_per_step = 0;
  // optional per-kernel-class timing (CUDA events on the launching stream; bench.py roofline)
  bool profiling = false;
  struct ProfRec { cudaEvent_t a, b; int kind; double work; };
  std::vector<ProfRec> prof;
};

enum ProfKind { PK_GEMM = 0, PK_ATTN = 1, PK_ASSEMBLE = 2, PK_DECODE_ATTN = 3, PK_GEMV = 4, PK_PLAN = 5, PK_COUNT = 6 };

namespace {

__global__ void fill_int_kernel(int* dst, int value, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = value;
}

template <typename T>
int dev_alloc(gar_engine* e, T** p, size_t count, bool zero = false) {
  void* q = nullptr;
  GAR_CUDA_CHECK(cudaMalloc(&q, count * sizeof(T)));
  if (zero) GAR_CUDA_CHECK(cudaMemset(q, 0, count * sizeof(T)));
  e->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return GAR_OK;
}

const void* W(gar_engine* e, const std::string& name, bool required = true) {
  auto it = e->w.find(name);
  if (it == e->w.end()) {
    if (required) gar_set_error("gar_engine: weight '%s' was never set", name.c_str());
    return nullptr;
  }
  return it->second;
}
#define NEEDW(var, name)                 \
  const void* var = W(e, name);          \
  if (!var) return GAR_ERR_STATE

std::string L(const char* prefix, int i, const char* suffix) {
  return std::string(prefix) + "." + std::to_string(i) + "." + suffix;
}

// Thread-local SM budget for the duration of one engine call: persistent kernels launched into a partition's stream
// size their grids to the partition (common.cuh gar_grid_sms).
struct SmScope {
  SmScope(gar_engine* e, void* s) {
    int lim = 0;
    for (int i = 0; i < 2; ++i)
      if (e && e->part_stream[i] && e->part_stream[i] == static_cast<cudaStream_t>(s)) lim = e->part_sms[i];
    gar_set_sm_limit(lim);
  }
  ~SmScope() { gar_set_sm_limit(0); }
};

struct ProfScope {
  gar_engine* e;
  cudaStream_t s;
  int idx = -1;
  ProfScope(gar_engine* e_, void* s_, int kind, double work) : e(e_), s(static_cast<cudaStream_t>(s_)) {
    if (!e->profiling) return;
    gar_engine::ProfRec r;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    r.kind = kind;
    r.work = work;
    cudaEventRecord(r.a, s);
    e->prof.push_back(r);
    idx = static_cast<int>(e->prof.size()) - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(e->prof[idx].b, s);
  }
};

int gemm(gar_engine* e, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K, int epi,
         const void* bias, const void* gamma, const void* res, int ldres, void* s) {
  e->launches += 1;
  ProfScope ps(e, s, M <= 16 ? PK_GEMV : PK_GEMM, M <= 16 ? 2.0 * N * static_cast<double>(K) : 2.0 * M * static_cast<double>(N) * K);
  if (e->use_skinny && M <= 16 && (epi == GAR_EPI_NONE || epi == GAR_EPI_RES))
    return gar_gemm_skinny_bf16(A, lda, B, ldb, C, ldc, M, N, K, epi, res, ldres, s);
  return gar_gemm_bf16(A, lda, B, ldb, C, ldc, M, N, K, epi, bias, gamma, res, ldres, 0, s);
}

// one Llama layer stack over `rows` rows of e->h (prefill: rows = S, one slot; decode: rows = batch)
int llm_layers(gar_engine* e, bf16* h, bf16* ln, int rows, bool decode, int slot, void* s, int nreq = 1) {
  const gar_model_desc& d = e->d;
  const int D = d.llm_dim, hq = d.llm_heads, hkv = d.llm_kv_heads, hd = d.llm_head_dim, I = d.llm_inter;
  const int qkv_w = (hq + 2 * hkv) * hd;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  NEEDW(cos_t, "llm.cos");
  NEEDW(sin_t, "llm.sin");
  for (int i = 0; i < d.llm_layers; ++i) {
    NEEDW(ln1, L("llm", i, "ln1"));
    NEEDW(wqkv, L("llm", i, "qkv.w"));
    NEEDW(wo, L("llm", i, "o.w"));
    NEEDW(ln2, L("llm", i, "ln2"));
    NEEDW(wgu, L("llm", i, "gu.w"));
    NEEDW(wdown, L("llm", i, "down.w"));
    const size_t slot_elems = static_cast<size_t>(hkv) * d.max_seq * hd;
    bf16* kc = e->kcache + static_cast<size_t>(i) * d.max_batch * slot_elems + (decode ? e->slot_base * slot_elems : 0);
    bf16* vc = e->vcache + static_cast<size_t>(i) * d.max_batch * slot_elems + (decode ? e->slot_base * slot_elems : 0);
    bf16* b_qkv = decode ? e->d_qkv : e->qkv_l;
    bf16* b_att = decode ? e->d_att : e->att_l;
    bf16* b_gu = decode ? e->d_gu : e->gu;
    bf16* b_act = decode ? e->d_act : e->act;
    const int* b_pos = e->pos + e->slot_base;
    TRY(gar_rmsnorm(h, D, ln1, ln, D, rows, D, d.rms_eps, s));
    TRY(gemm(e, ln, D, wqkv, D, b_qkv, qkv_w, rows, qkv_w, D, GAR_EPI_NONE, nullptr, nullptr, nullptr, 0, s));
    if (decode) {
      // decode: row b is sequence slot b, positions come from the device-side counters
      TRY(gar_llama_rope_kv(b_qkv, rows, hq, hkv, hd, cos_t, sin_t, 0, b_pos, nullptr, -1, kc, vc, d.max_seq, 0, s));
    } else {
      TRY(gar_llama_rope_kv(b_qkv, rows, hq, hkv, hd, cos_t, sin_t, 0, nullptr, nullptr, slot, kc, vc, d.max_seq,
                            nreq > 1 ? rows / nreq : 0, s));
    }
    e->launches += 2;
    if (decode) {
      ProfScope ps(e, s, PK_DECODE_ATTN, 0.0);
      int splits = e->dec_splits;
      if (splits <= 0) {  // ~4 CTAs per SM across (split, kv head, sequence)
        splits = (3 * gar_num_sms()) / (rows * hkv);  // one resident wave of 3 CTAs per SM
        splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
      }
      TRY(gar_decode_attn(b_qkv, qkv_w, kc, vc, b_pos, nullptr, rows, hq, hkv, hd, d.max_seq, splits, scale,
                          e->o_part, e->ml_part, e->tickets, b_att, hq * hd, s));
      e->launches += 2;
    } else {
      const bf16* kslot = kc + static_cast<size_t>(slot) * hkv * d.max_seq * hd;
      const bf16* vslot = vc + static_cast<size_t>(slot) * hkv * d.max_seq * hd;
      const int sl = rows / nreq;  // per-request sequence length; requests occupy consecutive slots
      // causal: half of the 4*S*S*hd*heads MACs-as-flops
      ProfScope ps(e, s, PK_ATTN, 2.0 * nreq * sl * static_cast<double>(sl) * hd * hq);
      TRY(gar_attn_fwd(b_qkv, qkv_w, static_cast<long long>(sl) * qkv_w, 0, kslot, vslot, hd,
                       static_cast<long long>(d.max_seq) * hd, nreq * hkv, 0, 0, 0, hkv, 1, b_att, hq * hd,
                       static_cast<long long>(sl) * hq * hd, 0, nreq, hq, hkv, hd, sl, sl, d.max_seq, 1, scale, s));
      e->launches += 1;
    }
    TRY(gemm(e, b_att, hq * hd, wo, hq * hd, h, D, rows, D, hq * hd, GAR_EPI_RES, nullptr, nullptr, h, D, s));
    TRY(gar_rmsnorm(h, D, ln2, ln, D, rows, D, d.rms_eps, s));
    // "llm.{i}.gu.w" rows are interleaved [32 gate | 32 up]: prefill fuses SwiGLU into the GEMM epilogue
    if (!decode && rows > 16) {
      TRY(gemm(e, ln, D, wgu, D, b_act, I, rows, 2 * I, D, GAR_EPI_SWIGLU, nullptr, nullptr, nullptr, 0, s));
    } else {
      TRY(gemm(e, ln, D, wgu, D, b_gu, 2 * I, rows, 2 * I, D, GAR_EPI_NONE, nullptr, nullptr, nullptr, 0, s));
      TRY(gar_swiglu(b_gu, b_act, rows, I, 1, s));
    }
    TRY(gemm(e, b_act, I, wdown, I, h, D, rows, D, I, GAR_EPI_RES, nullptr, nullptr, h, D, s));
    e->launches += 1;
  }
  return GAR_OK;
}

// `scratch`: rows x llm_dim normalised rows.  Prefill passes its own lnl, decode its private lndec: a decode group may
// run on a second stream while another group's prefill is in flight, so the two must never share scratch.
int lm_head(gar_engine* e, const bf16* hrows, int ldh, int rows, bf16* logits, bf16* scratch, void* s) {
  const gar_model_desc& d = e->d;
  NEEDW(wn, "llm.norm");
  NEEDW(wl, "llm.lm_head");
  TRY(gar_rmsnorm(hrows, ldh, wn, scratch, d.llm_dim, rows, d.llm_dim, d.rms_eps, s));
  e->launches += 1;
  return gemm(e, scratch, d.llm_dim, wl, d.llm_dim, logits, d.vocab_padded, rows, d.vocab_padded, d.llm_dim,
              GAR_EPI_NONE, nullptr, nullptr, nullptr, 0, s);
}

// KV splits of the decode attention: one resident wave of ~3 CTAs per SM across (split, kv head, sequence)
int decode_splits(const gar_engine* e, int batch) {
  int splits = e->dec_splits;
  if (splits <= 0) {
    // resident CTAs per SM of the decode attention kernel: 3 at head_dim 64, 2 at head_dim 128 (204 registers)
    const int per_sm = e->d.llm_head_dim >= 128 ? 2 : 3;
    splits = (per_sm * gar_num_sms()) / (batch * e->d.llm_kv_heads);
    splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
  }
  return splits;
}

// Decode step on the latency-first GEMV kernels: RMSNorm, SwiGLU and residual adds live inside them, the
// RMSNorm statistics travel as per-CTA partial sums of squares (7 launches per layer instead of 11).
int decode_step_gemv(gar_engine* e, int batch, void* s) {
  const gar_model_desc& d = e->d;
  const int D = d.llm_dim, hq = d.llm_heads, hkv = d.llm_kv_heads, hd = d.llm_head_dim, I = d.llm_inter;
  const int qkv_w = (hq + 2 * hkv) * hd;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  NEEDW(table, "llm.embed");
  NEEDW(cos_t, "llm.cos");
  NEEDW(sin_t, "llm.sin");
  TRY(gar_embed_gather_sumsq(table, e->next_tok + e->slot_base, e->hdec, e->sumsq_a, batch, D, s));
  e->launches += 1;
  int n_part = 1, rows = 0;
  const int splits = decode_splits(e, batch);
  for (int i = 0; i < d.llm_layers; ++i) {
    NEEDW(ln1, L("llm", i, "ln1"));
    NEEDW(wqkv, L("llm", i, "qkv.w"));
    NEEDW(wo, L("llm", i, "o.w"));
    NEEDW(ln2, L("llm", i, "ln2"));
    NEEDW(wgu, L("llm", i, "gu.w"));
    NEEDW(wdown, L("llm", i, "down.w"));
    const size_t slot_elems = static_cast<size_t>(hkv) * d.max_seq * hd;
    bf16* kc = e->kcache + (static_cast<size_t>(i) * d.max_batch + e->slot_base) * slot_elems;
    bf16* vc = e->vcache + (static_cast<size_t>(i) * d.max_batch + e->slot_base) * slot_elems;
    {
      ProfScope ps(e, s, PK_GEMV, 2.0 * qkv_w * static_cast<double>(D));
      TRY(gar_gemv_decode(e->hdec, D, wqkv, D, e->d_qkv, qkv_w, batch, qkv_w, D, 0, nullptr, 0, ln1, e->sumsq_a, n_part,
                          d.rms_eps, nullptr, nullptr, s));
    }
    {
      ProfScope ps(e, s, PK_DECODE_ATTN, 0.0);
      if (e->fuse_decode_rope && (hd == 64 || hd == 128)) {
        // RoPE of q and of the new K row, the KV append and the attention in one launch
        TRY(gar_decode_attn_rope(e->d_qkv, qkv_w, kc, vc, e->pos + e->slot_base, nullptr, batch, hq, hkv, hd, d.max_seq,
                                 splits, scale, cos_t, sin_t, e->o_part, e->ml_part, e->tickets, e->d_att, hq * hd, s));
        e->launches -= 1;
      } else {
        TRY(gar_llama_rope_kv(e->d_qkv, batch, hq, hkv, hd, cos_t, sin_t, 0, e->pos + e->slot_base, nullptr, -1, kc, vc,
                              d.max_seq, 0, s));
        TRY(gar_decode_attn(e->d_qkv, qkv_w, kc, vc, e->pos + e->slot_base, nullptr, batch, hq, hkv, hd, d.max_seq, splits,
                            scale, e->o_part, e->ml_part, e->tickets, e->d_att, hq * hd, s));
      }
    }
    {
      ProfScope ps(e, s, PK_GEMV, 2.0 * D * static_cast<double>(hq * hd));
      TRY(gar_gemv_decode(e->d_att, hq * hd, wo, hq * hd, e->hdec, D, batch, D, hq * hd, 1, e->hdec, D, nullptr, nullptr, 0,
                          0.f, e->sumsq_b, &rows, s));
    }
    {
      ProfScope ps(e, s, PK_GEMV, 2.0 * 2.0 * I * D);
      TRY(gar_gemv_decode(e->hdec, D, wgu, D, e->d_act, I, batch, 2 * I, D, 2, nullptr, 0, ln2, e->sumsq_b, D / rows,
                          d.rms_eps, nullptr, nullptr, s));
    }
    {
      ProfScope ps(e, s, PK_GEMV, 2.0 * D * static_cast<double>(I));
      TRY(gar_gemv_decode(e->d_act, I, wdown, I, e->hdec, D, batch, D, I, 1, e->hdec, D, nullptr, nullptr, 0, 0.f,
                          e->sumsq_a, &rows, s));
    }
    n_part = D / rows;
    e->launches += 6;  // qkv, RoPE+KV append, attention (+combine by ticket), o-proj, gate/up, down (fused RoPE: one less)
  }
  NEEDW(wn, "llm.norm");
  NEEDW(wl, "llm.lm_head");
  {
    ProfScope ps(e, s, PK_GEMV, 2.0 * d.vocab_padded * static_cast<double>(D));
    TRY(gar_gemv_decode(e->hdec, D, wl, D, e->logits + static_cast<size_t>(e->slot_base) * d.vocab_padded, d.vocab_padded,
                        batch, d.vocab_padded, D, 0, nullptr, 0, wn, e->sumsq_a, n_part, d.rms_eps, nullptr, nullptr, s));
  }
  TRY(gar_argmax_step_eos(e->logits + static_cast<size_t>(e->slot_base) * d.vocab_padded, d.vocab_padded, d.vocab, batch,
                          e->next_tok + e->slot_base, e->out_tokens + static_cast<size_t>(e->slot_base) * d.max_new_tokens,
                          d.max_new_tokens, e->step, e->pos + e->slot_base, e->finished + e->slot_base, e->eos_ids, e->n_eos,
                          1, 1, e->argmax_ws, s));
  e->launches += 3;
  return GAR_OK;
}

int decode_step(gar_engine* e, int batch, void* s) {
  const gar_model_desc& d = e->d;
  if (e->use_gemv && batch <= 8 && d.llm_dim % 256 == 0 && d.llm_inter % 256 == 0 &&
      (d.llm_heads * d.llm_head_dim) % 256 == 0 && d.llm_inter % 32 == 0 && d.vocab_padded % 64 == 0)
    return decode_step_gemv(e, batch, s);
  NEEDW(table, "llm.embed");
  TRY(gar_embed_gather(table, nullptr, e->next_tok + e->slot_base, e->hdec, batch, d.llm_dim, s));
  e->launches += 1;
  TRY(llm_layers(e, e->hdec, e->lndec, batch, true, 0, s));
  TRY(lm_head(e, e->hdec, d.llm_dim, batch, e->logits + static_cast<size_t>(e->slot_base) * d.vocab_padded, e->lndec, s));
  TRY(gar_argmax_step_eos(e->logits + static_cast<size_t>(e->slot_base) * d.vocab_padded, d.vocab_padded, d.vocab, batch,
                          e->next_tok + e->slot_base, e->out_tokens + static_cast<size_t>(e->slot_base) * d.max_new_tokens,
                          d.max_new_tokens, e->step, e->pos + e->slot_base, e->finished + e->slot_base, e->eos_ids, e->n_eos,
                          1, 1, e->argmax_ws, s));
  e->launches += 2;
  return GAR_OK;
}

}  // namespace

extern "C" {

int gar_engine_create(const gar_model_desc* desc, gar_engine** out) {
  if (!desc || !out) return GAR_ERR_ARG;
  const gar_model_desc& d = *desc;
  if (d.vit_dim % d.vit_heads || d.img_size % d.patch_size || d.vit_dim / d.vit_heads > 128 ||
      (d.llm_head_dim != 64 && d.llm_head_dim != 128)) {
    gar_set_error("gar_engine_create: unsupported head dims (vit %d <= 128 padded to 64/128, llm %d in {64,128})",
                  d.vit_dim / d.vit_heads, d.llm_head_dim);
    return GAR_ERR_SHAPE;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    gar_set_error("gar_engine_create: no CUDA device (this library has no CPU path)");
    return GAR_ERR_CUDA;
  }
  gar_engine* e = new gar_engine();
  e->d = d;
  e->grid = d.img_size / d.patch_size;
  e->npt = d.vit_use_cls ? 1 : 0;
  e->ntok = e->grid * e->grid + e->npt;
  e->hd_v = d.vit_dim / d.vit_heads;
  e->hdp_v = e->hd_v <= 64 ? 64 : 128;  // qkv/proj weights arrive zero-padded per head when hd_v != hdp_v
  e->dq_v = d.vit_heads * e->hdp_v;
  e->kpad = ((3 * d.patch_size * d.patch_size + 63) / 64) * 64;
  e->tpt = e->grid / d.pool_ratio;
  e->tokens_per_tile = e->tpt * e->tpt;
  e->dec_splits = 0;  // 0 = pick per batch so the split kernel fills the GPU
  const size_t T = d.max_tiles, NP = static_cast<size_t>(e->grid) * e->grid, NT = e->ntok;
  const size_t dv = d.vit_dim, D = d.llm_dim, S = d.max_seq, B = d.max_batch;
  const size_t PR = d.max_prefill_rows > d.max_seq ? d.max_prefill_rows : d.max_seq;  // rows of one prefill pass
  const size_t qkv_w = static_cast<size_t>(d.llm_heads + 2 * d.llm_kv_heads) * d.llm_head_dim;
  int rc = 0;
  rc |= dev_alloc(e, &e->a_img, T * NP * e->kpad, true);
  rc |= dev_alloc(e, &e->a_mask, T * NP * e->kpad, true);
  rc |= dev_alloc(e, &e->pe, T * NP * dv);
  rc |= dev_alloc(e, &e->x, T * NT * dv);
  rc |= dev_alloc(e, &e->ln, T * NT * dv);
  rc |= dev_alloc(e, &e->qkv_v, T * NT * 3 * e->dq_v);
  rc |= dev_alloc(e, &e->att_v, T * NT * e->dq_v);
  rc |= dev_alloc(e, &e->mlp_v, T * NT * d.vit_mlp);
  rc |= dev_alloc(e, &e->p1, T * NT * D);
  rc |= dev_alloc(e, &e->p2, T * NT * D);
  rc |= dev_alloc(e, &e->feats, T * e->tokens_per_tile * D);
  rc |= dev_alloc(e, &e->emb, PR * D);
  rc |= dev_alloc(e, &e->h, PR * D);
  rc |= dev_alloc(e, &e->lnl, PR * D);
  rc |= dev_alloc(e, &e->qkv_l, PR * qkv_w);
  rc |= dev_alloc(e, &e->d_qkv, B * qkv_w);
  rc |= dev_alloc(e, &e->d_att, B * d.llm_heads * d.llm_head_dim);
  rc |= dev_alloc(e, &e->d_gu, B * 2 * d.llm_inter);
  rc |= dev_alloc(e, &e->d_act, B * d.llm_inter);
  rc |= dev_alloc(e, &e->att_l, PR * d.llm_heads * d.llm_head_dim);
  rc |= dev_alloc(e, &e->gu, PR * 2 * d.llm_inter);
  rc |= dev_alloc(e, &e->act, PR * d.llm_inter);
  rc |= dev_alloc(e, &e->logits, B * d.vocab_padded, true);
  rc |= dev_alloc(e, &e->hdec, B * D);
  rc |= dev_alloc(e, &e->lndec, (B > 1 ? B : 1) * D);
  const size_t kv_elems = static_cast<size_t>(d.llm_layers) * B * d.llm_kv_heads * S * d.llm_head_dim;
  rc |= dev_alloc(e, &e->kcache, kv_elems, true);  // zero-filled: TMA tiles past the live length must be finite
  rc |= dev_alloc(e, &e->vcache, kv_elems, true);
  rc |= dev_alloc(e, &e->o_part, B * d.llm_heads * 32 * d.llm_head_dim);
  rc |= dev_alloc(e, &e->ml_part, B * d.llm_heads * 32 * 2);
  rc |= dev_alloc(e, &e->plan, PR + 8 * (B + 1));  // row plan + the per-request region-start table
  rc |= dev_alloc(e, &e->status, 2 * B, true);
  rc |= dev_alloc(e, &e->next_tok, B, true);
  rc |= dev_alloc(e, &e->out_tokens, B * d.max_new_tokens, true);
  rc |= dev_alloc(e, &e->pos, B, true);
  rc |= dev_alloc(e, &e->finished, B, true);
  rc |= dev_alloc(e, &e->step, 2, true);
  rc |= dev_alloc(e, &e->tickets, B * d.llm_kv_heads, true);
  rc |= dev_alloc(e, &e->argmax_ws, B * ((static_cast<size_t>(d.vocab) + 4095) / 4096) * 2, true);
  rc |= dev_alloc(e, &e->ids_scratch, S);
  rc |= dev_alloc(e, &e->sumsq_a, (D / 16 + 1) * 8, true);
  rc |= dev_alloc(e, &e->sumsq_b, (D / 16 + 1) * 8, true);
  if (cudaMallocHost(reinterpret_cast<void**>(&e->fin_host), 2 * B * sizeof(int)) != cudaSuccess) rc |= GAR_ERR_CUDA;
  if (cudaMallocHost(reinterpret_cast<void**>(&e->status_host), 2 * B * sizeof(int)) != cudaSuccess) rc |= GAR_ERR_CUDA;
  if (cudaMallocHost(reinterpret_cast<void**>(&e->out_host), B * static_cast<size_t>(d.max_new_tokens) * sizeof(int)) !=
      cudaSuccess)
    rc |= GAR_ERR_CUDA;
  if (cudaEventCreateWithFlags(&e->plan_ev, cudaEventDisableTiming) != cudaSuccess) rc |= GAR_ERR_CUDA;
  if (cudaEventCreateWithFlags(&e->out_ev, cudaEventDisableTiming) != cudaSuccess) rc |= GAR_ERR_CUDA;
  for (int i = 0; i < 2 && !rc; ++i)
    if (cudaEventCreateWithFlags(&e->fin_ev[i], cudaEventDisableTiming) != cudaSuccess) rc |= GAR_ERR_CUDA;
  if (rc) {
    gar_engine_destroy(e);
    return GAR_ERR_CUDA;
  }
  *out = e;
  return GAR_OK;
}

int gar_engine_destroy(gar_engine* e) {
  if (!e) return GAR_OK;
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  for (void* p : e->allocs) cudaFree(p);
  if (e->fin_host) cudaFreeHost(e->fin_host);
  if (e->status_host) cudaFreeHost(e->status_host);
  if (e->out_host) cudaFreeHost(e->out_host);
  if (e->plan_ev) cudaEventDestroy(e->plan_ev);
  if (e->out_ev) cudaEventDestroy(e->out_ev);
  for (int i = 0; i < 2; ++i)
    if (e->fin_ev[i]) cudaEventDestroy(e->fin_ev[i]);
  delete e;
  return GAR_OK;
}

int gar_engine_set_weight(gar_engine* e, const char* name, const void* dev_ptr) {
  if (!e || !name) return GAR_ERR_ARG;
  e->w[name] = dev_ptr;
  return GAR_OK;
}

int gar_engine_get_buffer(gar_engine* e, const char* name, void** dev_ptr, long long* bytes) {
  if (!e || !name || !dev_ptr) return GAR_ERR_ARG;
  const gar_model_desc& d = e->d;
  const std::string n(name);
  const long long T = d.max_tiles;
  long long b = 0;
  void* p = nullptr;
  if (n == "image_features") { p = e->feats; b = T * e->tokens_per_tile * d.llm_dim * 2; }
  else if (n == "inputs_embeds") { p = e->emb; b = static_cast<long long>(d.max_seq) * d.llm_dim * 2; }
  else if (n == "hidden") { p = e->h; b = static_cast<long long>(d.max_seq) * d.llm_dim * 2; }
  else if (n == "logits") { p = e->logits; b = static_cast<long long>(d.max_batch) * d.vocab_padded * 2; }
  else if (n == "patch_embed") { p = e->pe; b = T * e->grid * e->grid * d.vit_dim * 2; }
  else if (n == "vit_x") { p = e->x; b = T * e->ntok * d.vit_dim * 2; }
  else if (n == "kcache") { p = e->kcache; b = static_cast<long long>(d.llm_layers) * d.max_batch * d.llm_kv_heads * d.max_seq * d.llm_head_dim * 2; }
  else if (n == "vcache") { p = e->vcache; b = static_cast<long long>(d.llm_layers) * d.max_batch * d.llm_kv_heads * d.max_seq * d.llm_head_dim * 2; }
  else if (n == "out_tokens") { p = e->out_tokens; b = static_cast<long long>(d.max_batch) * d.max_new_tokens * 4; }
  else {
    gar_set_error("gar_engine_get_buffer: unknown buffer '%s'", name);
    return GAR_ERR_ARG;
  }
  *dev_ptr = p;
  if (bytes) *bytes = b;
  return GAR_OK;
}

// R2 + operand prep of R3/K1 for T tiles of ONE request, written at tile `tile0` of the im2col arenas: the requests of a
// batched pass are patchified straight from their own tensors (no concatenated copy of the pixel data is ever made).
int gar_engine_patchify_at(gar_engine* e, const void* pixel_values, const void* mask_values, int T, int tile0, void* s) {
  if (!e || !pixel_values || T <= 0 || tile0 < 0 || tile0 + T > e->d.max_tiles) {
    gar_set_error("gar_engine_patchify_at: tiles [%d,%d) outside [0,%d)", tile0, tile0 + T, e ? e->d.max_tiles : 0);
    return GAR_ERR_ARG;
  }
  const gar_model_desc& d = e->d;
  const size_t off = static_cast<size_t>(tile0) * e->grid * e->grid * e->kpad;
  TRY(gar_patchify(pixel_values, mask_values, e->a_img + off, mask_values ? e->a_mask + off : nullptr, T, d.img_size,
                   d.patch_size, e->kpad, d.prompt_numbers, s));
  e->launches += 1;
  return GAR_OK;
}

int gar_engine_vit_encode(gar_engine* e, const void* pixel_values, const void* mask_values, int T, void* s) {
  if (!e || !pixel_values || T <= 0 || T > e->d.max_tiles) {
    gar_set_error("gar_engine_vit_encode: T=%d outside [1,%d]", T, e ? e->d.max_tiles : 0);
    return GAR_ERR_ARG;
  }
  TRY(gar_engine_patchify_at(e, pixel_values, mask_values, T, 0, s));
  return gar_engine_vit_encode_patches(e, T, mask_values != nullptr, s);
}

// R3-R5 on the first T tiles of the im2col arenas (filled by gar_engine_patchify_at)
int gar_engine_vit_encode_patches(gar_engine* e, int T, int with_mask, void* s) {
  SmScope sm_scope(e, s);
  if (!e || T <= 0 || T > e->d.max_tiles) {
    gar_set_error("gar_engine_vit_encode_patches: T=%d outside [1,%d]", T, e ? e->d.max_tiles : 0);
    return GAR_ERR_ARG;
  }
  const gar_model_desc& d = e->d;
  const int dv = d.vit_dim, NP = e->grid * e->grid, NT = e->ntok, rows = T * NT, D = d.llm_dim;
  NEEDW(wp, "vit.patch_w");
  NEEDW(pos, "vit.pos");
  NEEDW(npw, "vit.norm_pre.w");
  NEEDW(npb, "vit.norm_pre.b");
  NEEDW(rope, "vit.rope");
  NEEDW(rope_cm, "vit.rope_cm");
  const void* cls = e->npt ? W(e, "vit.cls") : nullptr;
  if (e->npt && !cls) return GAR_ERR_STATE;
  // the two k=s=14 convs as GEMMs: x = bf16(conv(px)) ; x = bf16(x + bf16(conv(mask)))
  TRY(gemm(e, e->a_img, e->kpad, wp, e->kpad, e->pe, dv, T * NP, dv, e->kpad, GAR_EPI_NONE, nullptr, nullptr, nullptr,
           0, s));
  if (with_mask) {
    NEEDW(wm, "vit.mask_w");
    TRY(gemm(e, e->a_mask, e->kpad, wm, e->kpad, e->pe, dv, T * NP, dv, e->kpad, GAR_EPI_RES, nullptr, nullptr, e->pe,
             dv, s));
  }
  TRY(gar_vit_embed_ln(e->pe, cls, pos, npw, npb, e->x, T, NT, e->npt, dv, d.ln_eps, s));
  e->launches += 1;
  const float scale = 1.0f / sqrtf(static_cast<float>(e->hd_v));
  for (int i = 0; i < d.vit_depth; ++i) {
    NEEDW(l1w, L("vit", i, "ln1.w"));
    NEEDW(l1b, L("vit", i, "ln1.b"));
    NEEDW(wqkv, L("vit", i, "qkv.w"));
    NEEDW(bqkv, L("vit", i, "qkv.b"));
    NEEDW(wpr, L("vit", i, "proj.w"));
    NEEDW(bpr, L("vit", i, "proj.b"));
    NEEDW(g1, L("vit", i, "g1"));
    NEEDW(l2w, L("vit", i, "ln2.w"));
    NEEDW(l2b, L("vit", i, "ln2.b"));
    NEEDW(w1, L("vit", i, "fc1.w"));
    NEEDW(b1, L("vit", i, "fc1.b"));
    NEEDW(w2, L("vit", i, "fc2.w"));
    NEEDW(b2, L("vit", i, "fc2.b"));
    NEEDW(g2, L("vit", i, "g2"));
    TRY(gar_layernorm(e->x, dv, l1w, l1b, e->ln, dv, rows, dv, d.ln_eps, s));
    const int dq = e->dq_v, hdp = e->hdp_v;
    if (e->fuse_rope) {
      e->launches += 1;
      ProfScope ps(e, s, PK_GEMM, 2.0 * rows * 3.0 * dq * dv);
      TRY(gar_gemm_bf16_rope2d(e->ln, dv, wqkv, dv, e->qkv_v, 3 * dq, rows, 3 * dq, dv, bqkv, rope_cm, NT, e->npt, hdp,
                               2 * dq, s));
    } else {
      TRY(gemm(e, e->ln, dv, wqkv, dv, e->qkv_v, 3 * dq, rows, 3 * dq, dv, GAR_EPI_BIAS, bqkv, nullptr, nullptr, 0, s));
      TRY(gar_vit_rope(e->qkv_v, rope, rows, NT, e->npt, d.vit_heads, hdp, s));
      e->launches += 1;
    }
    {
    ProfScope ps(e, s, PK_ATTN, 4.0 * T * static_cast<double>(NT) * NT * dq);
    TRY(gar_attn_fwd(e->qkv_v, 3 * dq, static_cast<long long>(NT) * 3 * dq, 0, e->qkv_v, e->qkv_v, 3 * dq,
                     static_cast<long long>(NT) * 3 * dq, T, dq, 2 * dq, hdp, 1, 0, e->att_v, dq,
                     static_cast<long long>(NT) * dq, 0, T, d.vit_heads, d.vit_heads, hdp, NT, NT, NT, 0, scale, s));
    }
    TRY(gemm(e, e->att_v, dq, wpr, dq, e->x, dv, rows, dv, dq, GAR_EPI_BIAS_LS_RES, bpr, g1, e->x, dv, s));
    TRY(gar_layernorm(e->x, dv, l2w, l2b, e->ln, dv, rows, dv, d.ln_eps, s));
    TRY(gemm(e, e->ln, dv, w1, dv, e->mlp_v, d.vit_mlp, rows, d.vit_mlp, dv, GAR_EPI_BIAS_GELU, b1, nullptr, nullptr, 0,
             s));
    TRY(gemm(e, e->mlp_v, d.vit_mlp, w2, d.vit_mlp, e->x, dv, rows, dv, d.vit_mlp, GAR_EPI_BIAS_LS_RES, b2, g2, e->x, dv,
             s));
    e->launches += 3;
  }
  // projector on every token row (cls rows are computed and then skipped by the pool: 0.1% extra work)
  NEEDW(pw1, "proj.l1.w");
  NEEDW(pb1, "proj.l1.b");
  NEEDW(pw2, "proj.l2.w");
  NEEDW(pb2, "proj.l2.b");
  TRY(gemm(e, e->x, dv, pw1, dv, e->p1, D, rows, D, dv, GAR_EPI_BIAS_GELU, pb1, nullptr, nullptr, 0, s));
  TRY(gemm(e, e->p1, D, pw2, D, e->p2, D, rows, D, D, GAR_EPI_BIAS, pb2, nullptr, nullptr, 0, s));
  if (d.pool_ratio == 2) {
    TRY(gar_pool2x2(e->p2, e->feats, T, NT, e->npt, e->grid, D, s));
  } else {
    gar_set_error("gar_engine_vit_encode: projector_pooling_ratio %d not built (2 only)", d.pool_ratio);
    return GAR_ERR_SHAPE;
  }
  e->launches += 1;
  return GAR_OK;
}

int gar_engine_assemble(gar_engine* e, int slot, const long long* input_ids_dev, int S, int T, const float* rois,
                        const int* has_roi, float spatial_scale, int ncw, int nch, void* s) {
  if (!e || slot < 0 || slot >= e->d.max_batch || S <= 0 || S > e->d.max_seq) {
    gar_set_error("gar_engine_assemble: slot=%d S=%d outside capacity (max_batch=%d max_seq=%d)", slot, S,
                  e ? e->d.max_batch : 0, e ? e->d.max_seq : 0);
    return GAR_ERR_ARG;
  }
  if (e->replay_mode != GAR_REPLAY_FRAME && ncw * nch != T - 1) {  // reference assert in _merge, modeling_gar.py:250
    gar_set_error("%d != %d", ncw * nch, T - 1);
    return GAR_ERR_SHAPE;
  }
  const gar_model_desc& d = e->d;
  NEEDW(table, "llm.embed");
  e->launches += 2;
  e->plan_slot0 = -1;
  // algorithmic bytes: write S*D*2 + read one D-row per non-replay row + <=16 cells per region
  ProfScope ps(e, s, PK_ASSEMBLE, 2.0 * (2.0 * S - 0.0) * d.llm_dim);
  return gar_assemble_embeds(input_ids_dev, S, table, d.vocab, e->feats, T * e->tokens_per_tile, d.llm_dim,
                             e->crop_ids(), rois, has_roi, e->n_regions(), spatial_scale, ncw, nch, e->tpt,
                             d.image_token_id, e->replay_mode, e->emb, e->plan, e->status + 2 * slot, s);
}

int gar_engine_assemble_at(gar_engine* e, int slot, const long long* input_ids_dev, int S, int T, int feat_tile0,
                           int emb_row0, const float* rois, const int* has_roi, float spatial_scale, int ncw, int nch,
                           void* s) {
  const gar_model_desc& d = e->d;
  const long long pr = d.max_prefill_rows > d.max_seq ? d.max_prefill_rows : d.max_seq;
  if (!e || slot < 0 || slot >= d.max_batch || S <= 0 || S > d.max_seq || emb_row0 < 0 || emb_row0 + S > pr ||
      feat_tile0 < 0 || feat_tile0 + T > d.max_tiles) {
    gar_set_error("gar_engine_assemble_at: slot=%d S=%d T=%d tile0=%d row0=%d outside capacity", slot, S, T, feat_tile0,
                  emb_row0);
    return GAR_ERR_ARG;
  }
  if (e->replay_mode != GAR_REPLAY_FRAME && ncw * nch != T - 1) {  // reference assert in _merge, modeling_gar.py:250
    gar_set_error("%d != %d", ncw * nch, T - 1);
    return GAR_ERR_SHAPE;
  }
  NEEDW(table, "llm.embed");
  e->launches += 2;
  e->plan_slot0 = -1;
  ProfScope ps(e, s, PK_ASSEMBLE, 2.0 * (2.0 * S) * d.llm_dim);
  return gar_assemble_embeds(input_ids_dev, S, table, d.vocab,
                             e->feats + static_cast<size_t>(feat_tile0) * e->tokens_per_tile * d.llm_dim,
                             T * e->tokens_per_tile, d.llm_dim, e->crop_ids(), rois, has_roi, e->n_regions(),
                             spatial_scale, ncw, nch, e->tpt, d.image_token_id, e->replay_mode,
                             e->emb + static_cast<size_t>(emb_row0) * d.llm_dim, e->plan + emb_row0, e->status + 2 * slot, s);
}

// R6 row plan + validation for nreq equal-shape requests.  Depends only on input_ids / which boxes exist, so callers
// run it FIRST (before the ViT): the status words come back to the host right away and a malformed prompt raises
// before any feature is computed (the reference raises its ValueError before the language model runs,
// modeling_perception_lm.py:309-315).  Synchronises `stream` (cheap when the plan is the first thing enqueued).
// status_host: nreq status words (GAR_STATUS_* bits), may be NULL.
int gar_engine_plan_batch(gar_engine* e, int slot0, int nreq, const long long* const* input_ids_dev, int S, int T,
                          const int* has_roi, int* status_host, void* s) {
  if (!e) return GAR_ERR_ARG;
  const gar_model_desc& d = e->d;
  const long long pr = d.max_prefill_rows > d.max_seq ? d.max_prefill_rows : d.max_seq;
  if (nreq <= 0 || slot0 < 0 || slot0 + nreq > d.max_batch || S <= 0 || S > d.max_seq ||
      static_cast<long long>(nreq) * S > pr || T <= 0 || static_cast<long long>(nreq) * T > d.max_tiles || !has_roi) {
    gar_set_error("gar_engine_plan_batch: slot0=%d nreq=%d S=%d T=%d outside capacity", slot0, nreq, S, T);
    return GAR_ERR_ARG;
  }
  {
    ProfScope ps(e, s, PK_PLAN, 0.0);  // the plan kernel only (the status read-back below is host plumbing)
    TRY(gar_assemble_plan_batch(nreq, input_ids_dev, S, T * e->tokens_per_tile, e->crop_ids(), has_roi, e->n_regions(),
                                e->tpt, d.image_token_id, e->replay_mode, e->plan, e->status + 2 * slot0, s));
  }
  e->launches += (nreq + 7) / 8;
  e->plan_slot0 = slot0; e->plan_nreq = nreq; e->plan_S = S; e->plan_T = T;
  if (status_host) {
    int* st = e->status_host + 2 * slot0;
    GAR_CUDA_CHECK(cudaMemcpyAsync(st, e->status + 2 * slot0, sizeof(int) * 2 * nreq, cudaMemcpyDeviceToHost,
                                   static_cast<cudaStream_t>(s)));
    GAR_CUDA_CHECK(cudaEventRecord(e->plan_ev, static_cast<cudaStream_t>(s)));
    GAR_CUDA_CHECK(cudaEventSynchronize(e->plan_ev));
    for (int r = 0; r < nreq; ++r) status_host[r] = st[2 * r];
  }
  return GAR_OK;
}

// R6-R7 for nreq equal-shape requests of a batched pass: request r reads features from tile r*T, writes
// inputs_embeds rows [r*S, (r+1)*S) and status slot slot0+r.  One row launch for everything; the plan launch is skipped
// when gar_engine_plan_batch already ran for the same (slot0, nreq, S, T).
int gar_engine_assemble_batch(gar_engine* e, int slot0, int nreq, const long long* const* input_ids_dev, int S, int T,
                              const float* rois, const int* has_roi, float spatial_scale, const int* ncw,
                              const int* nch, void* s) {
  if (!e) return GAR_ERR_ARG;
  const gar_model_desc& d = e->d;
  const long long pr = d.max_prefill_rows > d.max_seq ? d.max_prefill_rows : d.max_seq;
  if (nreq <= 0 || slot0 < 0 || slot0 + nreq > d.max_batch || S <= 0 || S > d.max_seq ||
      static_cast<long long>(nreq) * S > pr || nreq * T > d.max_tiles) {
    gar_set_error("gar_engine_assemble_batch: slot0=%d nreq=%d S=%d T=%d outside capacity", slot0, nreq, S, T);
    return GAR_ERR_ARG;
  }
  for (int r = 0; r < nreq; ++r) {
    if (e->replay_mode != GAR_REPLAY_FRAME && ncw[r] * nch[r] != T - 1) {  // reference assert in _merge, modeling_gar.py:250
      gar_set_error("%d != %d", ncw[r] * nch[r], T - 1);
      return GAR_ERR_SHAPE;
    }
  }
  NEEDW(table, "llm.embed");
  const bool planned = e->plan_slot0 == slot0 && e->plan_nreq == nreq && e->plan_S == S && e->plan_T == T;
  e->plan_slot0 = -1;
  // algorithmic bytes: write S*D*2 + read one D-row per non-replay row + <=16 cells per region
  ProfScope ps(e, s, PK_ASSEMBLE, 2.0 * (2.0 * S) * d.llm_dim * nreq);
  if (!planned) {
    TRY(gar_assemble_plan_batch(nreq, input_ids_dev, S, T * e->tokens_per_tile, e->crop_ids(), has_roi, e->n_regions(),
                                e->tpt, d.image_token_id, e->replay_mode, e->plan, e->status + 2 * slot0, s));
    e->launches += (nreq + 7) / 8;
  }
  e->launches += 2 * ((nreq + 7) / 8);
  return gar_assemble_rows_batch(nreq, input_ids_dev, S, table, d.vocab, e->feats, T * e->tokens_per_tile, d.llm_dim,
                                 e->crop_ids(), rois, has_roi, e->n_regions(), spatial_scale, ncw, nch, e->tpt,
                                 d.image_token_id, e->replay_mode, e->emb, e->plan, s);
}

int gar_engine_prefill_batch(gar_engine* e, int slot0, int nreq, int S, void* s) {
  SmScope sm_scope(e, s);
  const gar_model_desc& d = e->d;
  const long long pr = d.max_prefill_rows > d.max_seq ? d.max_prefill_rows : d.max_seq;
  if (!e || nreq <= 0 || slot0 < 0 || slot0 + nreq > d.max_batch || S <= 0 || S + d.max_new_tokens > d.max_seq ||
      static_cast<long long>(nreq) * S > pr) {
    gar_set_error("gar_engine_prefill_batch: slot0=%d nreq=%d S=%d outside capacity (max_prefill_rows=%lld)", slot0, nreq,
                  S, pr);
    return GAR_ERR_ARG;
  }
  const int rows = nreq * S;
  GAR_CUDA_CHECK(cudaMemcpyAsync(e->h, e->emb, static_cast<size_t>(rows) * d.llm_dim * 2, cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(s)));
  TRY(llm_layers(e, e->h, e->lnl, rows, false, slot0, s, nreq));
  // last position of every request: rows (r+1)*S-1, i.e. a strided row view of h
  TRY(lm_head(e, e->h + static_cast<size_t>(S - 1) * d.llm_dim, S * d.llm_dim, nreq,
              e->logits + static_cast<size_t>(slot0) * d.vocab_padded, e->lnl, s));
  fill_int_kernel<<<1, 64, 0, static_cast<cudaStream_t>(s)>>>(e->pos + slot0, S, nreq);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

int gar_engine_prefill(gar_engine* e, int slot, int S, void* s) {
  SmScope sm_scope(e, s);
  if (!e || slot < 0 || slot >= e->d.max_batch || S <= 0 || S + e->d.max_new_tokens > e->d.max_seq) {
    gar_set_error("gar_engine_prefill: slot=%d S=%d (+%d new) exceeds max_seq=%d", slot, S,
                  e ? e->d.max_new_tokens : 0, e ? e->d.max_seq : 0);
    return GAR_ERR_ARG;
  }
  const gar_model_desc& d = e->d;
  GAR_CUDA_CHECK(cudaMemcpyAsync(e->h, e->emb, static_cast<size_t>(S) * d.llm_dim * 2, cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(s)));
  TRY(llm_layers(e, e->h, e->lnl, S, false, slot, s));
  // logits only for the last position (logits_to_keep=1 in HF generate's prefill)
  TRY(lm_head(e, e->h + static_cast<size_t>(S - 1) * d.llm_dim, d.llm_dim, 1,
              e->logits + static_cast<size_t>(slot) * d.vocab_padded, e->lnl, s));
  fill_int_kernel<<<1, 64, 0, static_cast<cudaStream_t>(s)>>>(e->pos + slot, S, 1);
  GAR_CUDA_CHECK(cudaGetLastError());
  return GAR_OK;
}

int gar_engine_prefill_all_logits(gar_engine* e, int slot, int S, void* logits_out, void* s) {
  SmScope sm_scope(e, s);
  if (!e || !logits_out || slot < 0 || slot >= e->d.max_batch || S <= 0 || S > e->d.max_seq) return GAR_ERR_ARG;
  const gar_model_desc& d = e->d;
  GAR_CUDA_CHECK(cudaMemcpyAsync(e->h, e->emb, static_cast<size_t>(S) * d.llm_dim * 2, cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(s)));
  TRY(llm_layers(e, e->h, e->lnl, S, false, slot, s));
  NEEDW(wn, "llm.norm");
  NEEDW(wl, "llm.lm_head");
  TRY(gar_rmsnorm(e->h, d.llm_dim, wn, e->lnl, d.llm_dim, S, d.llm_dim, d.rms_eps, s));
  e->launches += 1;
  return gemm(e, e->lnl, d.llm_dim, wl, d.llm_dim, logits_out, d.vocab_padded, S, d.vocab_padded, d.llm_dim,
              GAR_EPI_NONE, nullptr, nullptr, nullptr, 0, s);
}

// Greedy decode of KV slots [slot0, slot0 + batch).  Only one decode runs at a time, but it may run on its own
// stream while gar_engine_vit_encode / assemble / prefill for OTHER slots are in flight on another stream
// (decode touches only decode-private scratch, its slots' KV / logits / counters, and the read-only weights).
//
// Stop ids: HF's greedy loop ends a sequence at the first token in GenerationConfig.eos_token_id (an int or a list,
// demo/gar_with_mask.py:115-120).  Finished sequences are latched on the device; the host polls the latches every
// `eos_poll_steps` graph replays (async copy + event, one chunk always in flight) and stops launching once every
// sequence of the group has finished, so a 1024-token budget costs what the caption needs, not 1024 steps.
int gar_engine_generate_eos(gar_engine* e, int slot0, int batch, int max_new_tokens, const int* eos_ids, int n_eos,
                            int* out_tokens_host, int* steps_run, void* s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  SmScope sm_scope(e, s_);
  if (!e || batch <= 0 || slot0 < 0 || slot0 + batch > e->d.max_batch || max_new_tokens <= 0 ||
      max_new_tokens > e->d.max_new_tokens || !out_tokens_host || n_eos < 0 || n_eos > 4 || (n_eos > 0 && !eos_ids)) {
    gar_set_error("gar_engine_generate: slot0=%d batch=%d max_new_tokens=%d n_eos=%d outside capacity", slot0, batch,
                  max_new_tokens, n_eos);
    return GAR_ERR_ARG;
  }
  const gar_model_desc& d = e->d;
  e->slot_base = slot0;
  e->n_eos = 0;
  for (int i = 0; i < 4; ++i) e->eos_ids[i] = -1;
  for (int i = 0; i < n_eos; ++i)
    if (eos_ids[i] >= 0) e->eos_ids[e->n_eos++] = eos_ids[i];
  const std::vector<int> eos_key(e->eos_ids, e->eos_ids + 4);
  int* out_dev = e->out_tokens + static_cast<size_t>(slot0) * d.max_new_tokens;
  GAR_CUDA_CHECK(cudaMemsetAsync(e->step, 0, 2 * sizeof(int), s));
  GAR_CUDA_CHECK(cudaMemsetAsync(e->finished + slot0, 0, sizeof(int) * batch, s));
  GAR_CUDA_CHECK(cudaMemsetAsync(out_dev, 0xff, sizeof(int) * batch * d.max_new_tokens, s));
  // token 0 from the prefill logits (position bookkeeping already at S per slot)
  TRY(gar_argmax_step_eos(e->logits + static_cast<size_t>(slot0) * d.vocab_padded, d.vocab_padded, d.vocab, batch,
                          e->next_tok + slot0, out_dev, d.max_new_tokens, e->step, e->pos + slot0, e->finished + slot0,
                          e->eos_ids, e->n_eos, 0, 1, e->argmax_ws, s));
  e->launches += 2;
  const int steps = max_new_tokens - 1;
  int done = 0;  // decode steps launched
  if (steps > 0) {
    const bool poll = e->n_eos > 0 && e->eos_poll_steps > 0;
    int* fin[2] = {e->fin_host, e->fin_host + d.max_batch};
    bool pending[2] = {false, false};
    int cur = 0;
    // returns true when every sequence of the group had finished at the time of poll buffer `c`
    auto all_finished = [&](int c) -> bool {
      if (!pending[c]) return false;
      cudaEventSynchronize(e->fin_ev[c]);
      pending[c] = false;
      for (int b = 0; b < batch; ++b)
        if (!fin[c][b]) return false;
      return true;
    };
    auto post_poll = [&]() -> int {
      GAR_CUDA_CHECK(cudaMemcpyAsync(fin[cur], e->finished + slot0, sizeof(int) * batch, cudaMemcpyDeviceToHost, s));
      GAR_CUDA_CHECK(cudaEventRecord(e->fin_ev[cur], s));
      pending[cur] = true;
      cur ^= 1;
      return GAR_OK;
    };
    bool stop = false;
    if (poll) TRY(post_poll());  // the first token may already be a stop id: checked at the first chunk boundary
    if (!stop) {
      // step 1 eagerly (also the warm-up that sets every kernel attribute), the rest as graph replays
      const long long before = e->launches;
      TRY(decode_step(e, batch, s));
      e->launches_per_step = static_cast<int>(e->launches - before);
      done = 1;
    }
    if (!stop && steps > 1 && e->profiling) {
      for (int i = 1; i < steps; ++i) TRY(decode_step(e, batch, s));
      done = steps;
    } else if (!stop && steps > 1) {
      // keyed by the stream too: a graph captured inside one SM partition is replayed only there
      const long long key = (static_cast<long long>(slot0) << 32) | (static_cast<long long>(batch) << 8) |
                            (static_cast<long long>(reinterpret_cast<uintptr_t>(s) >> 4 & 0xff) << 40);
      auto it = e->graphs.find(key);
      if (it == e->graphs.end() || e->graph_eos[key] != eos_key) {
        if (s == nullptr || s == cudaStreamLegacy) {
          gar_set_error("gar_engine_generate: needs a non-default stream (CUDA graph capture)");
          return GAR_ERR_ARG;
        }
        if (it != e->graphs.end()) cudaGraphExecDestroy(it->second);
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        GAR_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const long long keep = e->launches;
        int rc = decode_step(e, batch, s);
        e->launches = keep;
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc) return rc;
        GAR_CUDA_CHECK(ce);
        GAR_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        e->graphs[key] = exec;
        e->graph_eos[key] = eos_key;
        it = e->graphs.find(key);
      }
      int i = 1;
      while (i < steps && !stop) {
        const int n = poll ? (steps - i < e->eos_poll_steps ? steps - i : e->eos_poll_steps) : steps - i;
        for (int j = 0; j < n; ++j) GAR_CUDA_CHECK(cudaGraphLaunch(it->second, s));
        i += n;
        done += n;
        if (poll && i < steps) {
          TRY(post_poll());
          stop = all_finished(cur);  // the poll posted one chunk earlier: the newest chunk keeps the GPU busy meanwhile
        }
      }
      e->launches += static_cast<long long>(e->launches_per_step) * (done - 1);
    }
  }
  e->steps_run = done;
  if (steps_run) *steps_run = done;
  int* tmp = e->out_host + static_cast<size_t>(slot0) * d.max_new_tokens;
  GAR_CUDA_CHECK(cudaMemcpyAsync(tmp, out_dev, static_cast<size_t>(batch) * d.max_new_tokens * sizeof(int),
                                 cudaMemcpyDeviceToHost, s));
  GAR_CUDA_CHECK(cudaEventRecord(e->out_ev, s));
  GAR_CUDA_CHECK(cudaEventSynchronize(e->out_ev));
  for (int b = 0; b < batch; ++b)
    for (int t = 0; t < max_new_tokens; ++t)
      out_tokens_host[b * max_new_tokens + t] = tmp[static_cast<size_t>(b) * d.max_new_tokens + t];
  return GAR_OK;
}

int gar_engine_generate_group(gar_engine* e, int slot0, int batch, int max_new_tokens, int eos, int* out_tokens_host,
                              void* s) {
  return gar_engine_generate_eos(e, slot0, batch, max_new_tokens, &eos, eos >= 0 ? 1 : 0, out_tokens_host, nullptr, s);
}

int gar_engine_generate(gar_engine* e, int batch, int max_new_tokens, int eos, int* out_tokens_host, void* s) {
  return gar_engine_generate_group(e, 0, batch, max_new_tokens, eos, out_tokens_host, s);
}

int gar_engine_status(gar_engine* e, int slot, int* status_bits) {
  if (!e || slot < 0 || slot >= e->d.max_batch || !status_bits) return GAR_ERR_ARG;
  int st[2] = {0, 0};
  GAR_CUDA_CHECK(cudaMemcpy(st, e->status + 2 * slot, sizeof(st), cudaMemcpyDeviceToHost));
  *status_bits = st[0];
  return GAR_OK;
}

long long gar_engine_launch_count(gar_engine* e) { return e ? e->launches : 0; }

// Splits the GPU's SMs into a decode partition (>= decode_sms SMs; the driver rounds to its granularity) and an encode
// partition (the rest) with one stream each (CUDA green contexts).  Work enqueued on the two streams runs on disjoint
// SMs, so the HBM-bound decode of one request group and the tensor-bound ViT + prefill of the next overlap for real
// instead of interleaving at kernel boundaries.  Fails (GAR_ERR_CUDA) when the driver cannot partition.
int gar_engine_partition(gar_engine* e, int decode_sms, void** encode_stream, int* encode_count, void** decode_stream,
                         int* decode_count) {
  if (!e || decode_sms <= 0 || !encode_stream || !decode_stream) return GAR_ERR_ARG;
  cudaStream_t small = nullptr, rest = nullptr;
  int ns = 0, nr = 0;
  TRY(gar_make_sm_partition(decode_sms, &small, &ns, &rest, &nr));
  e->part_stream[0] = rest; e->part_sms[0] = nr;
  e->part_stream[1] = small; e->part_sms[1] = ns;
  *encode_stream = rest;
  *decode_stream = small;
  if (encode_count) *encode_count = nr;
  if (decode_count) *decode_count = ns;
  return GAR_OK;
}

int gar_engine_set_replay(gar_engine* e, int replay_mode, const long long* crop_ids, int n) {
  if (!e || replay_mode < GAR_REPLAY_MERGED || replay_mode > GAR_REPLAY_FRAME || n < 0 || n > 8 || (n > 0 && !crop_ids)) {
    gar_set_error("gar_engine_set_replay: mode=%d n=%d (at most 8 crop-token ids)", replay_mode, n);
    return GAR_ERR_ARG;
  }
  e->replay_mode = replay_mode;
  e->replay_n = n;
  for (int i = 0; i < 8; ++i) e->replay_ids[i] = i < n ? crop_ids[i] : -1;
  e->plan_slot0 = -1;  // a plan made under another mode is not reusable
  return GAR_OK;
}

int gar_engine_set_option(gar_engine* e, const char* name, int value) {
  if (!e || !name) return GAR_ERR_ARG;
  const std::string n(name);
  if (n == "skinny_gemm") e->use_skinny = value;
  else if (n == "fuse_rope") e->fuse_rope = value;
  else if (n == "gemv_decode") e->use_gemv = value;
  else if (n == "fuse_decode_rope") e->fuse_decode_rope = value;
  else if (n == "decode_splits" && value >= 0 && value <= 32) e->dec_splits = value;
  else if (n == "eos_poll_steps" && value >= 0) e->eos_poll_steps = value;
  else if (n == "pdl" || n == "decode_attn_mma" || n == "attn_bkv" || n == "attn_poly" || n == "attn_persist" || n == "gemm_group_kb" || n == "gemm_group_force" || n == "gemm_2cta" ||
           n == "gemv_norm_smem" || n == "gemv_wide") {
    if (gar_set_global_option(name, value)) return GAR_ERR_ARG;
  }
  else {
    gar_set_error("gar_engine_set_option: unknown option '%s'=%d", name, value);
    return GAR_ERR_ARG;
  }
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  e->graphs.clear();
  return GAR_OK;
}

int gar_engine_profile(gar_engine* e, int enable) {
  if (!e) return GAR_ERR_ARG;
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e->prof.clear();
  e->profiling = enable != 0;
  return GAR_OK;
}

// per kernel class k (GAR_PROF_*): ms[k] = summed CUDA-event durations, work[k] = summed algorithmic
// work (flops for GEMM/attention, bytes for assemble), count[k] = launches.  Synchronises the device.
int gar_engine_profile_read(gar_engine* e, double* ms, double* work, long long* count) {
  if (!e || !ms || !work || !count) return GAR_ERR_ARG;
  GAR_CUDA_CHECK(cudaDeviceSynchronize());
  for (int k = 0; k < PK_COUNT; ++k) { ms[k] = 0; work[k] = 0; count[k] = 0; }
  for (auto& r : e->prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    ms[r.kind] += t;
    work[r.kind] += r.work;
    count[r.kind] += 1;
  }
  return GAR_OK;
}

}  // extern "C"
