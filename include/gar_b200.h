/* gar_b200.h -- C ABI of libgar_b200.so: the B200 (sm_100a) implementation of Grasp-Any-Region's
 * region-captioning inference hot path.
 *
 * The reference (Haochen-Wang409/Grasp-Any-Region) is pure Python and has no FFI; this is the
 * boundary a maintainer binds from `projects/grasp_any_region/hf_models/modeling_gar.py`
 * (see INTEGRATION.md for the ctypes stub).  Each entry point cites the reference code it replaces.
 *
 * Conventions: every pointer named *_dev / taking activations or weights is a CUDA device pointer to
 * bf16 data unless stated otherwise; `stream` is a cudaStream_t passed as void*; all functions
 * return 0 on success or a GAR_ERR_* code, with a message available from gar_last_error().
 * Nothing here falls back to the CPU: without a CUDA device every compute call fails.
 */
#ifndef GAR_B200_H
#define GAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAR_OK 0
#define GAR_ERR_ARG 1
#define GAR_ERR_CUDA 2
#define GAR_ERR_SHAPE 3
#define GAR_ERR_STATE 4

/* assemble status bits (gar_engine_status): the reference's Python exceptions */
#define GAR_STATUS_IMAGE_TOKEN_MISMATCH 1 /* ValueError modeling_perception_lm.py:309-315 */
#define GAR_STATUS_BAD_CROP_RUN 2         /* run != 256 contiguous tokens, modeling_gar.py:357-411 */
#define GAR_STATUS_MISSING_BBOX 4         /* KeyError bboxes[str(crop_token)], modeling_gar.py:367 */

const char* gar_last_error(void);
int gar_abi_version(void);
int gar_copy_d2d(void* dst, const void* src, long long bytes, void* stream);
/* process-wide kernel-selection switches for A/B runs: "pdl" (default 1), "decode_attn_mma" (default 1) */
int gar_set_global_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------
 * Per-kernel entry points (unit parity; the engine below is built from exactly these)
 * ---------------------------------------------------------------------------------------- */

/* GEMM epilogues, reproducing the bf16 rounding points of the torch ops they fuse */
#define GAR_EPI_NONE 0        /* C = bf16(A B^T)                                  nn.Linear(bias=False) */
#define GAR_EPI_BIAS 1        /* C = bf16(A B^T + bias)                           nn.Linear */
#define GAR_EPI_BIAS_GELU 2   /* C = bf16(gelu_erf(bf16(A B^T + bias)))           Linear -> GELU (timm Mlp, projector) */
#define GAR_EPI_BIAS_LS_RES 3 /* C = bf16(res + bf16(gamma * bf16(A B^T + bias))) EvaBlock x + gamma * f(x) */
#define GAR_EPI_RES 4         /* C = bf16(res + bf16(A B^T))                      Llama h + o_proj/down_proj */
#define GAR_EPI_SWIGLU 6      /* C[M,N/2] = bf16(bf16(silu(g)) * u), B rows interleaved [32 gate | 32 up] (LlamaMLP) */

/* C[M,N] = epi(A[M,K] * B[N,K]^T): tcgen05 / TMEM / TMA GEMM.  Replaces the cuBLAS GEMMs behind every
 * nn.Linear and both k=s=14 convs of the path (timm EvaBlock, modeling_perception_lm.py:85-92,
 * modeling_gar.py:326-328, HF LlamaDecoderLayer).  K % 64 == 0, N % 32 == 0, ld* % 8 == 0.
 * block_n_hint: 0 = auto, or 64/128/256. */
int gar_gemm_bf16(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K, int epi,
                  const void* bias, const void* gamma, const void* res, int ldres, int block_n_hint, void* stream);

/* Host-only (no CUDA call): the tile walk gar_gemm_bf16 uses for a launch of this geometry -- tile_m x block_n output
 * tiles (256-row tiles for the CTA-pair kernel), `fly` tiles in flight.  Column panels are walked in groups of
 * *n_group panels sized to keep the operands in flight L2-resident (global option gemm_group_kb, default 40 MB of
 * working set; 0 = one group); inside a group n runs fastest, then m.  If m_blk / n_blk are given, also maps `tile`. */
int gar_gemm_raster_probe(int M, int N, int K, int tile_m, int block_n, int fly, int tile, int* n_group, int* m_blk,
                          int* n_blk);

/* Same GEMM with timm's apply_rot_embed_cat fused after the bias (PE-ViT qkv projection): columns < rope_cols of
 * rows with (row % ntok) >= npt are rotated with the sin / cos values of token (row % ntok - npt).  `rope` is the
 * CHUNK-MAJOR form of the (ntok - npt, 2 * hd) table cat(sin[hd], cos[hd]) that gar_vit_rope takes:
 * rope[((k * hd / 8 + j) * (ntok - npt) + t) * 8 + i] = table[t][k * hd + 8 * j + i], k = 0 sin / 1 cos -- the 32 rows
 * of an epilogue warp then read contiguous memory. */
int gar_gemm_bf16_rope2d(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                         const void* bias, const void* rope, int ntok, int npt, int hd, int rope_cols, void* stream);

/* out[M<=16,N] = epi(X[M,K] W[N,K]^T), epi in {GAR_EPI_NONE, GAR_EPI_RES}: swap-AB tcgen05 weight-streaming GEMM with
 * cluster split-K reduced through distributed shared memory (the decode step's cuBLAS GEMV-like calls). */
int gar_gemm_skinny_bf16(const void* X, int ldx, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                         int epi, const void* res, int ldres, void* stream);

/* Decode-step linear layer, latency-first (mma.sync weight stream, no TMEM/cluster/TMA setup):
 * out[M<=8,N] = epi(norm?(X) W^T).  epi 0 none | 1 out = bf16(res + bf16(acc)), optionally emitting per-CTA sums of
 * squares of out into sumsq_out[N/rows_per_cta][8] | 2 SwiGLU on W rows interleaved [32 gate | 32 up], out [M,N/2].
 * norm_w != NULL applies LlamaRMSNorm to X on the fly using sumsq_in[n_part][8] (partials emitted by the producer).
 * Returns GAR_ERR_SHAPE if K % 256 or N granularity does not fit (use gar_gemm_skinny_bf16 then). */
int gar_gemv_decode(const void* X, int ldx, const void* W, int ldw, void* out, int ldo, int M, int N, int K, int epi,
                    const void* res, int ldres, const void* norm_w, const float* sumsq_in, int n_part, float eps,
                    float* sumsq_out, int* rows_per_cta, void* stream);
/* debug: per-CTA globaltimer stamps of the following gar_gemv_decode launches into dev_buf (int64[8 * grid]); NULL disables */
int gar_gemv_debug_timeline(void* dev_buf);
int gar_embed_gather_sumsq(const void* table, const int* ids32, void* out, float* sumsq, int rows, int D, void* stream);

/* mask-id decode + im2col of both stride-14 convs (modeling_gar.py:315-328; timm PatchEmbed) */
int gar_patchify(const void* pixel_values, const void* mask_values, void* a_img, void* a_mask, int T, int img,
                 int patch, int ldk, int prompt_numbers, void* stream);
/* cls concat + pos_embed add + norm_pre (timm Eva._pos_embed / norm_pre; modeling_perception_lm.py:197-198) */
int gar_vit_embed_ln(const void* patch, const void* cls, const void* pos, const void* w, const void* b, void* x, int T,
                     int ntok, int npt, int d, float eps, void* stream);
int gar_layernorm(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int rows, int d, float eps,
                  void* stream);
/* LlamaRMSNorm (HF modeling_llama.py:53-70) */
int gar_rmsnorm(const void* x, int ldx, const void* w, void* y, int ldy, int rows, int d, float eps, void* stream);
/* timm apply_rot_embed_cat on q,k of the fused qkv buffer, in place (prefix tokens untouched) */
int gar_vit_rope(void* qkv, const void* rope, int rows, int ntok, int npt, int heads, int hd, void* stream);
/* HF apply_rotary_pos_emb on q (in place) and k (into the KV cache) + V append (DynamicCache.update) */
/* seq_len > 0: rows hold nreq sequences of seq_len positions each, written to slots slot0, slot0+1, ... */
int gar_llama_rope_kv(void* qkv, int rows, int hq, int hkv, int hd, const void* cos_t, const void* sin_t, int pos0,
                      const int* pos_dev, const int* slot_dev, int slot0, void* kcache, void* vcache, int max_seq,
                      int seq_len, void* stream);
/* HF LlamaMLP act_fn(gate) * up on a fused buffer: rows = [gate(I) | up(I)], or with interleaved != 0
 * 32-column groups [gate | up] (the GAR_EPI_SWIGLU weight layout) */
int gar_swiglu(const void* gu, void* act, int rows, int inter, int interleaved, void* stream);
/* PerceptionLMAdaptiveAvgPooling 2x2 (+ cls drop), modeling_perception_lm.py:42-60,263-264 */
int gar_pool2x2(const void* x, void* y, int T, int ntok, int npt, int grid, int D, void* stream);
int gar_embed_gather(const void* table, const long long* ids64, const int* ids32, void* out, int rows, int D,
                     void* stream);
/* greedy argmax + device-side decode bookkeeping (HF GenerationMixin greedy step) */
/* step_ptr: 2 ints {step, ticket}; bump_step: the last block increments step (no extra launch);
 * partial_ws: batch*ceil(V/4096)*8 bytes for the two-stage reduction (NULL = single-stage scan). */
int gar_argmax_step(const void* logits, int ld, int V, int batch, int* next_tok, int* out_tokens, int out_ld,
                    int* step_ptr, int* pos, int* finished, int eos, int advance_pos, int bump_step, void* partial_ws,
                    void* stream);
/* same with up to four stop ids (GenerationConfig.eos_token_id may be a list); eos_ids is a HOST array */
int gar_argmax_step_eos(const void* logits, int ld, int V, int batch, int* next_tok, int* out_tokens, int out_ld,
                        int* step_ptr, int* pos, int* finished, const int* eos_ids, int n_eos, int advance_pos,
                        int bump_step, void* partial_ws, void* stream);
int gar_bump_step(int* step, void* stream);

/* Which reference replay loop the assemble kernels reproduce (replay_mode):
 *   GAR_REPLAY_MERGED      GARModel.generate, hf_models/modeling_gar.py:348-414: every crop token present is replayed
 *                          from the merged map of tiles 1..T-1 (thumbnail dropped), on the running tensor
 *   GAR_REPLAY_MERGED_LAST the vendored PLM's prepare_inputs_for_generation(feature_replay=True),
 *                          models/modeling/modeling_perception_lm.py:648-763: same map and boxes, but each replay is
 *                          spliced into the ORIGINAL embeddings (:736-742), so only the LAST crop token present keeps
 *                          its replay; the other runs keep embed_tokens(crop id)
 *   GAR_REPLAY_FRAME       prepare_inputs_for_generation(feature_replay_video=True), :765-852: crop id k is replayed
 *                          from tile (frame) k's own tpt x tpt map, no thumbnail drop, running tensor; ncw/nch ignored */
enum { GAR_REPLAY_MERGED = 0, GAR_REPLAY_MERGED_LAST = 1, GAR_REPLAY_FRAME = 2 };

/* Fused embed + image-feature scatter + RoI-aligned feature replay (modeling_gar.py:332-414).
 * rois: n_regions x [0,x1,y1,x2,y2] fp32 host array computed like modeling_gar.py:367-387;
 * plan_ws: S + 8 ints (device); status_dev: 2 ints (device): [status bits, n image tokens]. */
int gar_assemble_embeds(const long long* input_ids_dev, int S, const void* embed_table, int vocab,
                        const void* image_features, int n_image_rows, int D, const long long* crop_ids,
                        const float* rois, const int* has_roi, int n_regions, float spatial_scale, int ncw, int nch,
                        int tpt, long long image_token_id, int replay_mode, void* out, int* plan_ws, int* status_dev,
                        void* stream);

/* The same for nreq equal-length requests (one plan launch, then the replay and gather launches side by side: an HBM
 * stream over all requests).
 * input_ids_dev: HOST array of nreq device pointers; rois (nreq, n_regions, 5), has_roi (nreq, n_regions), ncw/nch
 * (nreq) host arrays; request r uses features image_features + r*n_image_rows*D and writes rows [r*S,(r+1)*S) of out
 * and plan_ws (nreq*S + 8*nreq ints) and status_dev[2r..2r+1]. */
int gar_assemble_embeds_batch(int nreq, const long long* const* input_ids_dev, int S, const void* embed_table,
                              int vocab, const void* image_features, int n_image_rows, int D,
                              const long long* crop_ids, const float* rois, const int* has_roi, int n_regions,
                              float spatial_scale, const int* ncw, const int* nch, int tpt, long long image_token_id,
                              int replay_mode, void* out, int* plan_ws, int* status_dev, void* stream);
/* Its two halves.  The plan (row kinds, image ordinals, crop bins, status bits) depends only on input_ids, so it can run
 * before the image features exist and its status word can be read early: the reference raises its ValueError
 * (modeling_perception_lm.py:309-315) before the language model runs, and so does a caller of these two. */
int gar_assemble_plan_batch(int nreq, const long long* const* input_ids_dev, int S, int n_image_rows,
                            const long long* crop_ids, const int* has_roi, int n_regions, int tpt,
                            long long image_token_id, int replay_mode, int* plan_ws, int* status_dev, void* stream);
int gar_assemble_rows_batch(int nreq, const long long* const* input_ids_dev, int S, const void* embed_table, int vocab,
                            const void* image_features, int n_image_rows, int D, const long long* crop_ids,
                            const float* rois, const int* has_roi, int n_regions, float spatial_scale, const int* ncw,
                            const int* nch, int tpt, long long image_token_id, int replay_mode, void* out,
                            const int* plan_ws, void* stream);

/* ---- request preprocessing on the device: uint8 image + uint8 region-id map -> bf16 tile tensors ----
 * One separable pass of ATen's native uint8 antialiased bicubic resize (what torchvision v2 `resize(..., BICUBIC)` runs on
 * uint8 CPU tensors; image_processing_perception_lm_fast.py:254-277, 296-340): out = clamp((2^(precision-1) +
 * sum_j weights[o][j] * src[idx_min[o] + j]) >> precision, 0, 255).  horizontal != 0 resizes x (dst (planes, in_h,
 * out_size)), else y (dst (planes, out_size, in_w)).  idx_min / idx_size (out_size ints) and weights (out_size x max_interp
 * int16) are DEVICE tables built on the host in double precision (preprocess_device.aa_tables). */
int gar_resize_aa_u8_pass(const void* src_u8, int planes, int in_h, int in_w, int horizontal, int out_size,
                          const int* idx_min, const int* idx_size, const short* weights, int max_interp, int precision,
                          void* dst_u8, void* stream);
/* uint8 canvas (3, nch*tile, ncw*tile) -> bf16 tiles (nch*ncw, 3, tile, tile), row-major over (tile row, tile column),
 * values ((v * fl(1/255)) - 0.5) / 0.5 in torch's fp32 op order (`_split` + rescale + normalise, :316-372) */
int gar_u8_to_tiles(const void* canvas_u8, int ncw, int nch, int tile, void* tiles_out, void* stream);
/* uint8 id map (H, W) -> global_mask_values (1 + ncw*nch, 3, tile, tile) bf16 = 2 v / 255 - 1, thumbnail first, NEAREST
 * through host-built DEVICE index tables (ys/xs of the thumbnail: `tile` ints each; of the canvas: nch*tile / ncw*tile) */
int gar_idmap_to_tiles(const void* idmap_u8, int H, int W, const int* ys_thumb, const int* xs_thumb, const int* ys_canvas,
                       const int* xs_canvas, int ncw, int nch, int tile, void* mask_values_out, void* stream);

/* softmax(QK^T*scale)V, head_dim 64, tcgen05 + TMEM (timm SDPA non-causal; flash-attn 2 causal GQA prefill).
 * All strides/offsets in bf16 elements.  K/V batch coordinate = b*kv_batch_mul + kv_head*kv_batch_head_mul. */
int gar_attn_fwd(const void* q, long long q_row_stride, long long q_batch_stride, int q_col0, const void* k,
                 const void* v, long long kv_row_stride, long long kv_batch_stride, int kv_batches, int k_col0,
                 int v_col0, int kv_head_cols, int kv_batch_mul, int kv_batch_head_mul, void* o, long long o_row_stride,
                 long long o_batch_stride, int o_col0, int batch, int heads, int kv_heads, int head_dim, int q_len,
                 int kv_len, int kv_rows_alloc, int causal, float scale, void* stream);
/* debug: clock64 stamps of one CTA's softmax phases into dev_buf (int64[8*kv_blocks]); NULL disables */
int gar_attn_debug_timeline(void* dev_buf);
/* q_len = 1 attention over the KV cache, split-KV + combine (flash-attn 2 decode + DynamicCache) */
/* tickets: batch*hkv zero-initialised ints -> the split kernel's last CTA combines (one launch); NULL -> separate combine */
int gar_decode_attn(const void* qkv, int ldq, const void* kcache, const void* vcache, const int* pos_dev,
                    const int* slot_dev, int batch, int hq, int hkv, int head_dim, int max_seq, int splits, float scale,
                    void* o_part_ws, void* ml_part_ws, int* tickets, void* out, int ldo, void* stream);

/* decode attention with HF apply_rotary_pos_emb + DynamicCache.update fused in (un-rotated q/k/v rows in `qkv`) */
int gar_decode_attn_rope(const void* qkv, int ldq, void* kcache, void* vcache, const int* pos_dev, const int* slot_dev,
                         int batch, int hq, int hkv, int head_dim, int max_seq, int splits, float scale,
                         const void* cos_t, const void* sin_t, void* o_part_ws, void* ml_part_ws, int* tickets,
                         void* out, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------
 * Engine: GARModel.generate() (modeling_gar.py:295-428) as a device-resident pipeline
 * ---------------------------------------------------------------------------------------- */
typedef struct gar_engine gar_engine;

typedef struct gar_model_desc {
  /* PE ViT (timm Eva) */
  int vit_dim, vit_depth, vit_heads, vit_mlp, vit_use_cls, img_size, patch_size;
  float ln_eps;
  /* projector */
  int pool_ratio;
  /* Llama */
  int llm_dim, llm_layers, llm_heads, llm_kv_heads, llm_head_dim, llm_inter, vocab, vocab_padded;
  float rms_eps;
  /* GAR */
  int prompt_numbers;
  long long image_token_id;
  long long crop_token_ids[8];
  /* capacity */
  int max_tiles, max_seq, max_batch, max_new_tokens;
  int max_prefill_rows; /* rows of one (batched) prefill pass; 0 = max_seq */
} gar_model_desc;

int gar_engine_create(const gar_model_desc* desc, gar_engine** out);
int gar_engine_destroy(gar_engine* e);
/* weights stay owned by the caller (torch tensors); names listed in DESIGN.md §4 */
int gar_engine_set_weight(gar_engine* e, const char* name, const void* dev_ptr);
/* internal buffers for parity taps: "image_features", "inputs_embeds", "logits", "mask_embeds", "vit_x" ... */
int gar_engine_get_buffer(gar_engine* e, const char* name, void** dev_ptr, long long* bytes);

/* R2-R5: pixel_values/global_mask_values (T,3,img,img) -> image features (T, tokens_per_tile, llm_dim) */
int gar_engine_vit_encode(gar_engine* e, const void* pixel_values, const void* mask_values, int T, void* stream);
/* The same in two steps, so that the requests of a batched pass are read from their own tensors (no concatenated copy):
 * patchify_at does R2 + the im2col of T tiles of one request into tiles [tile0, tile0+T) of the engine's operand arena;
 * vit_encode_patches then runs R3-R5 over the first T_total tiles. */
int gar_engine_patchify_at(gar_engine* e, const void* pixel_values, const void* mask_values, int T, int tile0,
                           void* stream);
int gar_engine_vit_encode_patches(gar_engine* e, int T_total, int with_mask, void* stream);
/* R6-R7 on the engine's image features -> inputs_embeds (S, llm_dim) */
int gar_engine_assemble(gar_engine* e, int slot, const long long* input_ids_dev, int S, int T, const float* rois,
                        const int* has_roi, float spatial_scale, int ncw, int nch, void* stream);
/* R6-R7 for request r of a batched pass: features start at tile feat_tile0, rows are written at emb_row0 */
int gar_engine_assemble_at(gar_engine* e, int slot, const long long* input_ids_dev, int S, int T, int feat_tile0,
                           int emb_row0, const float* rois, const int* has_roi, float spatial_scale, int ncw, int nch,
                           void* stream);
/* R6-R7 for all nreq requests of a batched pass at once (request r: features from tile r*T, rows [r*S,(r+1)*S),
 * status slot slot0+r); input_ids_dev is a HOST array of device pointers */
int gar_engine_assemble_batch(gar_engine* e, int slot0, int nreq, const long long* const* input_ids_dev, int S, int T,
                              const float* rois, const int* has_roi, float spatial_scale, const int* ncw,
                              const int* nch, void* stream);
/* R6 row plan + validation only, for the requests a following gar_engine_assemble_batch(slot0, nreq, S, T) will
 * assemble.  Needs only input_ids and has_roi (nreq x prompt_numbers), so it runs BEFORE gar_engine_vit_encode;
 * synchronises `stream` and returns one GAR_STATUS_* word per request in status_host (may be NULL): callers raise the
 * reference's ValueError / KeyError before any feature or logit is computed. */
int gar_engine_plan_batch(gar_engine* e, int slot0, int nreq, const long long* const* input_ids_dev, int S, int T,
                          const int* has_roi, int* status_host, void* stream);
/* R8 for nreq equal-length requests at once (rows = nreq*S of inputs_embeds, slots slot0..slot0+nreq-1): same math per
 * request as gar_engine_prefill, 1/nreq of the launches */
int gar_engine_prefill_batch(gar_engine* e, int slot0, int nreq, int S, void* stream);
/* R8: causal prefill of the engine's inputs_embeds into KV-cache slot; last-position logits -> logits[slot] */
int gar_engine_prefill(gar_engine* e, int slot, int S, void* stream);
/* same, teacher-forced: all-position logits (S, vocab_padded) into caller memory (forward(mode="loss")) */
int gar_engine_prefill_all_logits(gar_engine* e, int slot, int S, void* logits_out, void* stream);
/* R9: greedy decode of slots [0,batch): first token from the prefill logits, then max_new_tokens-1 steps
 * (CUDA-graph replay, no host sync inside).  out_tokens_host: batch x max_new_tokens int32 (-1 after EOS).
 * eos < 0 disables EOS. */
int gar_engine_generate(gar_engine* e, int batch, int max_new_tokens, int eos, int* out_tokens_host, void* stream);
/* same for KV slots [slot0, slot0+batch): lets one group decode on its own stream while another group's
 * ViT/assemble/prefill runs on a second stream (pipelined serving; bench.py) */
int gar_engine_generate_group(gar_engine* e, int slot0, int batch, int max_new_tokens, int eos, int* out_tokens_host,
                              void* stream);
/* R9 with HF's stop-id list: eos_ids (HOST array, n_eos <= 4; n_eos = 0 disables).  A sequence ends at the first stop id
 * (which is emitted, like HF); once every sequence of the group has ended the engine stops launching decode steps (the
 * device-side latches are polled every "eos_poll_steps" graph replays).  steps_run (may be NULL): decode steps launched. */
int gar_engine_generate_eos(gar_engine* e, int slot0, int batch, int max_new_tokens, const int* eos_ids, int n_eos,
                            int* out_tokens_host, int* steps_run, void* stream);
int gar_engine_status(gar_engine* e, int slot, int* status_bits);
long long gar_engine_launch_count(gar_engine* e);
/* SM partitioning for pipelined serving: creates two streams bound to disjoint SM sets (CUDA green contexts) -- a decode
 * partition of >= decode_sms SMs (driver granularity: multiples of 8) and an encode partition with the rest.  Passing
 * these streams to the stage / generate entry points confines their kernels to the partition; persistent kernels size
 * their grids to it.  Results are identical to the unpartitioned run.  GAR_ERR_CUDA if the driver cannot partition. */
int gar_engine_partition(gar_engine* e, int decode_sms, void** encode_stream, int* encode_sms, void** decode_stream,
                         int* decode_sms_out);
/* Which replay loop (GAR_REPLAY_*) and which crop-token ids (HOST array, n <= 8; NULL = the model's crop_tokens_ids)
 * the following plan / assemble calls use; (GAR_REPLAY_MERGED, NULL, 0) restores GARModel.generate's behaviour.  The
 * rois / has_roi arrays of those calls are indexed by this id list.  In GAR_REPLAY_FRAME the T tiles are T frames:
 * crop id k replays from tile k and the ncw x nch == T-1 check of the merged map does not apply. */
int gar_engine_set_replay(gar_engine* e, int replay_mode, const long long* crop_ids, int n);
/* options: "skinny_gemm" (1 = decode GEMMs use gar_gemm_skinny_bf16, default 1), "decode_splits" */
int gar_engine_set_option(gar_engine* e, const char* name, int value);
/* optional per-kernel-class CUDA-event timing for bench.py's roofline (decode runs ungraphed while on) */
#define GAR_PROF_GEMM 0
#define GAR_PROF_ATTN 1
#define GAR_PROF_ASSEMBLE 2
#define GAR_PROF_DECODE_ATTN 3
#define GAR_PROF_GEMV 4 /* decode weight-stream kernels; work = weight bytes */
#define GAR_PROF_PLAN 5 /* row plan + validation of the prompts (runs before the ViT) */
#define GAR_PROF_COUNT 6
int gar_engine_profile(gar_engine* e, int enable);
int gar_engine_profile_read(gar_engine* e, double* ms, double* work, long long* count);

#ifdef __cplusplus
}
#endif
#endif /* GAR_B200_H */
