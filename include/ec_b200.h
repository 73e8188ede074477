/*
 * libec_b200 -- C ABI of the B200-native captioner hot path (sm_100a only).
 *
 * The reference (hsp-iit/embodied-captioning) has no FFI layer: its boundary for this path is a Python object
 * protocol.  Each entry point below names the reference interface it stands behind;
 * the Python shims in `embodied_captioning_b200/captioner/` keep those names and call these symbols through
 * ctypes on torch-owned device pointers.
 *
 * Conventions: every function returns 0 (EC_OK) or a negative EC_ERR_* code; `ec_last_error()` returns a
 * thread-local message.  All pointers named *_dev are device pointers on the model's device; `stream` is a
 * cudaStream_t passed as void*.  Hot calls never allocate: the caller passes a workspace of at least
 * `ec_*_workspace_bytes()` bytes.  Nothing here falls back to the CPU; a non-sm_100 device is EC_ERR_DEVICE.
 */
#ifndef EC_B200_H_
#define EC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EC_OK 0
#define EC_ERR_INVALID (-1)
#define EC_ERR_CUDA (-2)
#define EC_ERR_UNSUPPORTED (-3)
#define EC_ERR_WORKSPACE (-4)
#define EC_ERR_STATE (-5)
#define EC_ERR_DEVICE (-6)
#define EC_ERR_TIMEOUT (-7)

#define EC_ABI_VERSION 1

/* Architecture hyper-parameters; mirrors the JSON registry read by
 * experimenting_env/captioner/models/coca/factory.py:27-57 (model_configs/coca_ViT-L-14.json). */
typedef struct ec_config {
  int32_t embed_dim, image_size, patch_size;
  int32_t vision_width, vision_layers, vision_head_dim;
  int32_t pooler_heads, n_queries;
  int32_t text_width, text_heads, text_layers;
  int32_t mm_layers, mm_heads;
  int32_t context_length, vocab_size;
  int32_t sot_token_id, eos_token_id, pad_token_id;
} ec_config;

typedef struct ec_model ec_model;

int ec_abi_version(void);
const char* ec_last_error(void);

/* ---- model lifetime: replaces `create_model_and_transforms(...)` (factory.py:378-432) + `load_checkpoint`
 * (factory.py:131-180).  Tensors are passed by their open_clip state-dict names as host fp32 arrays. */
int ec_model_create(const ec_config* cfg, int device, ec_model** out);
int ec_model_destroy(ec_model* m);
int ec_model_set_tensor(ec_model* m, const char* name, const float* host_data, int64_t numel);
int ec_model_finalize(ec_model* m, void* stream);   /* checks completeness, pre-computes the pooler queries */
int ec_model_set_option(ec_model* m, const char* key, int value);   /* "decode_graph": 0/1 */

/* ---- crop + image transform: replaces `self.preprocess(PIL.Image)` at coca.py:28 (upstream
 * image_transform_v2) applied to the crops cut at goal_exploration.py:95-99 / pseudolabeler.py:670-675 /
 * evaluate_finetuned_model.py:33-37.  rects are integer source rectangles already computed by the box policy. */
int ec_preprocess_crops(const uint8_t* frames_dev, int32_t n_frames, int32_t frame_h, int32_t frame_w,
                        int32_t frame_c, const int32_t* rects_dev, const int32_t* frame_idx_dev,
                        const int32_t* flags_dev, int32_t n, int32_t out_size, void* out_bf16_dev,
                        float* out_f32_dev, int32_t* status_dev, void* stream);

/* ---- image tower: replaces `CoCa._encode_image` (coca_model.py:152-155). x: bf16 [n,3,S,S];
 * latent: f32 [n,E] (L2-normalised); embs: f32 [n, n_queries-1, E]. */
size_t ec_encode_workspace_bytes(const ec_model* m, int32_t n);
int ec_encode_image(ec_model* m, const void* x_bf16_dev, int32_t n, float* latent_dev, float* embs_dev, void* ws_dev,
                    size_t ws_bytes, void* stream);

/* ---- greedy decoder: replaces `CoCa.generate(..., generation_type='top_k')` with top_k=1
 * (coca_model.py:205-333, called at coca.py:29 and evaluate_finetuned_model.py:43).
 * tokens: int32 [n, seq_len] (col 0 = SOT, PAD after EOS); logits (optional): f32 [seq_len-1, n, vocab],
 * row b of step t valid while crop b was still active; n_steps_host (optional, host int): number of
 * columns the reference loop would have produced (it stops once every crop has emitted EOS). */
size_t ec_decode_workspace_bytes(const ec_model* m, int32_t n, int32_t seq_len);
int ec_decode_greedy(ec_model* m, const float* embs_dev, int32_t n, int32_t seq_len, int32_t min_seq_len,
                     int32_t* tokens_dev, float* logits_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ---- teacher-forced logits: replaces `CoCa.forward(image, text, image_embs=...)` (coca_model.py:170-203)
 * for output_labels=False: text int32 [n, t] -> logits f32 [t, n, vocab] (step-major; the shim permutes). */
int ec_forward_logits(ec_model* m, const float* embs_dev, const int32_t* text_dev, int32_t n, int32_t t,
                      float* logits_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ---- per-instance disagreement: replaces the dict-comprehension over `_cosine_distance`
 * (projection_utils.py:743-754, 928-942; mode 0) and `mean_cosine_similarity`
 * (scripts/compute_cosine_sim.py:11-22; mode 1) grouped by instance key (pseudocaptioner.py:137). */
/* OR-ed into `mode`: drop exact-duplicate view rows of an instance first, i.e. the `unique(dim=0)` that
 * SemanticPointCloud.preprocess applies to every object's embedding set (projection_utils.py:574-578);
 * out_count then holds the number of DISTINCT views.  Local rows only (not the peer-memory form). */
#define EC_DISAGREEMENT_DEDUP 4
size_t ec_disagreement_workspace_bytes(int32_t n);
int ec_instance_disagreement(const float* emb_dev, int64_t ld_emb, const int64_t* keys_dev, int32_t n, int32_t d,
                             int32_t mode, int64_t* out_keys_dev, float* out_score_dev, int32_t* out_count_dev,
                             int32_t* out_m_dev, void* ws_dev, size_t ws_bytes, void* stream);
/* multi-GPU form: rows live in `world` peer-mapped buffers (row i in peers[i / rows_per_rank]); this rank
 * reduces the instances m with m % part_world == part_rank and writes 0 for the others. */
int ec_instance_disagreement_peers(const float* const* peers_dev, int32_t world, int32_t rows_per_rank,
                                   int64_t ld_emb, const int64_t* keys_dev, int32_t n, int32_t d, int32_t mode,
                                   int32_t part_rank, int32_t part_world, int64_t* out_keys_dev,
                                   float* out_score_dev, int32_t* out_count_dev, int32_t* out_m_dev, void* ws_dev,
                                   size_t ws_bytes, void* stream);

/* ---- caption-text encoder (MiniLM-class BERT sentence encoder): replaces `self.encoder.encode(caption)` of
 * `SentenceTransformer("all-MiniLM-L6-v2")` (agents/goal_exploration/goal_exploration.py:57,102-103;
 * detector/pseudolabeler.py:568,651,677) for WordPiece ids that are already on the device.  Weights are loaded by
 * HF `BertModel` state-dict names ("embeddings.word_embeddings.weight", "encoder.layer.0.attention.self.query.weight",
 * ...; "bert." / "0.auto_model." prefixes accepted).  ids int32 [n, seq] (row = [CLS] ... [SEP] pad...), lens int32 [n]
 * = number of real tokens per row (the attention mask is a prefix mask); out f32 [n, hidden] = L2-normalised masked
 * mean of the last hidden state. */
typedef struct ec_textenc_config {
  int32_t vocab_size, hidden, layers, heads, intermediate, max_positions, type_vocab;
  float layer_norm_eps;
} ec_textenc_config;
typedef struct ec_textenc ec_textenc;
int ec_textenc_create(const ec_textenc_config* cfg, int device, ec_textenc** out);
int ec_textenc_destroy(ec_textenc* t);
int ec_textenc_set_tensor(ec_textenc* t, const char* name, const float* host_data, int64_t numel);
int ec_textenc_finalize(ec_textenc* t);
size_t ec_textenc_workspace_bytes(const ec_textenc* t, int32_t n, int32_t seq);
int ec_textenc_encode(ec_textenc* t, const int32_t* ids_dev, const int32_t* lens_dev, int32_t n, int32_t seq,
                      float* out_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ---- caption frequency table per instance: replaces `PseudoCaptioner.compute_captions_frequency`
 * (captioner/pseudocaptioner.py:156-177; grouping key :137) on token-id rows.  tokens int32 [n, ld] (first t columns
 * compared), keys int64 [n].  One entry per distinct (key, caption): out_key, out_first (row of its first occurrence),
 * out_count; ordered by (key, first occurrence) = dict insertion order inside an instance.  Synchronises `stream`. */
size_t ec_caption_frequency_workspace_bytes(int32_t n);
int ec_caption_frequency(const int32_t* tokens_dev, int32_t ld_tokens, int32_t t, const int64_t* keys_dev, int32_t n,
                         int64_t* out_key_dev, int32_t* out_first_dev, int32_t* out_count_dev, int32_t* out_m_dev,
                         void* ws_dev, size_t ws_bytes, void* stream);

/* ---- disagreement map + exploration reward: replaces the score gather / centre write / six "inflate" writes of
 * `SemanticPointCloud.get_topdown_semantic` (projection_utils.py:694-703,760-785) and `get_reward`
 * (envs/env_habitat3.py:75-84).  points f32 [P,3] (x -> column, z -> row), point_keys int64 [P] (object id per
 * point), inst_keys/inst_scores [m] = key-sorted output of ec_instance_disagreement.  off_x = floor(lower_x/scale),
 * off_y = ceil(lower_z/scale) (host integers, projection_utils.py:694-699).  map f32 [h,w] is fully overwritten with
 * numpy-assignment semantics (last write wins, negative indices wrap, an assignment with any out-of-range index is
 * dropped whole); reward_dev (nullable) = map.sum() / 1000. */
size_t ec_disagreement_map_workspace_bytes(int32_t n_points, int32_t h, int32_t w);
int ec_disagreement_map(const float* points_dev, int32_t n_points, const int64_t* point_keys_dev,
                        const int64_t* inst_keys_dev, const float* inst_scores_dev, int32_t m, float scale,
                        int32_t off_x, int32_t off_y, int32_t h, int32_t w, float* map_dev, float* reward_dev,
                        void* ws_dev, size_t ws_bytes, void* stream);

/* ---- building blocks exported for unit tests and micro-benchmarks (same kernels the model path launches) */
int ec_gemm_bf16(const void* a_bf16_dev, int32_t lda, const void* w_bf16_dev, int32_t ldw, int32_t m, int32_t n,
                 int32_t k, const float* bias_dev, const float* residual_dev, int32_t ldr, void* out_dev, int32_t ldo,
                 int32_t epilogue /* 0 bf16, 1 bf16+gelu, 2 f32, 3 f32+residual */, int32_t use_simt_reference,
                 void* stream);
/* residual GEMM (epilogue 3) that also writes ln_out_bf16[m, n] = LayerNorm(out rows; gamma, beta, eps 1e-5) as soon as
 * a 128-row block of `out` is complete (the fused tail the model path uses for every LayerNorm but the first);
 * ln_counters: >= ceil(m/128) int32, zero before the first call (re-armed by the kernel); n <= 1024. */
int ec_gemm_bf16_ln(const void* a_bf16_dev, int32_t lda, const void* w_bf16_dev, int32_t ldw, int32_t m, int32_t n,
                    int32_t k, const float* bias_dev, const float* residual_dev, int32_t ldr, float* out_dev, int32_t ldo,
                    const float* ln_gamma_dev, const float* ln_beta_dev, void* ln_out_bf16_dev, int32_t ld_ln,
                    int32_t* ln_counters_dev, void* stream);
int ec_gemm_argmax(const void* a_bf16_dev, int32_t lda, const void* w_bf16_dev, int32_t ldw, int32_t m, int32_t n,
                   int32_t k, float* logits_dev /* nullable [m,n] */, int32_t* argmax_dev /* [m] */, void* ws_dev,
                   size_t ws_bytes, void* stream);
int ec_layernorm(const float* x_dev, int64_t ldx, const float* gamma_dev, const float* beta_dev, void* y_dev,
                 int64_t ldy, int32_t y_is_bf16, int32_t rows, int32_t width, void* stream);
int ec_attention(const void* q_dev, int64_t q_bs, int64_t q_rs, const void* k_dev, int64_t k_bs, int64_t k_rs,
                 const void* v_dev, int64_t v_bs, int64_t v_rs, void* o_dev, int64_t o_bs, int64_t o_rs,
                 int32_t batch, int32_t heads, int32_t sq, int32_t skv, int32_t head_dim, void* stream);
/* q/k/v as column slices of one fused [batch*seq, ld] bf16 buffer (the layout the image tower uses);
 * impl 0 = mma.sync flash kernel, 1 = tcgen05/TMEM kernels (head_dim 64, seq <= 258; 257 tokens run the block-wise kernel),
 * 2 = the whole-row tcgen05 kernel for every length (A/B) */
int ec_attention_fused_qkv(const void* qkv_dev, int64_t ld, int32_t q_col0, int32_t k_col0, int32_t v_col0,
                           void* out_dev, int64_t o_bs, int64_t o_rs, int32_t batch, int32_t heads, int32_t seq,
                           int32_t head_dim, int32_t impl, void* stream);
/* residual GEMM (epilogue 3) with deterministic split-K: every 128 x 128 output tile is computed by `split_k` (2..4, 0 =
 * choose) adjacent CTAs over disjoint K ranges; partial accumulators travel through ws_dev ((split_k - 1) * m * n
 * floats) and are added in slice order (bit-reproducible).  counters_dev: >= ceil(m/128) * ceil(n/128) int32, zero
 * before the first call (re-armed by the kernel).  Used by the decoder's c_proj GEMMs (mlp.c_proj of every
 * ResidualAttentionBlock: M = crops, K = 4 * width).  Needs tiles * split_k <= SM count and (k / 64) % split_k == 0. */
int ec_gemm_bf16_splitk(const void* a_bf16_dev, int32_t lda, const void* w_bf16_dev, int32_t ldw, int32_t m, int32_t n,
                        int32_t k, const float* bias_dev, const float* residual_dev, int32_t ldr, float* out_dev,
                        int32_t ldo, int32_t split_k, float* ws_dev, int32_t* counters_dev, void* stream);
/* grouped (block-diagonal) GEMM: out[m, n] = sum_k A[m, (n / group_cols) * k_per_group + k] * W[n, k]; epilogue 0 or 2.
 * The decoder's absorbed cross-attention uses it for the per-head projections on either side of
 * ec_cross_attention_absorbed (reference: the k/v in-projections of text_decoder.cross_attn.{l}.attn, an
 * nn.MultiheadAttention, coca_model.py:130-135 -> upstream MultimodalTransformer). */
int ec_gemm_bf16_grouped(const void* a_bf16_dev, int32_t lda, const void* w_bf16_dev, int32_t ldw, int32_t m, int32_t n,
                         int32_t k_per_group, int32_t group_cols, const float* bias_dev, void* out_dev, int32_t ldo,
                         int32_t epilogue, int32_t use_simt_reference, void* stream);
/* image tokens in the layout the absorbed cross-attention streams: per crop ceil(n_kv / 32) chunks of 32 rows, each row
 * `width` bf16 + 8 zero elements, rows past n_kv zero; normalize != 0 applies (x - mean) * rstd (eps 1e-5), i.e. the
 * ln_1_kv LayerNorm without its affine part (text_decoder.cross_attn.{l}.ln_1_kv; gamma / beta are folded into the
 * projection weights when the model is finalized).  x f32 [n, n_kv, width]. */
size_t ec_packed_image_tokens_bytes(int32_t n, int32_t n_kv, int32_t width);
int ec_pack_image_tokens(const float* x_dev, void* out_bf16_dev, int32_t n, int32_t n_kv, int32_t width,
                         int32_t normalize, void* stream);
/* single-query attention whose keys AND values are the rows of xn (the packed image tokens of a crop):
 * ctx[b, h, :] = sum_j softmax_j(qp[b, h, :] . xn[b, j, :]) xn[b, j, :].  qp, ctx bf16 [batch, heads * width];
 * xn = ec_pack_image_tokens output; width in {256, 512, 768}, heads <= 16; finished (nullable) uint8 [batch] skips rows. */
int ec_cross_attention_absorbed(const void* qp_dev, const void* xn_dev, void* ctx_dev, int32_t batch, int32_t heads,
                                int32_t width, int32_t n_kv, const uint8_t* finished_dev,
                                int32_t reverse /* walk the crops last to first: same result, other L2 reuse */, void* stream);
int ec_check_watchdog(void* stream);   /* EC_ERR_TIMEOUT if a tcgen05 / bulk-copy pipeline wait ever timed out */

/* ---- measurement hooks used by bench.py */
int64_t ec_launch_count(void);          /* kernels launched by this library so far (graph replays included) */
int ec_profile_gemm(int32_t enable);    /* time every non-captured tcgen05 GEMM launch with CUDA events */
int ec_profile_gemm_read(double* ms, double* flops, int64_t* launches);   /* sums since the last read; resets */
/* Stage clocks of the decoder's GEMM-sequence kernel (csrc/gemm_seq.cuh), recorded by CTA 0 in builds with
 * -DEC_SEQ_TIMING (zeros otherwise): [64 launches of a step][6 stages][16 events] SM cycles since kernel entry. */
int ec_seq_timing_read(int64_t* out_host, int32_t count);
/* The same for the stand-alone small-M GEMM kernel (-DEC_GEMM_TIMING): a ring of 8192 records x 10 (gemm_kernel.cuh). */
int ec_gemm_timing_read(int64_t* out_host, int32_t count);

#ifdef __cplusplus
}
#endif
#endif /* EC_B200_H_ */
