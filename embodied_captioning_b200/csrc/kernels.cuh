// Launch wrappers of the non-GEMM kernels of libec_b200 (all sm_100a, all asynchronous on `stream`).
#pragma once
#include "common.cuh"

namespace ec {

// ---- layernorm.cu -------------------------------------------------------------------------------
// y[r,:] = (x[r,:] - mean) * rstd * gamma + beta, fp32 statistics (eps 1e-5 everywhere in open_clip).
// x fp32 [rows, ldx]; y bf16 or fp32 [rows, ldy]; width % 4 == 0, width <= 1024.
int layernorm(const float* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy,
              bool y_is_bf16, int rows, int width, const int* skip_flag, cudaStream_t stream);

// Fold a LayerNorm's affine part into the nn.Linear that follows it:  LN(x) @ W^T + b = xhat @ (W * gamma)^T + (b + W @ beta)
// with xhat = (x - mean) * rstd.  W [n_out, k] bf16, gamma / beta [k], bias [n_out] -> w_fold bf16 [n_out, k], b_fold [n_out].
int fold_layernorm_into_linear(const bf16* w, const float* bias, const float* gamma, const float* beta, bf16* w_fold,
                               float* b_fold, int n_out, int k, cudaStream_t stream);
// out[i] = value
int fill_f32(float* out, int n, float value, cudaStream_t stream);

// Vision stem: x[b,0,:] = cls + pos[0]; x[b,1+p,:] = patch[b,p,:] + pos[1+p]; then ln_pre -> fp32 [B*tokens, W].
int vision_embed_ln_pre(const float* patch_emb, const float* cls, const float* pos, const float* gamma,
                        const float* beta, float* out, int batch, int tokens, int width, cudaStream_t stream);

// [N,3,S,S] bf16 -> [N*G*G, ld] bf16 patches in conv-weight (c,py,px) order (ld >= 3*P*P, pad untouched).
int im2col_patches(const bf16* img, bf16* out, int n, int size, int patch, int ld, cudaStream_t stream);

// L2-normalise rows: y = x / max(||x||, 1e-12)   (F.normalize, coca_model.py:154)
int l2_normalize_rows(const float* x, long long ldx, float* y, long long ldy, int rows, int width,
                      cudaStream_t stream);

// fp32 -> bf16 copy with strides (used to gather the pooled row x[:,0] for the projection GEMM)
int convert_rows_f32_to_bf16(const float* x, long long ldx, bf16* y, long long ldy, int rows, int width,
                             cudaStream_t stream);

// ---- attention.cu (encoder / pooler, tensor cores via mma.sync, non-causal) -------------------------
struct AttnArgs {
  const bf16* q; long long q_bs, q_rs;   // element strides: batch, row; head h starts at column h*D
  const bf16* k; long long k_bs, k_rs;
  const bf16* v; long long v_bs, v_rs;
  bf16* o; long long o_bs, o_rs;
  int batch, heads, sq, skv, head_dim;  // head_dim in {32, 64, 96}
  float scale;                          // 1/sqrt(head_dim)
  const int* kv_len = nullptr;          // optional per-batch key count (<= skv): padded text batches
};
int attention_full(const AttnArgs& a, cudaStream_t stream);

// ---- attention_sm100.cu (ViT blocks: tcgen05.mma with S, P and O in TMEM; q/k/v are column slices of one
// fused [batch*seq, ld] bf16 buffer; head_dim 64, seq <= 258) ----------------------------------------
bool attention_sm100_supported(int seq, int head_dim);
// seq == 257 runs the block-wise kernel (two 128-key blocks per query tile, the 257th query on a spare mma.sync warp);
// `force_v1` keeps the whole-row kernel for A/B runs
int attention_sm100_fused_qkv(const bf16* qkv, long long ld, int q_col0, int k_col0, int v_col0, bf16* out,
                              long long o_bs, long long o_rs, int batch, int heads, int seq, cudaStream_t stream,
                              bool force_v1 = false);
int attention_sm100_check_watchdog(cudaStream_t stream);

// ---- decode.cu (KV-cached greedy decoder, one new token per crop per step) -----------------------
struct DecodeState {          // device-resident step state (lets one CUDA graph serve every step)
  int cur_len;                // prefix length incl. SOT == index of the token being produced
  int n_active;               // rows still generating after the last select
  int done;                   // 1 when n_active == 0 (kernels early-out)
  int next_active;            // accumulator for the running select
};

// x[b,:] = tok_emb[tokens[b, cur_len-1], :] + pos_emb[cur_len-1, :]
int decode_embed(const int* tokens, int ld_tokens, const float* tok_emb, const float* pos_emb, float* x, int batch,
                 int width, const DecodeState* st, cudaStream_t stream);

// causal self-attention of the new token against the per-layer cache; appends k,v at position cur_len-1.
// qkv bf16 [B, 3C]; kcache/vcache bf16 [B, max_len, C]; out bf16 [B, C]
int decode_self_attention(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, int batch, int heads, int head_dim,
                          int max_len, const unsigned char* finished, const DecodeState* st, cudaStream_t stream);

// cross-attention of the new token over the image K/V cache: q bf16 [B,C]; kv bf16 [B, n_kv, 2C] (K | V)
int decode_cross_attention(const bf16* q, const bf16* kv, bf16* out, int batch, int heads, int head_dim, int n_kv,
                           const unsigned char* finished, const DecodeState* st, cudaStream_t stream);

// ---- cross_absorbed.cu: absorbed-projection cross-attention (keys = values = the normalised image tokens) ---------
// qp bf16 [B, heads*T] (absorbed queries W_k,h^T q_h / sqrt(d)), xn bf16 = the crop's image tokens in the PACKED layout
// of pack_image_tokens (shared by all layers), ctx bf16 [B, heads*T] (per-head softmax-weighted sums of the token rows).
// T in {256, 512, 768}, heads <= 16.
size_t packed_image_tokens_elems(int n, int n_kv, int width);
// x f32 [n, n_kv, T] -> packed bf16 (32-row chunks, rows of T + 8 elements, zero padding); normalize: (x - mean) * rstd
int pack_image_tokens(const float* x, bf16* out, int n, int n_kv, int width, bool normalize, cudaStream_t stream);
bool decode_cross_absorbed_supported(int width, int heads, int head_dim, int n_kv);
// reverse: walk the crops from the last to the first (alternate layers start on what the previous launch left in L2)
int decode_cross_absorbed(const bf16* qp, const bf16* xn, bf16* ctx, int batch, int heads, int width, int n_kv,
                          const unsigned char* finished, const DecodeState* st, bool reverse, cudaStream_t stream);
// wk_abs [heads*T, head_dim] = per-head transpose of the K projection rows wk [heads*head_dim, T], times `scale`
int absorb_k_weights(const bf16* wk, bf16* out, int heads, int head_dim, int width, float scale, cudaStream_t stream);
int decode_cross_absorbed_check_watchdog(cudaStream_t stream);

// reduce the arg-max partials of the vocab GEMM, apply the generate() token rules (coca_model.py:304-322)
int decode_select(const float* part_val, const int* part_idx, int n_parts, int* tokens, int ld_tokens,
                  unsigned char* finished, int batch, int seq_len, int eos, int pad, DecodeState* st,
                  cudaStream_t stream);
int decode_advance(DecodeState* st, bool teacher_forced, int batch, cudaStream_t stream);
int decode_init(int* tokens, int ld_tokens, unsigned char* finished, int batch, int seq_len, int sot, int pad,
                DecodeState* st, cudaStream_t stream);

// ---- preprocess.cu ------------------------------------------------------------------------------------
// Crop + Pillow-exact antialiased bicubic resize (shortest edge -> size) + centre crop + /255 + normalise.
// frames u8 [F, H, W, C] (C = 3 or 4, only the first 3 used); rects int32 [N,4] = x1,y1,x2,y2 source
// rectangle (pixels outside the frame read as 0, i.e. PIL crop semantics; numpy-slice policies clamp on
// the host); flags bit0 = swap R/B.  Outputs [N,3,size,size] bf16 and/or fp32 (either may be null).
int preprocess_crops(const unsigned char* frames, int n_frames, int frame_h, int frame_w, int frame_c,
                     const int* rects, const int* frame_idx, const int* flags, int n, int size, bf16* out_bf16,
                     float* out_f32, int* status, cudaStream_t stream);

// ---- disagreement.cu -------------------------------------------------------------------------------------
size_t disagreement_workspace_bytes(int n);
// mode 0: mean over all VxV of (1 - cos) (projection_utils.py:928-942); mode 1: mean over i<j of cos
// (compute_cosine_sim.py:11-22, NaN when V < 2).  Outputs sorted by key; *out_m = number of instances.
int instance_disagreement(const float* emb, long long ld_emb, const long long* keys, int n, int d, int mode,
                          long long* out_keys, float* out_score, int* out_count, int* out_m, void* ws,
                          size_t ws_bytes, cudaStream_t stream);
// same reduction with rows fetched from per-rank peer buffers: row i lives in peers[i / rows_per_rank]
int instance_disagreement_peers(const float* const* peers_dev, int world, int rows_per_rank, long long ld_emb,
                                const long long* keys, int n, int d, int mode, int part_rank, int part_world,
                                long long* out_keys, float* out_score, int* out_count, int* out_m, void* ws,
                                size_t ws_bytes, cudaStream_t stream);

// Caption frequency table (pseudocaptioner.py:156-177) over token-id rows: one output entry per distinct
// (instance key, caption): the key, the row index of its FIRST occurrence and how many rows share it; entries are
// ordered by (key, first occurrence), i.e. dict insertion order inside every instance.  Synchronises the stream.
size_t caption_frequency_workspace_bytes(int n);
int caption_frequency(const int* tokens, int ld_tokens, int t, const long long* keys, int n, long long* out_key,
                      int* out_first, int* out_count, int* out_m, void* ws, size_t ws_bytes, cudaStream_t stream);

// ---- rasterize.cu -----------------------------------------------------------------------------------
// Disagreement map + reward (projection_utils.py:694-703,760-785; env_habitat3.py:75-84).  points f32 [P,3]
// (x, y, z: x -> column, z -> row), point_keys int64 [P] = object id of each point, (inst_keys, inst_scores) =
// key-sorted output of instance_disagreement.  map f32 [H,W] is fully overwritten; reward (nullable) = sum/1000.
size_t disagreement_map_workspace_bytes(int n_points, int H, int W);
int disagreement_map(const float* points, int n_points, const long long* point_keys, const long long* inst_keys,
                     const float* inst_scores, int m, float scale, int off_x, int off_y, int H, int W, float* map,
                     float* reward, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace ec
