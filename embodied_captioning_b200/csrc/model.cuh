// Model object behind the opaque `ec_model` handle of include/ec_b200.h.
#pragma once
#include <vector>
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "kernels.cuh"
#include "../../include/ec_b200.h"

namespace ec {

struct Block {   // one open_clip ResidualAttentionBlock
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr, *lnkv_g = nullptr, *lnkv_b = nullptr;
  bf16 *w_qkv = nullptr, *w_out = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;     // nn.Linear layout [out, in]
  float *b_qkv = nullptr, *b_out = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
};

struct DecodeWs;
struct Prefetch;

struct GraphKey { const void* ws; const void* tokens; const void* logits; int n, seq_len, min_len, ld, tf, n_total; };

constexpr int kMaxChains = 8;
constexpr int kLnCountersEncoder = 4096;     // 128-row blocks of one image-tower call (1024 crops x 257 tokens = 2056)
constexpr int kLnCountersPerChain = 64;      // 128-row blocks of one decoder chain
constexpr int kSkCountersPerChain = 256;     // output tiles of one split-K GEMM of a decoder chain
constexpr int kSeqLaunchesPerStep = 64;      // GEMM-sequence launches of one decoder step (one per block), per chain

struct Model {
  ec_config cfg;
  int device;
  bool finalized = false;
  bool use_graph = false;
  bool fuse_ln = true;             // (ctor: off unless EC_FUSE_LN=1) LayerNorms computed by the residual GEMM's fused tail
  bool fuse_ln_encoder = false;    // image tower: LayerNorm tails inside the CTA-pair residual GEMMs (EC_FUSE_LN_ENC=1: on; measured slower)
  bool prefetch_weights = true;    // decoder GEMMs prefetch the next GEMM's weights into L2 (EC_PREFETCH_WEIGHTS=0: off)
  // cross-K/V prefill: the 12 ln_1_kv LayerNorms of the image tokens differ only in gamma / beta, so the tokens are
  // normalised ONCE and the affine parts live in pre-multiplied K/V projection weights (EC_SHARED_LNKV=0: off)
  bool shared_lnkv = true;
  std::vector<bf16*> wkv_fold;
  std::vector<float*> bkv_fold;
  // absorbed-projection cross-attention (cross_absorbed.cu): the decoder streams the normalised image tokens instead of
  // per-layer K / V caches; wk_abs[l] = per-head transposed, pre-scaled K projection [heads*T, 64] (option
  // "absorbed_cross" / EC_ABSORB=0: off -> K/V caches as in round 1)
  bool absorbed_cross = true;
  std::vector<bf16*> wk_abs;
  bool use_absorbed() const;
  float *unit_gamma = nullptr, *zero_beta = nullptr;
  int* ln_counters = nullptr;      // row-block arrival counters of the fused tails (zero between launches)
  int* sk_counters = nullptr;      // split-K tile counters of the decoder's FC2 GEMMs, kSkCountersPerChain per chain
  // decoder steps: the GEMMs + LayerNorms between two attention kernels as ONE persistent launch (gemm_seq.cuh; option
  // "decode_seq" / EC_SEQ=1; default OFF: bit-identical but measured slower, profiles/r02_gemm_seq.txt)
  bool use_seq = false;
  int* seq_counters = nullptr;     // kMaxChains x kSeqLaunchesPerStep x kGemmSeqCounterInts, zeroed at the top of every step
  bool use_tc_attention = false;   // tcgen05/TMEM attention for the ViT blocks (option "tc_attention")
  int patch_k = 0, patch_ld = 0;

  // vision tower
  bf16* w_conv1 = nullptr;
  float *cls_emb = nullptr, *vis_pos = nullptr, *ln_pre_g = nullptr, *ln_pre_b = nullptr;
  std::vector<Block> vblocks;
  // attentional pooler
  float* pool_query = nullptr;
  bf16 *w_pool_q = nullptr, *w_pool_kv = nullptr, *w_pool_out = nullptr, *w_vproj = nullptr, *pool_q_proj = nullptr;
  float *b_pool_qkv = nullptr, *b_pool_out = nullptr;
  float *ln_q_g = nullptr, *ln_q_b = nullptr, *ln_k_g = nullptr, *ln_k_b = nullptr, *ln_post_g = nullptr,
        *ln_post_b = nullptr;
  // text tower + multimodal decoder
  float *tok_emb = nullptr, *txt_pos = nullptr;
  std::vector<Block> tblocks, dself, dcross;
  float *ln_final_g = nullptr, *ln_final_b = nullptr;
  bf16* w_vocab = nullptr;

  std::vector<void*> owned;
  // CUDA-graph replay of the decoder step
  // decoder row chains: each owns a stream and a linear step graph; 0 = choose from the batch size
  int decode_chains = 0;
  cudaStream_t chain_streams[kMaxChains] = {};
  cudaEvent_t chain_done[kMaxChains] = {};
  cudaEvent_t chain_fork = nullptr;
  cudaGraphExec_t graph_exec[kMaxChains] = {};
  GraphKey graph_key[kMaxChains] = {};
  long long graph_nodes[kMaxChains] = {};

  Model(const ec_config& c, int dev);
  ~Model();
  bool decoder_fuse(int rows) const { return fuse_ln && ceil_div(rows, 128) <= kLnCountersPerChain; }
  int cfg_tokens() const { return (cfg.image_size / cfg.patch_size) * (cfg.image_size / cfg.patch_size) + 1; }

  int set_tensor(const char* name, const float* host, int64_t numel);
  int finalize(cudaStream_t stream);

  size_t encode_workspace_bytes(int n) const;
  int encode_image(const bf16* x, int n, float* latent, float* embs, void* ws, size_t ws_bytes, cudaStream_t s);

  size_t decode_workspace_bytes(int n, int seq_len) const;
  int decode_greedy(const float* embs, int n, int seq_len, int min_len, int* tokens, float* logits, void* ws,
                    size_t ws_bytes, cudaStream_t s);
  int forward_logits(const float* embs, const int* text, int n, int t, float* logits, void* ws, size_t ws_bytes,
                     cudaStream_t s);

 private:
  int upload_f32(const float* host, size_t n, float** dst);
  int upload_bf16(const float* host, int rows, int cols, bool src_transposed, bf16** dst, int total_rows, int ld,
                  int row_off);
  int set_block_tensor(Block& b, int width, const char* leaf, const char* name, const float* h, int64_t numel);
  int self_block(const Block& b, const DecodeWs& w, bf16* kc, bf16* vc, int n, int max_len, int heads, bool ln1_done,
                 const float* next_g, const float* next_b, Prefetch next_pf, cudaStream_t s);
  int cross_block(const Block& b, const DecodeWs& w, const bf16* kv, int layer, int n, bool ln1_done, const float* next_g,
                  const float* next_b, Prefetch next_pf, cudaStream_t s);
  int chain_rows(int n) const;
  int chain_step(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                 float* logits, bool teacher_forced, cudaStream_t s);
  int chain_step_seq(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                     float* logits, bool teacher_forced, cudaStream_t s);
  int run_stages(const GemmSeqStage* st, int ns, int* counters, const int* skip, Prefetch pf, int n, cudaStream_t s,
                 int tag = -1);
  int ln_linear(const DecodeWs& w, const float* g, const float* b, const bf16* W, const float* bias, int n, int N, void* out,
                int ldo, GemmEpilogue epi, const int* skip, Prefetch pf, cudaStream_t s);
  int vocab_head(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len, float* logits,
                 bool teacher_forced, cudaStream_t s);
  int decode_step(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                  float* logits, bool teacher_forced, cudaStream_t s);
  int run_chain(int c, const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                float* logits, bool teacher_forced, int steps, cudaStream_t cs);
  int prefill_cross_kv(const DecodeWs& w, const float* embs, int n, cudaStream_t s);
  int run_steps(const DecodeWs& w, int* tokens, int ld_tokens, int n, int seq_len, int min_len, float* logits,
                bool teacher_forced, int steps, cudaStream_t s);
};

}  // namespace ec

struct ec_model {
  ec::Model impl;
  ec_model(const ec_config& c, int dev) : impl(c, dev) {}
};
