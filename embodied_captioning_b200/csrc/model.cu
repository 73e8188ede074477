// CoCa captioner on B200: weight store + image tower + KV-cached greedy decoder orchestration.
// Mirrors the wiring of the reference `CoCa` module (coca_model.py:92-203) and its `generate()` top-k path
// (coca_model.py:205-333); the arithmetic of the towers is upstream open_clip's.
#include "model.cuh"
#include <cstring>
#include <cstdlib>
#include <string>

namespace ec {

// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
static inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;                                                          // round to nearest even
  return (uint16_t)(u >> 16);
}

struct Bump {     // workspace carving, 256-byte aligned
  unsigned char* base; size_t off = 0;
  explicit Bump(void* b) : base(reinterpret_cast<unsigned char*>(b)) {}
  template <typename T> T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(count * sizeof(T), 256);
    return p;
  }
};

int Model::upload_f32(const float* host, size_t n, float** dst) {
  if (*dst == nullptr) EC_CHECK_CUDA(cudaMalloc(dst, std::max<size_t>(n, 4) * sizeof(float)));
  EC_CHECK_CUDA(cudaMemcpy(*dst, host, n * sizeof(float), cudaMemcpyHostToDevice));
  return EC_OK;
}

// host [rows, cols] fp32 (optionally transposed source [cols, rows]) -> device bf16 [rows, ld] at row offset
int Model::upload_bf16(const float* host, int rows, int cols, bool src_transposed, bf16** dst, int total_rows, int ld,
                       int row_off) {
  if (*dst == nullptr) {
    EC_CHECK_CUDA(cudaMalloc(dst, (size_t)total_rows * ld * sizeof(bf16)));
    EC_CHECK_CUDA(cudaMemset(*dst, 0, (size_t)total_rows * ld * sizeof(bf16)));
  }
  std::vector<uint16_t> tmp((size_t)rows * ld, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c)
      tmp[(size_t)r * ld + c] = f32_to_bf16_bits(src_transposed ? host[(size_t)c * rows + r] : host[(size_t)r * cols + c]);
  EC_CHECK_CUDA(cudaMemcpy(*dst + (size_t)row_off * ld, tmp.data(), tmp.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  return EC_OK;
}

Model::Model(const ec_config& c, int dev) : cfg(c), device(dev) {
  vblocks.resize(c.vision_layers);
  tblocks.resize(c.text_layers);
  dself.resize(c.mm_layers);
  dcross.resize(c.mm_layers);
  patch_k = 3 * c.patch_size * c.patch_size;
  patch_ld = (int)align_up((size_t)patch_k, 8);
  const size_t n_cnt = kLnCountersEncoder + (size_t)kMaxChains * kLnCountersPerChain;
  if (cudaSetDevice(dev) == cudaSuccess && cudaMalloc(&ln_counters, n_cnt * sizeof(int)) == cudaSuccess) {
    cudaMemset(ln_counters, 0, n_cnt * sizeof(int));
    owned.push_back(ln_counters);
  } else {
    ln_counters = nullptr;
    fuse_ln = false;
  }
  const size_t n_sk = (size_t)kMaxChains * kSkCountersPerChain;
  if (cudaMalloc(&sk_counters, n_sk * sizeof(int)) == cudaSuccess) {
    cudaMemset(sk_counters, 0, n_sk * sizeof(int));
    owned.push_back(sk_counters);
  } else {
    sk_counters = nullptr;
  }
  // Measured on B200 (profiles/r01_fused_ln.txt): the fused tail is bit-exact but SLOWER -- one CTA normalises each
  // 128-row block (latency-bound row loop on the GEMM's critical path) where the stand-alone kernel spreads the rows
  // over every SM at the HBM roofline, and every tile pays two named barriers + a device-scope fence + an atomic.
  // Encoder 180 -> 220 ms, decoder 113 -> 150 ms at 1024 crops.  Opt-in only (EC_FUSE_LN=1).
  const size_t n_seq = (size_t)kMaxChains * kSeqLaunchesPerStep * kGemmSeqCounterInts;
  if (cudaMalloc(&seq_counters, n_seq * sizeof(int)) == cudaSuccess) {
    cudaMemset(seq_counters, 0, n_seq * sizeof(int));
    owned.push_back(seq_counters);
  } else {
    seq_counters = nullptr;
  }
  // Measured on B200 at 1024 crops (profiles/r02_gemm_seq.txt): bit-identical, ~80 graph nodes per step instead of ~280,
  // and SLOWER -- decoder alone 94.6 vs 86.6 ms, inside the bench loop 115.7 vs 109.5.  The hand-over itself is cheap
  // (0.4-1.1 us from "rows published" to the next stage's first operands), but publishing a tile costs the epilogue
  // + a GPU-scope release (2.3-3 us per tile), the LayerNorm tail 4-5 us against 3.65 us for the stand-alone launch,
  // and a lone CTA's GELU epilogue is 5 us either way.  Opt-in (EC_SEQ=1 / option "decode_seq").
  static const int env_seq = [] { const char* e = getenv("EC_SEQ"); return e ? atoi(e) : 0; }();
  use_seq = env_seq != 0 && seq_counters != nullptr;
  static const int env_pf = [] { const char* e = getenv("EC_PREFETCH_WEIGHTS"); return e ? atoi(e) : 1; }();
  prefetch_weights = env_pf != 0;
  static const int env_lnkv = [] { const char* e = getenv("EC_SHARED_LNKV"); return e ? atoi(e) : 1; }();
  shared_lnkv = env_lnkv != 0;
  static const int env_abs = [] { const char* e = getenv("EC_ABSORB"); return e ? atoi(e) : 1; }();
  absorbed_cross = env_abs != 0;
  static const int env_fuse = [] { const char* e = getenv("EC_FUSE_LN"); return e ? atoi(e) : 0; }();
  if (env_fuse == 0) fuse_ln = false;
  // image tower: the CTA-local tail (no atomics, contiguous tile schedule) also loses -- the epilogue warps that run it
  // delay the next tile's accumulator drain (bench 318.0 vs 302.9 ms/step) -> opt-in as well (EC_FUSE_LN_ENC=1)
  static const int env_fuse_enc = [] { const char* e = getenv("EC_FUSE_LN_ENC"); return e ? atoi(e) : 0; }();
  fuse_ln_encoder = env_fuse_enc != 0 && ln_counters != nullptr;
}

Model::~Model() {
  cudaSetDevice(device);
  for (void* p : owned) cudaFree(p);
  for (int c = 0; c < kMaxChains; ++c) {
    if (graph_exec[c]) cudaGraphExecDestroy(graph_exec[c]);
    if (chain_streams[c]) cudaStreamDestroy(chain_streams[c]);
    if (chain_done[c]) cudaEventDestroy(chain_done[c]);
  }
  if (chain_fork) cudaEventDestroy(chain_fork);
}

// ---------------------------------------------------------------------------------------------------
// weight loading by open_clip state-dict name
// ---------------------------------------------------------------------------------------------------
#define EC_EXPECT_NUMEL(expected)                                                                              \
  EC_REQUIRE(numel == (int64_t)(expected), EC_ERR_INVALID, "tensor %s: expected %lld elements, got %lld", name, \
             (long long)(expected), (long long)numel)

int Model::set_block_tensor(Block& b, int width, const char* leaf, const char* name, const float* h, int64_t numel) {
  auto vec = [&](float** dst, int n) -> int {
    EC_EXPECT_NUMEL(n);
    bool fresh = (*dst == nullptr);
    EC_TRY(upload_f32(h, n, dst));
    if (fresh) owned.push_back(*dst);
    return EC_OK;
  };
  auto mat = [&](bf16** dst, int rows, int cols) -> int {
    EC_EXPECT_NUMEL((int64_t)rows * cols);
    bool fresh = (*dst == nullptr);
    EC_TRY(upload_bf16(h, rows, cols, false, dst, rows, cols, 0));
    if (fresh) owned.push_back(*dst);
    return EC_OK;
  };
  const int W = width;
  if (!strcmp(leaf, "ln_1.weight")) return vec(&b.ln1_g, W);
  if (!strcmp(leaf, "ln_1.bias")) return vec(&b.ln1_b, W);
  if (!strcmp(leaf, "ln_1_kv.weight")) return vec(&b.lnkv_g, W);
  if (!strcmp(leaf, "ln_1_kv.bias")) return vec(&b.lnkv_b, W);
  if (!strcmp(leaf, "ln_2.weight")) return vec(&b.ln2_g, W);
  if (!strcmp(leaf, "ln_2.bias")) return vec(&b.ln2_b, W);
  if (!strcmp(leaf, "attn.in_proj_weight")) return mat(&b.w_qkv, 3 * W, W);
  if (!strcmp(leaf, "attn.in_proj_bias")) return vec(&b.b_qkv, 3 * W);
  if (!strcmp(leaf, "attn.out_proj.weight")) return mat(&b.w_out, W, W);
  if (!strcmp(leaf, "attn.out_proj.bias")) return vec(&b.b_out, W);
  if (!strcmp(leaf, "mlp.c_fc.weight")) return mat(&b.w_fc1, 4 * W, W);
  if (!strcmp(leaf, "mlp.c_fc.bias")) return vec(&b.b_fc1, 4 * W);
  if (!strcmp(leaf, "mlp.c_proj.weight")) return mat(&b.w_fc2, W, 4 * W);
  if (!strcmp(leaf, "mlp.c_proj.bias")) return vec(&b.b_fc2, W);
  set_error("unknown block tensor %s", name);
  return EC_ERR_INVALID;
}

int Model::set_tensor(const char* name, const float* h, int64_t numel) {
  EC_CHECK_CUDA(cudaSetDevice(device));
  const int W = cfg.vision_width, E = cfg.embed_dim, T = cfg.text_width, V = cfg.vocab_size;
  finalized = false;
  auto vec = [&](float** dst, int64_t n) -> int {
    EC_EXPECT_NUMEL(n);
    bool fresh = (*dst == nullptr);
    EC_TRY(upload_f32(h, (size_t)n, dst));
    if (fresh) owned.push_back(*dst);
    return EC_OK;
  };
  auto mat = [&](bf16** dst, int rows, int cols, bool transposed, int total_rows, int ld, int row_off) -> int {
    EC_EXPECT_NUMEL((int64_t)rows * cols);
    bool fresh = (*dst == nullptr);
    EC_TRY(upload_bf16(h, rows, cols, transposed, dst, total_rows, ld, row_off));
    if (fresh) owned.push_back(*dst);
    return EC_OK;
  };
  int layer = -1, consumed = 0;
  if (!strcmp(name, "logit_scale")) return EC_OK;     // unused by the captioning path (coca_model.py:137,197)
  if (sscanf(name, "visual.transformer.resblocks.%d.%n", &layer, &consumed) == 1 && consumed > 0) {
    EC_REQUIRE(layer >= 0 && layer < cfg.vision_layers, EC_ERR_INVALID, "layer index out of range in %s", name);
    return set_block_tensor(vblocks[layer], W, name + consumed, name, h, numel);
  }
  if (sscanf(name, "text.transformer.resblocks.%d.%n", &layer, &consumed) == 1 && consumed > 0) {
    EC_REQUIRE(layer >= 0 && layer < cfg.text_layers, EC_ERR_INVALID, "layer index out of range in %s", name);
    return set_block_tensor(tblocks[layer], T, name + consumed, name, h, numel);
  }
  if (sscanf(name, "text_decoder.resblocks.%d.%n", &layer, &consumed) == 1 && consumed > 0) {
    EC_REQUIRE(layer >= 0 && layer < cfg.mm_layers, EC_ERR_INVALID, "layer index out of range in %s", name);
    return set_block_tensor(dself[layer], T, name + consumed, name, h, numel);
  }
  if (sscanf(name, "text_decoder.cross_attn.%d.%n", &layer, &consumed) == 1 && consumed > 0) {
    EC_REQUIRE(layer >= 0 && layer < cfg.mm_layers, EC_ERR_INVALID, "layer index out of range in %s", name);
    return set_block_tensor(dcross[layer], T, name + consumed, name, h, numel);
  }
  if (!strcmp(name, "visual.conv1.weight")) return mat(&w_conv1, W, patch_k, false, W, patch_ld, 0);
  if (!strcmp(name, "visual.class_embedding")) return vec(&cls_emb, W);
  if (!strcmp(name, "visual.positional_embedding")) return vec(&vis_pos, (int64_t)(cfg_tokens()) * W);
  if (!strcmp(name, "visual.ln_pre.weight")) return vec(&ln_pre_g, W);
  if (!strcmp(name, "visual.ln_pre.bias")) return vec(&ln_pre_b, W);
  if (!strcmp(name, "visual.attn_pool.query")) return vec(&pool_query, (int64_t)cfg.n_queries * E);
  if (!strcmp(name, "visual.attn_pool.attn.q_proj_weight")) return mat(&w_pool_q, E, E, false, E, E, 0);
  if (!strcmp(name, "visual.attn_pool.attn.k_proj_weight")) return mat(&w_pool_kv, E, W, false, 2 * E, W, 0);
  if (!strcmp(name, "visual.attn_pool.attn.v_proj_weight")) return mat(&w_pool_kv, E, W, false, 2 * E, W, E);
  if (!strcmp(name, "visual.attn_pool.attn.in_proj_bias")) return vec(&b_pool_qkv, 3 * E);
  if (!strcmp(name, "visual.attn_pool.attn.out_proj.weight")) return mat(&w_pool_out, E, E, false, E, E, 0);
  if (!strcmp(name, "visual.attn_pool.attn.out_proj.bias")) return vec(&b_pool_out, E);
  if (!strcmp(name, "visual.attn_pool.ln_q.weight")) return vec(&ln_q_g, E);
  if (!strcmp(name, "visual.attn_pool.ln_q.bias")) return vec(&ln_q_b, E);
  if (!strcmp(name, "visual.attn_pool.ln_k.weight")) return vec(&ln_k_g, W);
  if (!strcmp(name, "visual.attn_pool.ln_k.bias")) return vec(&ln_k_b, W);
  if (!strcmp(name, "visual.ln_post.weight")) return vec(&ln_post_g, E);
  if (!strcmp(name, "visual.ln_post.bias")) return vec(&ln_post_b, E);
  if (!strcmp(name, "visual.proj")) return mat(&w_vproj, E, E, true, E, E, 0);          // pooled @ proj -> W = proj^T
  if (!strcmp(name, "text.token_embedding.weight")) return vec(&tok_emb, (int64_t)V * T);
  if (!strcmp(name, "text.positional_embedding")) return vec(&txt_pos, (int64_t)(cfg.context_length + 1) * T);
  if (!strcmp(name, "text.cls_emb") || !strcmp(name, "text.ln_final.weight") || !strcmp(name, "text.ln_final.bias") ||
      !strcmp(name, "text.text_projection"))
    return EC_OK;   // only feed the pooled text latent, which generate() never consumes
  if (!strcmp(name, "text_decoder.ln_final.weight")) return vec(&ln_final_g, T);
  if (!strcmp(name, "text_decoder.ln_final.bias")) return vec(&ln_final_b, T);
  if (!strcmp(name, "text_decoder.text_projection")) return mat(&w_vocab, V, T, true, V, T, 0);   // [T,V] -> [V,T]
  set_error("unknown tensor name %s", name);
  return EC_ERR_INVALID;
}

bool Model::use_absorbed() const {
  return absorbed_cross && shared_lnkv && !dcross.empty() && wk_abs.size() == dcross.size() && wk_abs[0] != nullptr &&
         decode_cross_absorbed_supported(cfg.text_width, cfg.mm_heads, cfg.text_width / std::max(cfg.mm_heads, 1),
                                         cfg.n_queries - 1);
}

static bool block_complete(const Block& b, bool cross) {
  return b.ln1_g && b.ln1_b && b.ln2_g && b.ln2_b && b.w_qkv && b.b_qkv && b.w_out && b.b_out && b.w_fc1 && b.b_fc1 &&
         b.w_fc2 && b.b_fc2 && (!cross || (b.lnkv_g && b.lnkv_b));
}

int Model::finalize(cudaStream_t stream) {
  EC_CHECK_CUDA(cudaSetDevice(device));
  for (size_t i = 0; i < vblocks.size(); ++i)
    EC_REQUIRE(block_complete(vblocks[i], false), EC_ERR_STATE, "visual block %zu incomplete", i);
  for (size_t i = 0; i < tblocks.size(); ++i)
    EC_REQUIRE(block_complete(tblocks[i], false), EC_ERR_STATE, "text block %zu incomplete", i);
  for (size_t i = 0; i < dself.size(); ++i) {
    EC_REQUIRE(block_complete(dself[i], false), EC_ERR_STATE, "decoder self block %zu incomplete", i);
    EC_REQUIRE(block_complete(dcross[i], true), EC_ERR_STATE, "decoder cross block %zu incomplete", i);
  }
  EC_REQUIRE(w_conv1 && cls_emb && vis_pos && ln_pre_g && ln_pre_b && pool_query && w_pool_q && w_pool_kv &&
                 b_pool_qkv && w_pool_out && b_pool_out && ln_q_g && ln_q_b && ln_k_g && ln_k_b && ln_post_g &&
                 ln_post_b && w_vproj && tok_emb && txt_pos && ln_final_g && ln_final_b && w_vocab,
             EC_ERR_STATE, "model tensors missing (load every open_clip CoCa tensor before finalize)");
  // the pooler queries do not depend on the image: q = ln_q(query) @ Wq^T + bq, computed once
  const int E = cfg.embed_dim, Q = cfg.n_queries;
  bf16* lnq = nullptr;
  EC_CHECK_CUDA(cudaMalloc(&lnq, (size_t)Q * E * sizeof(bf16)));
  if (!pool_q_proj) {
    EC_CHECK_CUDA(cudaMalloc(&pool_q_proj, (size_t)Q * E * sizeof(bf16)));
    owned.push_back(pool_q_proj);
  }
  int rc = layernorm(pool_query, E, ln_q_g, ln_q_b, lnq, E, true, Q, E, nullptr, stream);
  if (rc == EC_OK) {
    GemmArgs g{};
    g.A = lnq; g.lda = E; g.W = w_pool_q; g.ldw = E; g.M = Q; g.N = E; g.K = E; g.bias = b_pool_qkv;
    g.out = pool_q_proj; g.ldo = E;
    rc = gemm_bf16(g, EPI_BF16, stream);
  }
  // pre-multiplied cross K/V projections (rows [T, 3T) of in_proj) for the shared normalisation of the image tokens
  if (rc == EC_OK && shared_lnkv && !dcross.empty()) {
    const int T = cfg.text_width;
    if (unit_gamma == nullptr) {
      if (cudaMalloc(&unit_gamma, 2 * (size_t)T * sizeof(float)) != cudaSuccess) rc = EC_ERR_CUDA;
      else { owned.push_back(unit_gamma); zero_beta = unit_gamma + T; }
    }
    if (rc == EC_OK) rc = fill_f32(unit_gamma, T, 1.0f, stream);
    if (rc == EC_OK) rc = fill_f32(zero_beta, T, 0.0f, stream);
    wkv_fold.resize(dcross.size(), nullptr);
    bkv_fold.resize(dcross.size(), nullptr);
    for (size_t l = 0; l < dcross.size() && rc == EC_OK; ++l) {
      if (wkv_fold[l] == nullptr) {
        if (cudaMalloc(&wkv_fold[l], 2 * (size_t)T * T * sizeof(bf16)) != cudaSuccess ||
            cudaMalloc(&bkv_fold[l], 2 * (size_t)T * sizeof(float)) != cudaSuccess) { rc = EC_ERR_CUDA; break; }
        owned.push_back(wkv_fold[l]);
        owned.push_back(bkv_fold[l]);
      }
      const Block& b = dcross[l];
      rc = fold_layernorm_into_linear(b.w_qkv + (size_t)T * T, b.b_qkv + T, b.lnkv_g, b.lnkv_b, wkv_fold[l], bkv_fold[l],
                                      2 * T, T, stream);
    }
    // absorbed K projections: wk_abs[l][(h*T + i), d] = wkv_fold[l][(h*64 + d), i] / sqrt(64)
    const int hd = T / std::max(cfg.mm_heads, 1);
    if (rc == EC_OK && absorbed_cross && decode_cross_absorbed_supported(T, cfg.mm_heads, hd, cfg.n_queries - 1)) {
      wk_abs.resize(dcross.size(), nullptr);
      for (size_t l = 0; l < dcross.size() && rc == EC_OK; ++l) {
        if (wk_abs[l] == nullptr) {
          if (cudaMalloc(&wk_abs[l], (size_t)T * T * sizeof(bf16)) != cudaSuccess) { rc = EC_ERR_CUDA; break; }
          owned.push_back(wk_abs[l]);
        }
        rc = absorb_k_weights(wkv_fold[l], wk_abs[l], cfg.mm_heads, hd, T, 1.0f / sqrtf((float)hd), stream);
      }
    }
  }
  cudaError_t e = cudaStreamSynchronize(stream);
  cudaFree(lnq);
  EC_TRY(rc);
  EC_CHECK_CUDA(e);
  EC_TRY(gemm_check_watchdog(stream));
  finalized = true;
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// transformer block helpers
// ---------------------------------------------------------------------------------------------------
struct FusedLn {      // LayerNorm applied to the rows of a residual GEMM's output by the GEMM itself (gemm_kernel.cuh::ln_tail)
  const float* g = nullptr; const float* b = nullptr; bf16* out = nullptr; int ld = 0; int* counters = nullptr;
};

struct Prefetch { const void* ptr = nullptr; size_t bytes = 0; };     // the next GEMM's weights -> L2
struct SplitK { float* ws = nullptr; int* cnt = nullptr; };            // deterministic split-K scratch (gemm_sm100.cuh)

static int linear(const bf16* A, int lda, const bf16* W, int ldw, const float* bias, int M, int N, int K, void* out,
                  int ldo, GemmEpilogue epi, const float* resid, int ldr, const int* skip, cudaStream_t s,
                  const FusedLn* ln = nullptr, Prefetch pf = Prefetch(), SplitK sk = SplitK()) {
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.M = M; g.N = N; g.K = K; g.bias = bias;
  g.residual = resid; g.ldr = ldr; g.out = out; g.ldo = ldo; g.skip_flag = skip;
  g.prefetch_ptr = pf.ptr; g.prefetch_bytes = pf.bytes;
  if (ln != nullptr && ln->out != nullptr) {
    g.ln_gamma = ln->g; g.ln_beta = ln->b; g.ln_out = ln->out; g.ld_ln = ln->ld; g.ln_counters = ln->counters;
  } else if (epi == EPI_F32_RESID && sk.ws != nullptr && sk.cnt != nullptr &&
             ceil_div(M, 128) * ceil_div(N, 128) <= kSkCountersPerChain) {
    g.split_k = gemm_choose_split_k(M, N, K);
    g.sk_ws = sk.ws; g.sk_cnt = sk.cnt;
  }
  return gemm_bf16(g, epi, s);
}

// ---------------------------------------------------------------------------------------------------
// image tower (coca_model.py:152-155 -> upstream VisionTransformer + AttentionalPooler)
// ---------------------------------------------------------------------------------------------------
struct EncodeWs {
  bf16* patches; float* patch_emb; float* resid; bf16* ln; bf16* qkv; bf16* att; bf16* fc1;
  bf16* pool_kv; bf16* pool_att; float* pool_out; bf16* pooled_bf16; float* pooled_proj;
  size_t bytes;
};

static EncodeWs carve_encode(const Model& m, void* ws, int n) {
  const ec_config& c = m.cfg;
  const size_t G2 = (size_t)(c.image_size / c.patch_size) * (c.image_size / c.patch_size);
  const size_t Mv = (size_t)n * (G2 + 1);
  const size_t W = c.vision_width, E = c.embed_dim, Q = c.n_queries;
  Bump b(ws);
  EncodeWs w;
  w.patches = b.take<bf16>((size_t)n * G2 * m.patch_ld);
  w.patch_emb = b.take<float>((size_t)n * G2 * W);
  w.resid = b.take<float>(Mv * W);
  w.ln = b.take<bf16>(Mv * W);
  w.qkv = b.take<bf16>(Mv * 3 * W);
  w.att = b.take<bf16>(Mv * W);
  w.fc1 = b.take<bf16>(Mv * 4 * W);
  w.pool_kv = b.take<bf16>(Mv * 2 * E);
  w.pool_att = b.take<bf16>((size_t)n * Q * E);
  w.pool_out = b.take<float>((size_t)n * Q * E);
  w.pooled_bf16 = b.take<bf16>((size_t)n * E);
  w.pooled_proj = b.take<float>((size_t)n * E);
  w.bytes = b.off;
  return w;
}

size_t Model::encode_workspace_bytes(int n) const { return carve_encode(*this, nullptr, std::max(n, 1)).bytes; }

// ln_post over the pooler output, split like `pooled, tokens = x[:,0], x[:,1:]`:
// row (b, 0) -> bf16 pooled[b] (operand of the projection GEMM); row (b, q>0) -> f32 embs[b, q-1]
__global__ void __launch_bounds__(256) ln_post_split_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, bf16* __restrict__ pooled,
                                                            float* __restrict__ embs, int batch, int nq, int width) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)batch * nq) return;
  const int b = (int)(row / nq), q = (int)(row % nq);
  const float* xr = x + (size_t)row * width;
  float s = 0.f;
  for (int c = lane; c < width; c += 32) s += xr[c];
  const float mean = warp_sum(s) / (float)width;
  float v = 0.f;
  for (int c = lane; c < width; c += 32) { const float d = xr[c] - mean; v += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)width + 1e-5f);
  for (int c = lane; c < width; c += 32) {
    const float y = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if (q == 0) pooled[(size_t)b * width + c] = __float2bfloat16_rn(y);
    else if (embs != nullptr) embs[((size_t)b * (nq - 1) + (q - 1)) * width + c] = y;
  }
}

int Model::encode_image(const bf16* x, int n, float* latent, float* embs, void* ws, size_t ws_bytes, cudaStream_t s) {
  EC_REQUIRE(finalized, EC_ERR_STATE, "model not finalized");
  EC_REQUIRE(n > 0, EC_ERR_INVALID, "encode_image: n = %d", n);
  EC_CHECK_CUDA(cudaSetDevice(device));
  const EncodeWs w = carve_encode(*this, ws, n);
  EC_REQUIRE(ws_bytes >= w.bytes, EC_ERR_WORKSPACE, "encode workspace %zu < %zu", ws_bytes, w.bytes);
  const int G = cfg.image_size / cfg.patch_size, G2 = G * G, TV = G2 + 1;
  const int W = cfg.vision_width, E = cfg.embed_dim, Q = cfg.n_queries;
  const int Mv = n * TV;
  const int heads = W / cfg.vision_head_dim;

  EC_TRY(im2col_patches(x, w.patches, n, cfg.image_size, cfg.patch_size, patch_ld, s));
  EC_TRY(linear(w.patches, patch_ld, w_conv1, patch_ld, nullptr, n * G2, W, patch_k, w.patch_emb, W, EPI_F32, nullptr, 0,
                nullptr, s));
  EC_TRY(vision_embed_ln_pre(w.patch_emb, cls_emb, vis_pos, ln_pre_g, ln_pre_b, w.resid, n, TV, W, s));
  // Every LayerNorm except the first is computed by the residual GEMM that completes its input rows (fused tail).
  // fused only on the CTA-pair kernel, whose tail is local to one CTA (the 1-CTA kernel's atomic form is slower)
  const bool fuse = fuse_ln_encoder && gemm_uses_pair_kernel(Mv, W) && ceil_div(Mv, 128) <= kLnCountersEncoder;
  for (int l = 0; l < cfg.vision_layers; ++l) {
    const Block& b = vblocks[l];
    if (l == 0 || !fuse) EC_TRY(layernorm(w.resid, W, b.ln1_g, b.ln1_b, w.ln, W, true, Mv, W, nullptr, s));
    EC_TRY(linear(w.ln, W, b.w_qkv, W, b.b_qkv, Mv, 3 * W, W, w.qkv, 3 * W, EPI_BF16, nullptr, 0, nullptr, s));
    AttnArgs a{};
    a.q = w.qkv; a.k = w.qkv + W; a.v = w.qkv + 2 * W;
    a.q_bs = a.k_bs = a.v_bs = (long long)TV * 3 * W; a.q_rs = a.k_rs = a.v_rs = 3 * W;
    a.o = w.att; a.o_bs = (long long)TV * W; a.o_rs = W;
    a.batch = n; a.heads = heads; a.sq = TV; a.skv = TV; a.head_dim = cfg.vision_head_dim;
    a.scale = 1.0f / sqrtf((float)cfg.vision_head_dim);
    if (use_tc_attention && attention_sm100_supported(TV, cfg.vision_head_dim))
      EC_TRY(attention_sm100_fused_qkv(w.qkv, 3 * W, 0, W, 2 * W, w.att, (long long)TV * W, W, n, heads, TV, s));
    else
      EC_TRY(attention_full(a, s));
    FusedLn ln2{b.ln2_g, b.ln2_b, fuse ? w.ln : nullptr, W, ln_counters};
    EC_TRY(linear(w.att, W, b.w_out, W, b.b_out, Mv, W, W, w.resid, W, EPI_F32_RESID, w.resid, W, nullptr, s, &ln2));
    if (!fuse) EC_TRY(layernorm(w.resid, W, b.ln2_g, b.ln2_b, w.ln, W, true, Mv, W, nullptr, s));
    EC_TRY(linear(w.ln, W, b.w_fc1, W, b.b_fc1, Mv, 4 * W, W, w.fc1, 4 * W, EPI_BF16_GELU, nullptr, 0, nullptr, s));
    const bool last = (l + 1 == cfg.vision_layers);     // the pooler's ln_k follows the last block
    FusedLn nxt{last ? ln_k_g : vblocks[l + 1].ln1_g, last ? ln_k_b : vblocks[l + 1].ln1_b, fuse ? w.ln : nullptr, W,
                ln_counters};
    EC_TRY(linear(w.fc1, 4 * W, b.w_fc2, 4 * W, b.b_fc2, Mv, W, 4 * W, w.resid, W, EPI_F32_RESID, w.resid, W, nullptr, s,
                  &nxt));
  }
  // attentional pooler: k = v = ln_k(x), q precomputed
  if (!fuse) EC_TRY(layernorm(w.resid, W, ln_k_g, ln_k_b, w.ln, W, true, Mv, W, nullptr, s));
  EC_TRY(linear(w.ln, W, w_pool_kv, W, b_pool_qkv + E, Mv, 2 * E, W, w.pool_kv, 2 * E, EPI_BF16, nullptr, 0, nullptr, s));
  {
    const int hd = E / cfg.pooler_heads;
    AttnArgs a{};
    a.q = pool_q_proj; a.q_bs = 0; a.q_rs = E;
    a.k = w.pool_kv; a.v = w.pool_kv + E; a.k_bs = a.v_bs = (long long)TV * 2 * E; a.k_rs = a.v_rs = 2 * E;
    a.o = w.pool_att; a.o_bs = (long long)Q * E; a.o_rs = E;
    a.batch = n; a.heads = cfg.pooler_heads; a.sq = Q; a.skv = TV; a.head_dim = hd;
    a.scale = 1.0f / sqrtf((float)hd);
    EC_TRY(attention_full(a, s));
  }
  EC_TRY(linear(w.pool_att, E, w_pool_out, E, b_pool_out, n * Q, E, E, w.pool_out, E, EPI_F32, nullptr, 0, nullptr, s));
  {
    const long long rows = (long long)n * Q;
    ln_post_split_kernel<<<(unsigned)ceil_div<long long>(rows, 8), 256, 0, s>>>(w.pool_out, ln_post_g, ln_post_b,
                                                                               w.pooled_bf16, embs, n, Q, E);
    EC_POST_LAUNCH();
  }
  if (latent != nullptr) {
    EC_TRY(linear(w.pooled_bf16, E, w_vproj, E, nullptr, n, E, E, w.pooled_proj, E, EPI_F32, nullptr, 0, nullptr, s));
    EC_TRY(l2_normalize_rows(w.pooled_proj, E, latent, E, n, E, s));
  }
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// decoder (coca_model.py:205-333 greedy path, KV-cached)
// ---------------------------------------------------------------------------------------------------
struct DecodeWs {
  bf16* ln_kv; bf16* cross_kv; bf16* tk; bf16* tv; bf16* dk; bf16* dv;
  float* x; bf16* ln; bf16* qkv; bf16* att; bf16* fc1; bf16* q;
  bf16* qp; bf16* ctx;   // absorbed cross-attention: per-head absorbed queries / contexts [n, heads*T]
  float* part_val; int* part_idx; unsigned char* finished; DecodeState* st; int* tokens_tf;
  int* ln_cnt;           // fused-LayerNorm row-block counters of this chain
  float* sk_ws; int* sk_cnt;   // split-K partial tiles ([3][n][T] floats) and tile counters of this chain's FC2 GEMMs
  int* seq_cnt;          // GEMM-sequence counters of this chain (kSeqLaunchesPerStep x kGemmSeqCounterInts)
  size_t cache_layer;    // elements between the self-K/V caches of consecutive layers (whole batch)
  size_t cross_layer;    // elements between the cross-K/V caches of consecutive layers (whole batch)
  size_t bytes;
};

// rows [r0, ...) of every per-crop buffer: the view one decoder chain works on (layer strides stay global)
static DecodeWs slice_rows(const Model& m, const DecodeWs& w, int r0, int seq_len, int chain) {
  const ec_config& c = m.cfg;
  const size_t T = c.text_width, NK = c.n_queries - 1, L = seq_len, r = (size_t)r0;
  const size_t parts = (size_t)gemm_argmax_parts(c.vocab_size);
  DecodeWs v = w;
  v.ln_kv += m.use_absorbed() ? packed_image_tokens_elems(r0, (int)NK, (int)T) : r * NK * T;
  v.cross_kv += r * NK * 2 * T;
  v.tk += r * L * T; v.tv += r * L * T; v.dk += r * L * T; v.dv += r * L * T;
  v.x += r * T; v.ln += r * T; v.qkv += r * 3 * T; v.att += r * T; v.fc1 += r * 4 * T; v.q += r * T;
  v.qp += r * (size_t)c.mm_heads * T; v.ctx += r * (size_t)c.mm_heads * T;
  v.part_val += r * parts; v.part_idx += r * parts; v.finished += r; v.tokens_tf += r * L;
  v.st += chain;
  v.ln_cnt = m.ln_counters + kLnCountersEncoder + chain * kLnCountersPerChain;
  v.sk_ws += r * 3 * T;
  v.sk_cnt = m.sk_counters ? m.sk_counters + chain * kSkCountersPerChain : nullptr;
  v.seq_cnt = m.seq_counters ? m.seq_counters + (size_t)chain * kSeqLaunchesPerStep * kGemmSeqCounterInts : nullptr;
  return v;
}

static DecodeWs carve_decode(const Model& m, void* ws, int n, int seq_len) {
  const ec_config& c = m.cfg;
  const size_t T = c.text_width, NK = c.n_queries - 1;
  const size_t L = seq_len;   // cache length (positions 0 .. seq_len-2 are ever written)
  Bump b(ws);
  DecodeWs w;
  const bool absorbed = m.use_absorbed();     // no per-layer K / V caches: the decoder streams ln_kv itself (packed layout)
  w.ln_kv = b.take<bf16>(absorbed ? packed_image_tokens_elems(n, (int)NK, (int)T) : (size_t)n * NK * T);
  w.cross_kv = b.take<bf16>(absorbed ? 0 : (size_t)c.mm_layers * n * NK * 2 * T);
  w.tk = b.take<bf16>((size_t)c.text_layers * n * L * T);
  w.tv = b.take<bf16>((size_t)c.text_layers * n * L * T);
  w.dk = b.take<bf16>((size_t)c.mm_layers * n * L * T);
  w.dv = b.take<bf16>((size_t)c.mm_layers * n * L * T);
  w.x = b.take<float>((size_t)n * T);
  w.ln = b.take<bf16>((size_t)n * T);
  w.qkv = b.take<bf16>((size_t)n * 3 * T);
  w.att = b.take<bf16>((size_t)n * T);
  w.fc1 = b.take<bf16>((size_t)n * 4 * T);
  w.q = b.take<bf16>((size_t)n * T);
  w.qp = b.take<bf16>(absorbed ? (size_t)n * c.mm_heads * T : 0);
  w.ctx = b.take<bf16>(absorbed ? (size_t)n * c.mm_heads * T : 0);
  const size_t parts = (size_t)gemm_argmax_parts(c.vocab_size);
  w.part_val = b.take<float>((size_t)n * parts);
  w.part_idx = b.take<int>((size_t)n * parts);
  w.finished = b.take<unsigned char>((size_t)n);
  w.st = b.take<DecodeState>(kMaxChains);
  w.tokens_tf = b.take<int>((size_t)n * L);
  w.sk_ws = b.take<float>((size_t)n * 3 * T);
  w.sk_cnt = m.sk_counters;
  w.ln_cnt = m.ln_counters + kLnCountersEncoder;
  w.seq_cnt = m.seq_counters;
  w.cache_layer = (size_t)n * L * T;
  w.cross_layer = (size_t)n * NK * 2 * T;
  w.bytes = b.off;
  return w;
}

size_t Model::decode_workspace_bytes(int n, int seq_len) const {
  return carve_decode(*this, nullptr, std::max(n, 1), std::max(seq_len, 2)).bytes;
}

// w.ln = LayerNorm(w.x; g, b), then out = epi(w.ln W^T + bias).  (Round 2 also tried the LayerNorm as a PRE-PHASE of the
// consuming GEMM -- its epilogue warps normalise the rows during the launch's cold start and hand them to the TMA producers
// through row-block counters: bit-identical, 92.4 vs 85.8 ms per decode, and 1.4 ms slower even when switched off;
// removed, profiles/r02_gemm_timing.txt.)
int Model::ln_linear(const DecodeWs& w, const float* g, const float* b, const bf16* W, const float* bias, int n, int N,
                     void* out, int ldo, GemmEpilogue epi, const int* skip, Prefetch pf, cudaStream_t s) {
  const int T = cfg.text_width;
  EC_TRY(layernorm(w.x, T, g, b, w.ln, T, true, n, T, skip, s));
  return linear(w.ln, T, W, T, bias, n, N, T, out, ldo, epi, nullptr, 0, skip, s, nullptr, pf);
}

// `ln1_done`: w.ln already holds ln_1(x) (written by the previous block's fused tail); `next_g/next_b`: the LayerNorm
// that follows this block (next block's ln_1 or ln_final), fused into c_proj when `fuse_ln` is on.
int Model::self_block(const Block& b, const DecodeWs& w, bf16* kc, bf16* vc, int n, int max_len, int heads,
                      bool ln1_done, const float* next_g, const float* next_b, Prefetch next_pf, cudaStream_t s) {
  const int T = cfg.text_width;
  const int* skip = &w.st->done;
  const bool fuse = decoder_fuse(n);
  const size_t TT = (size_t)T * T * sizeof(bf16);
  const Prefetch none;
  // every GEMM pulls the next GEMM's weights into L2 (the 557 MB of decoder weights cycle through L2 once per step)
  if (!ln1_done)
    EC_TRY(ln_linear(w, b.ln1_g, b.ln1_b, b.w_qkv, b.b_qkv, n, 3 * T, w.qkv, 3 * T, EPI_BF16, skip,
                     prefetch_weights ? Prefetch{b.w_out, TT} : none, s));
  else
    EC_TRY(linear(w.ln, T, b.w_qkv, T, b.b_qkv, n, 3 * T, T, w.qkv, 3 * T, EPI_BF16, nullptr, 0, skip, s, nullptr,
                  prefetch_weights ? Prefetch{b.w_out, TT} : none));
  EC_TRY(decode_self_attention(w.qkv, kc, vc, w.att, n, heads, T / heads, max_len, w.finished, w.st, s));
  FusedLn ln2{b.ln2_g, b.ln2_b, fuse ? w.ln : nullptr, T, w.ln_cnt};
  EC_TRY(linear(w.att, T, b.w_out, T, b.b_out, n, T, T, w.x, T, EPI_F32_RESID, w.x, T, skip, s, &ln2,
                prefetch_weights ? Prefetch{b.w_fc1, 4 * TT} : none));
  if (!fuse)
    EC_TRY(ln_linear(w, b.ln2_g, b.ln2_b, b.w_fc1, b.b_fc1, n, 4 * T, w.fc1, 4 * T, EPI_BF16_GELU, skip,
                     prefetch_weights ? Prefetch{b.w_fc2, 4 * TT} : none, s));
  else
    EC_TRY(linear(w.ln, T, b.w_fc1, T, b.b_fc1, n, 4 * T, T, w.fc1, 4 * T, EPI_BF16_GELU, nullptr, 0, skip, s, nullptr,
                  prefetch_weights ? Prefetch{b.w_fc2, 4 * TT} : none));
  FusedLn nxt{next_g, next_b, fuse ? w.ln : nullptr, T, w.ln_cnt};
  EC_TRY(linear(w.fc1, 4 * T, b.w_fc2, 4 * T, b.b_fc2, n, T, 4 * T, w.x, T, EPI_F32_RESID, w.x, T, skip, s, &nxt,
                prefetch_weights ? next_pf : none, SplitK{w.sk_ws, w.sk_cnt}));
  return EC_OK;
}

// `layer` >= 0: absorbed-projection cross-attention with that layer's absorbed weights (kv unused); -1: K/V cache `kv`
int Model::cross_block(const Block& b, const DecodeWs& w, const bf16* kv, int layer, int n, bool ln1_done,
                       const float* next_g, const float* next_b, Prefetch next_pf, cudaStream_t s) {
  const int T = cfg.text_width, NK = cfg.n_queries - 1;
  const int* skip = &w.st->done;
  const bool fuse = decoder_fuse(n);
  const size_t TT = (size_t)T * T * sizeof(bf16);
  const Prefetch none;
  if (!ln1_done)
    EC_TRY(ln_linear(w, b.ln1_g, b.ln1_b, b.w_qkv, b.b_qkv, n, T, w.q, T, EPI_BF16, skip,
                     prefetch_weights ? Prefetch{b.w_out, TT} : none, s));   // q rows only
  else
    EC_TRY(linear(w.ln, T, b.w_qkv, T, b.b_qkv, n, T, T, w.q, T, EPI_BF16, nullptr, 0, skip, s, nullptr,
                  prefetch_weights ? Prefetch{b.w_out, TT} : none));   // q rows only
  if (layer >= 0) {
    // absorbed form (cross_absorbed.cu): q' = per-head W_k^T q / 8 (grouped GEMM, K = 64), one pass over the normalised
    // image tokens, then the per-head V projection of the contexts (grouped GEMM, K = T) -> w.att
    const int H = cfg.mm_heads, hd = T / H;
    GemmArgs gq{};
    gq.A = w.q; gq.lda = T; gq.W = wk_abs[layer]; gq.ldw = hd; gq.M = n; gq.N = H * T; gq.K = hd;
    gq.out = w.qp; gq.ldo = H * T; gq.skip_flag = skip; gq.group_cols = T;
    EC_TRY(gemm_bf16(gq, EPI_BF16, s));
    static const int env_rev = [] { const char* e = getenv("EC_XA_REVERSE"); return e ? atoi(e) : 1; }();
    EC_TRY(decode_cross_absorbed(w.qp, w.ln_kv, w.ctx, n, H, T, NK, w.finished, w.st, env_rev != 0 && (layer & 1), s));
    GemmArgs gv{};
    gv.A = w.ctx; gv.lda = H * T; gv.W = wkv_fold[layer] + (size_t)T * T; gv.ldw = T; gv.M = n; gv.N = T; gv.K = T;
    gv.bias = bkv_fold[layer] + T; gv.out = w.att; gv.ldo = T; gv.skip_flag = skip; gv.group_cols = hd; gv.bn_hint = 64;
    EC_TRY(gemm_bf16(gv, EPI_BF16, s));
  } else {
    EC_TRY(decode_cross_attention(w.q, kv, w.att, n, cfg.mm_heads, T / cfg.mm_heads, NK, w.finished, w.st, s));
  }
  FusedLn ln2{b.ln2_g, b.ln2_b, fuse ? w.ln : nullptr, T, w.ln_cnt};
  EC_TRY(linear(w.att, T, b.w_out, T, b.b_out, n, T, T, w.x, T, EPI_F32_RESID, w.x, T, skip, s, &ln2,
                prefetch_weights ? Prefetch{b.w_fc1, 4 * TT} : none));
  if (!fuse)
    EC_TRY(ln_linear(w, b.ln2_g, b.ln2_b, b.w_fc1, b.b_fc1, n, 4 * T, w.fc1, 4 * T, EPI_BF16_GELU, skip,
                     prefetch_weights ? Prefetch{b.w_fc2, 4 * TT} : none, s));
  else
    EC_TRY(linear(w.ln, T, b.w_fc1, T, b.b_fc1, n, 4 * T, T, w.fc1, 4 * T, EPI_BF16_GELU, nullptr, 0, skip, s, nullptr,
                  prefetch_weights ? Prefetch{b.w_fc2, 4 * TT} : none));
  FusedLn nxt{next_g, next_b, fuse ? w.ln : nullptr, T, w.ln_cnt};
  EC_TRY(linear(w.fc1, 4 * T, b.w_fc2, 4 * T, b.b_fc2, n, T, 4 * T, w.x, T, EPI_F32_RESID, w.x, T, skip, s, &nxt,
                prefetch_weights ? next_pf : none, SplitK{w.sk_ws, w.sk_cnt}));
  return EC_OK;
}

// one decoder step for the crops of one chain (`w` is the chain's row view): token at position cur_len-1 in,
// token cur_len out.  Chains only share the read-only step state and the `next_active` counter.
int Model::chain_step(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                      float* logits, bool teacher_forced, cudaStream_t s) {
  if (use_seq && w.seq_cnt != nullptr && cfg.text_layers + cfg.mm_layers > 0 && !decoder_fuse(n))
    return chain_step_seq(w, tokens, ld_tokens, n, n_total, seq_len, min_len, logits, teacher_forced, s);
  const int T = cfg.text_width;
  EC_TRY(decode_embed(tokens, ld_tokens, tok_emb, txt_pos, w.x, n, T, w.st, s));
  // block sequence: text tower, then (self, cross) pairs, then ln_final; each block's c_proj GEMM also produces the
  // LayerNorm the NEXT block starts with (fused tail), so only the very first LayerNorm of the step is a launch
  const bool fuse = decoder_fuse(n);
  for (int l = 0; l < cfg.text_layers; ++l) {
    const bool last_text = (l + 1 == cfg.text_layers);
    const Block& nb = last_text ? dself[0] : tblocks[l + 1];
    EC_TRY(self_block(tblocks[l], w, w.tk + l * w.cache_layer, w.tv + l * w.cache_layer, n, seq_len, cfg.text_heads,
                      fuse && l > 0, nb.ln1_g, nb.ln1_b, Prefetch{nb.w_qkv, 3 * (size_t)T * T * sizeof(bf16)}, s));
  }
  for (int l = 0; l < cfg.mm_layers; ++l) {
    const size_t TTb = (size_t)T * T * sizeof(bf16);
    EC_TRY(self_block(dself[l], w, w.dk + l * w.cache_layer, w.dv + l * w.cache_layer, n, seq_len, cfg.mm_heads,
                      fuse && (l > 0 || cfg.text_layers > 0), dcross[l].ln1_g, dcross[l].ln1_b,
                      Prefetch{dcross[l].w_qkv, TTb}, s));
    const bool last = (l + 1 == cfg.mm_layers);
    EC_TRY(cross_block(dcross[l], w, w.cross_kv + l * w.cross_layer, use_absorbed() ? l : -1, n, fuse,
                       last ? ln_final_g : dself[l + 1].ln1_g,
                       last ? ln_final_b : dself[l + 1].ln1_b,
                       last ? Prefetch{tblocks.empty() ? nullptr : tblocks[0].w_qkv, 3 * TTb} : Prefetch{dself[l + 1].w_qkv, 3 * TTb},
                       s));
  }
  if (!fuse) EC_TRY(layernorm(w.x, T, ln_final_g, ln_final_b, w.ln, T, true, n, T, &w.st->done, s));
  return vocab_head(w, tokens, ld_tokens, n, n_total, seq_len, min_len, logits, teacher_forced, s);
}

// vocabulary projection of ln_final(x) (in w.ln) with the arg-max epilogue, then the token selection
int Model::vocab_head(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                      float* logits, bool teacher_forced, cudaStream_t s) {
  const int T = cfg.text_width, V = cfg.vocab_size;
  GemmArgs g{};
  g.A = w.ln; g.lda = T; g.W = w_vocab; g.ldw = T; g.M = n; g.N = V; g.K = T;
  g.logits = logits; g.ld_logits = V; g.logits_step_stride = (long long)n_total * V;
  g.part_val = w.part_val; g.part_idx = w.part_idx; g.n_parts = gemm_argmax_parts(V);
  g.eos_col = teacher_forced ? -1 : cfg.eos_token_id; g.cur_len = &w.st->cur_len; g.min_len = min_len;
  g.skip_flag = &w.st->done;
  EC_TRY(gemm_bf16(g, EPI_ARGMAX, s));
  if (!teacher_forced)
    EC_TRY(decode_select(w.part_val, w.part_idx, g.n_parts, tokens, ld_tokens, w.finished, n, seq_len,
                         cfg.eos_token_id, cfg.pad_token_id, w.st, s));
  return EC_OK;
}

// A list of dependent GEMM stages (+ LayerNorm tails): one persistent launch when the shapes allow it, else the stages
// one after the other with stand-alone LayerNorms (same arithmetic, same results).
int Model::run_stages(const GemmSeqStage* st, int ns, int* counters, const int* skip, Prefetch pf, int n, cudaStream_t s,
                      int tag) {
  static const int env_max = [] { const char* e = getenv("EC_SEQ_MAX"); return e ? atoi(e) : 0; }();   // probe: stages per launch
  if (counters != nullptr && gemm_seq_supported(st, ns)) {
    if (env_max > 0 && env_max < ns) {
      // every group needs its own zeroed counter block: groups of one launch share `counters` at stage offsets
      for (int i = 0; i < ns; i += env_max) {
        const int cnt = std::min(env_max, ns - i);
        EC_TRY(gemm_seq(st + i, cnt, counters + (size_t)i * (kGemmSeqCounterInts / kGemmSeqMaxStages), skip, nullptr, 0, s,
                        i == 0 ? tag : -1));
      }
      return EC_OK;
    }
    return gemm_seq(st, ns, counters, skip, prefetch_weights ? pf.ptr : nullptr, prefetch_weights ? pf.bytes : 0, s, tag);
  }
  for (int i = 0; i < ns; ++i) {
    GemmArgs a = st[i].a;
    const float* g = a.ln_gamma; const float* b = a.ln_beta; bf16* lo = a.ln_out; const int ld = a.ld_ln;
    a.ln_gamma = a.ln_beta = nullptr; a.ln_out = nullptr; a.ld_ln = 0;
    a.skip_flag = skip;
    EC_TRY(gemm_bf16(a, st[i].epi, s));
    if (lo != nullptr)
      EC_TRY(layernorm(reinterpret_cast<const float*>(a.out), a.ldo, g, b, lo, ld, true, n, a.N, skip, s));
  }
  return EC_OK;
}

// The decoder step with GEMM sequences: per block one attention kernel and ONE persistent launch for everything up to
// the next attention -- [V-projection of the absorbed contexts] -> out_proj (+x) -> ln_2 -> FC1 + GELU -> FC2 (+x) ->
// the next block's ln_1 -> its QKV (self) or Q + absorbed Q' (cross).  ~80 graph nodes per step instead of ~280.
int Model::chain_step_seq(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                          float* logits, bool teacher_forced, cudaStream_t s) {
  const int T = cfg.text_width, NK = cfg.n_queries - 1;
  const int* skip = &w.st->done;
  struct Blk { const Block* b; int cross; int layer; bf16* kc; bf16* vc; int heads; };
  std::vector<Blk> blks;
  for (int l = 0; l < cfg.text_layers; ++l)
    blks.push_back(Blk{&tblocks[l], 0, l, w.tk + l * w.cache_layer, w.tv + l * w.cache_layer, cfg.text_heads});
  for (int l = 0; l < cfg.mm_layers; ++l) {
    blks.push_back(Blk{&dself[l], 0, l, w.dk + l * w.cache_layer, w.dv + l * w.cache_layer, cfg.mm_heads});
    blks.push_back(Blk{&dcross[l], 1, l, nullptr, nullptr, cfg.mm_heads});
  }
  const int nb = (int)blks.size();
  EC_REQUIRE(nb > 0 && nb <= kSeqLaunchesPerStep && blks[0].cross == 0, EC_ERR_UNSUPPORTED, "decoder without blocks");
  const bool absorbed = use_absorbed();
  const int H = cfg.mm_heads, hd = T / H;
  const size_t TTb = (size_t)T * T * sizeof(bf16);

  EC_TRY(decode_embed(tokens, ld_tokens, tok_emb, txt_pos, w.x, n, T, w.st, s));
  EC_CHECK_CUDA(cudaMemsetAsync(w.seq_cnt, 0, (size_t)nb * kGemmSeqCounterInts * sizeof(int), s));
  EC_TRY(layernorm(w.x, T, blks[0].b->ln1_g, blks[0].b->ln1_b, w.ln, T, true, n, T, skip, s));
  EC_TRY(linear(w.ln, T, blks[0].b->w_qkv, T, blks[0].b->b_qkv, n, 3 * T, T, w.qkv, 3 * T, EPI_BF16, nullptr, 0, skip, s,
                nullptr, prefetch_weights ? Prefetch{blks[0].b->w_out, TTb} : Prefetch()));

  auto stage = [&](const bf16* A, int lda, const bf16* W, int ldw, const float* bias, int N, int K, void* out, int ldo,
                   GemmEpilogue epi) {
    GemmSeqStage st{};
    st.a.A = A; st.a.lda = lda; st.a.W = W; st.a.ldw = ldw; st.a.bias = bias; st.a.M = n; st.a.N = N; st.a.K = K;
    st.a.out = out; st.a.ldo = ldo; st.epi = epi;
    if (epi == EPI_F32_RESID) { st.a.residual = reinterpret_cast<const float*>(out); st.a.ldr = ldo; }
    return st;
  };
  for (int i = 0; i < nb; ++i) {
    const Blk& k = blks[i];
    const Block& b = *k.b;
    GemmSeqStage st[kGemmSeqMaxStages];
    int ns = 0;
    if (!k.cross) {
      EC_TRY(decode_self_attention(w.qkv, k.kc, k.vc, w.att, n, k.heads, T / k.heads, seq_len, w.finished, w.st, s));
    } else if (absorbed) {
      static const int env_rev = [] { const char* e = getenv("EC_XA_REVERSE"); return e ? atoi(e) : 1; }();
      EC_TRY(decode_cross_absorbed(w.qp, w.ln_kv, w.ctx, n, H, T, NK, w.finished, w.st, env_rev != 0 && (k.layer & 1), s));
      st[ns] = stage(w.ctx, H * T, wkv_fold[k.layer] + (size_t)T * T, T, bkv_fold[k.layer] + T, T, T, w.att, T, EPI_BF16);
      st[ns].a.group_cols = hd; st[ns].a.bn_hint = 64;
      ++ns;
    } else {
      EC_TRY(decode_cross_attention(w.q, w.cross_kv + k.layer * w.cross_layer, w.att, n, H, hd, NK, w.finished, w.st, s));
    }
    st[ns] = stage(w.att, T, b.w_out, T, b.b_out, T, T, w.x, T, EPI_F32_RESID);
    st[ns].a.ln_gamma = b.ln2_g; st[ns].a.ln_beta = b.ln2_b; st[ns].a.ln_out = w.ln; st[ns].a.ld_ln = T;
    ++ns;
    st[ns++] = stage(w.ln, T, b.w_fc1, T, b.b_fc1, 4 * T, T, w.fc1, 4 * T, EPI_BF16_GELU);
    st[ns] = stage(w.fc1, 4 * T, b.w_fc2, 4 * T, b.b_fc2, T, 4 * T, w.x, T, EPI_F32_RESID);
    const bool last = (i + 1 == nb);
    const Block* nxt = last ? nullptr : blks[i + 1].b;
    st[ns].a.ln_gamma = last ? ln_final_g : nxt->ln1_g; st[ns].a.ln_beta = last ? ln_final_b : nxt->ln1_b;
    st[ns].a.ln_out = w.ln; st[ns].a.ld_ln = T;
    ++ns;
    Prefetch pf;
    if (!last) {
      const Blk& kn = blks[i + 1];
      if (!kn.cross) {
        st[ns++] = stage(w.ln, T, nxt->w_qkv, T, nxt->b_qkv, 3 * T, T, w.qkv, 3 * T, EPI_BF16);
      } else {
        st[ns++] = stage(w.ln, T, nxt->w_qkv, T, nxt->b_qkv, T, T, w.q, T, EPI_BF16);      // q rows only
        if (absorbed) {
          st[ns] = stage(w.q, T, wk_abs[kn.layer], hd, nullptr, H * T, hd, w.qp, H * T, EPI_BF16);
          st[ns].a.group_cols = T;
          ++ns;
        }
      }
      pf = Prefetch{nxt->w_out, TTb};       // the first weights of the next sequence
    } else {
      pf = Prefetch{blks[0].b->w_qkv, 3 * TTb};
    }
    EC_TRY(run_stages(st, ns, w.seq_cnt ? w.seq_cnt + (size_t)i * kGemmSeqCounterInts : nullptr, skip, pf, n, s, i));
  }
  return vocab_head(w, tokens, ld_tokens, n, n_total, seq_len, min_len, logits, teacher_forced, s);
}

// Rows per chain.  The batch can be cut into independent row chains; each chain owns a stream, a device-resident
// DecodeState and a LINEAR captured step graph replayed on that stream (measured on B200: parallel branches inside
// one graph double the dependent-kernel latency, 12.6 vs 6.3 us per level, whereas separate linear graphs on
// separate streams overlap almost perfectly).  Measured result for this decoder (profiles/r01_decoder_chains.txt):
// the step is a chain of ~250 DEPENDENT kernels whose individual latency (~10 us per GEMM level, set by each CTA
// streaming its own K panel at ~85 GB/s per SM) does not shrink with fewer rows, so chaining neither helps nor
// hurts at 1024 crops (122 ms with 1 chain, 125 ms with 8).  Default: one chain; the option stays for larger
// batches, where several chains share one latency chain.
int Model::chain_rows(int n) const {
  static const int env_chains = [] { const char* e = getenv("EC_DECODE_CHAINS"); return e ? atoi(e) : 0; }();
  int nc = decode_chains > 0 ? decode_chains : env_chains;
  if (nc <= 0) nc = 1;
  nc = std::max(1, std::min(nc, kMaxChains));
  const int per = ceil_div(ceil_div(n, nc), 128) * 128;     // whole GEMM row tiles per chain
  return std::max(per, 1);
}

int Model::decode_step(const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                       float* logits, bool teacher_forced, cudaStream_t s) {
  EC_TRY(chain_step(w, tokens, ld_tokens, n, n_total, seq_len, min_len, logits, teacher_forced, s));
  return decode_advance(w.st, teacher_forced, n, s);
}

int Model::prefill_cross_kv(const DecodeWs& w, const float* embs, int n, cudaStream_t s) {
  const int T = cfg.text_width, NK = cfg.n_queries - 1;
  const int rows = n * NK;
  if (shared_lnkv && (int)wkv_fold.size() == cfg.mm_layers) {
    // one normalisation of the image tokens for all layers; gamma / beta sit in the pre-multiplied weights (finalize)
    // absorbed form: the decoder attends over the normalised tokens themselves, staged in the packed layout
    if (use_absorbed()) return pack_image_tokens(embs, w.ln_kv, n, NK, T, true, s);
    EC_TRY(layernorm(embs, T, unit_gamma, zero_beta, w.ln_kv, T, true, rows, T, nullptr, s));
    for (int l = 0; l < cfg.mm_layers; ++l)
      EC_TRY(linear(w.ln_kv, T, wkv_fold[l], T, bkv_fold[l], rows, 2 * T, T, w.cross_kv + (size_t)l * rows * 2 * T,
                    2 * T, EPI_BF16, nullptr, 0, nullptr, s));
    return EC_OK;
  }
  for (int l = 0; l < cfg.mm_layers; ++l) {
    const Block& b = dcross[l];
    EC_TRY(layernorm(embs, T, b.lnkv_g, b.lnkv_b, w.ln_kv, T, true, rows, T, nullptr, s));
    EC_TRY(linear(w.ln_kv, T, b.w_qkv + (size_t)T * T, T, b.b_qkv + T, rows, 2 * T, T,
                  w.cross_kv + (size_t)l * rows * 2 * T, 2 * T, EPI_BF16, nullptr, 0, nullptr, s));
  }
  return EC_OK;
}

// `steps` decoder steps of one chain on its own stream `cs` (graph replay or eager launches)
int Model::run_chain(int c, const DecodeWs& w, int* tokens, int ld_tokens, int n, int n_total, int seq_len, int min_len,
                     float* logits, bool teacher_forced, int steps, cudaStream_t cs) {
  if (!use_graph) {
    for (int t = 0; t < steps; ++t)
      EC_TRY(decode_step(w, tokens, ld_tokens, n, n_total, seq_len, min_len, logits, teacher_forced, cs));
    return EC_OK;
  }
  // capture one step on the chain's private stream, replay it `steps` times; re-capture only when the launch set changes
  GraphKey key{w.x, tokens, logits, n, seq_len, min_len, ld_tokens, teacher_forced ? 1 : 0, n_total};
  if (!graph_exec[c] || memcmp(&key, &graph_key[c], sizeof(GraphKey)) != 0) {
    if (graph_exec[c]) { cudaGraphExecDestroy(graph_exec[c]); graph_exec[c] = nullptr; }
    cudaGraph_t graph = nullptr;
    EC_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const long long before = launch_count();
    int rc = decode_step(w, tokens, ld_tokens, n, n_total, seq_len, min_len, logits, teacher_forced, cs);
    graph_nodes[c] = launch_count() - before;
    note_launches(-graph_nodes[c]);                      // captured, not launched
    cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (rc != EC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    EC_CHECK_CUDA(ce);
    ce = cudaGraphInstantiate(&graph_exec[c], graph, 0);
    cudaGraphDestroy(graph);
    EC_CHECK_CUDA(ce);
    graph_key[c] = key;
  }
  for (int t = 0; t < steps; ++t) EC_CHECK_CUDA(cudaGraphLaunch(graph_exec[c], cs));
  note_launches(graph_nodes[c] * steps);
  return EC_OK;
}

// fork the caller's stream into the row chains, run `steps` steps on each, join
int Model::run_steps(const DecodeWs& w, int* tokens, int ld_tokens, int n, int seq_len, int min_len, float* logits,
                     bool teacher_forced, int steps, cudaStream_t s) {
  if (chain_fork == nullptr) {
    for (int c = 0; c < kMaxChains; ++c) {
      EC_CHECK_CUDA(cudaStreamCreateWithFlags(&chain_streams[c], cudaStreamNonBlocking));
      EC_CHECK_CUDA(cudaEventCreateWithFlags(&chain_done[c], cudaEventDisableTiming));
    }
    EC_CHECK_CUDA(cudaEventCreateWithFlags(&chain_fork, cudaEventDisableTiming));
  }
  const int per = chain_rows(n);
  const int nc = ceil_div(n, per);
  // step state of every chain (stream-ordered before the fork)
  for (int c = 0; c < nc; ++c) {
    const int r0 = c * per, nr = std::min(per, n - r0);
    EC_TRY(decode_init(teacher_forced ? nullptr : tokens + (size_t)r0 * ld_tokens, ld_tokens, w.finished + r0, nr, seq_len,
                       cfg.sot_token_id, cfg.pad_token_id, w.st + c, s));
  }
  EC_CHECK_CUDA(cudaEventRecord(chain_fork, s));
  int rc = EC_OK;
  for (int c = 0; c < nc && rc == EC_OK; ++c) {
    const int r0 = c * per, nr = std::min(per, n - r0);
    cudaStream_t cs = chain_streams[c];
    EC_CHECK_CUDA(cudaStreamWaitEvent(cs, chain_fork, 0));
    const DecodeWs v = slice_rows(*this, w, r0, seq_len, c);
    rc = run_chain(c, v, tokens + (size_t)r0 * ld_tokens, ld_tokens, nr, n, seq_len, min_len,
                   logits ? logits + (size_t)r0 * cfg.vocab_size : nullptr, teacher_forced, steps, cs);
    EC_CHECK_CUDA(cudaEventRecord(chain_done[c], cs));
    EC_CHECK_CUDA(cudaStreamWaitEvent(s, chain_done[c], 0));
  }
  return rc;
}

int Model::decode_greedy(const float* embs, int n, int seq_len, int min_len, int* tokens, float* logits, void* ws,
                         size_t ws_bytes, cudaStream_t s) {
  EC_REQUIRE(finalized, EC_ERR_STATE, "model not finalized");
  EC_REQUIRE(n > 0, EC_ERR_INVALID, "decode_greedy: n = %d", n);
  EC_REQUIRE(seq_len > min_len, EC_ERR_INVALID, "seq_len must be larger than min_seq_len");   // coca_model.py:228
  EC_REQUIRE(seq_len <= cfg.context_length + 1, EC_ERR_UNSUPPORTED, "seq_len %d exceeds the positional table (%d)",
             seq_len, cfg.context_length + 1);
  EC_CHECK_CUDA(cudaSetDevice(device));
  const DecodeWs w = carve_decode(*this, ws, n, seq_len);
  EC_REQUIRE(ws_bytes >= w.bytes, EC_ERR_WORKSPACE, "decode workspace %zu < %zu", ws_bytes, w.bytes);
  EC_TRY(prefill_cross_kv(w, embs, n, s));
  // Decode into the workspace's own token table and copy out at the end: the caller's output buffer is usually a
  // fresh allocation, and its address is part of the step graph's key -- a new address would re-capture and
  // re-instantiate the 258-node graph on every call.
  EC_TRY(run_steps(w, w.tokens_tf, seq_len, n, seq_len, min_len, logits, false, seq_len - 1, s));
  EC_CHECK_CUDA(cudaMemcpyAsync(tokens, w.tokens_tf, (size_t)n * seq_len * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return EC_OK;
}

int Model::forward_logits(const float* embs, const int* text, int n, int t, float* logits, void* ws, size_t ws_bytes,
                          cudaStream_t s) {
  EC_REQUIRE(finalized, EC_ERR_STATE, "model not finalized");
  EC_REQUIRE(n > 0 && t > 0 && logits != nullptr, EC_ERR_INVALID, "forward_logits: bad arguments");
  EC_REQUIRE(t <= cfg.context_length, EC_ERR_UNSUPPORTED, "text length %d exceeds context %d", t, cfg.context_length);
  EC_CHECK_CUDA(cudaSetDevice(device));
  const int L = t + 1;
  const DecodeWs w = carve_decode(*this, ws, n, L);
  EC_REQUIRE(ws_bytes >= w.bytes, EC_ERR_WORKSPACE, "decode workspace %zu < %zu", ws_bytes, w.bytes);
  EC_TRY(prefill_cross_kv(w, embs, n, s));
  // teacher forcing: the step engine reads its input token from a [n, L] table holding the given text
  EC_CHECK_CUDA(cudaMemsetAsync(w.tokens_tf, 0, (size_t)n * L * sizeof(int), s));
  EC_CHECK_CUDA(cudaMemcpy2DAsync(w.tokens_tf, (size_t)L * sizeof(int), text, (size_t)t * sizeof(int),
                                  (size_t)t * sizeof(int), n, cudaMemcpyDeviceToDevice, s));
  return run_steps(w, w.tokens_tf, L, n, L, 0, logits, true, t, s);
}

}  // namespace ec
