// extern "C" surface of libec_b200 (include/ec_b200.h).
#include "model.cuh"
#include <new>

using namespace ec;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int check_device(int device) {
  int count = 0;
  EC_CHECK_CUDA(cudaGetDeviceCount(&count));
  EC_REQUIRE(device >= 0 && device < count, EC_ERR_DEVICE, "device %d out of range (%d visible)", device, count);
  int major = 0;
  EC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  EC_REQUIRE(major == 10, EC_ERR_DEVICE, "libec_b200 is built for sm_100a only; device %d has compute capability %d.x",
             device, major);
  return EC_OK;
}

extern "C" {

int ec_abi_version(void) { return EC_ABI_VERSION; }
const char* ec_last_error(void) { return last_error(); }

int ec_model_create(const ec_config* c, int device, ec_model** out) {
  EC_REQUIRE(c != nullptr && out != nullptr, EC_ERR_INVALID, "ec_model_create: null argument");
  *out = nullptr;
  EC_TRY(check_device(device));
  EC_REQUIRE(c->embed_dim == c->text_width, EC_ERR_UNSUPPORTED, "CoCa decoder width must equal embed_dim");
  EC_REQUIRE(c->image_size % c->patch_size == 0, EC_ERR_INVALID, "image_size not a multiple of patch_size");
  EC_REQUIRE(c->vision_width % c->vision_head_dim == 0 && (c->vision_head_dim == 32 || c->vision_head_dim == 64),
             EC_ERR_UNSUPPORTED, "vision head_dim %d unsupported", c->vision_head_dim);
  EC_REQUIRE(c->embed_dim % c->pooler_heads == 0, EC_ERR_INVALID, "embed_dim not divisible by pooler heads");
  const int phd = c->embed_dim / c->pooler_heads;
  EC_REQUIRE(phd == 32 || phd == 64 || phd == 96, EC_ERR_UNSUPPORTED, "pooler head_dim %d unsupported", phd);
  EC_REQUIRE(c->text_width == c->text_heads * 64 && c->text_width == c->mm_heads * 64, EC_ERR_UNSUPPORTED,
             "text / decoder head_dim must be 64");
  EC_REQUIRE(c->vision_width <= 1024 && c->text_width <= 1024, EC_ERR_UNSUPPORTED, "widths above 1024 unsupported");
  EC_REQUIRE(c->vision_width % 8 == 0 && c->embed_dim % 8 == 0 && c->vocab_size % 8 == 0, EC_ERR_UNSUPPORTED,
             "widths and vocab must be multiples of 8");
  EC_REQUIRE(c->n_queries >= 2 && c->n_queries - 1 <= 260, EC_ERR_UNSUPPORTED, "n_queries %d unsupported", c->n_queries);
  EC_REQUIRE(c->context_length + 1 <= 128, EC_ERR_UNSUPPORTED, "context_length %d unsupported", c->context_length);
  EC_CHECK_CUDA(cudaSetDevice(device));
  ec_model* m = new (std::nothrow) ec_model(*c, device);
  EC_REQUIRE(m != nullptr, EC_ERR_INVALID, "out of host memory");
  *out = m;
  return EC_OK;
}

int ec_model_destroy(ec_model* m) {
  delete m;
  return EC_OK;
}

int ec_model_set_tensor(ec_model* m, const char* name, const float* host_data, int64_t numel) {
  EC_REQUIRE(m && name && host_data, EC_ERR_INVALID, "ec_model_set_tensor: null argument");
  return m->impl.set_tensor(name, host_data, numel);
}

int ec_model_finalize(ec_model* m, void* stream) {
  EC_REQUIRE(m, EC_ERR_INVALID, "null model");
  return m->impl.finalize(as_stream(stream));
}

int ec_model_set_option(ec_model* m, const char* key, int value) {
  EC_REQUIRE(m && key, EC_ERR_INVALID, "null argument");
  if (!strcmp(key, "decode_graph")) { m->impl.use_graph = value != 0; return EC_OK; }
  if (!strcmp(key, "decode_chains")) {
    EC_REQUIRE(value >= 0 && value <= kMaxChains, EC_ERR_INVALID, "decode_chains must be in [0, %d]", kMaxChains);
    m->impl.decode_chains = value;
    return EC_OK;
  }
  if (!strcmp(key, "decode_seq")) {
    const bool on = value != 0 && m->impl.seq_counters != nullptr;
    if (on != m->impl.use_seq) {      // another launch set: the captured step graphs are stale
      for (int c = 0; c < kMaxChains; ++c)
        if (m->impl.graph_exec[c]) { cudaGraphExecDestroy(m->impl.graph_exec[c]); m->impl.graph_exec[c] = nullptr; }
    }
    m->impl.use_seq = on;
    return EC_OK;
  }
  if (!strcmp(key, "tc_attention")) { m->impl.use_tc_attention = value != 0; return EC_OK; }
  // takes effect at the next ec_model_finalize (the absorbed weights are built there); switching it off works at once
  if (!strcmp(key, "absorbed_cross")) { m->impl.absorbed_cross = value != 0; return EC_OK; }
  set_error("unknown option %s", key);
  return EC_ERR_INVALID;
}

int ec_preprocess_crops(const uint8_t* frames, int32_t n_frames, int32_t fh, int32_t fw, int32_t fc,
                        const int32_t* rects, const int32_t* frame_idx, const int32_t* flags, int32_t n,
                        int32_t out_size, void* out_bf16, float* out_f32, int32_t* status, void* stream) {
  EC_REQUIRE(frames && rects && frame_idx, EC_ERR_INVALID, "ec_preprocess_crops: null argument");
  EC_REQUIRE(out_bf16 || out_f32, EC_ERR_INVALID, "ec_preprocess_crops: no output buffer");
  return preprocess_crops(frames, n_frames, fh, fw, fc, rects, frame_idx, flags, n, out_size,
                          reinterpret_cast<bf16*>(out_bf16), out_f32, status, as_stream(stream));
}

size_t ec_encode_workspace_bytes(const ec_model* m, int32_t n) { return m ? m->impl.encode_workspace_bytes(n) : 0; }

int ec_encode_image(ec_model* m, const void* x, int32_t n, float* latent, float* embs, void* ws, size_t ws_bytes,
                    void* stream) {
  EC_REQUIRE(m && x && ws, EC_ERR_INVALID, "ec_encode_image: null argument");
  return m->impl.encode_image(reinterpret_cast<const bf16*>(x), n, latent, embs, ws, ws_bytes, as_stream(stream));
}

size_t ec_decode_workspace_bytes(const ec_model* m, int32_t n, int32_t seq_len) {
  return m ? m->impl.decode_workspace_bytes(n, seq_len) : 0;
}

int ec_decode_greedy(ec_model* m, const float* embs, int32_t n, int32_t seq_len, int32_t min_seq_len, int32_t* tokens,
                     float* logits, void* ws, size_t ws_bytes, void* stream) {
  EC_REQUIRE(m && embs && tokens && ws, EC_ERR_INVALID, "ec_decode_greedy: null argument");
  return m->impl.decode_greedy(embs, n, seq_len, min_seq_len, tokens, logits, ws, ws_bytes, as_stream(stream));
}

int ec_forward_logits(ec_model* m, const float* embs, const int32_t* text, int32_t n, int32_t t, float* logits,
                      void* ws, size_t ws_bytes, void* stream) {
  EC_REQUIRE(m && embs && text && ws, EC_ERR_INVALID, "ec_forward_logits: null argument");
  return m->impl.forward_logits(embs, text, n, t, logits, ws, ws_bytes, as_stream(stream));
}

size_t ec_disagreement_workspace_bytes(int32_t n) { return disagreement_workspace_bytes(n); }

int ec_instance_disagreement(const float* emb, int64_t ld_emb, const int64_t* keys, int32_t n, int32_t d, int32_t mode,
                             int64_t* out_keys, float* out_score, int32_t* out_count, int32_t* out_m, void* ws,
                             size_t ws_bytes, void* stream) {
  EC_REQUIRE(out_keys && out_score && out_count && out_m, EC_ERR_INVALID, "ec_instance_disagreement: null output");
  EC_REQUIRE(n == 0 || (emb && keys && ws), EC_ERR_INVALID, "ec_instance_disagreement: null input");
  return instance_disagreement(emb, ld_emb, reinterpret_cast<const long long*>(keys), n, d, mode,
                               reinterpret_cast<long long*>(out_keys), out_score, out_count, out_m, ws, ws_bytes,
                               as_stream(stream));
}

int ec_instance_disagreement_peers(const float* const* peers, int32_t world, int32_t rows_per_rank, int64_t ld_emb,
                                   const int64_t* keys, int32_t n, int32_t d, int32_t mode, int32_t part_rank,
                                   int32_t part_world, int64_t* out_keys, float* out_score, int32_t* out_count,
                                   int32_t* out_m, void* ws, size_t ws_bytes, void* stream) {
  EC_REQUIRE(out_keys && out_score && out_count && out_m, EC_ERR_INVALID, "ec_instance_disagreement_peers: null output");
  EC_REQUIRE(part_world >= 1 && part_rank >= 0 && part_rank < part_world, EC_ERR_INVALID, "bad partition");
  return instance_disagreement_peers(peers, world, rows_per_rank, ld_emb, reinterpret_cast<const long long*>(keys), n, d,
                                     mode, part_rank, part_world, reinterpret_cast<long long*>(out_keys), out_score,
                                     out_count, out_m, ws, ws_bytes, as_stream(stream));
}

size_t ec_caption_frequency_workspace_bytes(int32_t n) { return caption_frequency_workspace_bytes(n); }

int ec_caption_frequency(const int32_t* tokens, int32_t ld_tokens, int32_t t, const int64_t* keys, int32_t n,
                         int64_t* out_key, int32_t* out_first, int32_t* out_count, int32_t* out_m, void* ws,
                         size_t ws_bytes, void* stream) {
  EC_REQUIRE(out_key && out_first && out_count && out_m, EC_ERR_INVALID, "ec_caption_frequency: null output");
  EC_REQUIRE(n == 0 || (tokens && keys && ws), EC_ERR_INVALID, "ec_caption_frequency: null input");
  return caption_frequency(tokens, ld_tokens, t, reinterpret_cast<const long long*>(keys), n,
                           reinterpret_cast<long long*>(out_key), out_first, out_count, out_m, ws, ws_bytes,
                           as_stream(stream));
}

size_t ec_disagreement_map_workspace_bytes(int32_t n_points, int32_t h, int32_t w) {
  return disagreement_map_workspace_bytes(n_points, h, w);
}

int ec_disagreement_map(const float* points, int32_t n_points, const int64_t* point_keys, const int64_t* inst_keys,
                        const float* inst_scores, int32_t m, float scale, int32_t off_x, int32_t off_y, int32_t h,
                        int32_t w, float* map, float* reward, void* ws, size_t ws_bytes, void* stream) {
  EC_REQUIRE(map && ws, EC_ERR_INVALID, "ec_disagreement_map: null output");
  EC_REQUIRE(n_points == 0 || (points && point_keys), EC_ERR_INVALID, "ec_disagreement_map: null points");
  EC_REQUIRE(m == 0 || (inst_keys && inst_scores), EC_ERR_INVALID, "ec_disagreement_map: null instance table");
  return disagreement_map(points, n_points, reinterpret_cast<const long long*>(point_keys),
                          reinterpret_cast<const long long*>(inst_keys), inst_scores, m, scale, off_x, off_y, h, w, map,
                          reward, ws, ws_bytes, as_stream(stream));
}

int ec_gemm_bf16(const void* a, int32_t lda, const void* w, int32_t ldw, int32_t m, int32_t n, int32_t k,
                 const float* bias, const float* residual, int32_t ldr, void* out, int32_t ldo, int32_t epilogue,
                 int32_t use_simt_reference, void* stream) {
  EC_REQUIRE(epilogue >= 0 && epilogue <= 3, EC_ERR_INVALID, "ec_gemm_bf16: epilogue %d", epilogue);
  GemmArgs g{};
  g.A = reinterpret_cast<const bf16*>(a); g.lda = lda; g.W = reinterpret_cast<const bf16*>(w); g.ldw = ldw;
  g.M = m; g.N = n; g.K = k; g.bias = bias; g.residual = residual; g.ldr = ldr; g.out = out; g.ldo = ldo;
  if (use_simt_reference) return gemm_reference_simt(g, (GemmEpilogue)epilogue, as_stream(stream));
  return gemm_bf16(g, (GemmEpilogue)epilogue, as_stream(stream));
}

int ec_gemm_bf16_ln(const void* a, int32_t lda, const void* w, int32_t ldw, int32_t m, int32_t n, int32_t k,
                    const float* bias, const float* residual, int32_t ldr, float* out, int32_t ldo, const float* ln_gamma,
                    const float* ln_beta, void* ln_out_bf16, int32_t ld_ln, int32_t* ln_counters, void* stream) {
  EC_REQUIRE(a && w && residual && out && ln_gamma && ln_beta && ln_out_bf16 && ln_counters, EC_ERR_INVALID,
             "ec_gemm_bf16_ln: null argument");
  GemmArgs g{};
  g.A = reinterpret_cast<const bf16*>(a); g.lda = lda; g.W = reinterpret_cast<const bf16*>(w); g.ldw = ldw;
  g.M = m; g.N = n; g.K = k; g.bias = bias; g.residual = residual; g.ldr = ldr; g.out = out; g.ldo = ldo;
  g.ln_gamma = ln_gamma; g.ln_beta = ln_beta; g.ln_out = reinterpret_cast<bf16*>(ln_out_bf16); g.ld_ln = ld_ln;
  g.ln_counters = ln_counters;
  return gemm_bf16(g, EPI_F32_RESID, as_stream(stream));
}

// arg-max of each logits row through the fused vocab-head epilogue (+ the select kernel's reduction)
__global__ void argmax_parts_kernel(const float* __restrict__ pv, const int* __restrict__ pi, int n_parts, int rows,
                                    int* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int p = 0; p < n_parts; ++p) {
    const float v = pv[(size_t)r * n_parts + p];
    const int i = pi[(size_t)r * n_parts + p];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
  out[r] = bi;
}

int ec_gemm_argmax(const void* a, int32_t lda, const void* w, int32_t ldw, int32_t m, int32_t n, int32_t k,
                   float* logits, int32_t* argmax, void* ws, size_t ws_bytes, void* stream) {
  EC_REQUIRE(a && w && argmax && ws, EC_ERR_INVALID, "ec_gemm_argmax: null argument");
  const int parts = gemm_argmax_parts(n);
  const size_t need = align_up((size_t)m * parts * sizeof(float), 256) + (size_t)m * parts * sizeof(int);
  EC_REQUIRE(ws_bytes >= need, EC_ERR_WORKSPACE, "ec_gemm_argmax: workspace %zu < %zu", ws_bytes, need);
  GemmArgs g{};
  g.A = reinterpret_cast<const bf16*>(a); g.lda = lda; g.W = reinterpret_cast<const bf16*>(w); g.ldw = ldw;
  g.M = m; g.N = n; g.K = k; g.logits = logits; g.ld_logits = n;
  g.part_val = reinterpret_cast<float*>(ws);
  g.part_idx = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws) + align_up((size_t)m * parts * sizeof(float), 256));
  g.n_parts = parts; g.eos_col = -1; g.min_len = 0;
  EC_TRY(gemm_bf16(g, EPI_ARGMAX, as_stream(stream)));
  argmax_parts_kernel<<<ceil_div(m, 128), 128, 0, as_stream(stream)>>>(g.part_val, g.part_idx, parts, m, argmax);
  EC_POST_LAUNCH();
  return EC_OK;
}

int ec_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                 int32_t y_is_bf16, int32_t rows, int32_t width, void* stream) {
  EC_REQUIRE(x && gamma && beta && y, EC_ERR_INVALID, "ec_layernorm: null argument");
  return layernorm(x, ldx, gamma, beta, y, ldy, y_is_bf16 != 0, rows, width, nullptr, as_stream(stream));
}

int ec_attention(const void* q, int64_t q_bs, int64_t q_rs, const void* k, int64_t k_bs, int64_t k_rs, const void* v,
                 int64_t v_bs, int64_t v_rs, void* o, int64_t o_bs, int64_t o_rs, int32_t batch, int32_t heads,
                 int32_t sq, int32_t skv, int32_t head_dim, void* stream) {
  EC_REQUIRE(q && k && v && o, EC_ERR_INVALID, "ec_attention: null argument");
  AttnArgs a{};
  a.q = reinterpret_cast<const bf16*>(q); a.q_bs = q_bs; a.q_rs = q_rs;
  a.k = reinterpret_cast<const bf16*>(k); a.k_bs = k_bs; a.k_rs = k_rs;
  a.v = reinterpret_cast<const bf16*>(v); a.v_bs = v_bs; a.v_rs = v_rs;
  a.o = reinterpret_cast<bf16*>(o); a.o_bs = o_bs; a.o_rs = o_rs;
  a.batch = batch; a.heads = heads; a.sq = sq; a.skv = skv; a.head_dim = head_dim;
  a.scale = 1.0f / sqrtf((float)head_dim);
  return attention_full(a, as_stream(stream));
}

int ec_check_watchdog(void* stream) {
  EC_TRY(decode_cross_absorbed_check_watchdog(as_stream(stream)));
  EC_TRY(gemm_check_watchdog(as_stream(stream)));
  return attention_sm100_check_watchdog(as_stream(stream));
}

int ec_attention_fused_qkv(const void* qkv, int64_t ld, int32_t q_col0, int32_t k_col0, int32_t v_col0, void* out,
                           int64_t o_bs, int64_t o_rs, int32_t batch, int32_t heads, int32_t seq, int32_t head_dim,
                           int32_t impl, void* stream) {
  EC_REQUIRE(qkv && out, EC_ERR_INVALID, "ec_attention_fused_qkv: null argument");
  const bf16* base = reinterpret_cast<const bf16*>(qkv);
  if (impl == 1 || impl == 2) {      // 2: the whole-row tcgen05 kernel also for 257 tokens (A/B against the block-wise one)
    EC_REQUIRE(attention_sm100_supported(seq, head_dim), EC_ERR_UNSUPPORTED,
               "tcgen05 attention supports head_dim 64 and seq <= 258 (got %d, %d)", head_dim, seq);
    return attention_sm100_fused_qkv(base, ld, q_col0, k_col0, v_col0, reinterpret_cast<bf16*>(out), o_bs, o_rs, batch,
                                     heads, seq, as_stream(stream), impl == 2);
  }
  AttnArgs a{};
  a.q = base + q_col0; a.k = base + k_col0; a.v = base + v_col0;
  a.q_bs = a.k_bs = a.v_bs = (long long)seq * ld; a.q_rs = a.k_rs = a.v_rs = ld;
  a.o = reinterpret_cast<bf16*>(out); a.o_bs = o_bs; a.o_rs = o_rs;
  a.batch = batch; a.heads = heads; a.sq = seq; a.skv = seq; a.head_dim = head_dim;
  a.scale = 1.0f / sqrtf((float)head_dim);
  return attention_full(a, as_stream(stream));
}

int ec_gemm_bf16_splitk(const void* a, int32_t lda, const void* w, int32_t ldw, int32_t m, int32_t n, int32_t k,
                        const float* bias, const float* residual, int32_t ldr, float* out, int32_t ldo, int32_t split_k,
                        float* ws, int32_t* counters, void* stream) {
  EC_REQUIRE(a && w && residual && out && ws && counters, EC_ERR_INVALID, "ec_gemm_bf16_splitk: null argument");
  GemmArgs g{};
  g.A = reinterpret_cast<const bf16*>(a); g.lda = lda; g.W = reinterpret_cast<const bf16*>(w); g.ldw = ldw;
  g.M = m; g.N = n; g.K = k; g.bias = bias; g.residual = residual; g.ldr = ldr; g.out = out; g.ldo = ldo;
  g.split_k = split_k > 0 ? split_k : 2;
  g.sk_ws = ws; g.sk_cnt = counters;
  return gemm_bf16(g, EPI_F32_RESID, as_stream(stream));
}

int ec_gemm_bf16_grouped(const void* a, int32_t lda, const void* w, int32_t ldw, int32_t m, int32_t n, int32_t k,
                         int32_t group_cols, const float* bias, void* out, int32_t ldo, int32_t epilogue,
                         int32_t use_simt_reference, void* stream) {
  EC_REQUIRE(epilogue == 0 || epilogue == 2, EC_ERR_INVALID, "ec_gemm_bf16_grouped: epilogue %d", epilogue);
  EC_REQUIRE(group_cols > 0, EC_ERR_INVALID, "ec_gemm_bf16_grouped: group_cols %d", group_cols);
  GemmArgs g{};
  g.A = reinterpret_cast<const bf16*>(a); g.lda = lda; g.W = reinterpret_cast<const bf16*>(w); g.ldw = ldw;
  g.M = m; g.N = n; g.K = k; g.bias = bias; g.out = out; g.ldo = ldo; g.group_cols = group_cols;
  if (use_simt_reference) return gemm_reference_simt(g, (GemmEpilogue)epilogue, as_stream(stream));
  return gemm_bf16(g, (GemmEpilogue)epilogue, as_stream(stream));
}

size_t ec_packed_image_tokens_bytes(int32_t n, int32_t n_kv, int32_t width) {
  return packed_image_tokens_elems(n, n_kv, width) * sizeof(bf16);
}

int ec_pack_image_tokens(const float* x, void* out, int32_t n, int32_t n_kv, int32_t width, int32_t normalize,
                         void* stream) {
  EC_REQUIRE(x && out, EC_ERR_INVALID, "ec_pack_image_tokens: null argument");
  return pack_image_tokens(x, reinterpret_cast<bf16*>(out), n, n_kv, width, normalize != 0, as_stream(stream));
}

int ec_cross_attention_absorbed(const void* qp, const void* xn, void* ctx, int32_t batch, int32_t heads, int32_t width,
                                int32_t n_kv, const uint8_t* finished, int32_t reverse, void* stream) {
  EC_REQUIRE(qp && xn && ctx, EC_ERR_INVALID, "ec_cross_attention_absorbed: null argument");
  return decode_cross_absorbed(reinterpret_cast<const bf16*>(qp), reinterpret_cast<const bf16*>(xn),
                               reinterpret_cast<bf16*>(ctx), batch, heads, width, n_kv, finished, nullptr, reverse != 0,
                               as_stream(stream));
}

int64_t ec_launch_count(void) { return (int64_t)launch_count(); }

int ec_profile_gemm(int32_t enable) {
  gemm_profile_enable(enable != 0);
  return EC_OK;
}

int ec_profile_gemm_read(double* ms, double* flops, int64_t* launches) {
  long long l = 0;
  EC_TRY(gemm_profile_read(ms, flops, &l));
  if (launches) *launches = l;
  return EC_OK;
}

int ec_gemm_timing_read(int64_t* out_host, int32_t count) {
  return gemm_timing_read(reinterpret_cast<long long*>(out_host), count);
}

int ec_seq_timing_read(int64_t* out_host, int32_t count) {
  return gemm_seq_timing_read(reinterpret_cast<long long*>(out_host), count);
}

}  // extern "C"
