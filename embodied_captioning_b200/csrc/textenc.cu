// Caption-text encoder on the device: the MiniLM-class BERT sentence encoder the reference uses
// to embed captions before the disagreement (`SentenceTransformer("all-MiniLM-L6-v2")`,
// agents/goal_exploration/goal_exploration.py:57,102-103; detector/pseudolabeler.py:568,651,677): WordPiece ids in,
// L2-normalised mean-pooled embedding out, so caption -> embedding -> disagreement never leaves the GPU.
//
// Architecture = HF `BertModel` (post-LN): embeddings (word + position + token type 0) -> LayerNorm(eps) ->
// L x { self-attention (heads x 32) -> dense + residual -> LayerNorm -> dense + GELU(erf) -> dense + residual ->
// LayerNorm } -> masked mean pooling -> L2 normalise (sentence-transformers Pooling + Normalize modules).
// GEMMs run on the tcgen05 kernels (bf16 operands, fp32 accumulate), the residual stream and every LayerNorm in fp32.
#include "model.cuh"
#include <cstring>
#include <string>

namespace ec {

struct TextLayer {
  bf16 *w_qkv = nullptr, *w_o = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;
  float *b_qkv = nullptr, *b_o = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};

struct TextEncoder {
  ec_textenc_config cfg;
  int device;
  bool finalized = false;
  float *word_emb = nullptr, *pos_emb = nullptr, *type_emb = nullptr, *emb_ln_g = nullptr, *emb_ln_b = nullptr;
  std::vector<TextLayer> layers;
  std::vector<void*> owned;
  ~TextEncoder() {
    cudaSetDevice(device);
    for (void* p : owned) cudaFree(p);
  }
};

static inline uint16_t te_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static int te_vec(TextEncoder& t, float** dst, const float* h, int64_t numel, int64_t expect, const char* name) {
  EC_REQUIRE(numel == expect, EC_ERR_INVALID, "tensor %s: expected %lld elements, got %lld", name, (long long)expect,
             (long long)numel);
  if (*dst == nullptr) {
    EC_CHECK_CUDA(cudaMalloc(dst, (size_t)std::max<int64_t>(expect, 4) * sizeof(float)));
    t.owned.push_back(*dst);
  }
  EC_CHECK_CUDA(cudaMemcpy(*dst, h, (size_t)expect * sizeof(float), cudaMemcpyHostToDevice));
  return EC_OK;
}

// host [rows, cols] fp32 -> device bf16 rows [row_off, row_off + rows) of a [total_rows, cols] matrix
static int te_mat(TextEncoder& t, bf16** dst, const float* h, int64_t numel, int rows, int cols, int total_rows,
                  int row_off, const char* name) {
  EC_REQUIRE(numel == (int64_t)rows * cols, EC_ERR_INVALID, "tensor %s: expected %lld elements, got %lld", name,
             (long long)rows * cols, (long long)numel);
  if (*dst == nullptr) {
    EC_CHECK_CUDA(cudaMalloc(dst, (size_t)total_rows * cols * sizeof(bf16)));
    t.owned.push_back(*dst);
  }
  std::vector<uint16_t> tmp((size_t)rows * cols);
  for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = te_bf16_bits(h[i]);
  EC_CHECK_CUDA(cudaMemcpy(*dst + (size_t)row_off * cols, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice));
  return EC_OK;
}

static int te_set_tensor(TextEncoder& t, const char* name, const float* h, int64_t numel) {
  EC_CHECK_CUDA(cudaSetDevice(t.device));
  const int H = t.cfg.hidden, I = t.cfg.intermediate;
  t.finalized = false;
  const char* n = name;
  if (!strncmp(n, "bert.", 5)) n += 5;                  // BertForX checkpoints; sentence-transformers saves bare names
  if (!strncmp(n, "0.auto_model.", 13)) n += 13;        // sentence-transformers module prefix
  if (!strcmp(n, "embeddings.word_embeddings.weight")) return te_vec(t, &t.word_emb, h, numel, (int64_t)t.cfg.vocab_size * H, name);
  if (!strcmp(n, "embeddings.position_embeddings.weight")) return te_vec(t, &t.pos_emb, h, numel, (int64_t)t.cfg.max_positions * H, name);
  if (!strcmp(n, "embeddings.token_type_embeddings.weight")) return te_vec(t, &t.type_emb, h, numel, (int64_t)t.cfg.type_vocab * H, name);
  if (!strcmp(n, "embeddings.LayerNorm.weight")) return te_vec(t, &t.emb_ln_g, h, numel, H, name);
  if (!strcmp(n, "embeddings.LayerNorm.bias")) return te_vec(t, &t.emb_ln_b, h, numel, H, name);
  if (!strcmp(n, "embeddings.position_ids") || !strncmp(n, "pooler.", 7)) return EC_OK;   // unused by the sentence encoder
  int l = -1, used = 0;
  if (sscanf(n, "encoder.layer.%d.%n", &l, &used) == 1 && used > 0) {
    EC_REQUIRE(l >= 0 && l < t.cfg.layers, EC_ERR_INVALID, "layer index out of range in %s", name);
    TextLayer& L = t.layers[l];
    const char* f = n + used;
    if (!strcmp(f, "attention.self.query.weight")) return te_mat(t, &L.w_qkv, h, numel, H, H, 3 * H, 0, name);
    if (!strcmp(f, "attention.self.key.weight")) return te_mat(t, &L.w_qkv, h, numel, H, H, 3 * H, H, name);
    if (!strcmp(f, "attention.self.value.weight")) return te_mat(t, &L.w_qkv, h, numel, H, H, 3 * H, 2 * H, name);
    if (!strcmp(f, "attention.self.query.bias") || !strcmp(f, "attention.self.key.bias") ||
        !strcmp(f, "attention.self.value.bias")) {
      EC_REQUIRE(numel == H, EC_ERR_INVALID, "tensor %s: expected %d elements", name, H);
      if (L.b_qkv == nullptr) {
        EC_CHECK_CUDA(cudaMalloc(&L.b_qkv, (size_t)3 * H * sizeof(float)));
        t.owned.push_back(L.b_qkv);
      }
      const int off = strstr(f, "query") ? 0 : (strstr(f, "key") ? H : 2 * H);
      EC_CHECK_CUDA(cudaMemcpy(L.b_qkv + off, h, (size_t)H * sizeof(float), cudaMemcpyHostToDevice));
      return EC_OK;
    }
    if (!strcmp(f, "attention.output.dense.weight")) return te_mat(t, &L.w_o, h, numel, H, H, H, 0, name);
    if (!strcmp(f, "attention.output.dense.bias")) return te_vec(t, &L.b_o, h, numel, H, name);
    if (!strcmp(f, "attention.output.LayerNorm.weight")) return te_vec(t, &L.ln1_g, h, numel, H, name);
    if (!strcmp(f, "attention.output.LayerNorm.bias")) return te_vec(t, &L.ln1_b, h, numel, H, name);
    if (!strcmp(f, "intermediate.dense.weight")) return te_mat(t, &L.w_fc1, h, numel, I, H, I, 0, name);
    if (!strcmp(f, "intermediate.dense.bias")) return te_vec(t, &L.b_fc1, h, numel, I, name);
    if (!strcmp(f, "output.dense.weight")) return te_mat(t, &L.w_fc2, h, numel, H, I, H, 0, name);
    if (!strcmp(f, "output.dense.bias")) return te_vec(t, &L.b_fc2, h, numel, H, name);
    if (!strcmp(f, "output.LayerNorm.weight")) return te_vec(t, &L.ln2_g, h, numel, H, name);
    if (!strcmp(f, "output.LayerNorm.bias")) return te_vec(t, &L.ln2_b, h, numel, H, name);
  }
  set_error("unknown text-encoder tensor %s", name);
  return EC_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------
constexpr int kTeMaxVec = 8;   // hidden <= 1024

// one warp per token row: v = LayerNorm(x) with fp32 statistics; writes fp32 (next residual) and bf16 (next GEMM operand)
__device__ __forceinline__ void te_ln_row(float4 (&v)[kTeMaxVec], int nvec, int lane, int width, float eps,
                                          const float* __restrict__ g, const float* __restrict__ b, float* yf, bf16* yb) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kTeMaxVec; ++i) if (lane + i * 32 < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)width;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kTeMaxVec; ++i) {
    if (lane + i * 32 < nvec) {
      const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + c * c) + (d * d + e * e);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)width + eps);
#pragma unroll
  for (int i = 0; i < kTeMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + c), bb = __ldg(reinterpret_cast<const float4*>(b) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gg.x + bb.x; o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
      o.z = (v[i].z - mean) * rstd * gg.z + bb.z; o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
      reinterpret_cast<float4*>(yf)[c] = o;
      uint2 pk; pk.x = pack_bf16x2(o.x, o.y); pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(yb)[c] = pk;
    }
  }
}

__global__ void __launch_bounds__(256) te_embed_ln_kernel(const int* __restrict__ ids, const float* __restrict__ word,
                                                          const float* __restrict__ pos, const float* __restrict__ type0,
                                                          const float* __restrict__ g, const float* __restrict__ b,
                                                          float* __restrict__ hf, bf16* __restrict__ hb, int rows,
                                                          int seq, int width, int vocab, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = width >> 2;
  int id = ids[row];
  id = min(max(id, 0), vocab - 1);
  const float4* w = reinterpret_cast<const float4*>(word + (size_t)id * width);
  const float4* p = reinterpret_cast<const float4*>(pos + (size_t)(row % seq) * width);
  const float4* t = reinterpret_cast<const float4*>(type0);
  float4 v[kTeMaxVec];
#pragma unroll
  for (int i = 0; i < kTeMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 a = __ldg(w + c), q = __ldg(p + c), r = __ldg(t + c);
      v[i] = make_float4((a.x + r.x) + q.x, (a.y + r.y) + q.y, (a.z + r.z) + q.z, (a.w + r.w) + q.w);   // HF order
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  te_ln_row(v, nvec, lane, width, eps, g, b, hf + (size_t)row * width, hb + (size_t)row * width);
}

__global__ void __launch_bounds__(256) te_ln_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                    const float* __restrict__ b, float* __restrict__ hf,
                                                    bf16* __restrict__ hb, int rows, int width, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = width >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * width);
  float4 v[kTeMaxVec];
#pragma unroll
  for (int i = 0; i < kTeMaxVec; ++i) {
    const int c = lane + i * 32;
    v[i] = (c < nvec) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  te_ln_row(v, nvec, lane, width, eps, g, b, hf + (size_t)row * width, hb + (size_t)row * width);
}

// masked mean over the first len[b] tokens (clamp 1e-9 like sentence-transformers Pooling), then F.normalize
__global__ void __launch_bounds__(128) te_pool_kernel(const float* __restrict__ h, const int* __restrict__ lens,
                                                      float* __restrict__ out, int seq, int width) {
  __shared__ float s_red[4];
  const int b = blockIdx.x;
  const int n = max(0, min(lens[b], seq));
  const float denom = fmaxf((float)n, 1e-9f);
  float acc[8];
  float ss = 0.f;
  int k = 0;
  for (int c = threadIdx.x; c < width; c += 128, ++k) {
    float a = 0.f;
    for (int t = 0; t < n; ++t) a += h[((size_t)b * seq + t) * width + c];
    a /= denom;
    acc[k] = a;
    ss += a * a;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float tot = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
  const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
  k = 0;
  for (int c = threadIdx.x; c < width; c += 128, ++k) out[(size_t)b * width + c] = acc[k] * inv;
}

struct TeWs { float* hf; bf16* hb; bf16* qkv; bf16* att; float* t; bf16* fc1; size_t bytes; };

static TeWs te_carve(const TextEncoder& e, void* ws, int n, int seq) {
  const size_t M = (size_t)n * seq, H = e.cfg.hidden, I = e.cfg.intermediate;
  unsigned char* b = reinterpret_cast<unsigned char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { unsigned char* p = b + off; off += align_up(bytes, 256); return p; };
  TeWs w;
  w.hf = reinterpret_cast<float*>(take(M * H * 4));
  w.hb = reinterpret_cast<bf16*>(take(M * H * 2));
  w.qkv = reinterpret_cast<bf16*>(take(M * 3 * H * 2));
  w.att = reinterpret_cast<bf16*>(take(M * H * 2));
  w.t = reinterpret_cast<float*>(take(M * H * 4));
  w.fc1 = reinterpret_cast<bf16*>(take(M * I * 2));
  w.bytes = off;
  return w;
}

static int te_linear(const bf16* A, int lda, const bf16* W, const float* bias, int M, int N, int K, void* out, int ldo,
                     GemmEpilogue epi, const float* resid, cudaStream_t s) {
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = K; g.M = M; g.N = N; g.K = K; g.bias = bias;
  g.residual = resid; g.ldr = N; g.out = out; g.ldo = ldo;
  return gemm_bf16(g, epi, s);
}

static int te_encode(TextEncoder& e, const int* ids, const int* lens, int n, int seq, float* out, void* ws,
                     size_t ws_bytes, cudaStream_t s) {
  EC_REQUIRE(e.finalized, EC_ERR_STATE, "text encoder not finalized");
  EC_REQUIRE(n > 0 && seq > 0 && seq <= e.cfg.max_positions, EC_ERR_INVALID, "text encoder: n = %d, seq = %d (max %d)", n,
             seq, e.cfg.max_positions);
  EC_CHECK_CUDA(cudaSetDevice(e.device));
  const TeWs w = te_carve(e, ws, n, seq);
  EC_REQUIRE(ws_bytes >= w.bytes, EC_ERR_WORKSPACE, "text encoder workspace %zu < %zu", ws_bytes, w.bytes);
  const int H = e.cfg.hidden, I = e.cfg.intermediate, M = n * seq, heads = e.cfg.heads, hd = H / heads;
  const float eps = e.cfg.layer_norm_eps;
  const int ln_grid = ceil_div(M, 8);
  te_embed_ln_kernel<<<ln_grid, 256, 0, s>>>(ids, e.word_emb, e.pos_emb, e.type_emb, e.emb_ln_g, e.emb_ln_b, w.hf, w.hb,
                                             M, seq, H, e.cfg.vocab_size, eps);
  EC_POST_LAUNCH();
  for (int l = 0; l < e.cfg.layers; ++l) {
    const TextLayer& L = e.layers[l];
    EC_TRY(te_linear(w.hb, H, L.w_qkv, L.b_qkv, M, 3 * H, H, w.qkv, 3 * H, EPI_BF16, nullptr, s));
    AttnArgs a{};
    a.q = w.qkv; a.k = w.qkv + H; a.v = w.qkv + 2 * H;
    a.q_bs = a.k_bs = a.v_bs = (long long)seq * 3 * H; a.q_rs = a.k_rs = a.v_rs = 3 * H;
    a.o = w.att; a.o_bs = (long long)seq * H; a.o_rs = H;
    a.batch = n; a.heads = heads; a.sq = seq; a.skv = seq; a.head_dim = hd;
    a.scale = 1.0f / sqrtf((float)hd);
    a.kv_len = lens;
    EC_TRY(attention_full(a, s));
    EC_TRY(te_linear(w.att, H, L.w_o, L.b_o, M, H, H, w.t, H, EPI_F32_RESID, w.hf, s));           // h + dense(attn)
    te_ln_kernel<<<ln_grid, 256, 0, s>>>(w.t, L.ln1_g, L.ln1_b, w.hf, w.hb, M, H, eps);
    EC_POST_LAUNCH();
    EC_TRY(te_linear(w.hb, H, L.w_fc1, L.b_fc1, M, I, H, w.fc1, I, EPI_BF16_GELU, nullptr, s));
    EC_TRY(te_linear(w.fc1, I, L.w_fc2, L.b_fc2, M, H, I, w.t, H, EPI_F32_RESID, w.hf, s));       // h + dense(gelu(...))
    te_ln_kernel<<<ln_grid, 256, 0, s>>>(w.t, L.ln2_g, L.ln2_b, w.hf, w.hb, M, H, eps);
    EC_POST_LAUNCH();
  }
  te_pool_kernel<<<n, 128, 0, s>>>(w.hf, lens, out, seq, H);
  EC_POST_LAUNCH();
  return EC_OK;
}

}  // namespace ec

struct ec_textenc {
  ec::TextEncoder impl;
};

using namespace ec;

extern "C" {

int ec_textenc_create(const ec_textenc_config* c, int device, ec_textenc** out) {
  EC_REQUIRE(c != nullptr && out != nullptr, EC_ERR_INVALID, "ec_textenc_create: null argument");
  *out = nullptr;
  int count = 0, major = 0;
  EC_CHECK_CUDA(cudaGetDeviceCount(&count));
  EC_REQUIRE(device >= 0 && device < count, EC_ERR_DEVICE, "device %d out of range (%d visible)", device, count);
  EC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  EC_REQUIRE(major == 10, EC_ERR_DEVICE, "libec_b200 is built for sm_100a only (device %d: %d.x)", device, major);
  EC_REQUIRE(c->hidden % c->heads == 0, EC_ERR_INVALID, "hidden not divisible by heads");
  const int hd = c->hidden / c->heads;
  EC_REQUIRE(hd == 32 || hd == 64 || hd == 96, EC_ERR_UNSUPPORTED, "text encoder head_dim %d unsupported", hd);
  EC_REQUIRE(c->hidden % 8 == 0 && c->hidden <= 1024 && c->intermediate % 8 == 0, EC_ERR_UNSUPPORTED,
             "text encoder widths must be multiples of 8, hidden <= 1024");
  EC_REQUIRE(c->layers > 0 && c->vocab_size > 0 && c->max_positions > 0 && c->type_vocab > 0, EC_ERR_INVALID,
             "text encoder: bad sizes");
  ec_textenc* t = new (std::nothrow) ec_textenc();
  EC_REQUIRE(t != nullptr, EC_ERR_INVALID, "out of host memory");
  t->impl.cfg = *c;
  t->impl.device = device;
  t->impl.layers.resize(c->layers);
  *out = t;
  return EC_OK;
}

int ec_textenc_destroy(ec_textenc* t) {
  delete t;
  return EC_OK;
}

int ec_textenc_set_tensor(ec_textenc* t, const char* name, const float* host_data, int64_t numel) {
  EC_REQUIRE(t && name && host_data, EC_ERR_INVALID, "ec_textenc_set_tensor: null argument");
  return te_set_tensor(t->impl, name, host_data, numel);
}

int ec_textenc_finalize(ec_textenc* t) {
  EC_REQUIRE(t, EC_ERR_INVALID, "null text encoder");
  TextEncoder& e = t->impl;
  EC_REQUIRE(e.word_emb && e.pos_emb && e.type_emb && e.emb_ln_g && e.emb_ln_b, EC_ERR_STATE,
             "text encoder embeddings incomplete");
  for (size_t i = 0; i < e.layers.size(); ++i) {
    const TextLayer& L = e.layers[i];
    EC_REQUIRE(L.w_qkv && L.b_qkv && L.w_o && L.b_o && L.w_fc1 && L.b_fc1 && L.w_fc2 && L.b_fc2 && L.ln1_g && L.ln1_b &&
                   L.ln2_g && L.ln2_b,
               EC_ERR_STATE, "text encoder layer %zu incomplete", i);
  }
  e.finalized = true;
  return EC_OK;
}

size_t ec_textenc_workspace_bytes(const ec_textenc* t, int32_t n, int32_t seq) {
  return t ? te_carve(t->impl, nullptr, std::max(n, 1), std::max(seq, 1)).bytes : 0;
}

int ec_textenc_encode(ec_textenc* t, const int32_t* ids, const int32_t* lens, int32_t n, int32_t seq, float* out,
                      void* ws, size_t ws_bytes, void* stream) {
  EC_REQUIRE(t && ids && lens && out && ws, EC_ERR_INVALID, "ec_textenc_encode: null argument");
  return te_encode(t->impl, ids, lens, n, seq, out, ws, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
