// KV-cached greedy decoder step kernels (one new token per crop per step; memory-bound).
// All kernels read the step index from a device-resident DecodeState so one captured CUDA graph serves
// every step, and early-out when every crop has finished (no per-token host synchronisation, unlike the
// reference loop at coca_model.py:294-327 which syncs on `mask.all()` each step).
#include "kernels.cuh"

namespace ec {

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_embed_kernel(const int* __restrict__ tokens, int ld_tokens,
                                                           const float* __restrict__ tok_emb,
                                                           const float* __restrict__ pos_emb, float* __restrict__ x,
                                                           int batch, int width, const DecodeState* __restrict__ st) {
  if (st->done) return;
  const int pos = st->cur_len - 1;
  const int nvec = width >> 2;
  const long long total = (long long)batch * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / nvec), c = (int)(i % nvec);
    const int tok = tokens[(size_t)b * ld_tokens + pos];
    const float4 e = __ldg(reinterpret_cast<const float4*>(tok_emb + (size_t)tok * width) + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos_emb + (size_t)pos * width) + c);
    reinterpret_cast<float4*>(x + (size_t)b * width)[c] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
}

int decode_embed(const int* tokens, int ld_tokens, const float* tok_emb, const float* pos_emb, float* x, int batch,
                 int width, const DecodeState* st, cudaStream_t stream) {
  EC_REQUIRE(width % 4 == 0, EC_ERR_UNSUPPORTED, "decode_embed: width %% 4 != 0");
  const long long total = (long long)batch * (width / 4);
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)num_sms() * 8);
  decode_embed_kernel<<<grid, 256, 0, stream>>>(tokens, ld_tokens, tok_emb, pos_emb, x, batch, width, st);
  EC_POST_LAUNCH();
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Self-attention of the new token: one warp per (crop, head).  head_dim = 64: lane owns dims 2*lane, 2*lane+1.
constexpr int kMaxCtx = 128;   // >= context_length + 1

template <int D>
__global__ void __launch_bounds__(128) decode_self_attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kcache,
                                                               bf16* __restrict__ vcache, bf16* __restrict__ out,
                                                               int batch, int heads, int max_len,
                                                               const unsigned char* __restrict__ finished,
                                                               const DecodeState* __restrict__ st) {
  if (st->done) return;
  __shared__ float s_q[4][D];
  __shared__ float s_p[4][kMaxCtx];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + w;
  if (pair >= batch * heads) return;
  const int b = pair / heads, h = pair % heads;
  if (finished[b]) return;
  const int pos = st->cur_len - 1;           // position of the new token; attends to 0..pos
  const int C = heads * D;
  const bf16* row = qkv + (size_t)b * 3 * C + h * D;
  bf16* kc = kcache + ((size_t)b * max_len) * C + h * D;
  bf16* vc = vcache + ((size_t)b * max_len) * C + h * D;
  // append k, v of the new token; stash q (pre-scaled) in shared memory
  const float scale = rsqrtf((float)D);
  for (int d = lane; d < D; d += 32) {
    s_q[w][d] = __bfloat162float(row[d]) * scale;
    kc[(size_t)pos * C + d] = row[C + d];
    vc[(size_t)pos * C + d] = row[2 * C + d];
  }
  __syncwarp();
  // scores: lane j handles positions j, j+32, ...
  float mx = -INFINITY;
  for (int j = lane; j <= pos; j += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(kc + (size_t)j * C);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < D / 8; ++c) {
      const uint4 u = kr[c];
      const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
      const float* q = &s_q[w][c * 8];
      acc += a0.x * q[0] + a0.y * q[1] + a1.x * q[2] + a1.y * q[3] + a2.x * q[4] + a2.y * q[5] + a3.x * q[6] + a3.y * q[7];
    }
    s_p[w][j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j <= pos; j += 32) {
    const float e = __expf(s_p[w][j] - mx);
    s_p[w][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  // out[d] = sum_j p_j V[j][d]; each lane owns D/32 consecutive dims
  constexpr int DPL = D / 32;
  float acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  for (int j = 0; j <= pos; ++j) {
    const float p = s_p[w][j];
    const bf16* vr = vc + (size_t)j * C + lane * DPL;
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] += p * __bfloat162float(vr[i]);
  }
  bf16* o = out + (size_t)b * C + h * D + lane * DPL;
#pragma unroll
  for (int i = 0; i < DPL; ++i) o[i] = __float2bfloat16_rn(acc[i] * inv);
}

int decode_self_attention(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, int batch, int heads, int head_dim,
                          int max_len, const unsigned char* finished, const DecodeState* st, cudaStream_t stream) {
  EC_REQUIRE(max_len <= kMaxCtx, EC_ERR_UNSUPPORTED, "decode: context %d > %d", max_len, kMaxCtx);
  EC_REQUIRE(head_dim == 64, EC_ERR_UNSUPPORTED, "decode self-attention: head_dim %d != 64", head_dim);
  const int grid = ceil_div(batch * heads, 4);
  decode_self_attn_kernel<64><<<grid, 128, 0, stream>>>(qkv, kcache, vcache, out, batch, heads, max_len, finished, st);
  EC_POST_LAUNCH();
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Cross-attention over the image K/V cache: one warp per (crop, head).  The cache row of key j is
// [K(C) | V(C)] bf16; head h owns a contiguous 128-byte segment of each half.  8 lanes cover one key row
// (16 B each), 4 rows per warp instruction, 4 independent instructions in flight per lane.
template <int D>
__global__ void __launch_bounds__(128) decode_cross_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kv,
                                                                bf16* __restrict__ out, int batch, int heads, int n_kv,
                                                                const unsigned char* __restrict__ finished,
                                                                const DecodeState* __restrict__ st) {
  static_assert(D == 64, "lane mapping assumes 64-dim heads");
  if (st->done) return;
  __shared__ float s_p[4][260];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + w;
  if (pair >= batch * heads) return;
  const int b = pair / heads, h = pair % heads;
  if (finished[b]) return;
  const int C = heads * D;
  const int sub = lane & 7, grp = lane >> 3;           // 16-byte slice of the head, row within a group of 4
  float qf[8];
  {
    const uint4 u = *reinterpret_cast<const uint4*>(q + (size_t)b * C + h * D + sub * 8);
    const float sc = rsqrtf((float)D);
    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
    qf[0] = a0.x * sc; qf[1] = a0.y * sc; qf[2] = a1.x * sc; qf[3] = a1.y * sc;
    qf[4] = a2.x * sc; qf[5] = a2.y * sc; qf[6] = a3.x * sc; qf[7] = a3.y * sc;
  }
  const bf16* kbase = kv + (size_t)b * n_kv * 2 * C + h * D + sub * 8;
  const size_t rstride = (size_t)2 * C;
  // ---- pass 1: scores
  float mx = -INFINITY;
  for (int j0 = 0; j0 < n_kv; j0 += 16) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i * 4 + grp;
      u[i] = (j < n_kv) ? __ldg(reinterpret_cast<const uint4*>(kbase + (size_t)j * rstride)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a0 = unpack_bf16x2(u[i].x), a1 = unpack_bf16x2(u[i].y), a2 = unpack_bf16x2(u[i].z), a3 = unpack_bf16x2(u[i].w);
      float acc = a0.x * qf[0] + a0.y * qf[1] + a1.x * qf[2] + a1.y * qf[3] + a2.x * qf[4] + a2.y * qf[5] +
                  a3.x * qf[6] + a3.y * qf[7];
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      const int j = j0 + i * 4 + grp;
      if (j < n_kv) {
        if (sub == 0) s_p[w][j] = acc;
        mx = fmaxf(mx, acc);
      }
    }
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < n_kv; j += 32) {
    const float e = __expf(s_p[w][j] - mx);
    s_p[w][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  // ---- pass 2: weighted sum of V
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const bf16* vbase = kbase + C;
  for (int j0 = 0; j0 < n_kv; j0 += 16) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i * 4 + grp;
      u[i] = (j < n_kv) ? __ldg(reinterpret_cast<const uint4*>(vbase + (size_t)j * rstride)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + i * 4 + grp;
      const float p = (j < n_kv) ? s_p[w][j] : 0.f;
      const float2 a0 = unpack_bf16x2(u[i].x), a1 = unpack_bf16x2(u[i].y), a2 = unpack_bf16x2(u[i].z), a3 = unpack_bf16x2(u[i].w);
      acc[0] += p * a0.x; acc[1] += p * a0.y; acc[2] += p * a1.x; acc[3] += p * a1.y;
      acc[4] += p * a2.x; acc[5] += p * a2.y; acc[6] += p * a3.x; acc[7] += p * a3.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (grp == 0) {
    uint4 o;
    o.x = pack_bf16x2(acc[0] * inv, acc[1] * inv); o.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
    o.z = pack_bf16x2(acc[4] * inv, acc[5] * inv); o.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(out + (size_t)b * C + h * D + sub * 8) = o;
  }
}

int decode_cross_attention(const bf16* q, const bf16* kv, bf16* out, int batch, int heads, int head_dim, int n_kv,
                           const unsigned char* finished, const DecodeState* st, cudaStream_t stream) {
  EC_REQUIRE(head_dim == 64, EC_ERR_UNSUPPORTED, "decode cross-attention: head_dim %d != 64", head_dim);
  EC_REQUIRE(n_kv <= 260, EC_ERR_UNSUPPORTED, "decode cross-attention: %d keys > 260", n_kv);
  const int grid = ceil_div(batch * heads, 4);
  decode_cross_attn_kernel<64><<<grid, 128, 0, stream>>>(q, kv, out, batch, heads, n_kv, finished, st);
  EC_POST_LAUNCH();
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Token selection: arg-max over the vocab-GEMM partials + generate() bookkeeping (coca_model.py:304-322).
__global__ void __launch_bounds__(256) decode_select_kernel(const float* __restrict__ part_val,
                                                            const int* __restrict__ part_idx, int n_parts,
                                                            int* __restrict__ tokens, int ld_tokens,
                                                            unsigned char* __restrict__ finished, int batch,
                                                            int seq_len, int eos, int pad, DecodeState* st) {
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  const int cur_len = st->cur_len;
  if (finished[b]) {                       // finished rows are padded (coca_model.py:305)
    if (lane == 0) tokens[(size_t)b * ld_tokens + cur_len] = pad;
    return;
  }
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int p = lane; p < n_parts; p += 32) {
    const float v = part_val[(size_t)b * n_parts + p];
    const int i = part_idx[(size_t)b * n_parts + p];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) {
    int tok = (cur_len + 1 == seq_len) ? eos : bi;    // forced EOS at the last slot (coca_model.py:317-318)
    tokens[(size_t)b * ld_tokens + cur_len] = tok;
    if (tok == eos || tok == pad) finished[b] = 1;    // mask of the next iteration (coca_model.py:304)
    else atomicAdd(&st->next_active, 1);
  }
}

int decode_select(const float* part_val, const int* part_idx, int n_parts, int* tokens, int ld_tokens,
                  unsigned char* finished, int batch, int seq_len, int eos, int pad, DecodeState* st,
                  cudaStream_t stream) {
  decode_select_kernel<<<ceil_div(batch, 8), 256, 0, stream>>>(part_val, part_idx, n_parts, tokens, ld_tokens, finished,
                                                              batch, seq_len, eos, pad, st);
  EC_POST_LAUNCH();
  return EC_OK;
}

__global__ void decode_advance_kernel(DecodeState* st, int teacher_forced, int batch) {
  if (st->done) return;
  st->cur_len += 1;
  const int act = teacher_forced ? batch : st->next_active;   // teacher forcing never finishes a row
  st->n_active = act;
  st->done = (act == 0) ? 1 : 0;
  st->next_active = 0;
}

int decode_advance(DecodeState* st, bool teacher_forced, int batch, cudaStream_t stream) {
  decode_advance_kernel<<<1, 1, 0, stream>>>(st, teacher_forced ? 1 : 0, batch);
  EC_POST_LAUNCH();
  return EC_OK;
}

__global__ void __launch_bounds__(256) decode_init_kernel(int* __restrict__ tokens, int ld_tokens,
                                                          unsigned char* __restrict__ finished, int batch, int seq_len,
                                                          int sot, int pad, DecodeState* st) {
  const long long total = (long long)batch * seq_len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / seq_len), c = (int)(i % seq_len);
    tokens[(size_t)b * ld_tokens + c] = (c == 0) ? sot : pad;
    if (c == 0) finished[b] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->cur_len = 1;
    st->n_active = batch;
    st->done = 0;
    st->next_active = 0;
  }
}

int decode_init(int* tokens, int ld_tokens, unsigned char* finished, int batch, int seq_len, int sot, int pad,
                DecodeState* st, cudaStream_t stream) {
  const long long total = (long long)batch * seq_len;
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), 1024);
  decode_init_kernel<<<grid, 256, 0, stream>>>(tokens, ld_tokens, finished, batch, seq_len, sot, pad, st);
  EC_POST_LAUNCH();
  return EC_OK;
}

}  // namespace ec
