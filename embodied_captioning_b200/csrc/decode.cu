// KV-cached greedy decoder step kernels (one new token per crop per step; memory-bound).
// All kernels read the step index from a device-resident DecodeState so one captured CUDA graph serves
// every step, and early-out when every crop has finished (no per-token host synchronisation, unlike the
// reference loop at coca_model.py:294-327 which syncs on `mask.all()` each step).  Every kernel is launched
// with programmatic dependent launch so launch latency and CTA scheduling overlap the previous kernel's tail.
#include "kernels.cuh"
#include <cstdlib>

#ifndef EC_SA_STREAM
#define EC_SA_STREAM 1
#endif

namespace ec {

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_embed_kernel(const int* __restrict__ tokens, int ld_tokens,
                                                           const float* __restrict__ tok_emb,
                                                           const float* __restrict__ pos_emb, float* __restrict__ x,
                                                           int batch, int width, const DecodeState* __restrict__ st) {
  pdl_trigger_top();
  pdl_wait();
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
  pdl_trigger_bottom();
}

int decode_embed(const int* tokens, int ld_tokens, const float* tok_emb, const float* pos_emb, float* x, int batch,
                 int width, const DecodeState* st, cudaStream_t stream) {
  EC_REQUIRE(width % 4 == 0, EC_ERR_UNSUPPORTED, "decode_embed: width %% 4 != 0");
  const long long total = (long long)batch * (width / 4);
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)num_sms() * 8);
  note_launches(1);
  EC_CHECK_CUDA(launch_small_kernel(decode_embed_kernel, dim3(grid), dim3(256), 0, stream, tokens, ld_tokens, tok_emb,
                              pos_emb, x, batch, width, st));
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Single-query attention over `n` cached keys for one (crop, head), executed by one warp.
// Key j lives at kbase + j*rstride (64 bf16 = one 128-byte line), value at vbase + j*rstride.  8 lanes cover a
// row (16 B each), 4 rows per warp instruction, 4 independent instructions (16 rows) in flight per lane.
constexpr int kMaxKeys = 260;   // >= n_queries-1 and >= context_length+1

__device__ __forceinline__ void attend_load(const bf16* kbase, const bf16* vbase, size_t rstride, int n, int j0, int grp,
                                            uint4 (&ku)[4], uint4 (&vu)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + i * 4 + grp;
    if (j < n) {
      ku[i] = *reinterpret_cast<const uint4*>(kbase + (size_t)j * rstride);
      vu[i] = *reinterpret_cast<const uint4*>(vbase + (size_t)j * rstride);
    } else {
      ku[i] = make_uint4(0, 0, 0, 0);
      vu[i] = make_uint4(0, 0, 0, 0);
    }
  }
}

__device__ __forceinline__ void warp_attend(const float (&qf)[8], const bf16* kbase, const bf16* vbase, size_t rstride,
                                            int n, int lane, float (&acc)[8], float& inv) {
  // Single pass with an online softmax.  K and V rows of a 16-key batch are requested together and the next
  // batch is prefetched into registers before the current one is reduced (software pipelining), so one HBM
  // round trip is exposed per kernel instead of one per batch.  All 32 lanes agree on the running max.
  const int grp = lane >> 3;
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  uint4 ku[4], vu[4], kn[4], vn[4];
  attend_load(kbase, vbase, rstride, n, 0, grp, ku, vu);
  for (int j0 = 0; j0 < n; j0 += 16) {
    attend_load(kbase, vbase, rstride, n, j0 + 16, grp, kn, vn);     // prefetch (all-masked past the end)
    float sc[4];
    float bm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a0 = unpack_bf16x2(ku[i].x), a1 = unpack_bf16x2(ku[i].y), a2 = unpack_bf16x2(ku[i].z), a3 = unpack_bf16x2(ku[i].w);
      float d = a0.x * qf[0] + a0.y * qf[1] + a1.x * qf[2] + a1.y * qf[3] + a2.x * qf[4] + a2.y * qf[5] +
                a3.x * qf[6] + a3.y * qf[7];
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      sc[i] = (j0 + i * 4 + grp < n) ? d : -INFINITY;
      bm = fmaxf(bm, sc[i]);
    }
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
    const float mn = fmaxf(m, bm);                 // finite: every batch holds at least one valid key
    const float corr = __expf(m - mn);             // exp(-inf) = 0 on the first batch
    m = mn;
    l *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= corr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float p = __expf(sc[i] - mn);          // 0 for masked rows
      l += p;
      const float2 a0 = unpack_bf16x2(vu[i].x), a1 = unpack_bf16x2(vu[i].y), a2 = unpack_bf16x2(vu[i].z), a3 = unpack_bf16x2(vu[i].w);
      acc[0] += p * a0.x; acc[1] += p * a0.y; acc[2] += p * a1.x; acc[3] += p * a1.y;
      acc[4] += p * a2.x; acc[5] += p * a2.y; acc[6] += p * a3.x; acc[7] += p * a3.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { ku[i] = kn[i]; vu[i] = vn[i]; }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
  inv = 1.0f / l;
}

__device__ __forceinline__ void load_q_slice(const bf16* q, float (&qf)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(q);
  const float sc = 0.125f;   // 1/sqrt(64)
  const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
  qf[0] = a0.x * sc; qf[1] = a0.y * sc; qf[2] = a1.x * sc; qf[3] = a1.y * sc;
  qf[4] = a2.x * sc; qf[5] = a2.y * sc; qf[6] = a3.x * sc; qf[7] = a3.y * sc;
}

__device__ __forceinline__ void store_out_slice(bf16* o, const float (&acc)[8], float inv) {
  uint4 v;
  v.x = pack_bf16x2(acc[0] * inv, acc[1] * inv); v.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
  v.z = pack_bf16x2(acc[4] * inv, acc[5] * inv); v.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
  *reinterpret_cast<uint4*>(o) = v;
}

// causal self-attention of the new token: appends its k, v to the cache, then attends to positions 0..pos
__global__ void __launch_bounds__(128) decode_self_attn_kernel(const bf16* __restrict__ qkv, bf16* kcache, bf16* vcache,
                                                               bf16* __restrict__ out, int batch, int heads, int max_len,
                                                               const unsigned char* __restrict__ finished,
                                                               const DecodeState* __restrict__ st) {
  constexpr int D = 64;
  pdl_trigger_top();
  pdl_wait();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + w;
  if (pair >= batch * heads) return;
  const int b = pair / heads, h = pair % heads;
  // the three control words in ONE round trip (as dependent early-outs they were three L2 latencies in a row)
  const int C = heads * D;
  const int sub = lane & 7, grp = lane >> 3;
  const bf16* row = qkv + (size_t)b * 3 * C + h * D + sub * 8;
  const int done = st->done, cur_len = st->cur_len;
  const unsigned char fin = finished[b];
  const uint4 knew = *reinterpret_cast<const uint4*>(row + C), vnew = *reinterpret_cast<const uint4*>(row + 2 * C);
  float qf[8];
  load_q_slice(row, qf);                      // the new token's q, k, v ride in the same round trip
  if (done || fin) return;
  const int pos = cur_len - 1;
  bf16* kc = kcache + ((size_t)b * max_len) * C + h * D + sub * 8;
  bf16* vc = vcache + ((size_t)b * max_len) * C + h * D + sub * 8;
  if (grp == 0) *reinterpret_cast<uint4*>(kc + (size_t)pos * C) = knew;
  if (grp == 1) *reinterpret_cast<uint4*>(vc + (size_t)pos * C) = vnew;
  __syncwarp();     // the appended row is read back by other lanes of this warp below
  float acc[8], inv;
  warp_attend(qf, kc, vc, (size_t)C, pos + 1, lane, acc, inv);
  pdl_trigger_bottom();
  if (grp == 0) store_out_slice(out + (size_t)b * C + h * D + sub * 8, acc, inv);
}


// Same computation with EIGHT LANES per (crop, head) instead of a warp: the 12 288 pairs of a 1024-crop step are 3 072
// warps -- one resident wave on 148 SMs instead of three to five -- and every lane keeps a whole batch of rows in
// flight.  The warp-per-pair kernel above is latency-bound, not bandwidth-bound: 15.6 us per launch under ncu even at
// the steps where it moves 8 MB (3 waves x 2 dependent round trips).  Rows 0..kR-1 of the cache are requested at kernel
// entry, before the step index is known (rows at or past `pos` are masked afterwards: stale but mapped data), together
// with the control words and the new token's q / k / v; the new token enters the softmax from registers.
#ifndef EC_SA_R
#define EC_SA_R 4
#endif
#ifndef EC_SA_THREADS
#define EC_SA_THREADS 128
#endif
#ifndef EC_SA_PREFETCH
#define EC_SA_PREFETCH 0
#endif
constexpr int kSaThreads = EC_SA_THREADS;
constexpr int kSaMinBlocks = (EC_SA_THREADS == 32) ? 21 : (EC_SA_THREADS == 96 ? 7 : 6);
constexpr int kSaR = EC_SA_R;      // rows per batch and lane group (two batches in flight)

// One-warp CTAs (4 pairs each): 12 288 pairs x 8 lanes = 664 threads per SM at <= 96 registers -- the whole grid is one
// resident wave (larger CTAs round the per-SM count down and leave a tail).  Batches of kSaR rows, the next batch is
// requested before the current one is reduced.  Measured in the 1024-crop decode (profiles/r02_decoder_experiments.txt):
// warp-per-pair 90.3 ms; eight lanes per pair with batches of 8 / 6 / 4 rows and no explicit prefetch 91.6 / 89.5 / 86.9.
#ifdef EC_SA_MAXNREG
__global__ void __maxnreg__(EC_SA_MAXNREG) decode_self_attn_g8_kernel(
#else
__global__ void __launch_bounds__(kSaThreads, kSaMinBlocks) decode_self_attn_g8_kernel(
#endif
    const bf16* __restrict__ qkv, bf16* kcache, bf16* vcache,
                                                                     bf16* __restrict__ out, int batch, int heads,
                                                                     int max_len, const unsigned char* __restrict__ finished,
                                                                     const DecodeState* __restrict__ st) {
  constexpr int D = 64;
  pdl_trigger_top();
  pdl_wait();
  const int n_pairs = batch * heads;
  const int pair_raw = (int)((blockIdx.x * (unsigned)kSaThreads + threadIdx.x) >> 3);
  const bool valid = pair_raw < n_pairs;
  const int pair = valid ? pair_raw : n_pairs - 1;          // out-of-range groups shadow the last pair (no stores)
  const int b = pair / heads, h = pair % heads;
  const int C = heads * D;
  const int sub = threadIdx.x & 7;
  const bf16* row = qkv + (size_t)b * 3 * C + h * D + sub * 8;
  bf16* kc = kcache + ((size_t)b * max_len) * C + h * D + sub * 8;
  bf16* vc = vcache + ((size_t)b * max_len) * C + h * D + sub * 8;
  // K / V cache rows are a stream (2.3 GB over the 24 layers at 1024 crops, each row read once per step): ld.global.cs /
  // st.global.cs mark them evict-first so that they do not flush the step's small hot set out of L2 (EC_SA_STREAM=0: plain)
  auto load_rows = [&](uint4 (&kk)[kSaR], uint4 (&vv)[kSaR], int j0, int limit) {
#pragma unroll
    for (int r = 0; r < kSaR; ++r) {
      if (j0 + r < limit) {
#if EC_SA_STREAM
        kk[r] = __ldcs(reinterpret_cast<const uint4*>(kc + (size_t)(j0 + r) * C));
        vv[r] = __ldcs(reinterpret_cast<const uint4*>(vc + (size_t)(j0 + r) * C));
#else
        kk[r] = *reinterpret_cast<const uint4*>(kc + (size_t)(j0 + r) * C);
        vv[r] = *reinterpret_cast<const uint4*>(vc + (size_t)(j0 + r) * C);
#endif
      }
    }
  };
  uint4 ku[kSaR], vu[kSaR], kn[kSaR], vn[kSaR];
#pragma unroll
  for (int r = 0; r < kSaR; ++r) { ku[r] = make_uint4(0, 0, 0, 0); vu[r] = ku[r]; kn[r] = ku[r]; vn[r] = ku[r]; }
  load_rows(ku, vu, 0, max_len);                             // speculative: before the step index is known
  const int done = st->done, cur_len = st->cur_len;
  const unsigned char fin = finished[b];
  const uint4 knew = *reinterpret_cast<const uint4*>(row + C), vnew = *reinterpret_cast<const uint4*>(row + 2 * C);
  float qf[8];
  load_q_slice(row, qf);
  if (done) return;                                          // warp-uniform
  const bool live = valid && !fin;                           // per lane group; shuffles below stay warp-wide
  const int pos = cur_len - 1;
  if (EC_SA_PREFETCH) load_rows(kn, vn, kSaR, pos);          // second batch, now that the length is known
  if (live) {
#if EC_SA_STREAM
    __stcs(reinterpret_cast<uint4*>(kc + (size_t)pos * C), knew);
    __stcs(reinterpret_cast<uint4*>(vc + (size_t)pos * C), vnew);
#else
    *reinterpret_cast<uint4*>(kc + (size_t)pos * C) = knew;
    *reinterpret_cast<uint4*>(vc + (size_t)pos * C) = vnew;
#endif
  }
  auto dot8 = [&](const uint4& u) {
    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
    float d = a0.x * qf[0] + a0.y * qf[1] + a1.x * qf[2] + a1.y * qf[3] + a2.x * qf[4] + a2.y * qf[5] + a3.x * qf[6] +
              a3.y * qf[7];
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    return d;
  };
  float acc[8];
  float m = dot8(knew), l = 1.f;                             // the new token itself: p = 1 at the running maximum
  {
    const float2 a0 = unpack_bf16x2(vnew.x), a1 = unpack_bf16x2(vnew.y), a2 = unpack_bf16x2(vnew.z), a3 = unpack_bf16x2(vnew.w);
    acc[0] = a0.x; acc[1] = a0.y; acc[2] = a1.x; acc[3] = a1.y; acc[4] = a2.x; acc[5] = a2.y; acc[6] = a3.x; acc[7] = a3.y;
  }
  for (int j0 = 0; j0 < pos; j0 += kSaR) {
    float sc[kSaR];
    float bm = -INFINITY;
#pragma unroll
    for (int r = 0; r < kSaR; ++r) {
      const float d = dot8(ku[r]);
      sc[r] = (j0 + r < pos) ? d : -INFINITY;
      bm = fmaxf(bm, sc[r]);
    }
    const float mn = fmaxf(m, bm);                           // m is finite from the start
    const float corr = __expf(m - mn);
    m = mn;
    l *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= corr;
#pragma unroll
    for (int r = 0; r < kSaR; ++r) {
      if (j0 + r < pos) {                                    // rows past pos may hold anything (NaN x 0 = NaN): skipped
        const float p = __expf(sc[r] - mn);
        l += p;
        const uint4 u = vu[r];
        const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
        acc[0] += p * a0.x; acc[1] += p * a0.y; acc[2] += p * a1.x; acc[3] += p * a1.y;
        acc[4] += p * a2.x; acc[5] += p * a2.y; acc[6] += p * a3.x; acc[7] += p * a3.y;
      }
    }
    if (EC_SA_PREFETCH) {
#pragma unroll
      for (int r = 0; r < kSaR; ++r) { ku[r] = kn[r]; vu[r] = vn[r]; }
      load_rows(kn, vn, j0 + 2 * kSaR, pos);                 // prefetch the batch after the next
    } else {
      load_rows(ku, vu, j0 + kSaR, pos);
    }
  }
  pdl_trigger_bottom();
  if (live) store_out_slice(out + (size_t)b * C + h * D + sub * 8, acc, 1.0f / l);
}

int decode_self_attention(const bf16* qkv, bf16* kcache, bf16* vcache, bf16* out, int batch, int heads, int head_dim,
                          int max_len, const unsigned char* finished, const DecodeState* st, cudaStream_t stream) {
  EC_REQUIRE(max_len <= kMaxKeys, EC_ERR_UNSUPPORTED, "decode: context %d > %d", max_len, kMaxKeys);
  EC_REQUIRE(head_dim == 64, EC_ERR_UNSUPPORTED, "decode self-attention: head_dim %d != 64", head_dim);
  static const int version = [] { const char* e = getenv("EC_SELFATTN"); return e ? atoi(e) : 3; }();
  note_launches(1);
  if (version == 1) {
    EC_CHECK_CUDA(launch_small_kernel(decode_self_attn_kernel, dim3(ceil_div(batch * heads, 4)), dim3(128), 0, stream, qkv, kcache,
                                vcache, out, batch, heads, max_len, finished, st));
  } else {
    EC_CHECK_CUDA(launch_small_kernel(decode_self_attn_g8_kernel, dim3(ceil_div(batch * heads, kSaThreads / 8)), dim3(kSaThreads),
                                0, stream, qkv, kcache, vcache, out, batch, heads, max_len, finished, st));
  }
  return EC_OK;
}

// cross-attention over the image K/V cache: the cache row of key j is [K(C) | V(C)] bf16
__global__ void __launch_bounds__(128) decode_cross_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kv,
                                                                bf16* __restrict__ out, int batch, int heads, int n_kv,
                                                                const unsigned char* __restrict__ finished,
                                                                const DecodeState* __restrict__ st) {
  constexpr int D = 64;
  pdl_trigger_top();
  pdl_wait();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + w;
  if (pair >= batch * heads) return;
  const int b = pair / heads, h = pair % heads;
  const int C = heads * D;
  const int sub = lane & 7;
  const int done = st->done;                 // control words and the query slice in one round trip
  const unsigned char fin = finished[b];
  float qf[8];
  load_q_slice(q + (size_t)b * C + h * D + sub * 8, qf);
  if (done || fin) return;
  const bf16* kbase = kv + (size_t)b * n_kv * 2 * C + h * D + sub * 8;
  float acc[8], inv;
  warp_attend(qf, kbase, kbase + C, (size_t)2 * C, n_kv, lane, acc, inv);
  pdl_trigger_bottom();
  if ((lane >> 3) == 0) store_out_slice(out + (size_t)b * C + h * D + sub * 8, acc, inv);
}

int decode_cross_attention(const bf16* q, const bf16* kv, bf16* out, int batch, int heads, int head_dim, int n_kv,
                           const unsigned char* finished, const DecodeState* st, cudaStream_t stream) {
  EC_REQUIRE(head_dim == 64, EC_ERR_UNSUPPORTED, "decode cross-attention: head_dim %d != 64", head_dim);
  EC_REQUIRE(n_kv <= kMaxKeys, EC_ERR_UNSUPPORTED, "decode cross-attention: %d keys > %d", n_kv, kMaxKeys);
  const int grid = ceil_div(batch * heads, 4);
  note_launches(1);
  EC_CHECK_CUDA(launch_small_kernel(decode_cross_attn_kernel, dim3(grid), dim3(128), 0, stream, q, kv, out, batch, heads, n_kv,
                              finished, st));
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Token selection: arg-max over the vocab-GEMM partials + generate() bookkeeping (coca_model.py:304-322).
__global__ void __launch_bounds__(256) decode_select_kernel(const float* __restrict__ part_val,
                                                            const int* __restrict__ part_idx, int n_parts,
                                                            int* __restrict__ tokens, int ld_tokens,
                                                            unsigned char* __restrict__ finished, int batch,
                                                            int seq_len, int eos, int pad, DecodeState* st) {
  pdl_trigger_top();
  pdl_wait();
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
  pdl_trigger_bottom();
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
  note_launches(1);
  EC_CHECK_CUDA(launch_small_kernel(decode_select_kernel, dim3(ceil_div(batch, 8)), dim3(256), 0, stream, part_val, part_idx,
                              n_parts, tokens, ld_tokens, finished, batch, seq_len, eos, pad, st));
  return EC_OK;
}

__global__ void decode_advance_kernel(DecodeState* st, int teacher_forced, int batch) {
  pdl_trigger_top();
  pdl_wait();
  if (st->done) return;
  st->cur_len += 1;
  const int act = teacher_forced ? batch : st->next_active;   // teacher forcing never finishes a row
  st->n_active = act;
  st->done = (act == 0) ? 1 : 0;
  st->next_active = 0;
}

int decode_advance(DecodeState* st, bool teacher_forced, int batch, cudaStream_t stream) {
  note_launches(1);
  EC_CHECK_CUDA(launch_small_kernel(decode_advance_kernel, dim3(1), dim3(1), 0, stream, st, teacher_forced ? 1 : 0, batch));
  return EC_OK;
}

__global__ void __launch_bounds__(256) decode_init_kernel(int* __restrict__ tokens, int ld_tokens,
                                                          unsigned char* __restrict__ finished, int batch, int seq_len,
                                                          int sot, int pad, DecodeState* st) {
  const long long total = (long long)batch * seq_len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / seq_len), c = (int)(i % seq_len);
    if (tokens != nullptr) tokens[(size_t)b * ld_tokens + c] = (c == 0) ? sot : pad;   // null: table preset (teacher forcing)
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
