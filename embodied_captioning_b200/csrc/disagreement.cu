// Per-object-instance multi-view disagreement.
//
//   group crops by instance key -> for every instance, normalise its V view embeddings and reduce.
//   Because sum_ij cos(x_i, x_j) = |sum_i x_i/|x_i||^2, the V x V Gram matrix of
//   `_cosine_distance` (projection_utils.py:928-942) collapses into a segmented sum of unit vectors:
//       mode 0 (A11): 1 - |s|^2 / V^2          (mean over all VxV of 1 - cos, diagonal included)
//       mode 1 (A12): (|s|^2 - V) / (V (V-1))  (mean over i<j of cos; NaN when V < 2, compute_cosine_sim.py:13-14)
//   so the kernel is HBM-bound: every embedding row is read exactly once (N*D*4 bytes).
//
// Grouping is exact integer work: a bitonic sort of (key, crop index) pairs, so rows of an instance are
// reduced in ascending global crop index -- the fp32 summation order, hence the score bits, do not depend
// on how many GPUs the crops were sharded over (SURVEY.md 8(e) "Determinism").
#include "kernels.cuh"
#include <climits>

namespace ec {

constexpr int kSortTile = 2048;      // elements sorted per CTA in shared memory
constexpr int kSortThreads = 1024;

struct KeyIdx { long long key; int idx; };

__device__ __forceinline__ bool pair_less(long long ka, int ia, long long kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

__global__ void __launch_bounds__(256) sort_fill_kernel(const long long* __restrict__ keys, int n, int p,
                                                        long long* __restrict__ skey, int* __restrict__ sidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p) return;
  skey[i] = (i < n) ? keys[i] : LLONG_MAX;
  sidx[i] = (i < n) ? i : INT_MAX;
}

// all compare-exchange steps with stride < kSortTile for merge sizes k in [k_lo, k_hi] done in shared memory
__global__ void __launch_bounds__(kSortThreads) sort_tile_kernel(long long* __restrict__ skey, int* __restrict__ sidx,
                                                                 int k_lo, int k_hi) {
  __shared__ long long s_key[kSortTile];
  __shared__ int s_idx[kSortTile];
  const int base = blockIdx.x * kSortTile;
  for (int t = threadIdx.x; t < kSortTile; t += kSortThreads) { s_key[t] = skey[base + t]; s_idx[t] = sidx[base + t]; }
  __syncthreads();
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    for (int j = min(k >> 1, kSortTile >> 1); j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kSortTile / 2; t += kSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // lower index of the pair
        const int l = i | j;
        const bool up = (((base + i) & k) == 0);
        const long long ka = s_key[i], kb = s_key[l];
        const int ia = s_idx[i], ib = s_idx[l];
        const bool swap = up ? pair_less(kb, ib, ka, ia) : pair_less(ka, ia, kb, ib);
        if (swap) { s_key[i] = kb; s_key[l] = ka; s_idx[i] = ib; s_idx[l] = ia; }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < kSortTile; t += kSortThreads) { skey[base + t] = s_key[t]; sidx[base + t] = s_idx[t]; }
}

// one global compare-exchange step (stride j >= kSortTile) of merge size k
__global__ void __launch_bounds__(256) sort_global_step_kernel(long long* __restrict__ skey, int* __restrict__ sidx,
                                                               int p, int k, int j) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p / 2) return;
  const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  const int l = i | j;
  const bool up = ((i & k) == 0);
  const long long ka = skey[i], kb = skey[l];
  const int ia = sidx[i], ib = sidx[l];
  const bool swap = up ? pair_less(kb, ib, ka, ia) : pair_less(ka, ia, kb, ib);
  if (swap) { skey[i] = kb; skey[l] = ka; sidx[i] = ib; sidx[l] = ia; }
}

// single-CTA ordered compaction of segment heads: seg_start[m] = first sorted position of instance m
__global__ void __launch_bounds__(1024) segment_heads_kernel(const long long* __restrict__ skey, int n,
                                                             int* __restrict__ seg_start, long long* __restrict__ out_keys,
                                                             int* __restrict__ out_m) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const int head = (i < n) && (i == 0 || skey[i] != skey[i - 1]);
    const unsigned bal = __ballot_sync(0xffffffffu, head);
    const int wprefix = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int q = 0; q < w; ++q) woff += s_warp[q];
    const int base = s_base;
    if (head) {
      const int m = base + woff + wprefix;
      seg_start[m] = i;
      out_keys[m] = skey[i];
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < 32; ++q) tot += s_warp[q];
      s_base = base + tot;
    }
    __syncthreads();
  }
  if (tid == 0) { *out_m = s_base; seg_start[s_base] = n; }
}

// ---- exact-duplicate view removal (`unique(dim=0)` of SemanticPointCloud.preprocess, projection_utils.py:574-578)
// Values are compared with `==` like torch.unique (so -0.0 == 0.0 and NaN never matches); a 64-bit hash of the
// canonicalised bit patterns prunes the V^2 comparisons inside a segment to (almost) only the true duplicates.
__global__ void __launch_bounds__(256) row_hash_kernel(const float* __restrict__ emb, long long ld_emb, int n, int d,
                                                       unsigned long long* __restrict__ hash) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* r = emb + (size_t)row * ld_emb;
  unsigned long long h = 0x9E3779B97F4A7C15ull * (unsigned long long)(lane + 1);
  for (int c = lane; c < d; c += 32) {
    unsigned int b = __float_as_uint(r[c]);
    if (b == 0x80000000u) b = 0u;                       // -0.0 == 0.0
    h = (h ^ ((unsigned long long)b + ((unsigned long long)c << 32))) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);     // per-lane streams are keyed by lane
  if (lane == 0) hash[row] = h;
}

// dup[r] = 1 when sorted position r repeats an earlier row of its segment; seg_ndup[m] = duplicates in segment m
__global__ void __launch_bounds__(256) segment_dup_kernel(const float* __restrict__ emb, long long ld_emb,
                                                          const unsigned long long* __restrict__ hash,
                                                          const int* __restrict__ sidx, const int* __restrict__ seg_start,
                                                          const int* __restrict__ m_ptr, int d, int* __restrict__ dup,
                                                          int* __restrict__ seg_ndup) {
  __shared__ int s_cnt;
  const int M = *m_ptr;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const int start = seg_start[m], end = seg_start[m + 1];
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int r = start + threadIdx.x; r < end; r += blockDim.x) {
      const int gi = sidx[r];
      const unsigned long long h = hash[gi];
      int is_dup = 0;
      for (int q = start; q < r && !is_dup; ++q) {
        const int gj = sidx[q];
        if (hash[gj] != h) continue;
        const float* a = emb + (size_t)gi * ld_emb;
        const float* b = emb + (size_t)gj * ld_emb;
        int same = 1;
        for (int c = 0; c < d && same; ++c) same = (a[c] == b[c]);
        is_dup = same;
      }
      dup[r] = is_dup;
      mine += is_dup;
    }
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) seg_ndup[m] = s_cnt;
    __syncthreads();
  }
}

constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegMaxVec = 8;   // D <= 1024

__global__ void __launch_bounds__(kSegThreads) segment_reduce_kernel(
    const float* __restrict__ emb, const float* const* __restrict__ peers, int rows_per_rank, long long ld_emb,
    const int* __restrict__ sidx, const int* __restrict__ seg_start, const int* __restrict__ m_ptr, int d, int mode,
    int part_rank, int part_world, const int* __restrict__ dup, const int* __restrict__ seg_ndup,
    float* __restrict__ out_score, int* __restrict__ out_count) {
  __shared__ float s_part[kSegWarps][kSegMaxVec * 128];
  __shared__ float s_red[kSegWarps];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nvec = d >> 2;
  const int M = *m_ptr;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const int start = seg_start[m], end = seg_start[m + 1];
    const int V = end - start - (dup != nullptr ? seg_ndup[m] : 0);      // distinct views when de-duplicating
    if (tid == 0) out_count[m] = V;
    if (part_world > 1 && (m % part_world) != part_rank) {
      if (tid == 0) out_score[m] = 0.f;      // owned by another rank; summed across ranks afterwards
      continue;
    }
    float4 acc[kSegMaxVec];
#pragma unroll
    for (int i = 0; i < kSegMaxVec; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = start + w; r < end; r += kSegWarps) {
      if (dup != nullptr && dup[r]) continue;          // warp-uniform
      const int gi = sidx[r];
      const float* row = (peers != nullptr) ? peers[gi / rows_per_rank] + (size_t)(gi % rows_per_rank) * ld_emb
                                            : emb + (size_t)gi * ld_emb;
      float4 v[kSegMaxVec];
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < kSegMaxVec; ++i) {
        const int c = lane + i * 32;
        v[i] = (c < nvec) ? *(reinterpret_cast<const float4*>(row) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
      ss = warp_sum(ss);
      const float inv = 1.0f / sqrtf(ss);
#pragma unroll
      for (int i = 0; i < kSegMaxVec; ++i) {
        acc[i].x += v[i].x * inv; acc[i].y += v[i].y * inv; acc[i].z += v[i].z * inv; acc[i].w += v[i].w * inv;
      }
    }
#pragma unroll
    for (int i = 0; i < kSegMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) reinterpret_cast<float4*>(s_part[w])[c] = acc[i];
    }
    __syncthreads();
    // fixed-order cross-warp sum, then |s|^2
    float sq = 0.f;
    for (int c = tid; c < d; c += kSegThreads) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < kSegWarps; ++q) s += s_part[q][c];
      sq += s * s;
    }
    sq = warp_sum(sq);
    if (lane == 0) s_red[w] = sq;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int q = 0; q < kSegWarps; ++q) tot += s_red[q];
      const float fv = (float)V;
      float score;
      if (mode == 0) score = 1.0f - tot / (fv * fv);
      else score = (V < 2) ? __int_as_float(0x7fc00000) : (tot - fv) / (fv * (fv - 1.0f));
      out_score[m] = score;
    }
    __syncthreads();
  }
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ascending bitonic sort of p (power of two, >= kSortTile) (key, idx) pairs, idx breaking ties
static void sort_pairs(long long* skey, int* sidx, int p, cudaStream_t stream) {
  sort_tile_kernel<<<p / kSortTile, kSortThreads, 0, stream>>>(skey, sidx, 2, kSortTile);
  note_launches(1);
  for (int k = kSortTile * 2; k <= p; k <<= 1) {
    for (int j = k >> 1; j >= kSortTile; j >>= 1) {
      sort_global_step_kernel<<<ceil_div(p / 2, 256), 256, 0, stream>>>(skey, sidx, p, k, j);
      note_launches(1);
    }
    sort_tile_kernel<<<p / kSortTile, kSortThreads, 0, stream>>>(skey, sidx, k, k);
    note_launches(1);
  }
}

size_t disagreement_workspace_bytes(int n) {
  const size_t p = (size_t)std::max(next_pow2(std::max(n, 1)), kSortTile);
  return align_up(p * sizeof(long long), 256) + align_up(p * sizeof(int), 256) + align_up((p + 1) * sizeof(int), 256) +
         align_up(p * sizeof(unsigned long long), 256) + align_up(p * sizeof(int), 256) + align_up((p + 1) * sizeof(int), 256);
}

int instance_disagreement_peers(const float* const* peers_dev, int world, int rows_per_rank, long long ld_emb,
                                const long long* keys, int n, int d, int mode, int part_rank, int part_world,
                                long long* out_keys, float* out_score, int* out_count, int* out_m, void* ws,
                                size_t ws_bytes, cudaStream_t stream);

static int run(const float* emb, const float* const* peers_dev, int rows_per_rank, long long ld_emb,
               const long long* keys, int n, int d, int mode, int part_rank, int part_world, long long* out_keys,
               float* out_score, int* out_count, int* out_m, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const bool dedup = (mode & EC_DISAGREEMENT_DEDUP) != 0;
  mode &= ~EC_DISAGREEMENT_DEDUP;
  EC_REQUIRE(mode == 0 || mode == 1, EC_ERR_INVALID, "disagreement: mode %d", mode);
  EC_REQUIRE(!dedup || peers_dev == nullptr, EC_ERR_UNSUPPORTED,
             "disagreement: duplicate-view removal needs the rows in local memory (all-gather them first)");
  EC_REQUIRE(d > 0 && d % 4 == 0 && d <= kSegMaxVec * 128, EC_ERR_UNSUPPORTED,
             "disagreement: embedding dim %d unsupported (multiple of 4, <= %d)", d, kSegMaxVec * 128);
  EC_REQUIRE(ld_emb % 4 == 0, EC_ERR_INVALID, "disagreement: row stride must be a multiple of 4 floats");
  if (n <= 0) {
    EC_CHECK_CUDA(cudaMemsetAsync(out_m, 0, sizeof(int), stream));
    return EC_OK;
  }
  EC_REQUIRE(ws_bytes >= disagreement_workspace_bytes(n), EC_ERR_WORKSPACE, "disagreement: workspace %zu < %zu",
             ws_bytes, disagreement_workspace_bytes(n));
  const int p = std::max(next_pow2(n), kSortTile);
  unsigned char* wsb = reinterpret_cast<unsigned char*>(ws);
  long long* skey = reinterpret_cast<long long*>(wsb);
  int* sidx = reinterpret_cast<int*>(wsb + align_up((size_t)p * sizeof(long long), 256));
  int* seg_start = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(sidx) + align_up((size_t)p * sizeof(int), 256));
  unsigned long long* hash = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(seg_start) +
                                                                   align_up((size_t)(p + 1) * sizeof(int), 256));
  int* dup = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(hash) + align_up((size_t)p * sizeof(unsigned long long), 256));
  int* seg_ndup = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(dup) + align_up((size_t)p * sizeof(int), 256));

  sort_fill_kernel<<<ceil_div(p, 256), 256, 0, stream>>>(keys, n, p, skey, sidx);
  note_launches(1);
  sort_pairs(skey, sidx, p, stream);
  segment_heads_kernel<<<1, 1024, 0, stream>>>(skey, n, seg_start, out_keys, out_m);
  const int grid = std::min(n, num_sms() * 4);
  if (dedup) {
    row_hash_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(emb, ld_emb, n, d, hash);
    segment_dup_kernel<<<grid, 256, 0, stream>>>(emb, ld_emb, hash, sidx, seg_start, out_m, d, dup, seg_ndup);
    note_launches(2);
  }
  segment_reduce_kernel<<<grid, kSegThreads, 0, stream>>>(emb, peers_dev, rows_per_rank, ld_emb, sidx, seg_start, out_m,
                                                          d, mode, part_rank, part_world, dedup ? dup : nullptr,
                                                          dedup ? seg_ndup : nullptr, out_score, out_count);
  EC_POST_LAUNCH();
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// Caption frequency table per instance: the exact-string histogram of
// `PseudoCaptioner.compute_captions_frequency` (captioner/pseudocaptioner.py:156-177) over token-id sequences.
// Rows are grouped by a 64-bit hash of (instance key, token row); every row of a group is then compared in full
// with the group's first row, so a hash collision is reported (EC_ERR_STATE) instead of silently merging captions.
__global__ void __launch_bounds__(256) caption_hash_kernel(const int* __restrict__ tokens, int ld, int t,
                                                           const long long* __restrict__ keys, int n, int p,
                                                           long long* __restrict__ skey, int* __restrict__ sidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p) return;
  if (i >= n) { skey[i] = LLONG_MAX; sidx[i] = INT_MAX; return; }
  unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)keys[i];
  h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
  for (int c = 0; c < t; ++c) {
    h = (h ^ (unsigned long long)(unsigned int)tokens[(size_t)i * ld + c]) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
  }
  skey[i] = (long long)(h >> 1);          // non-negative: real rows sort before the LLONG_MAX padding
  sidx[i] = i;
}

// heads of the (hash, idx)-sorted runs (head_of[r] = 1): representative = first occurrence (smallest row index);
// every other row is verified against its predecessor
__global__ void __launch_bounds__(256) caption_runs_kernel(const long long* __restrict__ skey,
                                                           const int* __restrict__ sidx, int n,
                                                           const int* __restrict__ tokens, int ld, int t,
                                                           const long long* __restrict__ keys,
                                                           int* __restrict__ head_of, int* __restrict__ collision) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const bool is_head = (r == 0) || (skey[r - 1] != skey[r]);
  head_of[r] = is_head ? 1 : 0;
  if (!is_head) {                                   // equality is transitive: comparing neighbours covers the run
    const int a = sidx[r], b = sidx[r - 1];
    bool same = keys[a] == keys[b];
    for (int c = 0; c < t && same; ++c) same = tokens[(size_t)a * ld + c] == tokens[(size_t)b * ld + c];
    if (!same) *collision = 1;
  }
}

// one entry per run head: (instance key, first row) + run length, compacted in sorted-run order
__global__ void __launch_bounds__(1024) caption_compact_kernel(const int* __restrict__ head_of,
                                                               const int* __restrict__ sidx, int n,
                                                               const long long* __restrict__ keys, int p2,
                                                               long long* __restrict__ gkey, int* __restrict__ gfirst,
                                                               int* __restrict__ run_len, int* __restrict__ out_m) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const int head = (i < n) && head_of[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, head);
    const int wprefix = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int q = 0; q < w; ++q) woff += s_warp[q];
    const int base = s_base;
    if (head) {
      const int m = base + woff + wprefix;
      int len = 1;
      while (i + len < n && head_of[i + len] == 0) ++len;
      gkey[m] = keys[sidx[i]];
      gfirst[m] = sidx[i];
      run_len[sidx[i]] = len;               // looked up by first-row index after the second sort
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < 32; ++q) tot += s_warp[q];
      s_base = base + tot;
    }
    __syncthreads();
  }
  const int m_total = s_base;
  for (int i = m_total + tid; i < p2; i += 1024) { gkey[i] = LLONG_MAX; gfirst[i] = INT_MAX; }
  if (tid == 0) *out_m = m_total;
}

__global__ void __launch_bounds__(256) caption_emit_kernel(const long long* __restrict__ gkey,
                                                           const int* __restrict__ gfirst,
                                                           const int* __restrict__ run_len, const int* __restrict__ m_ptr,
                                                           long long* __restrict__ out_key, int* __restrict__ out_first,
                                                           int* __restrict__ out_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *m_ptr) return;
  out_key[i] = gkey[i];
  out_first[i] = gfirst[i];
  out_count[i] = run_len[gfirst[i]];
}

size_t caption_frequency_workspace_bytes(int n) {
  const size_t p = (size_t)std::max(next_pow2(std::max(n, 1)), kSortTile);
  return 2 * (align_up(p * sizeof(long long), 256) + align_up(p * sizeof(int), 256)) + 2 * align_up(p * sizeof(int), 256) + 256;
}

int caption_frequency(const int* tokens, int ld_tokens, int t, const long long* keys, int n, long long* out_key,
                      int* out_first, int* out_count, int* out_m, void* ws, size_t ws_bytes, cudaStream_t stream) {
  EC_REQUIRE(t > 0 && ld_tokens >= t, EC_ERR_INVALID, "caption_frequency: bad token matrix shape");
  if (n <= 0) {
    EC_CHECK_CUDA(cudaMemsetAsync(out_m, 0, sizeof(int), stream));
    return EC_OK;
  }
  EC_REQUIRE(ws_bytes >= caption_frequency_workspace_bytes(n), EC_ERR_WORKSPACE, "caption_frequency: workspace %zu < %zu",
             ws_bytes, caption_frequency_workspace_bytes(n));
  const int p = std::max(next_pow2(n), kSortTile);
  unsigned char* b = reinterpret_cast<unsigned char*>(ws);
  long long* skey = reinterpret_cast<long long*>(b); b += align_up((size_t)p * sizeof(long long), 256);
  int* sidx = reinterpret_cast<int*>(b); b += align_up((size_t)p * sizeof(int), 256);
  long long* gkey = reinterpret_cast<long long*>(b); b += align_up((size_t)p * sizeof(long long), 256);
  int* gfirst = reinterpret_cast<int*>(b); b += align_up((size_t)p * sizeof(int), 256);
  int* head_of = reinterpret_cast<int*>(b); b += align_up((size_t)p * sizeof(int), 256);
  int* run_len = reinterpret_cast<int*>(b); b += align_up((size_t)p * sizeof(int), 256);
  int* collision = reinterpret_cast<int*>(b);
  EC_CHECK_CUDA(cudaMemsetAsync(collision, 0, sizeof(int), stream));
  caption_hash_kernel<<<ceil_div(p, 256), 256, 0, stream>>>(tokens, ld_tokens, t, keys, n, p, skey, sidx);
  sort_pairs(skey, sidx, p, stream);
  caption_runs_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(skey, sidx, n, tokens, ld_tokens, t, keys, head_of, collision);
  caption_compact_kernel<<<1, 1024, 0, stream>>>(head_of, sidx, n, keys, p, gkey, gfirst, run_len, out_m);
  // groups in (instance key, first occurrence) order = dict insertion order inside every instance
  sort_pairs(gkey, gfirst, p, stream);
  caption_emit_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(gkey, gfirst, run_len, out_m, out_key, out_first, out_count);
  note_launches(4);
  EC_CHECK_CUDA(cudaGetLastError());
  int hit = 0;
  EC_CHECK_CUDA(cudaMemcpyAsync(&hit, collision, sizeof(int), cudaMemcpyDeviceToHost, stream));
  EC_CHECK_CUDA(cudaStreamSynchronize(stream));
  EC_REQUIRE(hit == 0, EC_ERR_STATE, "caption_frequency: 64-bit hash collision between different captions");
  return EC_OK;
}

int instance_disagreement(const float* emb, long long ld_emb, const long long* keys, int n, int d, int mode,
                          long long* out_keys, float* out_score, int* out_count, int* out_m, void* ws,
                          size_t ws_bytes, cudaStream_t stream) {
  return run(emb, nullptr, 1, ld_emb, keys, n, d, mode, 0, 1, out_keys, out_score, out_count, out_m, ws, ws_bytes,
             stream);
}

int instance_disagreement_peers(const float* const* peers_dev, int world, int rows_per_rank, long long ld_emb,
                                const long long* keys, int n, int d, int mode, int part_rank, int part_world,
                                long long* out_keys, float* out_score, int* out_count, int* out_m, void* ws,
                                size_t ws_bytes, cudaStream_t stream) {
  EC_REQUIRE(peers_dev != nullptr && world > 0 && rows_per_rank > 0, EC_ERR_INVALID, "disagreement: bad peer table");
  EC_REQUIRE((long long)world * rows_per_rank >= n, EC_ERR_INVALID, "disagreement: peers hold fewer than n rows");
  return run(nullptr, peers_dev, rows_per_rank, ld_emb, keys, n, d, mode, part_rank, part_world, out_keys, out_score,
             out_count, out_m, ws, ws_bytes, stream);
}

}  // namespace ec
