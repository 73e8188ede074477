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

constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegMaxVec = 8;   // D <= 1024

__global__ void __launch_bounds__(kSegThreads) segment_reduce_kernel(
    const float* __restrict__ emb, const float* const* __restrict__ peers, int rows_per_rank, long long ld_emb,
    const int* __restrict__ sidx, const int* __restrict__ seg_start, const int* __restrict__ m_ptr, int d, int mode,
    int part_rank, int part_world, float* __restrict__ out_score, int* __restrict__ out_count) {
  __shared__ float s_part[kSegWarps][kSegMaxVec * 128];
  __shared__ float s_red[kSegWarps];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nvec = d >> 2;
  const int M = *m_ptr;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const int start = seg_start[m], end = seg_start[m + 1];
    const int V = end - start;
    if (tid == 0) out_count[m] = V;
    if (part_world > 1 && (m % part_world) != part_rank) {
      if (tid == 0) out_score[m] = 0.f;      // owned by another rank; summed across ranks afterwards
      continue;
    }
    float4 acc[kSegMaxVec];
#pragma unroll
    for (int i = 0; i < kSegMaxVec; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = start + w; r < end; r += kSegWarps) {
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

size_t disagreement_workspace_bytes(int n) {
  const size_t p = (size_t)std::max(next_pow2(std::max(n, 1)), kSortTile);
  return align_up(p * sizeof(long long), 256) + align_up(p * sizeof(int), 256) + align_up((p + 1) * sizeof(int), 256);
}

int instance_disagreement_peers(const float* const* peers_dev, int world, int rows_per_rank, long long ld_emb,
                                const long long* keys, int n, int d, int mode, int part_rank, int part_world,
                                long long* out_keys, float* out_score, int* out_count, int* out_m, void* ws,
                                size_t ws_bytes, cudaStream_t stream);

static int run(const float* emb, const float* const* peers_dev, int rows_per_rank, long long ld_emb,
               const long long* keys, int n, int d, int mode, int part_rank, int part_world, long long* out_keys,
               float* out_score, int* out_count, int* out_m, void* ws, size_t ws_bytes, cudaStream_t stream) {
  EC_REQUIRE(mode == 0 || mode == 1, EC_ERR_INVALID, "disagreement: mode %d", mode);
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

  sort_fill_kernel<<<ceil_div(p, 256), 256, 0, stream>>>(keys, n, p, skey, sidx);
  sort_tile_kernel<<<p / kSortTile, kSortThreads, 0, stream>>>(skey, sidx, 2, kSortTile);
  for (int k = kSortTile * 2; k <= p; k <<= 1) {
    for (int j = k >> 1; j >= kSortTile; j >>= 1)
      sort_global_step_kernel<<<ceil_div(p / 2, 256), 256, 0, stream>>>(skey, sidx, p, k, j);
    sort_tile_kernel<<<p / kSortTile, kSortThreads, 0, stream>>>(skey, sidx, k, k);
  }
  segment_heads_kernel<<<1, 1024, 0, stream>>>(skey, n, seg_start, out_keys, out_m);
  const int grid = std::min(n, num_sms() * 4);
  segment_reduce_kernel<<<grid, kSegThreads, 0, stream>>>(emb, peers_dev, rows_per_rank, ld_emb, sidx, seg_start, out_m,
                                                          d, mode, part_rank, part_world, out_score, out_count);
  EC_POST_LAUNCH();
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
