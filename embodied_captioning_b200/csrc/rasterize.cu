// Disagreement-map rasterisation + exploration reward.
//
// Restates the tail of `SemanticPointCloud.get_topdown_semantic` (experimenting_env/utils/projection_utils.py:
// 694-703 grid coordinates, :760-785 score gather + centre write + "inflate" writes) and `get_reward`
// (experimenting_env/envs/env_habitat3.py:75-84: map.sum() / 1000).  The reference's semantics are those of
// numpy fancy assignment, kept exactly:
//   * seven assignments in sequence: the centre cell, then the offsets (i, j) in {-1,0,1}^2 with i != j
//     (so only SIX neighbours: (-1,-1) and (1,1) are skipped together with (0,0));
//   * inside one assignment duplicates resolve to the LAST point in point order, later assignments overwrite
//     earlier ones -> every cell takes the value of the write with the highest (assignment, point) sequence number;
//   * negative indices in [-size, 0) wrap around (numpy indexing);
//   * `_safe_assign` (projection_utils.py:614-631): if ANY index of an assignment is out of range numpy raises
//     before writing anything, and the fallback cannot re-shape the torch-tensor values -> the whole assignment is
//     dropped.
// HBM-bound scatter: 16 B/point read, 7 x 8 B atomics into an L2-resident sequence map, one pass over the map.
#include "kernels.cuh"

namespace ec {

struct RasterWs { int* gx; int* gy; float* val; int* call_bad; unsigned long long* seq; double* partial; };

__device__ __forceinline__ void raster_offset(int call, int& di, int& dj) {
  // call 0 = centre; calls 1..6 = (i, j) loop order of projection_utils.py:773-775 with i != j
  const int oi[7] = {0, -1, -1, 0, 0, 1, 1};
  const int oj[7] = {0, 0, 1, -1, 1, -1, 0};
  di = oi[call]; dj = oj[call];
}

// per point: grid cell (projection_utils.py:694-699) and score of its object (binary search in the sorted keys)
__global__ void __launch_bounds__(256) raster_points_kernel(const float* __restrict__ points, int n,
                                                            const long long* __restrict__ point_keys,
                                                            const long long* __restrict__ inst_keys,
                                                            const float* __restrict__ inst_scores, int m, float scale,
                                                            int off_x, int off_y, int H, int W, RasterWs w) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float x = points[(size_t)p * 3 + 0], z = points[(size_t)p * 3 + 2];
  const int gx = (int)__fdiv_rn(x, scale) - off_x;          // (points[:,0] / map_scale).astype(int): truncation
  const int gy = (int)__fdiv_rn(z, scale) - off_y;
  const long long key = point_keys[p];
  int lo = 0, hi = m;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (inst_keys[mid] < key) lo = mid + 1; else hi = mid; }
  const float v = (lo < m && inst_keys[lo] == key) ? inst_scores[lo] : 0.f;
  const bool valid = v > 0.f;                                // cosine_distance_per_point > 0 (NaN is not valid)
  w.gx[p] = gx; w.gy[p] = gy; w.val[p] = valid ? v : 0.f;
  if (valid) {
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      int di, dj; raster_offset(c, di, dj);
      const int yy = gy + di, xx = gx + dj;
      if (yy < -H || yy >= H || xx < -W || xx >= W) w.call_bad[c] = 1;    // benign race: all writers store 1
    }
  }
}

__global__ void __launch_bounds__(256) raster_claim_kernel(int n, int H, int W, RasterWs w) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * 7) return;
  const int c = (int)(t / n), p = (int)(t % n);
  if (!(w.val[p] > 0.f) || w.call_bad[c]) return;
  int di, dj; raster_offset(c, di, dj);
  int yy = w.gy[p] + di, xx = w.gx[p] + dj;
  if (yy < 0) yy += H;
  if (xx < 0) xx += W;
  atomicMax(&w.seq[(size_t)yy * W + xx], (unsigned long long)t + 1ull);
}

__global__ void __launch_bounds__(256) raster_write_kernel(int n, int H, int W, RasterWs w, float* __restrict__ map) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * 7) return;
  const int c = (int)(t / n), p = (int)(t % n);
  const float v = w.val[p];
  if (!(v > 0.f) || w.call_bad[c]) return;
  int di, dj; raster_offset(c, di, dj);
  int yy = w.gy[p] + di, xx = w.gx[p] + dj;
  if (yy < 0) yy += H;
  if (xx < 0) xx += W;
  const size_t cell = (size_t)yy * W + xx;
  if (w.seq[cell] == (unsigned long long)t + 1ull) map[cell] = v;     // the last write in reference order wins
}

constexpr int kRasterBlocks = 256;

// deterministic two-stage sum in double: contiguous chunk per block, fixed tree inside the block
__global__ void __launch_bounds__(256) raster_sum_kernel(const float* __restrict__ map, long long cells,
                                                         double* __restrict__ partial) {
  __shared__ double s[256];
  const long long per = (cells + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * per, e = min(cells, b + per);
  double acc = 0.0;
  for (long long i = b + threadIdx.x; i < e; i += 256) acc += (double)map[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

__global__ void raster_reward_kernel(const double* __restrict__ partial, int nb, float* __restrict__ reward) {
  double t = 0.0;
  for (int i = 0; i < nb; ++i) t += partial[i];
  *reward = (float)(t / 1000.0);                                        // env_habitat3.py:79
}

size_t disagreement_map_workspace_bytes(int n_points, int H, int W) {
  const size_t n = (size_t)std::max(n_points, 1);
  return 3 * align_up(n * 4, 256) + 256 + align_up((size_t)H * W * 8, 256) + align_up(kRasterBlocks * 8, 256);
}

int disagreement_map(const float* points, int n_points, const long long* point_keys, const long long* inst_keys,
                     const float* inst_scores, int m, float scale, int off_x, int off_y, int H, int W, float* map,
                     float* reward, void* ws, size_t ws_bytes, cudaStream_t stream) {
  EC_REQUIRE(H > 0 && W > 0 && scale > 0.f, EC_ERR_INVALID, "disagreement_map: bad grid %d x %d, scale %g", H, W, scale);
  EC_REQUIRE((long long)n_points * 7 < (1ll << 40), EC_ERR_UNSUPPORTED, "disagreement_map: too many points");
  EC_REQUIRE(ws_bytes >= disagreement_map_workspace_bytes(n_points, H, W), EC_ERR_WORKSPACE,
             "disagreement_map: workspace %zu < %zu", ws_bytes, disagreement_map_workspace_bytes(n_points, H, W));
  const size_t n = (size_t)std::max(n_points, 1);
  unsigned char* b = reinterpret_cast<unsigned char*>(ws);
  RasterWs w;
  w.gx = reinterpret_cast<int*>(b); b += align_up(n * 4, 256);
  w.gy = reinterpret_cast<int*>(b); b += align_up(n * 4, 256);
  w.val = reinterpret_cast<float*>(b); b += align_up(n * 4, 256);
  w.call_bad = reinterpret_cast<int*>(b); b += 256;
  w.seq = reinterpret_cast<unsigned long long*>(b); b += align_up((size_t)H * W * 8, 256);
  w.partial = reinterpret_cast<double*>(b);
  const long long cells = (long long)H * W;
  EC_CHECK_CUDA(cudaMemsetAsync(w.call_bad, 0, 256, stream));
  EC_CHECK_CUDA(cudaMemsetAsync(w.seq, 0, (size_t)cells * 8, stream));
  EC_CHECK_CUDA(cudaMemsetAsync(map, 0, (size_t)cells * 4, stream));
  if (n_points > 0 && m > 0) {
    raster_points_kernel<<<ceil_div(n_points, 256), 256, 0, stream>>>(points, n_points, point_keys, inst_keys,
                                                                      inst_scores, m, scale, off_x, off_y, H, W, w);
    const unsigned grid = (unsigned)ceil_div<long long>((long long)n_points * 7, 256);
    raster_claim_kernel<<<grid, 256, 0, stream>>>(n_points, H, W, w);
    raster_write_kernel<<<grid, 256, 0, stream>>>(n_points, H, W, w, map);
    note_launches(3);
  }
  if (reward != nullptr) {
    raster_sum_kernel<<<kRasterBlocks, 256, 0, stream>>>(map, cells, w.partial);
    raster_reward_kernel<<<1, 1, 0, stream>>>(w.partial, kRasterBlocks, reward);
    note_launches(2);
  }
  EC_CHECK_CUDA(cudaGetLastError());
  return EC_OK;
}

}  // namespace ec
