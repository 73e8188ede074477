// LayerNorm and small row-wise kernels (HBM-bound, one warp per row, 128-bit accesses).
#include "kernels.cuh"

namespace ec {

constexpr int kLnMaxVec = 8;   // 8 float4 per lane -> width <= 1024

__device__ __forceinline__ void ln_stats(const float4 (&v)[kLnMaxVec], int nvec, int lane, int width, float& mean,
                                         float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i)
    if (lane + i * 32 < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) / (float)width;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i)
    if (lane + i * 32 < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  const float var = warp_sum(q) / (float)width;
  rstd = 1.0f / sqrtf(var + 1e-5f);
}

template <bool kBf16Out>
__device__ __forceinline__ void ln_store(const float4 (&v)[kLnMaxVec], int nvec, int lane, float mean, float rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, void* yrow) {
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (kBf16Out) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(yrow)[c] = pk;
      } else {
        reinterpret_cast<float4*>(yrow)[c] = o;
      }
    }
  }
}

// Throughput variant (image tower: 263 k rows per launch, HBM-bound): few registers, four blocks per SM.
template <bool kBf16Out>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long ldx,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, void* __restrict__ y,
                                                        long long ldy, int rows, int width,
                                                        const int* __restrict__ skip_flag) {
  pdl_trigger_top();
  pdl_wait();
  if (skip_flag != nullptr && *skip_flag != 0) return;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = width >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
  float4 v[kLnMaxVec];
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    v[i] = (c < nvec) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float mean, rstd;
  ln_stats(v, nvec, lane, width, mean, rstd);
  pdl_trigger_bottom();
  void* yrow = kBf16Out ? (void*)(reinterpret_cast<bf16*>(y) + (size_t)row * ldy)
                        : (void*)(reinterpret_cast<float*>(y) + (size_t)row * ldy);
  ln_store<kBf16Out>(v, nvec, lane, mean, rstd, gamma, beta, yrow);
}

// Latency variant (decoder steps: ~1 k rows, a 2.5 us kernel on the critical path of the step).
template <bool kBf16Out>
__global__ void __launch_bounds__(256) layernorm_small_kernel(const float* __restrict__ x, long long ldx,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, void* __restrict__ y,
                                                              long long ldy, int rows, int width,
                                                              const int* __restrict__ skip_flag) {
  pdl_trigger_top();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nvec = width >> 2;
  // gamma / beta first, and before the dependency wait (they are parameters, not the previous kernel's output): in the
  // decoder steps they are HBM-cold every time (557 MB of weights + 9.6 GB of cross K/V pass through L2 per step), and
  // loading them after the row statistics put that miss, plus one L2 round trip per vector, on the critical path
  // of a 3 us kernel.
  float4 g[kLnMaxVec], b[kLnMaxVec];
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      b[i] = __ldg(reinterpret_cast<const float4*>(beta) + c);
    }
  }
  pdl_wait();
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
  float4 v[kLnMaxVec];
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    v[i] = (c < nvec) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (skip_flag != nullptr && *skip_flag != 0) return;      // after the row loads are in flight
  float mean, rstd;
  ln_stats(v, nvec, lane, width, mean, rstd);
  pdl_trigger_bottom();
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 o;
      o.x = (v[i].x - mean) * rstd * g[i].x + b[i].x;
      o.y = (v[i].y - mean) * rstd * g[i].y + b[i].y;
      o.z = (v[i].z - mean) * rstd * g[i].z + b[i].z;
      o.w = (v[i].w - mean) * rstd * g[i].w + b[i].w;
      if (kBf16Out) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + (size_t)row * ldy)[c] = pk;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)row * ldy)[c] = o;
      }
    }
  }
}

int layernorm(const float* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy,
              bool y_is_bf16, int rows, int width, const int* skip_flag, cudaStream_t stream) {
  EC_REQUIRE(width % 4 == 0 && width <= kLnMaxVec * 128 && width > 0, EC_ERR_UNSUPPORTED,
             "layernorm width %d unsupported (need multiple of 4, <= %d)", width, kLnMaxVec * 128);
  EC_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, EC_ERR_INVALID, "layernorm strides must be multiples of 4");
  if (rows <= 0) return EC_OK;
  const int wpb = 8;
  const int grid = ceil_div(rows, wpb);
  note_launches(1);
  const bool small = rows <= 8192;      // one wave of blocks: latency, not bandwidth, is what counts
  if (y_is_bf16)
    EC_CHECK_CUDA(launch_small_kernel(small ? layernorm_small_kernel<true> : layernorm_kernel<true>, dim3(grid), dim3(wpb * 32), 0,
                                stream, x, ldx, gamma, beta, y, ldy, rows, width, skip_flag));
  else
    EC_CHECK_CUDA(launch_small_kernel(small ? layernorm_small_kernel<false> : layernorm_kernel<false>, dim3(grid), dim3(wpb * 32),
                                0, stream, x, ldx, gamma, beta, y, ldy, rows, width, skip_flag));
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
// One warp per output row of the folded linear (model-load time, not on the hot path).
__global__ void __launch_bounds__(256) fold_ln_linear_kernel(const bf16* __restrict__ w, const float* __restrict__ bias,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, bf16* __restrict__ w_fold,
                                                             float* __restrict__ b_fold, int n_out, int k) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_out) return;
  float acc = 0.f;
  for (int c = lane; c < k; c += 32) {
    const float wv = __bfloat162float(w[(size_t)row * k + c]);
    acc += wv * beta[c];
    w_fold[(size_t)row * k + c] = __float2bfloat16_rn(wv * gamma[c]);
  }
  acc = warp_sum(acc);
  if (lane == 0) b_fold[row] = (bias != nullptr ? bias[row] : 0.f) + acc;
}

int fold_layernorm_into_linear(const bf16* w, const float* bias, const float* gamma, const float* beta, bf16* w_fold,
                               float* b_fold, int n_out, int k, cudaStream_t stream) {
  if (n_out <= 0) return EC_OK;
  fold_ln_linear_kernel<<<ceil_div(n_out, 8), 256, 0, stream>>>(w, bias, gamma, beta, w_fold, b_fold, n_out, k);
  EC_POST_LAUNCH();
  return EC_OK;
}

__global__ void fill_f32_kernel(float* out, int n, float value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = value;
}

int fill_f32(float* out, int n, float value, cudaStream_t stream) {
  if (n <= 0) return EC_OK;
  fill_f32_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(out, n, value);
  EC_POST_LAUNCH();
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vision_embed_ln_pre_kernel(const float* __restrict__ patch_emb,
                                                                  const float* __restrict__ cls,
                                                                  const float* __restrict__ pos,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  float* __restrict__ out, int batch, int tokens,
                                                                  int width) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)batch * tokens) return;
  const int b = (int)(row / tokens), t = (int)(row % tokens);
  const int nvec = width >> 2;
  const float4* src = (t == 0) ? reinterpret_cast<const float4*>(cls)
                               : reinterpret_cast<const float4*>(patch_emb + ((size_t)b * (tokens - 1) + (t - 1)) * width);
  const float4* pr = reinterpret_cast<const float4*>(pos + (size_t)t * width);
  float4 v[kLnMaxVec];
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 a = src[c], p = __ldg(pr + c);
      v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float mean, rstd;
  ln_stats(v, nvec, lane, width, mean, rstd);
  ln_store<false>(v, nvec, lane, mean, rstd, gamma, beta, out + (size_t)row * width);
}

int vision_embed_ln_pre(const float* patch_emb, const float* cls, const float* pos, const float* gamma,
                        const float* beta, float* out, int batch, int tokens, int width, cudaStream_t stream) {
  EC_REQUIRE(width % 4 == 0 && width <= kLnMaxVec * 128, EC_ERR_UNSUPPORTED, "vision width %d unsupported", width);
  const long long rows = (long long)batch * tokens;
  if (rows <= 0) return EC_OK;
  const int wpb = 8;
  vision_embed_ln_pre_kernel<<<(unsigned)ceil_div<long long>(rows, wpb), wpb * 32, 0, stream>>>(
      patch_emb, cls, pos, gamma, beta, out, batch, tokens, width);
  EC_POST_LAUNCH();
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_kernel(const bf16* __restrict__ img, bf16* __restrict__ out, int n,
                                                     int size, int patch, int ld) {
  const int grid_w = size / patch;
  const int kdim = 3 * patch * patch;
  const long long total = (long long)n * grid_w * grid_w * kdim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % kdim);
    const long long prow = i / kdim;
    const int pidx = (int)(prow % (grid_w * grid_w));
    const int b = (int)(prow / (grid_w * grid_w));
    const int c = k / (patch * patch), rem = k % (patch * patch);
    const int py = rem / patch, px = rem % patch;
    const int gy = pidx / grid_w, gx = pidx % grid_w;
    out[prow * ld + k] = img[(((size_t)b * 3 + c) * size + (gy * patch + py)) * size + gx * patch + px];
  }
}

// two pixels (one 32-bit word) per thread: patch rows are runs of `patch` contiguous pixels in both layouts and every
// offset is even when `patch` and `ld` are
__global__ void __launch_bounds__(256) im2col_x2_kernel(const uint32_t* __restrict__ img, uint32_t* __restrict__ out, int n,
                                                        int size, int patch, int ld) {
  const int grid_w = size / patch;
  const int hp = patch >> 1;                      // words per patch row
  const int kwords = 3 * patch * hp;              // words per output row
  const long long total = (long long)n * grid_w * grid_w * kwords;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(i % kwords);
    const long long prow = i / kwords;
    const int pidx = (int)(prow % (grid_w * grid_w));
    const int b = (int)(prow / (grid_w * grid_w));
    const int c = kw / (patch * hp), rem = kw % (patch * hp);
    const int py = rem / hp, pw = rem % hp;
    const int gy = pidx / grid_w, gx = pidx % grid_w;
    out[(prow * ld) / 2 + kw] = img[((((size_t)b * 3 + c) * size + (gy * patch + py)) * size + gx * patch) / 2 + pw];
  }
}

int im2col_patches(const bf16* img, bf16* out, int n, int size, int patch, int ld, cudaStream_t stream) {
  EC_REQUIRE(size % patch == 0 && ld >= 3 * patch * patch, EC_ERR_INVALID, "im2col: bad geometry");
  if (n <= 0) return EC_OK;
  if (patch % 2 == 0 && ld % 2 == 0 && size % 2 == 0 && (reinterpret_cast<uintptr_t>(img) & 3u) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
    const long long words = (long long)n * (size / patch) * (size / patch) * 3 * patch * (patch / 2);
    const int grid2 = (int)std::min<long long>(ceil_div<long long>(words, 256), (long long)num_sms() * 16);
    im2col_x2_kernel<<<grid2, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(img), reinterpret_cast<uint32_t*>(out), n,
                                                size, patch, ld);
    EC_POST_LAUNCH();
    return EC_OK;
  }
  const long long total = (long long)n * (size / patch) * (size / patch) * 3 * patch * patch;
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)num_sms() * 16);
  im2col_kernel<<<grid, 256, 0, stream>>>(img, out, n, size, patch, ld);
  EC_POST_LAUNCH();
  return EC_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ y,
                                                     long long ldy, int rows, int width) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ldx;
  float s = 0.f;
  for (int c = lane; c < width; c += 32) s += xr[c] * xr[c];
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane; c < width; c += 32) y[(size_t)row * ldy + c] = xr[c] * inv;
}

int l2_normalize_rows(const float* x, long long ldx, float* y, long long ldy, int rows, int width,
                      cudaStream_t stream) {
  if (rows <= 0) return EC_OK;
  l2norm_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(x, ldx, y, ldy, rows, width);
  EC_POST_LAUNCH();
  return EC_OK;
}

__global__ void __launch_bounds__(256) cvt_rows_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y,
                                                       long long ldy, int rows, int width) {
  const long long total = (long long)rows * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / width;
    const int c = (int)(i % width);
    y[r * ldy + c] = __float2bfloat16_rn(x[r * ldx + c]);
  }
}

int convert_rows_f32_to_bf16(const float* x, long long ldx, bf16* y, long long ldy, int rows, int width,
                             cudaStream_t stream) {
  if (rows <= 0) return EC_OK;
  const long long total = (long long)rows * width;
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), (long long)num_sms() * 16);
  cvt_rows_kernel<<<grid, 256, 0, stream>>>(x, ldx, y, ldy, rows, width);
  EC_POST_LAUNCH();
  return EC_OK;
}

}  // namespace ec
