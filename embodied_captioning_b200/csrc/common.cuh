// Shared device/host helpers for libec_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/ec_b200.h"

namespace ec {

// ----------------------------------------------------------------------------- error plumbing
// status codes: the EC_OK / EC_ERR_* macros of include/ec_b200.h
void set_error(const char* fmt, ...);
const char* last_error();

#define EC_CHECK_CUDA(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ::ec::set_error("%s:%d CUDA error %s (%s)", __FILE__, __LINE__, cudaGetErrorName(_e),      \
                      cudaGetErrorString(_e));                                                   \
      return EC_ERR_CUDA;                                                                  \
    }                                                                                            \
  } while (0)

#define EC_REQUIRE(cond, code, ...)                                                              \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      ::ec::set_error(__VA_ARGS__);                                                              \
      return (code);                                                                             \
    }                                                                                            \
  } while (0)

#define EC_TRY(expr)                                                                             \
  do {                                                                                           \
    int _rc = (expr);                                                                            \
    if (_rc != EC_OK) return _rc;                                                          \
  } while (0)

int num_sms();

// kernel-launch accounting (bench.py reports how many of our kernels ran inside the timed region)
void note_launches(long long n);
long long launch_count();

#define EC_POST_LAUNCH()                       \
  do {                                         \
    ::ec::note_launches(1);                    \
    EC_CHECK_CUDA(cudaGetLastError());         \
  } while (0)

// ----------------------------------------------------------------------------- device helpers
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() exact form (coca_model.py:65): 0.5 x (1 + erf(x / sqrt(2)))
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ec
