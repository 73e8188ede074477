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

// ----------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the decoder step is launched with the programmatic-stream-serialization attribute: its
// CTAs may be scheduled (and run their prologue) while the previous kernel drains; `pdl_wait()` then blocks
// until the previous grid has completed and its writes are visible.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // EC_PDL=1 enables (round-1 measurement: 150 ms vs 137 ms decode with it on, so default off)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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
  // nn.GELU() exact form (coca_model.py:65): 0.5 x (1 + erf(x / sqrt(2))).  erf by Abramowitz-Stegun 7.1.26
  // (|abs err| < 1.5e-7, far below the bf16 rounding of the stored result) -- a third of erff()'s instructions,
  // which matters because the fc1 epilogue evaluates it 128 x 256 times per tile next to the MMA main loop.
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * __expf(-z * z);
  const float erf_v = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_v);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ec
