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
int current_device();     // cudaGetDevice, -1 on error (kernel attributes such as the dynamic smem limit are per device)

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
// Where the trigger sits: at the top of a kernel the dependent grid is released as soon as every CTA of this grid
// has started (its CTAs then compete for SM slots with this grid's later waves); at the bottom (default) only the
// launch latency and the dependent's prologue overlap this grid's tail.
#ifdef EC_PDL_EARLY
__device__ __forceinline__ void pdl_trigger_top() { pdl_launch_dependents(); }
__device__ __forceinline__ void pdl_trigger_bottom() {}
#else
__device__ __forceinline__ void pdl_trigger_top() {}
__device__ __forceinline__ void pdl_trigger_bottom() { pdl_launch_dependents(); }
#endif

// EC_PDL: 0 = off, 1 = every kernel of the step (measured inside the step graph in round 1: 127.8 ms with, 121.5 ms
// without), 2 = only the small non-GEMM kernels (LayerNorm, attention, embed / select), whose CTAs can co-reside with
// the tail of the GEMM before them (a GEMM CTA holds 231 KB of shared memory: two GEMMs can never overlap on an SM).
int pdl_mode();
inline bool pdl_enabled() { return pdl_mode() == 1; }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                            cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// GEMM launches (gemm_host.cu)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t stream, Args... args) {
  return launch_kernel_pdl(pdl_mode() == 1, kernel, grid, block, smem, stream, args...);
}
// the small kernels between the GEMMs
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_small_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                              cudaStream_t stream, Args... args) {
  return launch_kernel_pdl(pdl_mode() >= 1, kernel, grid, block, smem, stream, args...);
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
  // nn.GELU() exact (erf) form (coca_model.py:65): 0.5 x (1 + erf(x / sqrt(2))).
  //
  // Evaluated as x * Phi(x) with Phi(-|x|) = 0.5 erfc(|x|/sqrt 2) = 2^Q(z): Q is a degree-7 minimax fit of
  // log2(erfc(z)) - 1 on z in [0, 4.2] (scripts of the fit: DESIGN.md section 3; max |gelu error| 5.6e-7, max
  // relative error 3.3e-6 -- three orders below the bf16 rounding of the stored result).  One MUFU (ex2) and
  // ten FMA-pipe instructions per element: the fc1 epilogue evaluates this 128 x 256 times per tile next to
  // an 8192-cycle MMA main loop, so the instruction count is what bounds that GEMM.
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.2f);
  float q = fmaf(-2.074870827e-05f, z, 4.926680704e-04f);
  q = fmaf(q, z, -5.263294093e-03f);
  q = fmaf(q, z, 3.403032571e-02f);
  q = fmaf(q, z, -1.526207328e-01f);
  q = fmaf(q, z, -9.169294834e-01f);
  q = fmaf(q, z, -1.628114700e+00f);
  q = fmaf(q, z, -9.999952912e-01f);
  float w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(q));
  return (x >= 0.f) ? fmaf(-x, w, x) : x * w;
}

// ---- packed-pair fp32 (Blackwell FFMA2 / FMUL2): two elements per instruction
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_erf for two elements at once: same polynomial, the Horner chain runs on the packed-pair pipe
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const float z0 = fminf(fabsf(x0) * 0.70710678118654752440f, 4.2f);
  const float z1 = fminf(fabsf(x1) * 0.70710678118654752440f, 4.2f);
  const uint64_t z = pack_f32x2(z0, z1);
  uint64_t q = fma_f32x2(pack_f32x2(-2.074870827e-05f, -2.074870827e-05f), z, pack_f32x2(4.926680704e-04f, 4.926680704e-04f));
  q = fma_f32x2(q, z, pack_f32x2(-5.263294093e-03f, -5.263294093e-03f));
  q = fma_f32x2(q, z, pack_f32x2(3.403032571e-02f, 3.403032571e-02f));
  q = fma_f32x2(q, z, pack_f32x2(-1.526207328e-01f, -1.526207328e-01f));
  q = fma_f32x2(q, z, pack_f32x2(-9.169294834e-01f, -9.169294834e-01f));
  q = fma_f32x2(q, z, pack_f32x2(-1.628114700e+00f, -1.628114700e+00f));
  q = fma_f32x2(q, z, pack_f32x2(-9.999952912e-01f, -9.999952912e-01f));
  float q0, q1, w0, w1;
  unpack_f32x2(q, q0, q1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(q1));
  x0 = (x0 >= 0.f) ? fmaf(-x0, w0, x0) : x0 * w0;
  x1 = (x1 >= 0.f) ? fmaf(-x1, w1, x1) : x1 * w1;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ec
