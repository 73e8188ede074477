// tcgen05 / TMEM / TMA GEMM for sm_100a -- see gemm_sm100.cuh for the contract.
#include "gemm_sm100.cuh"
#include <cuda.h>
#include <mutex>
#include <vector>
#include <algorithm>

namespace ec {

// ============================================================================================
// PTX wrappers (sm_100a)
// ============================================================================================
namespace ptx {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx

// ============================================================================================
// Kernel configuration
// ============================================================================================
constexpr int kBM = 128;            // UMMA M (one CTA, cta_group::1)
constexpr int kBK = 64;             // 64 bf16 = 128 B = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;        // two warps per TMEM lane quadrant (column halves)
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = (kFirstEpiWarp + kEpiWarps) * 32;   // 384
constexpr long long kWatchdogCycles = 4000000000LL;          // ~2 s at 2 GHz

template <int BN> struct TileCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;                   // two accumulator stages
  static constexpr int kBarBytes = 1024;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024 /*align slack*/;
};

__device__ int g_gemm_watchdog = 0;

__device__ __forceinline__ bool bounded_wait(uint32_t bar, uint32_t parity, int code) {
  if (ptx::mbar_try_wait(bar, parity)) return true;
  long long t0 = clock64();
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) {
      atomicExch(&g_gemm_watchdog, code);
      return false;
    }
  }
  return true;
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);       // start address      bits [0,14)
  d |= (uint64_t)1 << 16;                             // LBO (unused for swizzled K-major) bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                   // SBO = 1024 B       bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B
  return d;
}

template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4)                       // C format F32
         | (1u << 7)                     // A format BF16
         | (1u << 10)                    // B format BF16
         | (0u << 15) | (0u << 16)       // A, B K-major
         | ((uint32_t)(BN >> 3) << 17)   // N >> 3
         | ((uint32_t)(kBM >> 4) << 24); // M >> 4
}

struct EpiParams {
  int M, N;
  const float* bias;
  const float* residual; int ldr;
  void* out; int ldo;
  float* logits; long long ld_logits; long long logits_step_stride;
  float* part_val; int* part_idx; int n_parts;
  int eos_col; const int* cur_len; int min_len;
  const int* skip_flag;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    EpiParams p, int K, int tiles_m, int tiles_n) {
  using Cfg = TileCfg<BN>;
  if (p.skip_flag != nullptr && *p.skip_flag != 0) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + Cfg::kStages * Cfg::kABytes;
  const uint32_t bar0 = smem_base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!bounded_wait(empty_bar(stage), phase ^ 1u, 1)) goto done;
          ptx::mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          ptx::tma_load_2d(smem_a0 + stage * Cfg::kABytes, &tmap_a, kb * kBK, m_blk * kBM, full_bar(stage));
          ptx::tma_load_2d(smem_b0 + stage * Cfg::kBBytes, &tmap_w, kb * kBK, n_blk * BN, full_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        if (!bounded_wait(tempty_bar(acc), acc_phase ^ 1u, 2)) goto done;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!bounded_wait(full_bar(stage), phase, 3)) goto done;
          ptx::tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_a0 + stage * Cfg::kABytes);
          const uint64_t bdesc = make_smem_desc(smem_b0 + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            ptx::umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));          // frees the smem stage when the MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(tfull_bar(acc));              // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================================================================== epilogue warps
    const int e = warp - kFirstEpiWarp;
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int half = e >> 2;                  // column half of the tile
    constexpr int kColsPerWarp = BN / 2;
    constexpr int kChunks = kColsPerWarp / 32;
    int acc = 0; uint32_t acc_phase = 0;
    int cur_len = 0;
    if (EPI == EPI_ARGMAX) cur_len = (p.cur_len != nullptr) ? *p.cur_len : (1 << 30);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      if (!bounded_wait(tfull_bar(acc), acc_phase, 4)) break;
      ptx::tc_fence_after();
      const int grow = m_blk * kBM + quad * 32 + lane;
      const bool row_ok = grow < p.M;
      float best_v = -INFINITY; int best_i = 0;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int col_in_tile = half * kColsPerWarp + c * 32;
        const int gcol0 = n_blk * BN + col_in_tile;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + col_in_tile), r);
        ptx::tmem_ld_wait();
        if (gcol0 >= p.N) continue;           // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (gcol0 + j < p.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
        }
        if (EPI == EPI_BF16 || EPI == EPI_BF16_GELU) {
          if (EPI == EPI_BF16_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          if (row_ok) {
            bf16* orow = reinterpret_cast<bf16*>(p.out) + (size_t)grow * p.ldo + gcol0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (gcol0 + j < p.N) {
                uint4 q;
                q.x = pack_bf16x2(v[j], v[j + 1]); q.y = pack_bf16x2(v[j + 2], v[j + 3]);
                q.z = pack_bf16x2(v[j + 4], v[j + 5]); q.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(orow + j) = q;
              }
            }
          }
        } else if (EPI == EPI_F32 || EPI == EPI_F32_RESID) {
          if (row_ok) {
            float* orow = reinterpret_cast<float*>(p.out) + (size_t)grow * p.ldo + gcol0;
            const float* rrow = (EPI == EPI_F32_RESID) ? p.residual + (size_t)grow * p.ldr + gcol0 : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (gcol0 + j < p.N) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (EPI == EPI_F32_RESID) {
                  const float4 rr = *reinterpret_cast<const float4*>(rrow + j);
                  o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                }
                *reinterpret_cast<float4*>(orow + j) = o;
              }
            }
          }
        } else {  // EPI_ARGMAX
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = gcol0 + j;
              if (col < p.N) {
                float x = v[j];
                if (col == p.eos_col && cur_len < p.min_len) x = -INFINITY;
                v[j] = x;
                if (x > best_v) { best_v = x; best_i = col; }
              }
            }
            if (p.logits != nullptr) {
              float* lrow = p.logits + (p.cur_len != nullptr ? (size_t)(cur_len - 1) * p.logits_step_stride : 0) +
                            (size_t)grow * p.ld_logits + gcol0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (gcol0 + j < p.N)
                  *reinterpret_cast<float4*>(lrow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
            }
          }
        }
      }
      if (EPI == EPI_ARGMAX && row_ok) {
        const size_t slot = (size_t)grow * p.n_parts + (size_t)(n_blk * 2 + half);
        p.part_val[slot] = best_v;
        p.part_idx[slot] = best_i;
      }
      // release the accumulator stage back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ============================================================================================
// Host side: tensor maps + launch
// ============================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128B swizzle
static int make_tmap(CUtensorMap* out, const bf16* base, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  EC_REQUIRE(fn != nullptr, EC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  EC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0, EC_ERR_INVALID, "GEMM operand not 16-byte aligned");
  EC_REQUIRE(ld % 8 == 0, EC_ERR_INVALID, "GEMM operand row stride %d not a multiple of 8 elements", ld);
  // descriptors are pure functions of (base, shape, stride, box): memoise them, encoding costs microseconds
  struct Key { const void* p; int r, c, l, b; };
  struct Entry { Key k; CUtensorMap m; };
  static std::mutex mu;
  static std::vector<Entry> cache[256];
  const size_t hsh = ((reinterpret_cast<uintptr_t>(base) >> 8) ^ (size_t)rows * 131u ^ (size_t)box_rows) & 255u;
  {
    std::lock_guard<std::mutex> lk(mu);
    for (const Entry& e : cache[hsh])
      if (e.k.p == base && e.k.r == rows && e.k.c == cols && e.k.l == ld && e.k.b == box_rows) {
        *out = e.m;
        return EC_OK;
      }
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EC_REQUIRE(r == CUDA_SUCCESS, EC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache[hsh].size() < 64) cache[hsh].push_back(Entry{Key{base, rows, cols, ld, box_rows}, *out});
  }
  return EC_OK;
}

int gemm_argmax_parts(int N) { return ceil_div(N, 256) * 2; }

template <int BN, int EPI>
static int launch_tile(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  CUtensorMap ta, tw;
  EC_TRY(make_tmap(&ta, a.A, a.M, a.K, a.lda, kBM));
  EC_TRY(make_tmap(&tw, a.W, a.N, a.K, a.ldw, BN));
  EpiParams p;
  p.M = a.M; p.N = a.N; p.bias = a.bias; p.residual = a.residual; p.ldr = a.ldr; p.out = a.out; p.ldo = a.ldo;
  p.logits = a.logits; p.ld_logits = a.ld_logits; p.logits_step_stride = a.logits_step_stride; p.part_val = a.part_val; p.part_idx = a.part_idx;
  p.n_parts = a.n_parts; p.eos_col = a.eos_col; p.cur_len = a.cur_len; p.min_len = a.min_len;
  p.skip_flag = a.skip_flag;
  const int tiles_m = ceil_div(a.M, kBM), tiles_n = ceil_div(a.N, BN);
  const int grid = std::min(tiles_m * tiles_n, num_sms());
  static bool attr_set = false;   // per template instantiation
  if (!attr_set) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_set = true;
  }
  gemm_tcgen05_kernel<BN, EPI><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(ta, tw, p, a.K, tiles_m, tiles_n);
  EC_POST_LAUNCH();
  return EC_OK;
}

template <int EPI>
static int launch_epi(const GemmArgs& a, cudaStream_t stream) {
  if constexpr (EPI == EPI_ARGMAX) {
    return launch_tile<256, EPI>(a, stream);
  } else {
  const int sms = num_sms();
  const int tm = ceil_div(a.M, kBM);
  if (tm * ceil_div(a.N, 256) >= sms) return launch_tile<256, EPI>(a, stream);
  if (tm * ceil_div(a.N, 128) >= sms / 2 || a.N >= 1024) return launch_tile<128, EPI>(a, stream);
  return launch_tile<64, EPI>(a, stream);
  }
}

// ---- optional per-launch timing (CUDA events on the launching stream) for the roofline line of bench.py
struct ProfSlot { cudaEvent_t beg, end; double flops; };
static bool g_prof_on = false;
static std::vector<ProfSlot> g_prof_slots;
static size_t g_prof_used = 0;

void gemm_profile_enable(bool on) { g_prof_on = on; g_prof_used = 0; }

int gemm_profile_read(double* ms, double* flops, long long* launches) {
  double tms = 0.0, tf = 0.0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    EC_CHECK_CUDA(cudaEventSynchronize(g_prof_slots[i].end));
    float t = 0.f;
    EC_CHECK_CUDA(cudaEventElapsedTime(&t, g_prof_slots[i].beg, g_prof_slots[i].end));
    tms += t;
    tf += g_prof_slots[i].flops;
  }
  if (ms) *ms = tms;
  if (flops) *flops = tf;
  if (launches) *launches = (long long)g_prof_used;
  g_prof_used = 0;
  return EC_OK;
}

static int gemm_dispatch(const GemmArgs& a, GemmEpilogue epi, cudaStream_t stream);

int gemm_bf16(const GemmArgs& a, GemmEpilogue epi, cudaStream_t stream) {
  if (!g_prof_on) return gemm_dispatch(a, epi, stream);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
    return gemm_dispatch(a, epi, stream);
  if (g_prof_used == g_prof_slots.size()) {
    ProfSlot s{};
    EC_CHECK_CUDA(cudaEventCreate(&s.beg));
    EC_CHECK_CUDA(cudaEventCreate(&s.end));
    g_prof_slots.push_back(s);
  }
  ProfSlot& s = g_prof_slots[g_prof_used];
  EC_CHECK_CUDA(cudaEventRecord(s.beg, stream));
  const int rc = gemm_dispatch(a, epi, stream);
  EC_CHECK_CUDA(cudaEventRecord(s.end, stream));
  if (rc == EC_OK) { s.flops = 2.0 * a.M * (double)a.N * a.K; ++g_prof_used; }
  return rc;
}

static int gemm_dispatch(const GemmArgs& a, GemmEpilogue epi, cudaStream_t stream) {
  EC_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, EC_ERR_INVALID, "GEMM with empty dimension M=%d N=%d K=%d", a.M, a.N, a.K);
  // K may be ragged (TMA zero-fills the box beyond K); only row strides need 16-byte alignment
  EC_REQUIRE(a.N % 8 == 0, EC_ERR_UNSUPPORTED, "GEMM needs N to be a multiple of 8 (N=%d)", a.N);
  EC_REQUIRE(a.lda >= a.K && a.ldw >= a.K, EC_ERR_INVALID, "GEMM row strides shorter than K");
  EC_REQUIRE(a.A && a.W, EC_ERR_INVALID, "GEMM null operand");
  if (epi != EPI_ARGMAX) {
    EC_REQUIRE(a.out != nullptr && a.ldo % 8 == 0, EC_ERR_INVALID, "GEMM output null or ldo %% 8 != 0");
    EC_REQUIRE((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0, EC_ERR_INVALID, "GEMM output not 16-byte aligned");
  } else {
    EC_REQUIRE(a.part_val && a.part_idx && a.n_parts == gemm_argmax_parts(a.N), EC_ERR_INVALID,
               "arg-max GEMM needs part buffers with n_parts=%d", gemm_argmax_parts(a.N));
    EC_REQUIRE(a.logits == nullptr || a.ld_logits % 4 == 0, EC_ERR_INVALID, "logits row stride must be a multiple of 4");
  }
  if (epi == EPI_F32_RESID)
    EC_REQUIRE(a.residual != nullptr && a.ldr % 4 == 0, EC_ERR_INVALID, "residual epilogue needs a residual pointer");
  if (a.bias) EC_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15u) == 0, EC_ERR_INVALID, "bias not 16-byte aligned");
  switch (epi) {
    case EPI_BF16: return launch_epi<EPI_BF16>(a, stream);
    case EPI_BF16_GELU: return launch_epi<EPI_BF16_GELU>(a, stream);
    case EPI_F32: return launch_epi<EPI_F32>(a, stream);
    case EPI_F32_RESID: return launch_epi<EPI_F32_RESID>(a, stream);
    case EPI_ARGMAX: return launch_epi<EPI_ARGMAX>(a, stream);
  }
  set_error("unknown GEMM epilogue %d", (int)epi);
  return EC_ERR_INVALID;
}

int gemm_check_watchdog(cudaStream_t stream) {
  EC_CHECK_CUDA(cudaStreamSynchronize(stream));
  int v = 0;
  EC_CHECK_CUDA(cudaMemcpyFromSymbol(&v, g_gemm_watchdog, sizeof(int)));
  if (v != 0) {
    int zero = 0;
    cudaMemcpyToSymbol(g_gemm_watchdog, &zero, sizeof(int));
    set_error("tcgen05 GEMM pipeline wait timed out (role code %d)", v);
    return EC_ERR_TIMEOUT;
  }
  return EC_OK;
}

// ============================================================================================
// SIMT cross-check kernel (tests only)
// ============================================================================================
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw, int M, int N,
                                 int K, EpiParams p, int epi) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= N || row >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)row * lda + k]) * __bfloat162float(W[(size_t)col * ldw + k]);
  if (p.bias) acc += p.bias[col];
  if (epi == EPI_BF16) reinterpret_cast<bf16*>(p.out)[(size_t)row * p.ldo + col] = __float2bfloat16_rn(acc);
  else if (epi == EPI_BF16_GELU) reinterpret_cast<bf16*>(p.out)[(size_t)row * p.ldo + col] = __float2bfloat16_rn(gelu_erf(acc));
  else if (epi == EPI_F32) reinterpret_cast<float*>(p.out)[(size_t)row * p.ldo + col] = acc;
  else if (epi == EPI_F32_RESID)
    reinterpret_cast<float*>(p.out)[(size_t)row * p.ldo + col] = acc + p.residual[(size_t)row * p.ldr + col];
  else if (p.logits) p.logits[(size_t)row * p.ld_logits + col] = acc;
}

int gemm_reference_simt(const GemmArgs& a, GemmEpilogue epi, cudaStream_t stream) {
  EpiParams p;
  p.M = a.M; p.N = a.N; p.bias = a.bias; p.residual = a.residual; p.ldr = a.ldr; p.out = a.out; p.ldo = a.ldo;
  p.logits = a.logits; p.ld_logits = a.ld_logits; p.logits_step_stride = a.logits_step_stride; p.part_val = a.part_val; p.part_idx = a.part_idx;
  p.n_parts = a.n_parts; p.eos_col = a.eos_col; p.cur_len = a.cur_len; p.min_len = a.min_len; p.skip_flag = nullptr;
  dim3 grid(ceil_div(a.N, 128), a.M);
  gemm_simt_kernel<<<grid, 128, 0, stream>>>(a.A, a.lda, a.W, a.ldw, a.M, a.N, a.K, p, (int)epi);
  EC_POST_LAUNCH();
  return EC_OK;
}

}  // namespace ec
