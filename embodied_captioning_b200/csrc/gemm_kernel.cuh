// Device side of the tcgen05 / TMEM / TMA GEMMs (sm_100a).  See gemm_sm100.cuh for the contract.
//
// Two kernels share the epilogue code:
//   gemm_tcgen05_pair_kernel  CTA pairs (cluster of 2), tcgen05.mma.cta_group::2, UMMA 256 x 256 x 16: the large GEMMs
//   gemm_tcgen05_kernel       one CTA per tile, cta_group::1, UMMA 128 x BN x 16: small-M GEMMs (decoder steps, pooler)
// Warp roles in both:
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx); in the 1-CTA kernel it
//                 also initialises the barriers and issues the first ring of loads before the CTA-wide sync
//   warp 1      : MMA issuer     (warp-uniform loop, one elected lane issues tcgen05.mma / tcgen05.commit)
//   warp 2      : TMEM allocator; then L2 prefetch of the next GEMM's weights (decoder steps)
//   warp 3      : 1-CTA kernel, single-tile launches: second MMA issuer (odd k-blocks, second TMEM stage); else idle
//   warps 4..   : 8 or 16 epilogue warps (tcgen05.ld 32x32b -> bias / GELU / residual / arg-max in registers ->
//                 per-warp swizzled shared-memory transpose -> coalesced global row segments)
//   warp 12     : 1-CTA kernel only: second TMA producer (odd k-blocks)
// Persistent CTAs (1 per SM) walk the tile list; two TMEM accumulator stages overlap epilogue(i) with mainloop(i+1).
#pragma once
#include "gemm_sm100.cuh"
#include <cuda.h>

namespace ec {

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
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// the same load with an L2 eviction policy (createpolicy)
__device__ __forceinline__ void tma_load_2d_hint(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
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
// ---- cta_group::2 (CTA-pair) forms
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // relaxed: the only data this arrival publishes is "my tcgen05.ld of the accumulator stage has completed", which
  // tcgen05.wait::ld + tcgen05.fence::before_thread_sync order already.  The release.cluster form costs an ERRBAR
  // that waits for every earlier global store of the tile to drain (7 % of the stall samples of the fc1 GEMM).
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier of the pair's leader CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[BN rows: BN/2 per CTA]; issued by the leader CTA only
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs retired) on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

// one lane of a converged warp (the instruction must be reached by all 32 lanes)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

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
constexpr int kSecondProducerWarp = kFirstEpiWarp + kEpiWarps;   // 1-CTA kernel only: warp 12 loads the odd k-blocks
constexpr int kThreads1 = kThreads + 32;                          // 416
constexpr int kEpiStageBytes = 32 * 32 * 4;                  // one 32x32 fp32 chunk per epilogue warp
constexpr long long kWatchdogCycles = 4000000000LL;          // ~2 s at 2 GHz
#ifdef EC_NO_DUAL_ISSUE
constexpr bool kDualIssue = false;
#else
constexpr bool kDualIssue = true;
#endif

template <int BN> struct TileCfg {
  static constexpr int kStages = (BN >= 192) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (BN == 192) ? 512 : 2 * BN;   // two accumulator stages (allocations are powers of 2)
  static constexpr int kBarBytes = 1024;
  // BN = 192 (the decoder's FC1 + GELU): four more staging buffers for the warps that help with the epilogue
  static constexpr int kHelpWarps = (BN == 192) ? 4 : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + (kEpiWarps + kHelpWarps) * kEpiStageBytes + kBarBytes + 1024 /*align*/;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

__device__ int g_gemm_watchdog = 0;

// Stage clocks of CTA 0 of the 1-CTA kernel (-DEC_GEMM_TIMING builds; scripts/gemm_timing_probe.py): a ring of records
// {N, K, BN * 8 + EPI, -, cycles since kernel entry at: skip word loaded, prologue done, first operands landed, last
// MMA committed, accumulator ready, tile stored}.
constexpr int kGemmTimingSlots = 8192;
__device__ long long g_gemm_timing[kGemmTimingSlots][12];
__device__ unsigned int g_gemm_timing_next = 0;
#ifdef EC_GEMM_TIMING
#define GEMM_T(ev) do { if (blockIdx.x == 0 && lane == 0) g_gemm_timing[t_slot][4 + (ev)] = clock64() - t_entry; } while (0)
#else
#define GEMM_T(ev) do { } while (0)
#endif

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
  // fused LayerNorm tail (EPI_F32_RESID only, ln_out != nullptr): see ln_tail()
  const float* ln_g; const float* ln_b; bf16* ln_out; int ld_ln; int* ln_cnt; int ln_tiles_n;
  // L2 prefetch of the NEXT GEMM's weights (decoder steps: weights cycle through L2 once per step, so every K panel
  // would otherwise arrive at HBM latency); issued by warp 2 once it has allocated the tensor memory
  const unsigned char* pf_ptr; unsigned long long pf_bytes;
  int grp_cols;     // grouped GEMM: output columns per group (0 = plain GEMM); see GemmArgs::group_cols
  // Deterministic split-K (EPI_F32_RESID, one tile per CTA): the K range of every output tile is cut into `split_k`
  // slices run by adjacent CTAs.  Slices 1.. store their raw fp32 accumulators in sk_ws[slice - 1][M][N] and count
  // themselves on sk_cnt[tile]; slice 0 waits for them and adds them IN SLICE ORDER in its residual epilogue.
  int split_k; float* sk_ws; int* sk_cnt;
  // 1-CTA kernel, opt-in: W tiles loaded with the L2 evict-first policy (the 557 MB of decoder weights are a stream
  // that flushes the step's small hot set -- kernel parameters, tensor maps, biases, activations, code -- out of the
  // 126 MB L2).  The same hint on the image-token and K / V-cache streams pays (cross_absorbed.cu, decode.cu); on the
  // weights it measured +0.4 ms per decode.
  int w_evict_first;
};

__device__ __forceinline__ float4 ld_cg_f4(const float* p);

// One 32 x 32 chunk.  kFull = the whole chunk lies inside [M, N]: no per-element bound predicates (the interior
// of the matrix, i.e. nearly every chunk; they were a third of the epilogue's instructions).
template <int BN, int EPI, bool kFull>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, uint32_t taddr, int row0, int gcol0, int lane,
                                               uint32_t stg, int cur_len, float& best_v, int& best_i,
                                               uint32_t acc2_off = 0, int sk_slice = 0) {
  const int cc = lane & 7;
  const int gcol = gcol0 + cc * 4;
  // split-K: slice > 0 stores raw accumulators (no bias, no residual); slice 0 adds the other slices' partials
  const bool sk_raw = (EPI == EPI_F32_RESID) && sk_slice > 0;
  const int sk_parts = (EPI == EPI_F32_RESID && sk_slice == 0 && p.split_k > 1) ? p.split_k - 1 : 0;
  // The residual may alias the output (in-place residual stream): every element is read and written by the same
  // thread.  All eight loads of the chunk are issued before the TMEM load and the transpose so that their latency
  // is hidden (prefetching a whole chunk further ahead measured no better: 822 vs 842 TFLOP/s on out_proj).
  // (Requesting them even earlier -- while the K loop runs -- measured WORSE: the K loop is bound by the SM's L2 ingest
  // and the extra 32 KB per CTA slowed it from 1.7 to 2.1 us, profiles/r02_gemm_timing.txt.)
  float4 rsd[8];
  if (EPI == EPI_F32_RESID) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int grow = row0 + it * 4 + (lane >> 3);
      rsd[it] = (!sk_raw && (kFull || (grow < p.M && gcol < p.N)))
                    ? *reinterpret_cast<const float4*>(p.residual + (size_t)grow * p.ldr + gcol)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int s = 0; s < sk_parts; ++s) {                 // fixed order: bit-reproducible
      const float* part = p.sk_ws + (size_t)s * p.M * p.N;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int grow = row0 + it * 4 + (lane >> 3);
        if (kFull || (grow < p.M && gcol < p.N)) {
          const float4 t = ld_cg_f4(part + (size_t)grow * p.N + gcol);
          rsd[it].x += t.x; rsd[it].y += t.y; rsd[it].z += t.z; rsd[it].w += t.w;
        }
      }
    }
  }
  uint32_t r[32];
  ptx::tmem_ld_32x32(taddr, r);
  ptx::tmem_ld_wait();
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (acc2_off != 0) {      // dual-issuer mode: the odd k-blocks were accumulated in the second TMEM stage
    ptx::tmem_ld_32x32(taddr + acc2_off, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
  }
  if (p.bias != nullptr && !sk_raw) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (kFull || gcol0 + j < p.N) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
  }
  if (EPI == EPI_BF16_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
  }
  if (EPI == EPI_ARGMAX) {
    if (kFull || row0 + lane < p.M) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = gcol0 + j;
        if (kFull || col < p.N) {
          float x = v[j];
          if (col == p.eos_col && cur_len < p.min_len) x = -INFINITY;
          v[j] = x;
          if (x > best_v) { best_v = x; best_i = col; }
        }
      }
    }
    if (p.logits == nullptr) return;      // warp-uniform: nothing to store
  }
  if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_GELU) {
    // ---- bf16 output: pack in registers, transpose 32 rows x 64 B through shared memory (16-byte unit u of row r
    //      lives at unit u ^ ((r >> 1) & 3): conflict-free both ways), then 4 lanes x 16 B = one 64-byte row segment
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t a0 = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]), a1 = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
      const uint32_t a2 = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]), a3 = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
      ptx::st_shared_v4(stg + (uint32_t)lane * 64u + (uint32_t)((g ^ ((lane >> 1) & 3)) * 16), __uint_as_float(a0),
                        __uint_as_float(a1), __uint_as_float(a2), __uint_as_float(a3));
    }
    __syncwarp();
    const int cu = lane & 3;
    const int gc = gcol0 + cu * 8;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + (lane >> 2);
      const int grow = row0 + rr;
      const float4 o = ptx::ld_shared_v4(stg + (uint32_t)rr * 64u + (uint32_t)((cu ^ ((rr >> 1) & 3)) * 16));
      if (kFull || (grow < p.M && gc < p.N))
        *reinterpret_cast<float4*>(reinterpret_cast<bf16*>(p.out) + (size_t)grow * p.ldo + gc) = o;
    }
    __syncwarp();                           // staging buffer is reused by the next chunk
  } else {
    // ---- fp32 output: transpose through shared memory: lane = row  ->  8 lanes x 16 B = one 128-byte row segment
    //      (16-byte unit u of row r lives at unit u ^ (r & 7): conflict-free in both directions)
#pragma unroll
    for (int g = 0; g < 8; ++g)
      ptx::st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) * 16), v[4 * g], v[4 * g + 1],
                        v[4 * g + 2], v[4 * g + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      const int grow = row0 + rr;
      float4 o = ptx::ld_shared_v4(stg + (uint32_t)rr * 128u + (uint32_t)((cc ^ (rr & 7)) * 16));
      if (kFull || (grow < p.M && gcol < p.N)) {
        if (EPI == EPI_F32 || EPI == EPI_F32_RESID) {
          if (EPI == EPI_F32_RESID) {
            o.x += rsd[it].x; o.y += rsd[it].y; o.z += rsd[it].z; o.w += rsd[it].w;
          }
          if (sk_raw)
            *reinterpret_cast<float4*>(p.sk_ws + (size_t)(sk_slice - 1) * p.M * p.N + (size_t)grow * p.N + gcol) = o;
          else
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)grow * p.ldo + gcol) = o;
        } else {  // EPI_ARGMAX with logits requested
          float* lbase = p.logits + (p.cur_len != nullptr ? (size_t)(cur_len - 1) * p.logits_step_stride : 0);
          *reinterpret_cast<float4*>(lbase + (size_t)grow * p.ld_logits + gcol) = o;
        }
      }
    }
    __syncwarp();                           // staging buffer is reused by the next chunk
  }
}

// Epilogue of one 128 x BN accumulator tile for one epilogue warp: its 32-row TMEM lane quadrant (`quad`), the
// column half `half`, 32 columns at a time: tcgen05.ld -> bias / GELU / arg-max in registers -> swizzled
// shared-memory transpose -> coalesced global row segments.  `row0` = first global row of the slab.
template <int BN, int EPI, int PARTS = 2>
__device__ __forceinline__ void epilogue_tile(const EpiParams& p, uint32_t tmem_acc, int row0, int n_blk, int quad,
                                              int half, int lane, uint32_t stg, int cur_len, uint32_t acc2_off = 0,
                                              int sk_slice = 0) {
  static_assert(EPI != EPI_ARGMAX || PARTS == 2, "arg-max partial slots are laid out for two column halves");
  constexpr int kColsPerWarp = BN / PARTS;
  constexpr int kChunks = kColsPerWarp / 32;
  float best_v = -INFINITY; int best_i = 0;
  const bool rows_full = row0 + 32 <= p.M;
#pragma unroll 1
  for (int c = 0; c < kChunks; ++c) {
    const int col_in_tile = half * kColsPerWarp + c * 32;
    const int gcol0 = n_blk * BN + col_in_tile;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)col_in_tile;
    if (gcol0 >= p.N) continue;           // warp-uniform: nothing of this chunk is inside the matrix
    if (rows_full && gcol0 + 32 <= p.N)
      epilogue_chunk<BN, EPI, true>(p, taddr, row0, gcol0, lane, stg, cur_len, best_v, best_i, acc2_off, sk_slice);
    else
      epilogue_chunk<BN, EPI, false>(p, taddr, row0, gcol0, lane, stg, cur_len, best_v, best_i, acc2_off, sk_slice);
  }
  if (EPI == EPI_ARGMAX && row0 + lane < p.M) {
    const size_t slot = (size_t)(row0 + lane) * p.n_parts + (size_t)(n_blk * 2 + half);
    p.part_val[slot] = best_v;
    p.part_idx[slot] = best_i;
  }
}

// The same for an explicit set of 32-column chunks of the tile (bit c of `mask` = columns [32 c, 32 c + 32)): the GELU
// epilogue of a lone 128 x 192 tile shared by THREE warps per TMEM lane quadrant (see gemm_tcgen05_kernel).
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile_chunks(const EpiParams& p, uint32_t tmem_acc, int row0, int n_blk, int quad,
                                                     uint32_t mask, int lane, uint32_t stg, uint32_t acc2_off) {
  static_assert(EPI == EPI_BF16 || EPI == EPI_BF16_GELU, "chunk-wise epilogue: bf16 outputs only");
  float best_v = 0.f; int best_i = 0;
  const bool rows_full = row0 + 32 <= p.M;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    if (((mask >> c) & 1u) == 0u) continue;
    const int col_in_tile = c * 32;
    const int gcol0 = n_blk * BN + col_in_tile;
    const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)col_in_tile;
    if (gcol0 >= p.N) continue;
    if (rows_full && gcol0 + 32 <= p.N)
      epilogue_chunk<BN, EPI, true>(p, taddr, row0, gcol0, lane, stg, 0, best_v, best_i, acc2_off, 0);
    else
      epilogue_chunk<BN, EPI, false>(p, taddr, row0, gcol0, lane, stg, 0, best_v, best_i, acc2_off, 0);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused LayerNorm tail of the residual GEMMs.  The LayerNorm that follows out_proj / c_proj needs whole rows, which
// are spread over `tiles_n` CTAs.  Every CTA counts its finished tile on a per-128-row counter (stores -> fence ->
// atomicAdd); the CTA whose increment completes the row block re-reads those 128 rows (L2-hot: they were written
// microseconds ago), normalises them and writes the bf16 operand of the next GEMM.  That removes the separate
// LayerNorm launch (one of every seven kernels of a decoder step) and its cold HBM read of the residual stream
// in the image tower.  Called by all `n_warps` epilogue warps of the CTA with `bar_id` a free named barrier.
constexpr int kLnTailMaxVec = 8;     // row width <= 1024

__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void ln_tail(const EpiParams& p, int row_base, int e, int n_warps, int lane, int bar_id,
                                        volatile int* s_flag) {
  // all epilogue warps of this CTA have stored their part of the tile
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(n_warps * 32) : "memory");
  if (e == 0 && lane == 0) {
    __threadfence();
    const int old = atomicAdd(p.ln_cnt + row_base / kBM, 1);
    const int last = (old == p.ln_tiles_n - 1) ? 1 : 0;
    if (last) p.ln_cnt[row_base / kBM] = 0;          // re-armed for the next launch
    *s_flag = last;
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(n_warps * 32) : "memory");
  if (*s_flag == 0) return;
  __threadfence();
  const int width = p.N, nvec = width >> 2;
  for (int r = e; r < kBM; r += n_warps) {
    const int row = row_base + r;
    if (row >= p.M) break;
    const float* xr = reinterpret_cast<const float*>(p.out) + (size_t)row * p.ldo;
    float4 v[kLnTailMaxVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnTailMaxVec; ++i) {
      const int c = lane + i * 32;
      v[i] = (c < nvec) ? ld_cg_f4(xr + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)width;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnTailMaxVec; ++i) {
      if (lane + i * 32 < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)width + 1e-5f);      // same arithmetic as layernorm.cu
    bf16* yr = p.ln_out + (size_t)row * p.ld_ln;
#pragma unroll
    for (int i = 0; i < kLnTailMaxVec; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g) + c), b = __ldg(reinterpret_cast<const float4*>(p.ln_b) + c);
        uint2 pk;
        pk.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
        pk.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        reinterpret_cast<uint2*>(yr)[c] = pk;
      }
    }
  }
}

// Same-CTA form of the tail (CTA-pair kernel): the tile schedule gives each cluster WHOLE row blocks (all n-tiles of a
// 256-row block, consecutively), so the CTA that finishes the last n-tile wrote every column of its 128 rows itself:
// no counter, no fence, no atomic -- one named barrier among the epilogue warps, then the stand-alone LayerNorm
// arithmetic over rows that are still L2-hot, two rows per warp iteration so that the loads of a pair overlap.
__device__ __forceinline__ void ln_tail_local(const EpiParams& p, int row_base, int e, int n_warps, int lane, int bar_id) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(n_warps * 32) : "memory");
  const int width = p.N, nvec = width >> 2;
  for (int r = 2 * e; r < kBM; r += 2 * n_warps) {
    const int row = row_base + r;
    if (row >= p.M) break;
    const bool two = (row + 1 < p.M);
    const float* x0 = reinterpret_cast<const float*>(p.out) + (size_t)row * p.ldo;
    const float* x1 = x0 + p.ldo;
    float4 v0[kLnTailMaxVec], v1[kLnTailMaxVec];
#pragma unroll
    for (int i = 0; i < kLnTailMaxVec; ++i) {
      const int c = lane + i * 32;
      v0[i] = (c < nvec) ? ld_cg_f4(x0 + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      v1[i] = (c < nvec && two) ? ld_cg_f4(x1 + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float4 (&v)[kLnTailMaxVec] = h ? v1 : v0;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kLnTailMaxVec; ++i)
        if (lane + i * 32 < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mean = warp_sum(s) / (float)width;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < kLnTailMaxVec; ++i) {
        if (lane + i * 32 < nvec) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
          q += (a * a + b * b) + (c * c + d * d);
        }
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)width + 1e-5f);      // same arithmetic as layernorm.cu
      bf16* yr = p.ln_out + (size_t)(row + h) * p.ld_ln;
#pragma unroll
      for (int i = 0; i < kLnTailMaxVec; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g) + c), b = __ldg(reinterpret_cast<const float4*>(p.ln_b) + c);
          uint2 pk;
          pk.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
          pk.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
          reinterpret_cast<uint2*>(yr)[c] = pk;
        }
      }
    }
  }
}

// this CTA's share of [ptr, ptr + bytes) -> L2, 4 KB per bulk-prefetch instruction
__device__ __forceinline__ void l2_prefetch_slice(const unsigned char* ptr, unsigned long long bytes, int lane) {
  constexpr unsigned long long kChunk = 4096;
  const unsigned long long n_chunks = (bytes + kChunk - 1) / kChunk;
  for (unsigned long long c = (unsigned long long)blockIdx.x * 32 + lane; c < n_chunks; c += (unsigned long long)gridDim.x * 32) {
    const unsigned long long off = c * kChunk;
    const unsigned int sz = (unsigned int)((bytes - off < kChunk ? bytes - off : kChunk) & ~15ull);
    if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr + off), "r"(sz) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Hand-over of rows between CTAs through counters in global memory (GEMM sequences, gemm_seq.cuh)
__device__ __forceinline__ int seq_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// release: everything this thread wrote -- and, through the named barrier before it, everything the CTA's epilogue
// warps wrote -- is visible at GPU scope before the count
__device__ __forceinline__ void seq_red_release(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all lanes poll the same word (one broadcast L2 request per iteration); ONE warp per CTA polls a given counter
__device__ __forceinline__ bool seq_wait_counter(const int* cnt, int target) {
  if (seq_ld_acquire(cnt) >= target) return true;
  const long long t0 = clock64();
  while (seq_ld_acquire(cnt) < target) {
    if (clock64() - t0 > kWatchdogCycles) { atomicExch(&g_gemm_watchdog, 6); return false; }
  }
  return true;
}

// LayerNorm of ONE row by one warp (same arithmetic as layernorm.cu): the row (L2: written microseconds ago by the
// block's tiles) and gamma / beta (L2: prefetched at kernel entry) are requested together.
__device__ __forceinline__ void ln_row_warp(const float* xr, int width, const float* ln_g, const float* ln_b, bf16* yr,
                                            int lane) {
  const int nvec = width >> 2;
  float4 v[kLnTailMaxVec];
#pragma unroll
  for (int i = 0; i < kLnTailMaxVec; ++i) {
    const int c = lane + i * 32;
    v[i] = (c < nvec) ? ld_cg_f4(xr + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // gamma / beta: one 128-byte line per lane into L1 while the row is in flight (no registers held)
  if (lane * 32 < width) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ln_g + lane * 32));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ln_b + lane * 32));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnTailMaxVec; ++i)
    if (lane + i * 32 < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)width;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnTailMaxVec; ++i) {
    if (lane + i * 32 < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)width + 1e-5f);      // same arithmetic as layernorm.cu
#pragma unroll
  for (int i = 0; i < kLnTailMaxVec; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(ln_g) + c), b = __ldg(reinterpret_cast<const float4*>(ln_b) + c);
      uint2 pk;
      pk.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      pk.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      reinterpret_cast<uint2*>(yr)[c] = pk;
    }
  }
}
__device__ __forceinline__ void seq_ln_row(const EpiParams& p, int row, int lane) {
  ln_row_warp(reinterpret_cast<const float*>(p.out) + (size_t)row * p.ldo, p.N, p.ln_g, p.ln_b, p.ln_out + (size_t)row * p.ld_ln, lane);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads1, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    EpiParams p, int K, int tiles_m, int tiles_n) {
  using Cfg = TileCfg<BN>;
#ifdef EC_GEMM_TIMING
  const long long t_entry = clock64();
  __shared__ unsigned int s_t_slot;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    s_t_slot = atomicAdd(&g_gemm_timing_next, 1u) % kGemmTimingSlots;
    g_gemm_timing[s_t_slot][0] = p.N; g_gemm_timing[s_t_slot][1] = K; g_gemm_timing[s_t_slot][2] = BN * 8 + EPI;
    g_gemm_timing[s_t_slot][3] = p.M; g_gemm_timing[s_t_slot][10] = 0; g_gemm_timing[s_t_slot][11] = 0;
  }
#endif
  pdl_trigger_top();
  // The dependency wait comes first: without programmatic dependent launch (the default -- it measured slower inside
  // the step graphs) it is a no-op, and the skip flag's L2 round trip then overlaps the prologue instead of
  // sitting between it and the first TMA load of a 6 us kernel.
  pdl_wait();
  const int skip_word = (p.skip_flag != nullptr) ? *p.skip_flag : 0;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + Cfg::kStages * Cfg::kABytes;
  const uint32_t smem_epi = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar0 = smem_epi + kEpiWarps * kEpiStageBytes;
#ifdef EC_GEMM_TIMING
  long long t_skip = 0;
  if (skip_word >= 0) t_skip = clock64() - t_entry;      // after the skip word's round trip
#endif
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // split-K (host guarantees: EPI_F32_RESID, one tile per CTA, k-blocks divisible by split_k): tile index =
  // (m_blk * tiles_n + n_blk) * S + slice, so the slices of an output tile run on adjacent CTAs
  const int S = (EPI == EPI_F32_RESID && p.split_k > 1) ? p.split_k : 1;
  const int num_tiles = tiles_m * tiles_n * S;
  const int num_kb = ((K + kBK - 1) / kBK) / S;            // k-blocks per slice
  const int kb0 = ((int)blockIdx.x % S) * num_kb;          // first k-block of this CTA's slice (S > 1: one tile per CTA)

  // ---- prologue.  Warp 0 initialises the barriers and at once issues the loads of the first ring-full of
  // k-blocks: they need the barriers but neither the TMEM allocation nor the CTA-wide sync, and in a 6 us kernel
  // those ~0.3 us are worth having.
  // They do NOT wait for the skip word either (measured with -DEC_GEMM_TIMING: its round trip ends 0.6 us after kernel
  // entry -- the first parameter fetch of a launch is cold -- and the first operands only landed at 1.6-2.0 us): a
  // skipped launch drains them before it leaves.
  const int pre = ((int)blockIdx.x < num_tiles) ? (num_kb < Cfg::kStages ? num_kb : Cfg::kStages) : 0;
  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_a);
      ptx::prefetch_tmap(&tmap_w);
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
    __syncwarp();
    const int m_blk = ((int)blockIdx.x / S) / tiles_n, n_blk = ((int)blockIdx.x / S) % tiles_n;
    const int a_k0 = (p.grp_cols > 0) ? (n_blk * BN / p.grp_cols) * K : 0;     // grouped GEMM: this tile's column slab of A
    const uint64_t w_policy = p.w_evict_first ? ptx::policy_evict_first() : ptx::policy_evict_normal();
    for (int g = 0; g < pre; ++g) {               // warp-uniform loop, elected lane issues (see the MMA warp)
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(full_bar(g), Cfg::kStageBytes);
        ptx::tma_load_2d(smem_a0 + g * Cfg::kABytes, &tmap_a, a_k0 + (kb0 + g) * kBK, m_blk * kBM, full_bar(g));
        ptx::tma_load_2d_hint(smem_b0 + g * Cfg::kBBytes, &tmap_w, (kb0 + g) * kBK, n_blk * BN, full_bar(g), w_policy);
      }
      __syncwarp();
    }
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
#ifdef EC_GEMM_TIMING
  const unsigned int t_slot = s_t_slot;
  if (warp == 0 && blockIdx.x == 0 && lane == 0) g_gemm_timing[t_slot][4] = t_skip;
  if (warp == 0) GEMM_T(1);
#endif

  const bool skip = skip_word != 0;
  // Dual-issuer mode (one tile per CTA, i.e. the small-M GEMMs of the decoder steps).  Measured: the K loop of a
  // lone CTA costs ~0.3 us per k-block whatever the ring depth or the tile width (scripts/ingest_*_probe.py) -- the
  // time ONE thread needs to issue four tcgen05.mma + a commit -- so K = 3072 is 14 us however few rows there are.
  // Warp 3 becomes a second issuer: warp 1 takes the even k-blocks into accumulator stage 0, warp 3 the odd ones
  // into stage 1, and the epilogue adds the two stages (fixed assignment: bit-reproducible).
  const bool dual = (num_tiles <= (int)gridDim.x) && (num_kb >= 2) && kDualIssue;

  if (skip) {
    // the prologue's loads must have landed before the CTA may exit
    if (warp == 0)
      for (int g = 0; g < pre; ++g) bounded_wait(full_bar(g), 0u, 1);
  } else if (warp == 2) {
    if (p.pf_ptr != nullptr) l2_prefetch_slice(p.pf_ptr, p.pf_bytes, lane);
  } else if (warp == 3 || (warp == 1 && dual)) {
    if (dual && blockIdx.x < num_tiles) {
      const int w = (warp == 1) ? 0 : 1;
      constexpr uint32_t idesc = make_idesc<BN>();
      const uint64_t adesc0 = make_smem_desc(smem_a0), bdesc0 = make_smem_desc(smem_b0);
      const uint32_t tmem_d = tmem_base + (uint32_t)(w * BN);
      for (int kb = w; kb < num_kb; kb += 2) {
        const int stage = kb % Cfg::kStages;
        const uint32_t phase = (uint32_t)((kb / Cfg::kStages) & 1);
        if (!bounded_wait(full_bar(stage), phase, 3)) goto done;
        ptx::tc_fence_after();
        if (kb == 0) GEMM_T(2);
        const uint64_t adesc = adesc0 + (uint64_t)((stage * Cfg::kABytes) >> 4);
        const uint64_t bdesc = bdesc0 + (uint64_t)((stage * Cfg::kBBytes) >> 4);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            ptx::umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb > w || k > 0) ? 1u : 0u);
          ptx::umma_commit(empty_bar(stage));
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::umma_commit(tfull_bar(w));
      __syncwarp();
      if (w == 0) GEMM_T(3);
    }
  } else if (warp == 0 || warp == kSecondProducerWarp) {
    // ===================================================================== TMA producers (warp-uniform loops, one
    // elected lane issues).  Two warps, even / odd k-blocks: with no loads and no MMAs at all the ring handshake
    // alone cost 0.07 us per k-block (try_wait -> elect -> expect_tx -> issue -> syncwarp is one dependent chain
    // of shared-memory sync ops), and the loads doubled it (scripts/launch_floor_probe.py).  The first kStages
    // k-blocks find the ring empty and skip the wait.
    {
      const int w = (warp == 0) ? 0 : 1;
      const uint64_t w_policy = p.w_evict_first ? ptx::policy_evict_first() : ptx::policy_evict_normal();
      int g = 0;                                  // running k-block index of this CTA over all its tiles
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = (tile / S) / tiles_n, n_blk = (tile / S) % tiles_n;
        const int a_k0 = (p.grp_cols > 0) ? (n_blk * BN / p.grp_cols) * K : 0;
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          if ((g & 1) != w || g < pre) continue;      // the first `pre` k-blocks were issued in the prologue
          const int stage = g % Cfg::kStages;
          if (g >= Cfg::kStages) {
            const uint32_t phase = (uint32_t)((g / Cfg::kStages) & 1);
            if (!bounded_wait(empty_bar(stage), phase ^ 1u, 1)) goto done;
          }
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            ptx::tma_load_2d(smem_a0 + stage * Cfg::kABytes, &tmap_a, a_k0 + (kb0 + kb) * kBK, m_blk * kBM, full_bar(stage));
            ptx::tma_load_2d_hint(smem_b0 + stage * Cfg::kBBytes, &tmap_w, (kb0 + kb) * kBK, n_blk * BN, full_bar(stage), w_policy);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // The whole warp walks the loop (warp-uniform control flow and addresses, so the descriptors live in uniform
    // registers) and ONE elected lane issues the tcgen05 instructions.  With a single-lane branch around the loop
    // ptxas spent ~25 instructions per UMMA on election + register broadcasts, ~110 per k-block, and the tensor
    // pipe waited on this one warp: 85 % active for the plain epilogue, 67 % next to the GELU epilogue warps.
    {
      constexpr uint32_t idesc = make_idesc<BN>();
      const uint64_t adesc0 = make_smem_desc(smem_a0), bdesc0 = make_smem_desc(smem_b0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        if (!bounded_wait(tempty_bar(acc), acc_phase ^ 1u, 2)) goto done;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!bounded_wait(full_bar(stage), phase, 3)) goto done;
          ptx::tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)((stage * Cfg::kABytes) >> 4);
          const uint64_t bdesc = bdesc0 + (uint64_t)((stage * Cfg::kBBytes) >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
              ptx::umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                             (kb > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit(empty_bar(stage));          // frees the smem stage when the MMAs retire
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit(tfull_bar(acc));   // accumulator ready for the epilogue
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + kEpiWarps) {
    // ===================================================================== epilogue warps
    const int e = warp - kFirstEpiWarp;
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int half = e >> 2;                  // column half of the tile
    const uint32_t stg = smem_epi + (uint32_t)e * kEpiStageBytes;
    int acc = 0; uint32_t acc_phase = 0;
    int cur_len = 0;
    if (EPI == EPI_ARGMAX) cur_len = (p.cur_len != nullptr) ? *p.cur_len : (1 << 30);
    // Pull this warp's bias segment into L1 while the K loop runs.  In the decoder steps the parameters are
    // HBM-cold every time (weights and cross K/V stream through L2 between two uses), and the epilogue's bias load
    // sat behind the accumulator wait: one DRAM latency on the critical path of every small GEMM.
    if (p.bias != nullptr && (int)blockIdx.x < num_tiles) {
      const int c = (((int)blockIdx.x / S) % tiles_n) * BN + half * (BN / 2) + lane * 4;
      if (lane < BN / 8 && c < p.N) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        asm volatile("" ::"f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w));
      }
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = (tile / S) / tiles_n, n_blk = (tile / S) % tiles_n, slice = tile % S;
      if (!bounded_wait(tfull_bar(acc), acc_phase, 4)) break;
      if (dual && !bounded_wait(tfull_bar(1), 0u, 4)) break;      // second issuer's accumulator (stage 1)
      ptx::tc_fence_after();
      if (e == 0) GEMM_T(4);
      if (S > 1 && slice == 0) {
        // the other slices of this tile (adjacent, co-resident CTAs) have stored their partial accumulators
        volatile int* cnt = p.sk_cnt + tile / S;
        const int target = (S - 1) * kEpiWarps;
        const long long t0 = clock64();
        while (*cnt < target) {
          if (clock64() - t0 > kWatchdogCycles) { atomicExch(&g_gemm_watchdog, 5); break; }
        }
        __threadfence();
      }
      if constexpr (EPI == EPI_BF16_GELU && BN == 192) {
        // a lone tile (decoder FC1): warps 0..3 take chunks 2 and 5 of their lane quadrant once their own roles are
        // done (below), this warp the other two of its half -- 2 instead of 3 chunks on the critical path (the GELU
        // epilogue of a 128 x 192 tile on 8 warps was 3.4 of the launch's 7.7 us, profiles/r02_gemm_timing.txt)
        if (dual)
          epilogue_tile_chunks<BN, EPI>(p, tmem_base, m_blk * kBM + quad * 32, n_blk, quad, half ? 0x18u : 0x03u, lane, stg,
                                        (uint32_t)BN);
        else
          epilogue_tile<BN, EPI>(p, tmem_base + (uint32_t)(acc * BN), m_blk * kBM + quad * 32, n_blk, quad, half, lane, stg,
                                 cur_len, 0u, slice);
      } else {
        epilogue_tile<BN, EPI>(p, tmem_base + (uint32_t)(acc * BN), m_blk * kBM + quad * 32, n_blk, quad, half, lane, stg,
                               cur_len, dual ? (uint32_t)BN : 0u, slice);
      }
      if (S > 1) {
        __syncwarp();
        if (lane == 0) {
          if (slice > 0) {
            __threadfence();                                    // partial tile visible before the count
            atomicAdd(p.sk_cnt + tile / S, 1);
          } else {
            // every epilogue warp of slice 0 has passed the wait once the count reaches target + kEpiWarps: re-arm
            const int old = atomicAdd(p.sk_cnt + tile / S, 1);
            if (old == (S - 1) * kEpiWarps + kEpiWarps - 1) p.sk_cnt[tile / S] = 0;
          }
        }
      }
      if (e == 0) GEMM_T(5);
      // release the accumulator stage back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      if (EPI == EPI_F32_RESID && p.ln_out != nullptr)
        ln_tail(p, m_blk * kBM, e, kEpiWarps, lane, 1, reinterpret_cast<volatile int*>(tmem_slot_ptr + 1));
    }
    pdl_trigger_bottom();    // per-CTA trigger: the first epilogue warp that has stored its last tile
  }
  if constexpr (EPI == EPI_BF16_GELU && BN == 192) {
    // epilogue helpers: warp w (producer, issuers, allocator -- their loops are over) owns TMEM lane quadrant w
    if (!skip && dual && warp < 4 && (int)blockIdx.x < num_tiles) {
      const int m_blk = (int)blockIdx.x / tiles_n, n_blk = (int)blockIdx.x % tiles_n;
      if (!bounded_wait(tfull_bar(0), 0u, 4) || !bounded_wait(tfull_bar(1), 0u, 4)) goto done;
      ptx::tc_fence_after();
      epilogue_tile_chunks<BN, EPI>(p, tmem_base, m_blk * kBM + warp * 32, n_blk, warp, 0x24u, lane,
                                    smem_epi + (uint32_t)(kEpiWarps + warp) * kEpiStageBytes, (uint32_t)BN);
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
// CTA-pair kernel: tcgen05.mma.cta_group::2, UMMA 256 x BN x 16 over two SMs of one cluster
// ============================================================================================
// Each CTA of the pair stages its own 128 rows of A and HALF of the W tile (BN/2 rows) per k-block -- 32 KB instead
// of 48 KB per 128 x 256 x 64 of math for BN = 256 -- which is what bounds the 1-CTA kernel (L2 -> SM bytes and the
// shared-memory operand reads of the tensor pipe).  The leader CTA (cluster rank 0) issues every MMA; both CTAs' TMA
// loads credit the leader's `full` barrier; the leader's tcgen05.commit multicasts the `empty` / `accumulator full`
// arrivals to both CTAs; both CTAs' epilogue warps release the accumulator stage on the leader's barrier.
// bytes of transpose staging per epilogue warp: 32 rows x 64 B for the packed-bf16 epilogues, 32 x 128 B otherwise
template <int EPI> __host__ __device__ constexpr int kPairEpiStageBytes() {
  return (EPI == EPI_BF16 || EPI == EPI_BF16_GELU) ? 32 * 64 : kEpiStageBytes;
}

template <int BN> struct PairCfg {
  static constexpr int kABytes = kBM * kBK * 2;                // this CTA's 128 rows
  static constexpr int kBBytes = (BN / 2) * kBK * 2;           // this CTA's half of the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 6 : 8;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarBytes = 1024;
  template <int EPI, int EW> static constexpr int smem_bytes() {
    return kStages * kStageBytes + EW * kPairEpiStageBytes<EPI>() + kBarBytes + 1024 /*align*/;
  }
  static_assert(kStages * kStageBytes + kEpiWarps * kEpiStageBytes + kBarBytes + 1024 <= 227 * 1024,
                "shared memory budget exceeded");
};

template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc_pair() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * kBM) >> 4) << 24);
}

// EW = epilogue warps per CTA: 8 (two column halves per TMEM lane quadrant) or 16 (four column quarters; bf16
// epilogues only -- the GELU epilogue is latency-bound with two warps per scheduler: tensor pipe 63 % active)
template <int BN, int EPI, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((kFirstEpiWarp + EW) * 32, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         EpiParams p, int K, int tiles_m, int tiles_n) {
  using Cfg = PairCfg<BN>;
  pdl_trigger_top();

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // identical offset in both CTAs
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + Cfg::kStages * Cfg::kABytes;
  const uint32_t smem_epi = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar0 = smem_epi + EW * kPairEpiStageBytes<EPI>();
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int num_tiles = tiles_m * tiles_n;               // 256-row tiles
  const int num_kb = (K + kBK - 1) / kBK;
  // Tile schedule of this cluster: strided over the tile list, or -- with the fused LayerNorm tail -- a contiguous
  // run of whole 256-row blocks (all their n-tiles), so that every row a CTA normalises was written by that CTA.
  const bool contig = (EPI == EPI_F32_RESID) && (p.ln_out != nullptr);
  int t_first = cluster_id, t_step = n_clusters, t_count = (num_tiles - cluster_id + n_clusters - 1) / n_clusters;
  if (contig) {
    const int mb_per = (tiles_m + n_clusters - 1) / n_clusters;
    t_first = cluster_id * mb_per * tiles_n; t_step = 1;
    t_count = max(0, min(mb_per * tiles_n, num_tiles - t_first));
  }
  if (t_count < 0) t_count = 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);                    // leader's: its own expect_tx arrival (+ bytes of both CTAs)
      ptx::mbar_init(empty_bar(s), 1);                   // one multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);                   // one multicast commit
      ptx::mbar_init(tempty_bar(s), 2 * EW);      // leader's: the epilogue warps of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                               // peer barriers initialised before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  pdl_wait();
  const bool skip = (p.skip_flag != nullptr && *p.skip_flag != 0);

  if (skip) {
    // fall through to the teardown
  } else if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs; warp-uniform
    // loop, one elected lane issues)
    {
      int stage = 0; uint32_t phase = 0;
      for (int ti = 0, tile = t_first; ti < t_count; ++ti, tile += t_step) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        const int a_row = m_blk * (2 * kBM) + (int)rank * kBM;
        const int b_row = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!bounded_wait(empty_bar(stage), phase ^ 1u, 1)) goto done;
          const uint32_t lead_full = ptx::mapa_shared(full_bar(stage), 0);
          if (ptx::elect_one()) {
            if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            ptx::tma_load_2d_pair(smem_a0 + stage * Cfg::kABytes, &tmap_a, kb * kBK, a_row, lead_full);
            ptx::tma_load_2d_pair(smem_b0 + stage * Cfg::kBBytes, &tmap_w, kb * kBK, b_row, lead_full);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA; warp-uniform
    // loop, one elected lane issues -- see the 1-CTA kernel)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_pair<BN>();
      const uint64_t adesc0 = make_smem_desc(smem_a0), bdesc0 = make_smem_desc(smem_b0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int ti = 0, tile = t_first; ti < t_count; ++ti, tile += t_step) {
        if (!bounded_wait(tempty_bar(acc), acc_phase ^ 1u, 2)) goto done;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!bounded_wait(full_bar(stage), phase, 3)) goto done;
          ptx::tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)((stage * Cfg::kABytes) >> 4);
          const uint64_t bdesc = bdesc0 + (uint64_t)((stage * Cfg::kBBytes) >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k)
              ptx::umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
            ptx::umma_commit_pair(empty_bar(stage), (uint16_t)3);     // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(tfull_bar(acc), (uint16_t)3);   // accumulator ready in both CTAs
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================================================================== epilogue warps (both CTAs)
    const int e = warp - kFirstEpiWarp;
    const int quad = warp & 3;
    const int half = e >> 2;
    const uint32_t stg = smem_epi + (uint32_t)e * kPairEpiStageBytes<EPI>();
    int acc = 0; uint32_t acc_phase = 0;
    int cur_len = 0;
    if (EPI == EPI_ARGMAX) cur_len = (p.cur_len != nullptr) ? *p.cur_len : (1 << 30);
    for (int ti = 0, tile = t_first; ti < t_count; ++ti, tile += t_step) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      if (!bounded_wait(tfull_bar(acc), acc_phase, 4)) break;
      ptx::tc_fence_after();
      epilogue_tile<BN, EPI, EW / 4>(p, tmem_base + (uint32_t)(acc * BN), m_blk * (2 * kBM) + (int)rank * kBM + quad * 32,
                                     n_blk, quad, half, lane, stg, cur_len);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(tempty_bar(acc), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      if (contig && n_blk == tiles_n - 1) ln_tail_local(p, m_blk * (2 * kBM) + (int)rank * kBM, e, EW, lane, 1);
    }
    pdl_trigger_bottom();
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();          // the peer may still be read by the leader's MMAs / signalled by its commits
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace ec
