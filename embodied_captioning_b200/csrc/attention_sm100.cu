// ViT self-attention on the 5th-gen tensor cores (sm_100a): S = Q K^T and O = P V as tcgen05.mma with the
// accumulators, the probabilities and the output all living in TMEM.
//
// One persistent CTA per SM walks (crop, head) pairs.  Per pair: K and V of the head (<= 272 keys x 64) are staged
// once by TMA (128B swizzle); every 128-query tile then runs
//     QK^T  : tcgen05.mma  M=128, N=n_mma (<=256), K=64   A = Q tile (smem, K-major), B = K (smem, K-major) -> S (TMEM fp32)
//     softmax: one thread per query row (tcgen05.ld 32x32b: no shuffles), exp2 with the scale folded in,
//              P written back IN PLACE over S as packed bf16 (tcgen05.st), the row sum kept in registers
//     PV    : tcgen05.mma  M=128, N=64, K=n_mma         A = P (TMEM), B = V (smem, MN-major)            -> O (TMEM fp32,
//              columns [128,192) of the same buffer, free once the softmax has consumed S)
//     epilogue: O / rowsum -> bf16 -> global (row of 128 B per thread)
// Two S/P/O buffers (2 x 256 TMEM columns) and two softmax warpgroups ping-pong so the tensor core computes
// QK^T of tile i+1 while the CUDA cores run the softmax of tile i.  ViT-L/14 has 257 tokens: keys 0..255 go
// through the tensor core, the 257th key (and up to 15 more) is folded in by the row's own thread in fp32.
//
//   warp 0: TMA producer | warp 1: MMA issuer | warp 2: TMEM allocator | warps 4-7: softmax/epilogue group 0
//   | warps 8-11: group 1
#include "kernels.cuh"
#include <cuda.h>
#include <mutex>
#include <cstdlib>

namespace ec {

namespace aptx {
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
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one lane of a converged warp (reached by all 32 lanes)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Blackwell packed-pair fp32 pipe (FFMA2 / FADD2) and three-input max (FMNMX3): half the instructions per element
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
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
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
}  // namespace aptx

constexpr int kAD = 64;                  // head dim
constexpr int kATile = 128;              // query rows per tile (UMMA M)
constexpr int kAKeysMma = 256;           // keys handled by the tensor core (UMMA N limit)
constexpr int kAKeysTail = 16;           // rows of the TMA box that carries the keys beyond 256
constexpr int kATailMax = 2;             // extra keys (beyond 256) folded in by the row's own thread in fp32
constexpr int kAThreads = 12 * 32;
constexpr int kAKBytes = (kAKeysMma + kAKeysTail) * kAD * 2;     // 34816: K (or V) of one head, 272 rows x 128 B
constexpr int kAQBytes = kATile * kAD * 2;                        // 16384
constexpr int kASmemBytes = 2 * 2 * kAKBytes + 2 * kAQBytes + 1024 + 1024;
constexpr int kATmemCols = 512;          // two S/P/O buffers of 256 columns
constexpr long long kAWatchdog = 4000000000LL;

__device__ int g_attn_watchdog = 0;

__device__ __forceinline__ bool await(uint32_t bar, uint32_t parity, int code) {
  if (aptx::mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!aptx::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kAWatchdog) { atomicExch(&g_attn_watchdog, code); return false; }
  }
  return true;
}

// K-major SW128 operand (rows of 128 B, 8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major SW128 operand: 64 contiguous N elements (128 B) per K row, 8 K rows per 1024-B atom (SBO), one N chunk
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_qk(int n) {       // bf16 x bf16 -> f32, M=128, N=n, both K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kATile >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_pv() {            // M=128, N=64, A K-major (TMEM), B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(kAD >> 3) << 17) | ((uint32_t)(kATile >> 4) << 24);
}

struct AttnSm100Args {
  bf16* out; long long o_bs, o_rs;      // out[b*o_bs + s*o_rs + h*64 + d]
  int batch, heads, seq;                // seq = queries = keys (<= 272)
  int q_col0, k_col0, v_col0;           // first column of q / k / v inside the fused row
  float scale_log2;                     // (1/sqrt(64)) * log2(e)
};

// kMask = false: every key column the tensor core produces is a real key (n_valid_mma == n_mma, e.g. 257 tokens),
// so the two softmax passes carry no per-element bound checks (a third of their instructions; the softmax warps are
// what bounds this kernel: two warps per scheduler, issue slots 43 % busy, tensor pipe 14 %).
template <bool kMask>
__global__ void __launch_bounds__(kAThreads, 1)
attention_sm100_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16,
                       AttnSm100Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto k_smem = [&](int buf) { return base + (uint32_t)buf * 2u * kAKBytes; };
  auto v_smem = [&](int buf) { return base + (uint32_t)buf * 2u * kAKBytes + kAKBytes; };
  auto q_smem = [&](int buf) { return base + 4u * kAKBytes + (uint32_t)buf * kAQBytes; };
  const uint32_t bar0 = base + 4u * kAKBytes + 2u * kAQBytes;
  // barriers (8 B each): kv_full[2] kv_empty[2] q_full[2] q_empty[2] s_full[2] p_ready[2] o_full[2] s_empty[2]
  auto BAR = [&](int which, int i) { return bar0 + 8u * (uint32_t)(which * 2 + i); };
  enum { KV_FULL = 0, KV_EMPTY, Q_FULL, Q_EMPTY, S_FULL, P_READY, O_FULL, S_EMPTY };
  const uint32_t tmem_slot = bar0 + 8u * 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = a.batch * a.heads;
  const int n_qtiles = (a.seq + kATile - 1) / kATile;
  const int n_mma = min((a.seq + 15) & ~15, kAKeysMma);          // keys through the tensor core (multiple of 16)
  const int n_valid_mma = min(a.seq, kAKeysMma);
  const int n_tail = a.seq - n_valid_mma;                         // 0..16 keys handled in fp32 by the row threads
  const bool tail_box = a.seq > kAKeysMma;

  if (warp == 0 && lane == 0) { aptx::prefetch_tmap(&tm128); aptx::prefetch_tmap(&tm16); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      aptx::mbar_init(BAR(KV_FULL, i), 1);
      aptx::mbar_init(BAR(KV_EMPTY, i), n_qtiles);    // one commit per PV of the pair (each issued by its softmax group)
      aptx::mbar_init(BAR(Q_FULL, i), 1);
      aptx::mbar_init(BAR(Q_EMPTY, i), n_tail > 0 ? 1 + 128 : 1);   // QK^T commit (+ the row threads' tail-key dots)
      aptx::mbar_init(BAR(S_FULL, i), 1);
      aptx::mbar_init(BAR(P_READY, i), 128);
      aptx::mbar_init(BAR(O_FULL, i), 1);
      aptx::mbar_init(BAR(S_EMPTY, i), 128);
    }
    aptx::fence_barrier_init();
  }
  if (warp == 2) { aptx::tmem_alloc(tmem_slot, kATmemCols); aptx::tmem_relinquish(); }
  aptx::tc_fence_before();
  __syncthreads();
  aptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================================================ TMA producer (warp-uniform loop, one elected
    // lane issues: a single-lane branch costs ~25 election/broadcast instructions per bulk-tensor instruction)
    {
      uint32_t kv_it = 0, q_it = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const int b = pair / a.heads, h = pair % a.heads;
        const int row0 = b * a.seq;
        const int kvb = kv_it & 1;
        if (!await(BAR(KV_EMPTY, kvb), ((kv_it >> 1) & 1) ^ 1u, 1)) goto done;
        const int kv_boxes = (n_mma + kATile - 1) / kATile;
        if (aptx::elect_one()) {
          aptx::mbar_expect_tx(BAR(KV_FULL, kvb), (uint32_t)(2 * (kv_boxes * kAQBytes + (tail_box ? kAKeysTail * 128 : 0))));
          for (int i = 0; i < kv_boxes; ++i) {
            aptx::tma_load_2d(k_smem(kvb) + i * kAQBytes, &tm128, a.k_col0 + h * kAD, row0 + i * kATile, BAR(KV_FULL, kvb));
            aptx::tma_load_2d(v_smem(kvb) + i * kAQBytes, &tm128, a.v_col0 + h * kAD, row0 + i * kATile, BAR(KV_FULL, kvb));
          }
          if (tail_box) {
            aptx::tma_load_2d(k_smem(kvb) + 2 * kAQBytes, &tm16, a.k_col0 + h * kAD, row0 + kAKeysMma, BAR(KV_FULL, kvb));
            aptx::tma_load_2d(v_smem(kvb) + 2 * kAQBytes, &tm16, a.v_col0 + h * kAD, row0 + kAKeysMma, BAR(KV_FULL, kvb));
          }
        }
        __syncwarp();
        ++kv_it;
        for (int t = 0; t < n_qtiles; ++t, ++q_it) {
          const int qb = q_it & 1;
          if (!await(BAR(Q_EMPTY, qb), ((q_it >> 1) & 1) ^ 1u, 2)) goto done;
          if (aptx::elect_one()) {
            aptx::mbar_expect_tx(BAR(Q_FULL, qb), kAQBytes);
            aptx::tma_load_2d(q_smem(qb), &tm128, a.q_col0 + h * kAD, row0 + t * kATile, BAR(Q_FULL, qb));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ QK^T issuer (warp-uniform loop, one elected lane).
    // The P.V MMAs of a tile are issued by the tile's own softmax group (one handoff less in the per-tile chain).
    {
      const uint32_t id_qk = idesc_qk(n_mma);
      uint32_t kv_it = 0, it = 0;            // `it` counts q tiles globally: buffers alternate with it & 1
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const int kvb = kv_it & 1;
        if (!await(BAR(KV_FULL, kvb), (kv_it >> 1) & 1, 3)) goto done;
        for (int t = 0; t < n_qtiles; ++t, ++it) {
          const int sb = it & 1;
          if (!await(BAR(S_EMPTY, sb), ((it >> 1) & 1) ^ 1u, 4)) goto done;
          if (!await(BAR(Q_FULL, sb), (it >> 1) & 1, 6)) goto done;
          aptx::tc_fence_after();
          const uint32_t t_s = tmem_base + (uint32_t)(sb * 256);
          const uint64_t qdesc = desc_kmajor(q_smem(sb)), kdesc = desc_kmajor(k_smem(kvb));
          if (aptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kAD / 16; ++k)
              aptx::umma_ss(t_s, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), id_qk, k > 0 ? 1u : 0u);
            aptx::umma_commit(BAR(S_FULL, sb));
            aptx::umma_commit(BAR(Q_EMPTY, sb));
          }
          __syncwarp();
        }
        ++kv_it;
      }
    }
  } else if (warp >= 4) {
    // ============================================================ softmax + epilogue (one thread per query row)
    const int grp = (warp - 4) >> 2;          // buffer / tile parity this warpgroup serves
    const int quad = warp & 3;
    const int r = quad * 32 + lane;           // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    uint32_t kv_it = 0, it = 0;
    const uint32_t id_pv = idesc_pv();
    // all 128 rows of P are in TMEM -> the group's first warp issues O = P V and publishes O_FULL / frees K,V
    auto group_issue_pv = [&](int sb, int kvb) {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      if (quad == 0) {
        aptx::tc_fence_after();
        const uint32_t t_p = tmem_base + (uint32_t)(sb * 256);
        const uint32_t t_o = t_p + 128;
        const uint64_t vdesc0 = desc_mnmajor(v_smem(kvb));
        if (aptx::elect_one()) {
          for (int k = 0; k < n_mma / 16; ++k)
            aptx::umma_ts(t_o, t_p + (uint32_t)(k * 8), vdesc0 + (uint64_t)(k * (2048 >> 4)), id_pv, k > 0 ? 1u : 0u);
          aptx::umma_commit(BAR(O_FULL, sb));
          aptx::umma_commit(BAR(KV_EMPTY, kvb));
        }
        __syncwarp();
      }
    };
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++kv_it) {
      const int b = pair / a.heads, h = pair % a.heads;
      const int kvb = kv_it & 1;
      bool kv_waited = false;
      for (int t = 0; t < n_qtiles; ++t, ++it) {
        if ((int)(it & 1) != grp) continue;
        const int sb = grp;
        const uint32_t par = (it >> 1) & 1;
        const int qrow = t * kATile + r;
        const bool row_ok = qrow < a.seq;
        if (t * kATile + quad * 32 >= a.seq) {
          // no valid query row in this warp's 32-row slab (last tile of a 257-token sequence): keep the barrier
          // protocol in lock-step, skip all the arithmetic; the P/O rows of these lanes are never stored
          if (n_tail > 0) {
            if (!await(BAR(Q_FULL, sb), par, 8)) goto done;
            aptx::mbar_arrive(BAR(Q_EMPTY, sb));
          }
          if (!await(BAR(S_FULL, sb), par, 9)) goto done;
          group_issue_pv(sb, kvb);
          if (!await(BAR(O_FULL, sb), par, 10)) goto done;
          aptx::mbar_arrive(BAR(S_EMPTY, sb));
          continue;
        }
        // ---- tail keys (beyond the 256 the tensor core handles): fp32 dot products by the row's own thread
        float s_tail[kATailMax] = {0.f, 0.f};
        if (n_tail > 0) {
          if (!kv_waited) { if (!await(BAR(KV_FULL, kvb), (kv_it >> 1) & 1, 7)) goto done; kv_waited = true; }
          if (!await(BAR(Q_FULL, sb), par, 8)) goto done;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 uq = aptx::ld_shared_u4(q_smem(sb) + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) * 16));
            const float2 q0 = unpack_bf16x2(uq.x), q1 = unpack_bf16x2(uq.y), q2 = unpack_bf16x2(uq.z), q3 = unpack_bf16x2(uq.w);
#pragma unroll
            for (int j = 0; j < kATailMax; ++j) {
              if (j < n_tail) {
                const uint4 u = aptx::ld_shared_u4(k_smem(kvb) + 2u * kAQBytes + (uint32_t)j * 128u + (uint32_t)((c ^ (j & 7)) * 16));
                const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                s_tail[j] += q0.x * f0.x + q0.y * f0.y + q1.x * f1.x + q1.y * f1.y + q2.x * f2.x + q2.y * f2.y +
                             q3.x * f3.x + q3.y * f3.y;
              }
            }
          }
        }
        if (n_tail > 0) aptx::mbar_arrive(BAR(Q_EMPTY, sb));   // this thread no longer needs the Q tile
        // ---- S ready
        if (!await(BAR(S_FULL, sb), par, 9)) goto done;
        aptx::tc_fence_after();
        const uint32_t t_s = tmem_base + lane_addr + (uint32_t)(sb * 256);
        // Both passes over the S row keep one tcgen05.ld in flight while the previous 32 columns are processed.
        float mx = -INFINITY;
        {
          uint32_t va[32], vb[32];
          aptx::tmem_ld32(t_s, va);
          for (int c0 = 0; c0 < n_mma; c0 += 64) {
            aptx::tmem_ld_wait();
            const bool has_b = c0 + 32 < n_mma;
            if (has_b) aptx::tmem_ld32(t_s + (uint32_t)(c0 + 32), vb);
            if (!kMask) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) mx = aptx::max3(mx, __uint_as_float(va[j]), __uint_as_float(va[j + 1]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < n_valid_mma) mx = fmaxf(mx, __uint_as_float(va[j]));
            }
            if (has_b) {
              aptx::tmem_ld_wait();
              if (c0 + 64 < n_mma) aptx::tmem_ld32(t_s + (uint32_t)(c0 + 64), va);
              if (!kMask) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) mx = aptx::max3(mx, __uint_as_float(vb[j]), __uint_as_float(vb[j + 1]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c0 + 32 + j < n_valid_mma) mx = fmaxf(mx, __uint_as_float(vb[j]));
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < kATailMax; ++j)
          if (j < n_tail) mx = fmaxf(mx, s_tail[j]);
        const float ms = mx * a.scale_log2;
        float sum = 0.f;
        {
          uint32_t va[32], vb[32];
          const uint64_t sc2 = aptx::pack_f32x2(a.scale_log2, a.scale_log2), nms2 = aptx::pack_f32x2(-ms, -ms);
          uint64_t sum2 = aptx::pack_f32x2(0.f, 0.f);
          auto emit = [&](const uint32_t (&v)[32], int c0) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float e0, e1;
              aptx::unpack_f32x2(aptx::fma_f32x2(aptx::pack_f32x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, nms2),
                                 e0, e1);
              const float p0 = (!kMask || c0 + j < n_valid_mma) ? aptx::ex2(e0) : 0.f;
              const float p1 = (!kMask || c0 + j + 1 < n_valid_mma) ? aptx::ex2(e1) : 0.f;
              sum2 = aptx::add_f32x2(sum2, aptx::pack_f32x2(p0, p1));
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
            aptx::tmem_st16(t_s + (uint32_t)(c0 >> 1), pk);   // P overwrites S columns that were already consumed
          };
          aptx::tmem_ld32(t_s, va);
          for (int c0 = 0; c0 < n_mma; c0 += 64) {
            aptx::tmem_ld_wait();
            const bool has_b = c0 + 32 < n_mma;
            if (has_b) aptx::tmem_ld32(t_s + (uint32_t)(c0 + 32), vb);
            emit(va, c0);
            if (has_b) {
              aptx::tmem_ld_wait();
              if (c0 + 64 < n_mma) aptx::tmem_ld32(t_s + (uint32_t)(c0 + 64), va);
              emit(vb, c0 + 32);
            }
          }
          float s_lo, s_hi;
          aptx::unpack_f32x2(sum2, s_lo, s_hi);
          sum = s_lo + s_hi;
        }
        // tail keys: their P.V contribution is formed now, so every direct read of the pair's K/V tiles by this
        // thread happens before its P_READY arrive (the last PV commit of the pair then safely frees the tiles)
        float o[kAD];
#pragma unroll
        for (int j = 0; j < kAD; ++j) o[j] = 0.f;
#pragma unroll
        for (int j = 0; j < kATailMax; ++j) {
          if (j >= n_tail) break;
          const float p = aptx::ex2(fmaf(s_tail[j], a.scale_log2, -ms));
          sum += p;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = aptx::ld_shared_u4(v_smem(kvb) + 2u * kAQBytes + (uint32_t)j * 128u + (uint32_t)((c ^ (j & 7)) * 16));
            const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            o[c * 8 + 0] += p * f0.x; o[c * 8 + 1] += p * f0.y; o[c * 8 + 2] += p * f1.x; o[c * 8 + 3] += p * f1.y;
            o[c * 8 + 4] += p * f2.x; o[c * 8 + 5] += p * f2.y; o[c * 8 + 6] += p * f3.x; o[c * 8 + 7] += p * f3.y;
          }
        }
        aptx::tmem_st_wait();
        aptx::tc_fence_before();
        group_issue_pv(sb, kvb);
        // ---- O ready
        if (!await(BAR(O_FULL, sb), par, 10)) goto done;
        aptx::tc_fence_after();
        {
          uint32_t v[32];
          aptx::tmem_ld32(t_s + 128u, v);
          aptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] += __uint_as_float(v[j]);
          aptx::tmem_ld32(t_s + 160u, v);
          aptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) o[32 + j] += __uint_as_float(v[j]);
        }
        aptx::tc_fence_before();
        aptx::mbar_arrive(BAR(S_EMPTY, sb));             // buffer free for the QK^T of tile it+2
        if (row_ok) {
          const float inv = 1.0f / sum;
          bf16* orow = a.out + (size_t)b * a.o_bs + (size_t)qrow * a.o_rs + (size_t)h * kAD;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 u;
            u.x = pack_bf16x2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv); u.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
            u.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv); u.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(orow + c * 8) = u;
          }
        }
      }
    }
  }
done:
  aptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { aptx::tc_fence_after(); aptx::tmem_dealloc(tmem_base, kATmemCols); }
}


// ===================================================================================================
// v2: the 257-token specialisation (ViT-L/14, ViT-B/16 image towers: 1 class token + 256 patches).
//
// What bounded v1 (ncu, profiles/r01_ncu_full_final.csv: tensor pipe 18 %, MUFU 27 %, issue 27 %): a 128-query tile
// needs a 256-column fp32 S buffer, TMEM holds two of them, so each of the two softmax groups runs its tile's chain
// QK^T -> row max (first read of S) -> exp (second read of S) -> P.V -> O read strictly in series, and the 257th
// query costs a third, almost empty, tile per (crop, head).  v2 changes the decomposition:
//   * keys in three blocks (96 + 96 + 80 staged rows, the last one masked beyond key 256).  A block's S (96 columns) is
//     read from TMEM ONCE into registers and the S buffer handed back at once (S_FREE), so the next block's QK^T runs
//     while this block is exponentiated; P goes to one of two separate 48-column buffers, O is a single 64-column
//     accumulator (tcgen05.mma accumulates over the blocks);
//   * online softmax with a LAZY reference: block 0's row max is the reference for the whole row unless a later block
//     exceeds it by more than 2^8 (then the warp rescales O in TMEM and the running sum: rare, exact either way --
//     probabilities up to 256 are harmless in bf16 / fp32 and the final 1 / l uses the same reference);
//   * group g always owns query tile g of the pair (queries 128 g .. 128 g + 127): no third tile.  The 257th query is
//     one row: a spare warp runs it on the fp32 pipe straight from the staged K / V;
//   * one MMA-issuing warp PER GROUP (the in-order hazards on a group's TMEM columns need one issuer per group; each
//     warp blocks on its own group's barriers in their fixed order.  Measured alternatives: one warp polling both
//     groups answered every event ~1 k cycles late (0.77 ms per layer); the group's first softmax warp issuing the
//     MMAs itself put its own softmax work on everybody's critical path (1.04 ms)).
// TMEM columns of group g (base 256 g): [0,96) S block, [96,144) P (even blocks), [144,192) P (odd blocks), [192,256) O.
//
//   warp 0: TMA producer | warps 1-2: MMA issuers of group 0 / 1 (warp 2 also owns the TMEM allocation)
//   | warp 3: 257th query | warps 4-7: softmax group 0 | warps 8-11: softmax group 1
namespace av2 {

constexpr int kSeq = 257;
constexpr int kThreads = 12 * 32;
constexpr int kBlk = 96;                                 // keys of blocks 0 and 1
constexpr int kBlkLast = 80;                             // staged rows 192..271 (keys 192..256 + 15 rows of padding)
constexpr int kKvBytes = 2 * kAKBytes + 2048;            // K (272 rows) | V (272 rows) | 16-row box holding query 256
constexpr int kTailPartFloats = 100;                      // a warp's partial of the 257th query: m, l, (pad), acc[64]; + its 32 probabilities
constexpr int kSmemBytes = 2 * kKvBytes + 4 * kAQBytes + 8192 + 1024;   // + barriers, 257th-query partials, alignment
constexpr float kRescaleGap = 8.0f;                      // log2 units a later block's max may exceed the reference by
static_assert(kKvBytes % 1024 == 0 && kAKBytes % 1024 == 0 && (kBlk * 128) % 1024 == 0, "SW128 operands need 1024-byte aligned bases");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
static_assert(2 * kBlk + kBlkLast == kAKeysMma + kAKeysTail, "the three key blocks cover the staged rows exactly");

#ifdef EC_ATTN_TIMING
#define EC_TICK(var) const long long var = clock64()
#else
#define EC_TICK(var)
#endif

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// 2^x for two arguments on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 trick,
// a degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5 -- the result is rounded to bf16,
// 2^-9, right after), n added into the exponent field.  x is clamped at -126 (everything below is 0 in bf16 terms anyway).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  aptx::unpack_f32x2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f); x1 = fmaxf(x1, -126.0f);
  const uint64_t xc = aptx::pack_f32x2(x0, x1);
  const uint64_t magic = aptx::pack_f32x2(12582912.0f, 12582912.0f), nmagic = aptx::pack_f32x2(-12582912.0f, -12582912.0f);
  const uint64_t t = aptx::add_f32x2(xc, magic);                       // low mantissa bits of each half: round(x)
  const uint64_t r = aptx::add_f32x2(t, nmagic);
  const uint64_t f = aptx::fma_f32x2(r, aptx::pack_f32x2(-1.0f, -1.0f), xc);
  uint64_t p = aptx::fma_f32x2(aptx::pack_f32x2(0.05517151206731796f, 0.05517151206731796f), f,
                               aptx::pack_f32x2(0.2426111102104187f, 0.2426111102104187f));
  p = aptx::fma_f32x2(p, f, aptx::pack_f32x2(0.6932610273361206f, 0.6932610273361206f));
  p = aptx::fma_f32x2(p, f, aptx::pack_f32x2(0.9999280571937561f, 0.9999280571937561f));
  float t0, t1, q0, q1;
  aptx::unpack_f32x2(t, t0, t1);
  aptx::unpack_f32x2(p, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

#ifndef EC_ATTN_POLY
#define EC_ATTN_POLY 0      // pairs out of every four whose exponentials run on the FMA pipe instead of the MUFU
#endif

// exp of 32 score columns against the row's reference: P (bf16) -> 16 TMEM columns, running sum in `sum2`.
// EC_ATTN_POLY of every four pairs can be evaluated with the polynomial instead of the MUFU (the split FlashAttention-4
// uses).  Measured on B200 at 1024 crops: 0.604 ms per layer with 0, 0.596 with 1, 0.613 with 2 of 4 -- the MUFU pipe
// (46 % busy under ncu) is not what bounds this kernel, the per-warp dependency chains are -- so the default is 0.
__device__ __forceinline__ void exp_chunk(const uint32_t (&v)[32], uint64_t sc2, uint64_t nms2, uint64_t& sum2, uint32_t taddr) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const uint64_t e2 = aptx::fma_f32x2(aptx::pack_f32x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, nms2);
    float p0, p1;
    if (((j >> 1) & 3) < EC_ATTN_POLY) {
      exp2_poly2(e2, p0, p1);
    } else {
      float e0, e1;
      aptx::unpack_f32x2(e2, e0, e1);
      p0 = aptx::ex2(e0); p1 = aptx::ex2(e1);
    }
    sum2 = aptx::add_f32x2(sum2, aptx::pack_f32x2(p0, p1));
    pk[j >> 1] = pack_bf16x2(p0, p1);
  }
  aptx::tmem_st16(taddr, pk);
}
__device__ __forceinline__ float max_chunk(const uint32_t (&v)[32]) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    m0 = aptx::max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
    m1 = aptx::max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
  }
  return fmaxf(m0, m1);
}

__global__ void __launch_bounds__(kThreads, 1)
attention257_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16,
                    const __grid_constant__ CUtensorMap tm_out, AttnSm100Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto k_smem = [&](int buf) { return base + (uint32_t)buf * kKvBytes; };
  auto v_smem = [&](int buf) { return base + (uint32_t)buf * kKvBytes + kAKBytes; };
  auto qt_smem = [&](int buf) { return base + (uint32_t)buf * kKvBytes + 2u * kAKBytes; };
  auto q_smem = [&](int grp, int buf) { return base + 2u * kKvBytes + (uint32_t)(grp * 2 + buf) * kAQBytes; };
  const uint32_t bar0 = base + 2u * kKvBytes + 4u * kAQBytes;
  // barriers (8 B each)
  auto KV_FULL = [&](int b) { return bar0 + 8u * (uint32_t)b; };
  auto KV_EMPTY = [&](int b) { return bar0 + 8u * (uint32_t)(2 + b); };
  auto Q_FULL = [&](int g, int b) { return bar0 + 8u * (uint32_t)(4 + g * 2 + b); };
  auto Q_EMPTY = [&](int g, int b) { return bar0 + 8u * (uint32_t)(8 + g * 2 + b); };
  auto S_FULL = [&](int g) { return bar0 + 8u * (uint32_t)(12 + g); };
  auto S_FREE = [&](int g) { return bar0 + 8u * (uint32_t)(14 + g); };
  auto P_READY = [&](int g) { return bar0 + 8u * (uint32_t)(16 + g); };
  auto O_FREE = [&](int g) { return bar0 + 8u * (uint32_t)(18 + g); };
  auto PV_DONE = [&](int g, int b) { return bar0 + 8u * (uint32_t)(20 + g * 2 + b); };     // P.V of an even / odd block completed
  auto TAIL_FULL = [&](int b) { return bar0 + 8u * (uint32_t)(24 + b); };    // the 8 partials of the 257th query are written
  auto TAIL_EMPTY = [&](int b) { return bar0 + 8u * (uint32_t)(26 + b); };   // ... and merged
  auto OUT_READY = [&](int g) { return bar0 + 8u * (uint32_t)(28 + g); };    // the tile's output rows are staged in shared memory
  auto P_TAKEN = [&](int g) { return bar0 + 8u * (uint32_t)(30 + g); };      // the issuer has consumed a P_READY phase
  const uint32_t tmem_slot = bar0 + 8u * 32u;
  float* tail_part = reinterpret_cast<float*>(smem_raw + (bar0 + 512u - smem_u32(smem_raw)));   // [2][8][kTailPartFloats]
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int heads = a.heads;
  const int n_pairs = a.batch * heads;
  const int grid = (int)gridDim.x;

  if (warp == 0 && lane == 0) { aptx::prefetch_tmap(&tm128); aptx::prefetch_tmap(&tm16); aptx::prefetch_tmap(&tm_out); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      aptx::mbar_init(KV_FULL(i), 1);
      aptx::mbar_init(KV_EMPTY(i), 11);        // last P.V commit of either group + 8 softmax warps + the merging warp (257th query)
      aptx::mbar_init(TAIL_FULL(i), 8);
      aptx::mbar_init(TAIL_EMPTY(i), 1);
      for (int g = 0; g < 2; ++g) {
        aptx::mbar_init(Q_FULL(g, i), 1);
        aptx::mbar_init(Q_EMPTY(g, i), 1);     // the tile's output has left the buffer (it doubles as the store staging)
        aptx::mbar_init(PV_DONE(g, i), 1);
      }
      aptx::mbar_init(S_FULL(i), 1);
      aptx::mbar_init(S_FREE(i), 128);
      aptx::mbar_init(P_READY(i), 128);
      aptx::mbar_init(O_FREE(i), 128);
      aptx::mbar_init(OUT_READY(i), 128);
      aptx::mbar_init(P_TAKEN(i), 1);
    }
    aptx::fence_barrier_init();
  }
  if (warp == 2) { aptx::tmem_alloc(tmem_slot, kATmemCols); aptx::tmem_relinquish(); }
  aptx::tc_fence_before();
  __syncthreads();
  aptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================================================ TMA producer
    uint32_t it = 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += grid, ++it) {
      const int b = pair / heads, h = pair % heads;
      const int row0 = b * kSeq;
      const int buf = it & 1;
      const uint32_t par_empty = ((it >> 1) & 1) ^ 1u;
      if (!await(KV_EMPTY(buf), par_empty, 1)) goto done;
      if (aptx::elect_one()) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the 257th-query warps read these tiles with generic loads
        aptx::mbar_expect_tx(KV_FULL(buf), (uint32_t)kKvBytes);
        const int kc = a.k_col0 + h * kAD, vc = a.v_col0 + h * kAD;
        aptx::tma_load_2d(k_smem(buf), &tm128, kc, row0, KV_FULL(buf));
        aptx::tma_load_2d(k_smem(buf) + kAQBytes, &tm128, kc, row0 + 128, KV_FULL(buf));
        aptx::tma_load_2d(k_smem(buf) + 2 * kAQBytes, &tm16, kc, row0 + 256, KV_FULL(buf));
        aptx::tma_load_2d(v_smem(buf), &tm128, vc, row0, KV_FULL(buf));
        aptx::tma_load_2d(v_smem(buf) + kAQBytes, &tm128, vc, row0 + 128, KV_FULL(buf));
        aptx::tma_load_2d(v_smem(buf) + 2 * kAQBytes, &tm16, vc, row0 + 256, KV_FULL(buf));
        aptx::tma_load_2d(qt_smem(buf), &tm16, a.q_col0 + h * kAD, row0 + 256, KV_FULL(buf));
      }
      __syncwarp();
      for (int g = 0; g < 2; ++g) {
        if (!await(Q_EMPTY(g, buf), par_empty, 2)) goto done;
        if (aptx::elect_one()) {
          aptx::mbar_expect_tx(Q_FULL(g, buf), kAQBytes);
          aptx::tma_load_2d(q_smem(g, buf), &tm128, a.q_col0 + h * kAD, row0 + g * kATile, Q_FULL(g, buf));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ============================================================ MMA issuer of group g: its group's events come in a
    // fixed order -- S_FREE(0) P_READY(0) S_FREE(1) P_READY(1) S_FREE(2) P_READY(2) per tile -- so plain blocking waits do
    const int g = warp - 1;
    const uint32_t tg = tmem_base + (uint32_t)(g * 256);
    const uint32_t id_qk = idesc_qk(kBlk), id_qk_last = idesc_qk(kBlkLast), id_pv = idesc_pv();
    auto issue_qk = [&](uint32_t q_addr, uint32_t k_addr, uint32_t idesc) {
      const uint64_t qdesc = desc_kmajor(q_addr), kdesc = desc_kmajor(k_addr);
#pragma unroll
      for (int k = 0; k < kAD / 16; ++k)
        aptx::umma_ss(tg, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc, k > 0 ? 1u : 0u);
    };
    uint32_t it = 0;
    if ((int)blockIdx.x < n_pairs) {
      if (!await(KV_FULL(0), 0, 3)) goto done;
      if (!await(Q_FULL(g, 0), 0, 4)) goto done;
      aptx::tc_fence_after();
      if (aptx::elect_one()) {
        issue_qk(q_smem(g, 0), k_smem(0), id_qk);
        aptx::umma_commit(S_FULL(g));
      }
      __syncwarp();
    }
    for (int pair = blockIdx.x; pair < n_pairs; pair += grid, ++it) {
      const int buf = it & 1, nbuf = buf ^ 1;
      const bool has_next = pair + grid < n_pairs;
      const uint32_t npar = ((it + 1) >> 1) & 1;
      bool next_issued = false;
#pragma unroll
      for (int blk = 0; blk < 3; ++blk) {
        const uint32_t n = 3u * it + (uint32_t)blk;
        // ---- S of block `blk` is in registers -> QK^T of the next block (or of the next pair's first one)
        if (!await(S_FREE(g), n & 1u, 5)) goto done;
        aptx::tc_fence_after();
        if (blk < 2) {
          if (aptx::elect_one()) {
            issue_qk(q_smem(g, buf), k_smem(buf) + (uint32_t)((blk + 1) * kBlk * 128), blk == 1 ? id_qk_last : id_qk);
            aptx::umma_commit(S_FULL(g));
          }
          __syncwarp();
        } else if (has_next && aptx::mbar_try_wait(KV_FULL(nbuf), npar) && aptx::mbar_try_wait(Q_FULL(g, nbuf), npar)) {
          // (not staged yet: issued after this tile's last P.V instead, so that P.V is never held up)
          if (aptx::elect_one()) {
            issue_qk(q_smem(g, nbuf), k_smem(nbuf), id_qk);
            aptx::umma_commit(S_FULL(g));
          }
          __syncwarp();
          next_issued = true;
        }
        // ---- P of block `blk` is in TMEM -> O (+)= P V
        if (!await(P_READY(g), n & 1u, 7)) goto done;
        if (lane == 0) aptx::mbar_arrive(P_TAKEN(g));              // the group may publish the next block's P
#ifdef EC_ATTN_TIMING
        const long long t_obs = clock64();
#endif
        if (blk == 0 && !await(O_FREE(g), (it & 1) ^ 1u, 6)) goto done;    // the previous tile's epilogue has read O
        aptx::tc_fence_after();
        if (aptx::elect_one()) {
          const uint64_t vdesc0 = desc_mnmajor(v_smem(buf) + (uint32_t)(blk * kBlk * 128));
          const uint32_t t_p = tg + 96u + (n & 1u) * 48u;
          const int ksteps = (blk == 2) ? kBlkLast / 16 : kBlk / 16;
          for (int k = 0; k < ksteps; ++k)
            aptx::umma_ts(tg + 192u, t_p + (uint32_t)(k * 8), vdesc0 + (uint64_t)(k * (2048 >> 4)), id_pv, (k > 0 || blk > 0) ? 1u : 0u);
          aptx::umma_commit(PV_DONE(g, (int)(n & 1u)));
          if (blk == 2) aptx::umma_commit(KV_EMPTY(buf));
#ifdef EC_ATTN_TIMING
          long long* ts = reinterpret_cast<long long*>(smem_raw + (bar0 + 7168u - smem_u32(smem_raw))) + (g * 2 + (int)(n & 1u)) * 2;
          ts[0] = t_obs; ts[1] = clock64();
#endif
        }
        __syncwarp();
      }
      // ---- the tile's output rows are staged in its Q buffer -> one bulk tensor store; the buffer goes back to the
      // producer (for pair it + 2) once the store has read it
      if (!await(OUT_READY(g), it & 1u, 17)) goto done;
      if (aptx::elect_one()) {
        const int b = pair / heads, h = pair % heads;
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(&tm_out), "r"(h * kAD), "r"(b * kSeq + g * kATile), "r"(q_smem(g, buf)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        aptx::mbar_arrive(Q_EMPTY(g, buf));
      }
      __syncwarp();
      if (has_next && !next_issued) {
        if (!await(KV_FULL(nbuf), npar, 8)) goto done;
        if (!await(Q_FULL(g, nbuf), npar, 9)) goto done;
        aptx::tc_fence_after();
        if (aptx::elect_one()) {
          issue_qk(q_smem(g, nbuf), k_smem(nbuf), id_qk);
          aptx::umma_commit(S_FULL(g));
        }
        __syncwarp();
      }
    }
    if (aptx::elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // this lane's bulk stores have landed
    __syncwarp();
  } else if (warp == 3) {
    // ============================================================ the 257th query, merging warp.  The row's attention
    // is computed on the fp32 pipe straight from the staged K / V: every softmax warp takes 32 keys in the slot where it
    // would otherwise wait for its tile's last P.V (scores: a lane per key; context: a lane per pair of feature columns)
    // and leaves a partial (m, l, acc[64]); this warp adds key 256 and merges.  (Measured: the whole row on this one
    // warp made its scheduler's two softmax warps ~1.2 k cycles per (crop, head) slower than the other six, and both
    // groups wait for their slowest warp three times per tile; as a 16 x 272 x 64 mma.sync flash attention it took
    // 13 k cycles -- legacy HMMA at ~48 cycles each.)
    const float sc = a.scale_log2;
    uint32_t it = 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += grid, ++it) {
      const int b = pair / heads, h = pair % heads;
      const int buf = it & 1;
      if (!await(KV_FULL(buf), (it >> 1) & 1, 10)) goto done;
      // key 256: this lane's two feature columns of q, k and v (row 0 of each 16-row box: the swizzle term is 0)
      uint32_t wq, wk, wv;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wq) : "r"(qt_smem(buf) + (uint32_t)(lane * 4)));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wk) : "r"(k_smem(buf) + 256u * 128u + (uint32_t)(lane * 4)));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv) : "r"(v_smem(buf) + 256u * 128u + (uint32_t)(lane * 4)));
      __threadfence_block();
      __syncwarp();
      if (lane == 0) aptx::mbar_arrive(KV_EMPTY(buf));
      const float2 fq = unpack_bf16x2(wq), fk = unpack_bf16x2(wk), fv = unpack_bf16x2(wv);
      const float t256 = warp_sum(fmaf(fq.x, fk.x, fq.y * fk.y));
      if (!await(TAIL_FULL(buf), (it >> 1) & 1, 14)) goto done;
      const float* pp = tail_part + buf * 8 * kTailPartFloats;
      float m = t256;
#pragma unroll
      for (int w = 0; w < 8; ++w) m = fmaxf(m, pp[w * kTailPartFloats]);
      const float p256 = aptx::ex2((t256 - m) * sc);
      float l = p256, o0 = p256 * fv.x, o1 = p256 * fv.y;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float aw = aptx::ex2((pp[w * kTailPartFloats] - m) * sc);
        l = fmaf(pp[w * kTailPartFloats + 1], aw, l);
        o0 = fmaf(pp[w * kTailPartFloats + 4 + 2 * lane], aw, o0);
        o1 = fmaf(pp[w * kTailPartFloats + 5 + 2 * lane], aw, o1);
      }
      __syncwarp();
      if (lane == 0) aptx::mbar_arrive(TAIL_EMPTY(buf));
      const float inv = 1.0f / l;
      bf16* orow = a.out + (size_t)b * a.o_bs + (size_t)(kSeq - 1) * a.o_rs + (size_t)h * kAD + 2 * lane;
      *reinterpret_cast<uint32_t*>(orow) = pack_bf16x2(o0 * inv, o1 * inv);
    }
  } else if (warp >= 4) {
    // ============================================================ softmax + epilogue: one thread per query row
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t tg = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * 256);
    const float sc = a.scale_log2;
    const uint64_t sc2 = aptx::pack_f32x2(sc, sc);
    auto wait_pv = [&](uint32_t n) { return await(PV_DONE(grp, (int)(n & 1u)), (n >> 1) & 1u, 13); };   // P.V of block n complete
#ifdef EC_ATTN_TIMING
    long long tm_s = 0, tm_ld = 0, tm_exp = 0, tm_bp = 0, tm_epi_wait = 0, tm_epi = 0, tm_total = clock64(), t_arr = 0, tm_h1 = 0, tm_h2 = 0, tm_h3 = 0;
#endif
    uint32_t it = 0;
    for (int pair = blockIdx.x; pair < n_pairs; pair += grid, ++it) {
      const int b = pair / heads, h = pair % heads;
      const int buf = it & 1;
      float m_ref = 0.f, l = 0.f;
#pragma unroll
      for (int blk = 0; blk < 3; ++blk) {
        const uint32_t n = 3u * it + (uint32_t)blk;           // running block number: S_FULL phase, P buffer parity
        EC_TICK(c0);
        if (!await(S_FULL(grp), n & 1u, 11)) goto done;
        EC_TICK(c1);
        aptx::tc_fence_after();
        uint32_t v0[32], v1[32], v2[32];
        aptx::tmem_ld32(tg, v0);
        aptx::tmem_ld32(tg + 32u, v1);
        if (blk < 2) {
          aptx::tmem_ld32(tg + 64u, v2);
        } else {
          uint32_t vt[16];
          tmem_ld16(tg + 64u, vt);
          aptx::tmem_ld_wait();
          v2[0] = vt[0];                                        // column 64 of the last block = key 256; 65..79 are padding rows
        }
        aptx::tmem_ld_wait();
        aptx::tc_fence_before();
        aptx::mbar_arrive(S_FREE(grp));                         // S is in registers: the next QK^T may overwrite the buffer
        EC_TICK(c2);
        float mx = fmaxf(max_chunk(v0), max_chunk(v1));
        mx = (blk < 2) ? fmaxf(mx, max_chunk(v2)) : fmaxf(mx, __uint_as_float(v2[0]));
        const float x_max = mx * sc;
        if (blk == 0) {
          m_ref = x_max;
        } else {
          const bool grow = x_max > m_ref + kRescaleGap;
          if (__any_sync(0xffffffffu, grow)) {
            // rare: rescale this warp's rows of O and the running sums to the new reference (factor 1 for the other rows)
            if (!wait_pv(n - 1u)) goto done;
            aptx::tc_fence_after();
            const float f = grow ? aptx::ex2(m_ref - x_max) : 1.0f;
            uint32_t o[32];
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              aptx::tmem_ld32(tg + 192u + (uint32_t)(hlf * 32), o);
              aptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
              tmem_st32(tg + 192u + (uint32_t)(hlf * 32), o);
            }
            l *= f;
            if (grow) m_ref = x_max;
          }
        }
        if (n > 1u && !wait_pv(n - 2u)) goto done;             // this P buffer was last read by the P.V of block n - 2
        const uint64_t nms2 = aptx::pack_f32x2(-m_ref, -m_ref);
        uint64_t sum2 = aptx::pack_f32x2(0.f, 0.f);
        const uint32_t tp = tg + 96u + (n & 1u) * 48u;
        exp_chunk(v0, sc2, nms2, sum2, tp);
        exp_chunk(v1, sc2, nms2, sum2, tp + 16u);
        float s_lo, s_hi;
        if (blk < 2) {
          exp_chunk(v2, sc2, nms2, sum2, tp + 32u);
          aptx::unpack_f32x2(sum2, s_lo, s_hi);
        } else {
          const float pt = aptx::ex2(fmaf(__uint_as_float(v2[0]), sc, -m_ref));
          const uint32_t pk[8] = {pack_bf16x2(pt, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};     // key 256 | rows 257..271 (padding)
          tmem_st8(tp + 32u, pk);
          aptx::unpack_f32x2(sum2, s_lo, s_hi);
          s_lo += pt;
        }
        l += s_lo + s_hi;
        aptx::tmem_st_wait();
        EC_TICK(c3);
        // Back-pressure: block n's P is only published once the issuer has consumed the P_READY phase of block n - 1 --
        // it waits by parity (two unseen completions would look like none) and a fast warp must not arrive twice on one
        // phase.  In steady state that happened a whole block ago.
        if (n > 0u && !await(P_TAKEN(grp), (n - 1u) & 1u, 18)) goto done;
        aptx::tc_fence_before();
        aptx::mbar_arrive(P_READY(grp));
        EC_TICK(c4);
#ifdef EC_ATTN_TIMING
        t_arr = c4;
#endif
#ifdef EC_ATTN_TIMING
        tm_s += c1 - c0; tm_ld += c2 - c1; tm_exp += c3 - c2; tm_bp += c4 - c3;
#endif
      }
      // ---- this warp's 32 keys of the 257th query (the tile's last P.V is in flight meanwhile)
      {
        const int sw = warp - 4;
        const uint32_t kvpar = (it >> 1) & 1u;
        if (!await(KV_FULL(buf), kvpar, 15)) goto done;            // long complete: acquires the staged tiles for generic loads
        const uint32_t kb = k_smem(buf), vb = v_smem(buf), qb = qt_smem(buf);
        const int key = sw * 32 + lane;
        // scores: a lane per key; all sixteen 16-byte loads are issued before the first multiply
        uint4 uq[8], uk[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) uq[c] = aptx::ld_shared_u4(qb + (uint32_t)(c * 16));
#pragma unroll
        for (int c = 0; c < 8; ++c) uk[c] = aptx::ld_shared_u4(kb + (uint32_t)key * 128u + (uint32_t)((c ^ (key & 7)) * 16));
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float2 q0 = unpack_bf16x2(uq[c].x), q1 = unpack_bf16x2(uq[c].y), q2 = unpack_bf16x2(uq[c].z), q3 = unpack_bf16x2(uq[c].w);
          const float2 k0 = unpack_bf16x2(uk[c].x), k1 = unpack_bf16x2(uk[c].y), k2 = unpack_bf16x2(uk[c].z), k3 = unpack_bf16x2(uk[c].w);
          s0 = fmaf(q0.x, k0.x, s0); s1 = fmaf(q0.y, k0.y, s1); s2 = fmaf(q1.x, k1.x, s2); s3 = fmaf(q1.y, k1.y, s3);
          s0 = fmaf(q2.x, k2.x, s0); s1 = fmaf(q2.y, k2.y, s1); s2 = fmaf(q3.x, k3.x, s2); s3 = fmaf(q3.y, k3.y, s3);
        }
        const float t = (s0 + s1) + (s2 + s3);
        const float m_w = warp_max(t);
        const float p = aptx::ex2((t - m_w) * sc);
        const float l_w = warp_sum(p);
        // context: half-warp hf takes the keys of parity hf, lane (hf, d4) the four feature columns 4 d4 .. 4 d4 + 3
        if (!await(TAIL_EMPTY(buf), kvpar ^ 1u, 16)) goto done;    // the merge of pair it - 2 has read this slot
        float* pp = tail_part + (buf * 8 + sw) * kTailPartFloats;
        pp[68 + lane] = p;                                          // probabilities [32] behind the 68-float partial
        __syncwarp();
        const int hf = lane >> 4, d4 = lane & 15;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const uint32_t vrow = vb + (uint32_t)(sw * 32 + hf) * 128u + (uint32_t)((d4 & 1) * 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int kk = 2 * j + hf;                               // key inside the warp's 32 (sw * 32 is a multiple of 8)
          uint32_t w0, w1;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1)
                       : "r"(vrow + (uint32_t)(2 * j) * 128u + (uint32_t)(((d4 >> 1) ^ (kk & 7)) * 16)));
          const float pj = pp[68 + kk];
          const float2 f0 = unpack_bf16x2(w0), f1 = unpack_bf16x2(w1);
          a0 = fmaf(pj, f0.x, a0); a1 = fmaf(pj, f0.y, a1); a2 = fmaf(pj, f1.x, a2); a3 = fmaf(pj, f1.y, a3);
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16); a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
        a2 += __shfl_xor_sync(0xffffffffu, a2, 16); a3 += __shfl_xor_sync(0xffffffffu, a3, 16);
        __threadfence_block();       // keep the arrives below the last shared-memory read of the staged tiles
        if (lane == 0) { pp[0] = m_w; pp[1] = l_w; }
        if (hf == 0) *reinterpret_cast<float4*>(pp + 4 + 4 * d4) = make_float4(a0, a1, a2, a3);
        __syncwarp();
        if (lane == 0) { aptx::mbar_arrive(TAIL_FULL(buf)); aptx::mbar_arrive(KV_EMPTY(buf)); }
      }
      // ---- epilogue: o = O / l, staged in the tile's (now dead) Q buffer and written by one bulk tensor store
      const float inv = 1.0f / l;
      EC_TICK(e0);
      if (!wait_pv(3u * it + 2u)) goto done;
      EC_TICK(e1);
      aptx::tc_fence_after();
      uint32_t oa[32], ob[32];
      aptx::tmem_ld32(tg + 192u, oa);
      aptx::tmem_ld32(tg + 224u, ob);
      aptx::tmem_ld_wait();
      aptx::tc_fence_before();
      aptx::mbar_arrive(O_FREE(grp));                     // the accumulator is in registers
      const uint32_t stage = q_smem(grp, buf) + (uint32_t)r * 128u;       // every QK^T of the tile completed before its last P.V
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)((c ^ (r & 7)) * 16)),
                     "r"(pack_bf16x2(__uint_as_float(oa[8 * c]) * inv, __uint_as_float(oa[8 * c + 1]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(oa[8 * c + 2]) * inv, __uint_as_float(oa[8 * c + 3]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(oa[8 * c + 4]) * inv, __uint_as_float(oa[8 * c + 5]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(oa[8 * c + 6]) * inv, __uint_as_float(oa[8 * c + 7]) * inv)) : "memory");
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(((4 + c) ^ (r & 7)) * 16)),
                     "r"(pack_bf16x2(__uint_as_float(ob[8 * c]) * inv, __uint_as_float(ob[8 * c + 1]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(ob[8 * c + 2]) * inv, __uint_as_float(ob[8 * c + 3]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(ob[8 * c + 4]) * inv, __uint_as_float(ob[8 * c + 5]) * inv)),
                     "r"(pack_bf16x2(__uint_as_float(ob[8 * c + 6]) * inv, __uint_as_float(ob[8 * c + 7]) * inv)) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic stores -> the bulk store's async-proxy read
      aptx::mbar_arrive(OUT_READY(grp));                                 // the group's MMA warp issues the store
#ifdef EC_ATTN_TIMING
      { const long long e2 = clock64(); tm_epi_wait += e1 - e0; tm_epi += e2 - e1;
        const long long* ts = reinterpret_cast<const long long*>(smem_raw + (bar0 + 7168u - smem_u32(smem_raw))) + (grp * 2 + (int)((3u * it + 2u) & 1u)) * 2;
        tm_h1 += ts[0] - t_arr; tm_h2 += ts[1] - ts[0]; tm_h3 += e1 - ts[1]; }
#endif
    }
#ifdef EC_ATTN_TIMING
    if (blockIdx.x == 3 && lane == 0 && it > 0)
      printf("attention257 warp %d: %u pairs, per pair: S wait %lld, ld %lld, max+exp %lld, back-pressure %lld, last-PV wait %lld, "
             "epilogue %lld, total %lld cycles; last P.V: my arrive -> issuer saw P_READY %lld, -> issued %lld, -> I saw PV_DONE %lld\n",
             warp, it, tm_s / it, tm_ld / it, tm_exp / it, tm_bp / it, tm_epi_wait / it,
             tm_epi / it, (clock64() - tm_total) / it, tm_h1 / it, tm_h2 / it, tm_h3 / it);
#endif
  }
done:
  aptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { aptx::tc_fence_after(); aptx::tmem_dealloc(tmem_base, kATmemCols); }
}

}  // namespace av2

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(CUtensorMap* out, const bf16* base, long long rows, int cols, long long ld, int box_rows) {
  static PFN_encodeTiledA fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledA>(p);
  });
  EC_REQUIRE(fn != nullptr, EC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)kAD, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EC_REQUIRE(r == CUDA_SUCCESS, EC_ERR_CUDA, "cuTensorMapEncodeTiled (attention) failed with CUresult %d", (int)r);
  return EC_OK;
}

bool attention_sm100_supported(int seq, int head_dim) {
  return head_dim == kAD && seq >= 1 && seq <= kAKeysMma + kATailMax;
}

int attention_sm100_fused_qkv(const bf16* qkv, long long ld, int q_col0, int k_col0, int v_col0, bf16* out,
                              long long o_bs, long long o_rs, int batch, int heads, int seq, cudaStream_t stream,
                              bool force_v1) {
  EC_REQUIRE(attention_sm100_supported(seq, kAD), EC_ERR_UNSUPPORTED, "tcgen05 attention: seq %d unsupported", seq);
  EC_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15u) == 0, EC_ERR_INVALID, "tcgen05 attention: alignment");
  EC_REQUIRE(o_rs % 8 == 0 && o_bs % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, EC_ERR_INVALID,
             "tcgen05 attention: output alignment");
  const long long rows = (long long)batch * seq;
  const int cols = std::max(std::max(q_col0, k_col0), v_col0) + heads * kAD;
  EC_REQUIRE(cols <= ld, EC_ERR_INVALID, "tcgen05 attention: column slices exceed the row stride");
  CUtensorMap tm128, tm16;
  EC_TRY(encode_map(&tm128, qkv, rows, cols, ld, kATile));
  EC_TRY(encode_map(&tm16, qkv, rows, cols, ld, kAKeysTail));
  static bool attr_done[64] = {};
  const int dev_ = current_device();
  EC_REQUIRE(dev_ >= 0, EC_ERR_DEVICE, "no current CUDA device");
  bool& attr = attr_done[dev_];
  if (!attr) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(attention_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kASmemBytes));
    EC_CHECK_CUDA(cudaFuncSetAttribute(attention_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kASmemBytes));
    attr = true;
  }
  AttnSm100Args a;
  a.out = out; a.o_bs = o_bs; a.o_rs = o_rs; a.batch = batch; a.heads = heads; a.seq = seq;
  a.q_col0 = q_col0; a.k_col0 = k_col0; a.v_col0 = v_col0;
  a.scale_log2 = 0.125f * 1.4426950408889634f;
  const int grid = std::min(batch * heads, num_sms());
  // 257 tokens (ViT-L/14, ViT-B/16): the block-wise kernel (EC_ATTN_V2=0 or impl 2 of ec_attention_fused_qkv keep v1)
  static const int env_v2 = [] { const char* e = getenv("EC_ATTN_V2"); return e ? atoi(e) : 1; }();
  if (seq == av2::kSeq && env_v2 != 0 && !force_v1 && o_bs == (long long)seq * o_rs) {     // (one 2-D map covers the output)
    static bool attr2[64] = {};
    if (!attr2[dev_]) {
      EC_CHECK_CUDA(cudaFuncSetAttribute(av2::attention257_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, av2::kSmemBytes));
      attr2[dev_] = true;
    }
    CUtensorMap tm_out;
    EC_TRY(encode_map(&tm_out, out, rows, heads * kAD, o_rs, kATile));
    av2::attention257_kernel<<<grid, av2::kThreads, av2::kSmemBytes, stream>>>(tm128, tm16, tm_out, a);
    EC_POST_LAUNCH();
    return EC_OK;
  }
  const int n_mma = std::min((seq + 15) & ~15, kAKeysMma);
  if (std::min(seq, kAKeysMma) == n_mma)
    attention_sm100_kernel<false><<<grid, kAThreads, kASmemBytes, stream>>>(tm128, tm16, a);
  else
    attention_sm100_kernel<true><<<grid, kAThreads, kASmemBytes, stream>>>(tm128, tm16, a);
  EC_POST_LAUNCH();
  return EC_OK;
}

int attention_sm100_check_watchdog(cudaStream_t stream) {
  EC_CHECK_CUDA(cudaStreamSynchronize(stream));
  int v = 0;
  EC_CHECK_CUDA(cudaMemcpyFromSymbol(&v, g_attn_watchdog, sizeof(int)));
  if (v != 0) {
    int zero = 0;
    cudaMemcpyToSymbol(g_attn_watchdog, &zero, sizeof(int));
    set_error("tcgen05 attention pipeline wait timed out (role code %d)", v);
    return EC_ERR_TIMEOUT;
  }
  return EC_OK;
}

}  // namespace ec
