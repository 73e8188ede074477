// Absorbed-projection cross-attention of the decoder step (one new token per crop attends over the crop's image
// tokens; reference: text_decoder.cross_attn.{l} of coca_model.py:130-135 -> upstream MultimodalTransformer, i.e.
// nn.MultiheadAttention(q = ln_1(x), k = v = ln_1_kv(image_embs))).
//
// The keys and values of all 12 layers are projections of the SAME normalised image tokens X^ (the ln_1_kv affine
// parts are folded into the projection weights, model.cu::finalize), so instead of caching K and V per layer
// (2 x 255 x 768 bf16 per crop and layer, streamed every step at the HBM roofline) the projections are absorbed into
// the query and the output (the MLA form):
//
//     scores_h = q_h . (W_k,h x^_j + b_k,h) / 8  =  (W_k,h^T q_h / 8) . x^_j  + const      -> q'_h = W_k,h^T q_h / 8
//     out_h    = sum_j p_hj (W_v,h x^_j + b_v,h) =  W_v,h (sum_j p_hj x^_j) + b_v,h        -> c_h  = sum_j p_hj x^_j
//
// (the k-bias is constant over j and drops out of the softmax; sum_j p_hj = 1).  Two small grouped GEMMs produce q'
// ([crops, heads x T]) and apply W_v to c; this kernel is the part in between: per crop it streams X^ ([n_kv, T]
// bf16, ONE tensor for all layers: half the bytes of K and V) once through shared memory and computes
// S = Q' X^T (heads x n_kv), an online softmax, and C = P X^ (heads x T) on the tensor cores (mma.sync m16n8k16 with the
// <= 16 heads as the M tile -- the shape is 16 x 255 x 768 per crop, far too thin for a tcgen05 tile).
//
// One CTA per SM, persistent over crops (crop = blockIdx.x + i * gridDim.x).  X^ lives in HBM in the "packed" layout
// pack_image_tokens() writes: per crop ceil(n_kv / 32) chunks of 32 rows, each row T bf16 + 16 B of padding (so that
// ldmatrix on the staged copy is bank-conflict free), rows past n_kv zero.  A chunk is therefore ONE contiguous
// 32 x (2T + 16)-byte block and arrives through a 4-deep ring of shared-memory stages with a single cp.async.bulk per
// chunk completing on an mbarrier (measured: one copy per key ROW -- 32 small bulk copies per chunk issued by one warp
// -- made the copy issue the critical path, 127 us per launch at 1024 crops); the ring runs ahead across crop
// boundaries.  (The kernel reads "32" wherever the chunk size kKC appears in this comment.)
//
// Warp-specialised, 20 warps (<= 12 heads; 21 for 16), every role on its own chunk at any moment:
//   * producer warp: one bulk copy per chunk, carrying the L2 evict-first policy (the tokens are a 0.4 GB stream);
//   * SCORE group, 8 warps: warp w contracts its T/8 slice of the feature dimension for all 32 keys of a chunk and
//     stores a partial 16 x 32 tile.  Its Q' fragments stay in registers for the whole crop and come straight from
//     global memory (the next crop's are requested while this crop streams): no shared memory is spent on Q';
//   * SOFTMAX group, 3 warps (12 rows x 8 threads; 4 warps for 16 heads): sums the 8 partial tiles, online softmax,
//     publishes P (bf16), the rescale factors and the running normaliser in one of two buffers;
//   * CONTEXT group, 8 warps: warp w owns T/8 output columns, C = corr * C + P X^ with ldmatrix.trans fragments of the
//     same stage.
// mbarriers: full / empty per stage (released by the 16 MMA warps), sfull / sempty for the partial tiles, pfull / pempty
// per P buffer.  Measured on B200 at 1024 crops (profiles/r02_xattn.txt, r02_gemm_seq.txt, r02_l2_policies.txt): all
// warps doing both phases in lock-step with two __syncthreads per chunk ran 90 us per launch with 8 warps (latency-bound:
// issue slots 16 % busy) and 96 us with 16 warps (shared-memory pipe 69 % busy: twice the partial-tile traffic); score +
// softmax in one group next to a context group 92 us (the context group waited on the score group's serial S -> softmax
// chain); the three groups with a 3-stage ring and 21 warps 84.5 us under ncu (80 registers per thread: the context
// warps serialised LDSM.T -> 2 HMMA through one register quad, and a deeper ring changed nothing); 20 warps = 96
// registers 82.8 us, and THEN the fourth ring stage (bought by taking Q' out of shared memory and single-buffering the
// partial tiles) pays: 84.0 -> 79.9 us per launch in scripts/xattn_probe.py, decode 83.7 -> 82.0 ms.  Eight stages of
// 16 keys: 90.1 us (twice the hand-overs).
#include "kernels.cuh"

namespace ec {

namespace xa {

#ifndef EC_XA_KC
#define EC_XA_KC 32
#endif
constexpr int kKC = EC_XA_KC;    // keys per chunk (16 or 32)
#ifndef EC_XA_RING_KEYS
#define EC_XA_RING_KEYS 128
#endif
constexpr int kNS = EC_XA_RING_KEYS / kKC;   // ring depth (chunks): the ring holds 128 keys, all but one chunk of it in flight
#ifndef EC_XA_SP_BUFS
#define EC_XA_SP_BUFS 1
#endif
constexpr int kSB = EC_XA_SP_BUFS;           // partial-score buffers between the score and the softmax group (1 or 2)
constexpr int kKPT = kKC / 8;    // keys per softmax thread
constexpr int kPrefetch = 0;     // L2 prefetch ahead of the ring: measured WORSE (125 vs 92 us, 667 vs 426 MB of DRAM reads)
constexpr int kGroupWarps = 8;   // warps of the score group; the context group has as many
// softmax group: SW warps = 4 SW rows x 8 threads, 4 keys of the chunk per thread.  SW = 3 serves up to 12 heads (every
// CoCa decoder) with 20 warps per CTA = 5 per scheduler, which lifts the register cap from 80 to 96 per thread: the
// context warps then keep several LDSM.T quads in flight instead of serialising LDSM.T -> 2 HMMA through one quad.
template <int SW> constexpr int kThreadsOf = (2 * kGroupWarps + SW + 1) * 32;     // + the producer warp
constexpr int kSpLd = kKC + 8;   // floats per row of a partial-score tile (+ 8 pad: conflict-free float2 stores)
constexpr int kPLd = kKC + 8;    // bf16 per row of the probability tile (conflict-free ldmatrix)

template <int T> struct Cfg {
  static constexpr int kRowBytes = T * 2 + 16;
  static constexpr int kStageBytes = kKC * kRowBytes;
  static constexpr int kSpBufBytes = kGroupWarps * 16 * kSpLd * 4;
  static constexpr int kSpBytes = kSB * kSpBufBytes;
  static constexpr int kPBytes = 2 * 16 * kPLd * 2;    // double-buffered between the softmax and the context group
  static constexpr int kStatBytes = 2 * 2 * 16 * 4;    // corr[2][16], l[2][16]
  static constexpr int kBarBytes = 256;               // full/empty[kNS], sfull/sempty[2], pfull/pempty[2]
  static_assert((2 * kNS + 8) * 8 <= kBarBytes, "barrier block too small");
  static constexpr int kSmem = kNS * kStageBytes + kSpBytes + kPBytes + kStatBytes + kBarBytes + 128;
  static_assert(kSmem <= 227 * 1024, "shared memory budget exceeded");
  static_assert(T % (kGroupWarps * 16) == 0, "feature width must split into 16-wide k-steps per warp");
  static_assert(kKC == 32 || kKC == 16, "the softmax maps 8 threads onto the keys of a chunk");
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ int g_xa_watchdog = 0;
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { atomicExch(&g_xa_watchdog, 1); return false; }
  }
  return true;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the same copy with an L2 eviction policy (createpolicy): the image tokens are a 0.4 GB stream that is read once per
// launch -- evict-first keeps it from flushing the step's small hot data (kernel parameters, tensor maps, biases,
// activations, code) out of the 126 MB L2
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ldg_nc_pinned(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg_u8_pinned(const unsigned char* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int T, int SW>
__global__ void __launch_bounds__(kThreadsOf<SW>, 1)
cross_absorbed_kernel(const bf16* __restrict__ qp, const bf16* __restrict__ xn, bf16* __restrict__ ctx, int n, int heads,
                      int n_kv, const unsigned char* __restrict__ finished, const DecodeState* __restrict__ st, int reverse) {
  using C = Cfg<T>;
  constexpr int DPW = T / kGroupWarps;   // feature columns per warp (of the score and of the context group)
  constexpr int KS = DPW / 16;           // k-steps of the score phase
  constexpr int NT = DPW / 8;            // n-tiles of the context phase
  constexpr float kLog2e = 1.4426950408889634f;
  pdl_trigger_top();
  pdl_wait();
  if (st != nullptr && st->done) return;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage0 = base;
  float* sp = reinterpret_cast<float*>(gen + kNS * C::kStageBytes);                       // [kSB][8 warps][16][kSpLd]
  uint8_t* pb_gen = reinterpret_cast<uint8_t*>(sp) + C::kSpBytes;                          // P[2][16][kPLd] bf16
  const uint32_t pb = base + kNS * C::kStageBytes + C::kSpBytes;
  float* stat = reinterpret_cast<float*>(pb_gen + C::kPBytes);                             // corr[2][16], l[2][16]
  const uint32_t bar0 = pb + C::kPBytes + C::kStatBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kNS + s); };
  auto sfull_bar = [&](int b) { return bar0 + 8u * (2 * kNS + b); };
  auto sempty_bar = [&](int b) { return bar0 + 8u * (2 * kNS + 2 + b); };
  auto pfull_bar = [&](int b) { return bar0 + 8u * (2 * kNS + 4 + b); };
  auto pempty_bar = [&](int b) { return bar0 + 8u * (2 * kNS + 6 + b); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int nchunks = (n_kv + kKC - 1) / kKC;

  // Traversal order.  The image tokens are the same tensor for every layer, and what the previous layer's launch read
  // last is still in L2 (126 MB of the 0.4 GB): odd layers walk the crops backwards (`reverse`), so that every launch
  // starts on the part of the tensor the one before it finished on.  `b` below is a position in the walk.
  auto crop_of = [&](int b) { return (reverse & 1) ? n - 1 - b : b; };
  auto next_active = [&](int b) {
    if (finished != nullptr)
      while (b < n && finished[crop_of(b)]) b += (int)gridDim.x;
    return b;
  };

  // ---- prologue: barriers; Q' rows past `heads` are zeroed once (0 x NaN = NaN; key rows past n_kv are zero in HBM)
  if (tid == 0) {
    for (int s = 0; s < kNS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2 * kGroupWarps);      // every score warp + every context warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sfull_bar(b), kGroupWarps);           // every score warp stored its partial tile
      mbar_init(sempty_bar(b), 1);                    // the softmax group consumed the partial tiles
      mbar_init(pfull_bar(b), 1);                     // the softmax group published P / corr / l
      mbar_init(pempty_bar(b), kGroupWarps);          // every context warp consumed them
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int first = next_active((int)blockIdx.x);
  if (first >= n) return;

  // NOTE on the `__threadfence_block()` before every consumer-side mbarrier arrive: ptxas schedules the SYNCS.ARRIVE
  // above the warp's last LDSM / HMMA instructions otherwise (seen in the SASS), and the stage was then refilled while
  // it was still being read -- about one corrupted crop in 300 on B200.
  constexpr int kSoftWarps = SW;
  if (warp == 2 * kGroupWarps + kSoftWarps) {
    // =============================================================== producer warp: one bulk copy per chunk / per Q'.
    // The ring holds two chunks in flight (~97 KB per SM, 14 MB on the chip) -- not enough to cover the HBM latency at
    // full bandwidth -- so the chunks further ahead are pulled into L2 first (cp.async.bulk.prefetch.L2, kPrefetch
    // chunks ahead of the copy cursor, across crop boundaries): the copies into shared memory then hit L2.
    int gi = 0;
    const bool evict_first = (reverse & 2) != 0;
    const uint64_t policy = policy_evict_first();
    int pf_crop = first, pf_chunk = 0;
    auto prefetch_next = [&]() {
      if (pf_crop >= n) return;
      if (lane == 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(xn) +
                     ((size_t)crop_of(pf_crop) * nchunks + pf_chunk) * C::kStageBytes), "r"((uint32_t)C::kStageBytes) : "memory");
      if (++pf_chunk == nchunks) { pf_chunk = 0; pf_crop = next_active(pf_crop + (int)gridDim.x); }
    };
    if (kPrefetch > 0) for (int i = 0; i < kPrefetch; ++i) prefetch_next();
    for (int cb = first, cand = 0, cand_fin = 0; cb < n; cb = cand_fin ? next_active(cand) : cand) {
      // the next candidate's finished[] flag is requested now and looked at when this crop is done (see the score group)
      cand = cb + (int)gridDim.x;
      cand_fin = (finished != nullptr && cand < n) ? (int)ldg_u8_pinned(finished + crop_of(cand)) : 0;
      for (int c = 0; c < nchunks; ++c, ++gi) {
        const int stage = gi % kNS;
        if (kPrefetch > 0) prefetch_next();
        if (!mbar_wait(empty_bar(stage), (uint32_t)(((gi / kNS) & 1) ^ 1))) return;
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // consumers' generic-proxy reads -> async-proxy write
          mbar_expect_tx(full_bar(stage), (uint32_t)C::kStageBytes);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(xn) + ((size_t)crop_of(cb) * nchunks + c) * C::kStageBytes;
          if (evict_first) bulk_g2s_hint(stage0 + stage * C::kStageBytes, src, (uint32_t)C::kStageBytes, full_bar(stage), policy);
          else bulk_g2s(stage0 + stage * C::kStageBytes, src, (uint32_t)C::kStageBytes, full_bar(stage));
        }
        __syncwarp();
      }
    }
  } else if (warp < kGroupWarps) {
    // =============================================================== score group: partial S = Q' X^T per warp
    // The crop's absorbed queries: the A fragments of Q' (16 x T, this warp's DPW columns) stay in registers for the
    // whole crop.  They come straight from global memory (L2: the grouped GEMM before this kernel wrote them) in the
    // m16n8k16 fragment layout -- lane (g, q4) holds rows g / g + 8, columns 2 q4 and 8 + 2 q4 of every 16-wide k-step
    // -- and the NEXT crop's are requested while this crop's chunks stream, so no shared memory is spent on Q' (it
    // buys a fourth ring stage) and the crop boundary costs no round trip.  Rows past `heads` are zero.
    auto load_q = [&](uint32_t (&dst)[KS][4], int b) {
      const uint32_t* r_lo = reinterpret_cast<const uint32_t*>(qp + ((size_t)crop_of(b) * heads + g) * T + warp * DPW) + q4;
      const uint32_t* r_hi = r_lo + (size_t)8 * T / 2;
      const bool lo = g < heads, hi = g + 8 < heads;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        // volatile asm: a plain __ldg is sunk to its first use (ncu: the score warps spent 14 % of their time on that
        // round trip at every crop boundary); these stay where they are written, one crop ahead of their use
        dst[ks][0] = lo ? ldg_nc_pinned(r_lo + ks * 8) : 0u;
        dst[ks][1] = hi ? ldg_nc_pinned(r_hi + ks * 8) : 0u;
        dst[ks][2] = lo ? ldg_nc_pinned(r_lo + ks * 8 + 4) : 0u;
        dst[ks][3] = hi ? ldg_nc_pinned(r_hi + ks * 8 + 4) : 0u;
      }
    };
    int gu = 0;
    uint32_t qa[KS][4], qn[KS][4];
    load_q(qn, first);
    for (int cb = first; cb < n;) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) { qa[ks][0] = qn[ks][0]; qa[ks][1] = qn[ks][1]; qa[ks][2] = qn[ks][2]; qa[ks][3] = qn[ks][3]; }
      // Which crop comes next depends on finished[] -- a dependent round trip at every crop boundary (ncu: 14 % of the
      // score warps' time): the flag of the next candidate is requested here and looked at after the first chunk.
      const int cand = cb + (int)gridDim.x;
      uint32_t cand_fin = 0;
      if (finished != nullptr && cand < n) cand_fin = ldg_u8_pinned(finished + crop_of(cand));
      int cb_next = cand;

      for (int c = 0; c < nchunks; ++c, ++gu) {
        if (c == 1 || (c == 0 && nchunks == 1)) {
          if (cand_fin != 0) cb_next = next_active(cand);
          if (cb_next < n) load_q(qn, cb_next);
        }
        const int stage = gu % kNS, sb = gu % kSB;
        const uint32_t sbase = stage0 + stage * C::kStageBytes;
        if (!mbar_wait(full_bar(stage), (uint32_t)((gu / kNS) & 1))) return;
        float s[kKC / 8][4];
#pragma unroll
        for (int i = 0; i < kKC / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
          for (int np = 0; np < kKC / 16; ++np) {
            uint32_t b[4];
            ldsm_x4(b, sbase + (np * 16 + (lane & 7) + (lane >> 4) * 8) * C::kRowBytes +
                           (warp * DPW + ks * 16 + ((lane >> 3) & 1) * 8) * 2);
            mma16816(s[2 * np], qa[ks], b[0], b[1]);
            mma16816(s[2 * np + 1], qa[ks], b[2], b[3]);
          }
        }
        if (!mbar_wait(sempty_bar(sb), (uint32_t)(((gu / kSB) & 1) ^ 1))) return;    // softmax consumed chunk gu - kSB
        {
          float* my = sp + (sb * kGroupWarps + warp) * 16 * kSpLd;
#pragma unroll
          for (int nt = 0; nt < kKC / 8; ++nt) {
            *reinterpret_cast<float2*>(my + g * kSpLd + nt * 8 + 2 * q4) = make_float2(s[nt][0], s[nt][1]);
            *reinterpret_cast<float2*>(my + (g + 8) * kSpLd + nt * 8 + 2 * q4) = make_float2(s[nt][2], s[nt][3]);
          }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(sfull_bar(sb));
          mbar_arrive(empty_bar(stage));
        }
      }
      cb = cb_next;
    }
  } else if (warp < 2 * kGroupWarps) {
    // =============================================================== context group: C = corr * C + P X^
    const int cw = warp - kGroupWarps;
    int gu = 0;
    for (int cb = first, cand = 0, cand_fin = 0; cb < n; cb = cand_fin ? next_active(cand) : cand) {
      // the next candidate's finished[] flag is requested now and looked at when this crop is done (see the score group)
      cand = cb + (int)gridDim.x;
      cand_fin = (finished != nullptr && cand < n) ? (int)ldg_u8_pinned(finished + crop_of(cand)) : 0;
      float acc[NT][4];
#pragma unroll
      for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
      float i_lo = 0.f, i_hi = 0.f;
      for (int c = 0; c < nchunks; ++c, ++gu) {
        const int stage = gu % kNS, buf = gu & 1;
        const uint32_t sbase = stage0 + stage * C::kStageBytes;
        if (!mbar_wait(full_bar(stage), (uint32_t)((gu / kNS) & 1))) return;
        if (!mbar_wait(pfull_bar(buf), (uint32_t)((gu >> 1) & 1))) return;
        const float c_lo = stat[buf * 16 + g], c_hi = stat[buf * 16 + g + 8];
        if (c == nchunks - 1) { i_lo = 1.0f / stat[32 + buf * 16 + g]; i_hi = 1.0f / stat[32 + buf * 16 + g + 8]; }
#pragma unroll
        for (int i = 0; i < NT; ++i) { acc[i][0] *= c_lo; acc[i][1] *= c_lo; acc[i][2] *= c_hi; acc[i][3] *= c_hi; }
#pragma unroll
        for (int kk = 0; kk < kKC / 16; ++kk) {
          uint32_t pa[4];
          ldsm_x4(pa, pb + (((buf * 16) + (lane & 15)) * kPLd + kk * 16 + (lane >> 4) * 8) * 2);
#pragma unroll
          for (int np = 0; np < NT / 2; ++np) {
            uint32_t b[4];
            ldsm_x4_t(b, sbase + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::kRowBytes +
                             (cw * DPW + np * 16 + (lane >> 4) * 8) * 2);
            mma16816(acc[2 * np], pa, b[0], b[1]);
            mma16816(acc[2 * np + 1], pa, b[2], b[3]);
          }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) {                                    // this warp's reads of the stage and of P[buf] are done
          mbar_arrive(empty_bar(stage));
          mbar_arrive(pempty_bar(buf));
        }
      }
      // ---- normalise and store the crop's context rows (heads x T)
      bf16* o_lo = ctx + ((size_t)crop_of(cb) * heads + g) * T + cw * DPW + 2 * q4;
      bf16* o_hi = ctx + ((size_t)crop_of(cb) * heads + g + 8) * T + cw * DPW + 2 * q4;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (g < heads) *reinterpret_cast<uint32_t*>(o_lo + nt * 8) = pack_bf16x2(acc[nt][0] * i_lo, acc[nt][1] * i_lo);
        if (g + 8 < heads) *reinterpret_cast<uint32_t*>(o_hi + nt * 8) = pack_bf16x2(acc[nt][2] * i_hi, acc[nt][3] * i_hi);
      }
    }
  } else {
    // =============================================================== softmax group (128 threads): thread (r, c8) owns
    // keys 4 c8 .. 4 c8 + 3 of row r; sums the 8 partial tiles, online softmax, publishes P (bf16), corr and l
    const int t = tid - 2 * kGroupWarps * 32;
    const int r = t >> 3, c8 = t & 7;
    int gu = 0;
    for (int cb = first, cand = 0, cand_fin = 0; cb < n; cb = cand_fin ? next_active(cand) : cand) {
      // the next candidate's finished[] flag is requested now and looked at when this crop is done (see the score group)
      cand = cb + (int)gridDim.x;
      cand_fin = (finished != nullptr && cand < n) ? (int)ldg_u8_pinned(finished + crop_of(cand)) : 0;
      float m_run = -INFINITY, l_run = 0.f;      // softmax state of row r (replicated over its 8 threads)
      for (int c = 0; c < nchunks; ++c, ++gu) {
        const int buf = gu & 1, sb = gu % kSB;
        if (!mbar_wait(sfull_bar(sb), (uint32_t)((gu / kSB) & 1))) return;
        float v[kKPT];
#pragma unroll
        for (int i = 0; i < kKPT; ++i) v[i] = 0.f;
        const float* src = sp + (size_t)sb * kGroupWarps * 16 * kSpLd + r * kSpLd + kKPT * c8;
#pragma unroll
        for (int w = 0; w < kGroupWarps; ++w) {
          if constexpr (kKPT == 4) {
            const float4 x = *reinterpret_cast<const float4*>(src + w * 16 * kSpLd);
            v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w;
          } else {
            const float2 x = *reinterpret_cast<const float2*>(src + w * 16 * kSpLd);
            v[0] += x.x; v[1] += x.y;
          }
        }
        const int key0 = c * kKC + kKPT * c8;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kKPT; ++i) { v[i] = (key0 + i < n_kv) ? v[i] : -INFINITY; mx = fmaxf(mx, v[i]); }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float m_new = fmaxf(m_run, mx);            // finite: chunk 0 always holds a valid key
        const float corr = ex2((m_run - m_new) * kLog2e);
        uint32_t pk[kKPT / 2];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kKPT / 2; ++i) {
          pk[i] = pack_bf16x2(ex2((v[2 * i] - m_new) * kLog2e), ex2((v[2 * i + 1] - m_new) * kLog2e));
          const float2 a = unpack_bf16x2(pk[i]);         // the normaliser sums the ROUNDED probabilities the MMA uses
          sum += a.x + a.y;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        l_run = l_run * corr + sum;
        m_run = m_new;
        if (!mbar_wait(pempty_bar(buf), (uint32_t)(((gu >> 1) & 1) ^ 1))) return;   // context group released P[buf]
        {
          uint32_t* dst = reinterpret_cast<uint32_t*>(pb_gen + ((buf * 16 + r) * kPLd + kKPT * c8) * 2);
#pragma unroll
          for (int i = 0; i < kKPT / 2; ++i) dst[i] = pk[i];
        }
        if (c8 == 0) { stat[buf * 16 + r] = corr; stat[32 + buf * 16 + r] = l_run; }
        __threadfence_block();
        asm volatile("bar.sync 1, %0;" ::"n"(SW * 32) : "memory");     // partials consumed; P, corr, l written
        if (t == 0) {
          mbar_arrive(sempty_bar(sb));
          mbar_arrive(pfull_bar(buf));
        }
      }
    }
  }
  pdl_trigger_bottom();
}

// Packed image tokens (see the header): out row (crop, j) at ((crop * chunks * 32 + j) * (T + 8)) bf16; optionally
// normalised (x - mean) * rstd with eps 1e-5 (the shared ln_1_kv normalisation; the affine parts live in the folded
// projection weights).  One warp per packed row; rows j >= n_kv and the 8 padding columns are written as zeros.
__global__ void __launch_bounds__(256) pack_tokens_kernel(const float* __restrict__ x, bf16* __restrict__ out, int n,
                                                          int n_kv, int T, int normalize) {
  const int lane = threadIdx.x & 31;
  const int rows_per_crop = ((n_kv + kKC - 1) / kKC) * kKC;
  const long long total = (long long)n * rows_per_crop;
  const int ld = T + 8;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < total;
       row += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(row / rows_per_crop), j = (int)(row % rows_per_crop);
    bf16* o = out + (size_t)row * ld;
    if (lane < 4) reinterpret_cast<uint32_t*>(o + T)[lane] = 0u;
    if (j >= n_kv) {
      for (int c = lane * 2; c < T; c += 64) *reinterpret_cast<uint32_t*>(o + c) = 0u;
      continue;
    }
    const float* xr = x + ((size_t)b * n_kv + j) * T;
    float mean = 0.f, rstd = 1.f;
    if (normalize) {
      float s = 0.f;
      for (int c = lane * 2; c < T; c += 64) { const float2 v = *reinterpret_cast<const float2*>(xr + c); s += v.x + v.y; }
      mean = warp_sum(s) / (float)T;
      float q = 0.f;
      for (int c = lane * 2; c < T; c += 64) {
        const float2 v = *reinterpret_cast<const float2*>(xr + c);
        q += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean);
      }
      rstd = 1.0f / sqrtf(warp_sum(q) / (float)T + 1e-5f);
    }
    for (int c = lane * 2; c < T; c += 64) {
      const float2 v = *reinterpret_cast<const float2*>(xr + c);
      *reinterpret_cast<uint32_t*>(o + c) = pack_bf16x2((v.x - mean) * rstd, (v.y - mean) * rstd);
    }
  }
}

// wk_abs[(h * T + i) * D + d] = w_k[(h * D + d) * T + i] * scale     (w_k: [heads * D, T] rows of the folded K projection)
__global__ void __launch_bounds__(256) absorb_k_kernel(const bf16* __restrict__ wk, bf16* __restrict__ out, int heads,
                                                       int hd, int T, float scale) {
  const long long total = (long long)heads * T * hd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % hd);
    const long long hi = i / hd;
    const int col = (int)(hi % T), h = (int)(hi / T);
    out[i] = __float2bfloat16_rn(__bfloat162float(wk[((size_t)h * hd + d) * T + col]) * scale);
  }
}

}  // namespace xa

bool decode_cross_absorbed_supported(int width, int heads, int head_dim, int n_kv) {
  return (width == 256 || width == 512 || width == 768) && heads * head_dim == width && heads <= 16 && head_dim == 64 &&
         n_kv >= 1;
}

template <int T, int SW>
static int launch_xa_sw(const bf16* qp, const bf16* xn, bf16* ctx, int n, int heads, int n_kv, const unsigned char* finished,
                     const DecodeState* st, int reverse, cudaStream_t stream) {
  using C = xa::Cfg<T>;
  static bool attr_done[64] = {};
  const int dev = current_device();
  EC_REQUIRE(dev >= 0 && dev < 64, EC_ERR_DEVICE, "no current CUDA device");
  if (!attr_done[dev]) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(xa::cross_absorbed_kernel<T, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    attr_done[dev] = true;
  }
  // EC_XA_MAX_CTAS (probe): fewer CTAs than SMs leave room for another row chain's GEMMs next to this stream
  static const int env_cap = [] { const char* e = getenv("EC_XA_MAX_CTAS"); return e ? atoi(e) : 0; }();
  int grid = std::min(n, num_sms());
  if (env_cap > 0) grid = std::min(grid, env_cap);
  note_launches(1);
  EC_CHECK_CUDA(launch_small_kernel(xa::cross_absorbed_kernel<T, SW>, dim3(grid), dim3(xa::kThreadsOf<SW>), (size_t)C::kSmem, stream,
                                    qp, xn, ctx, n, heads, n_kv, finished, st, reverse));
  return EC_OK;
}

template <int T>
static int launch_xa(const bf16* qp, const bf16* xn, bf16* ctx, int n, int heads, int n_kv, const unsigned char* finished,
                     const DecodeState* st, int reverse, cudaStream_t stream) {
  static const int env_sw = [] { const char* e = getenv("EC_XA_SOFT_WARPS"); return e ? atoi(e) : 3; }();
  static const int env_ev = [] { const char* e = getenv("EC_XA_EVICT_FIRST"); return e ? atoi(e) : 1; }();
  reverse = (reverse & 1) | (env_ev ? 2 : 0);
  if (heads <= 12 && env_sw == 3) return launch_xa_sw<T, 3>(qp, xn, ctx, n, heads, n_kv, finished, st, reverse, stream);
  return launch_xa_sw<T, 4>(qp, xn, ctx, n, heads, n_kv, finished, st, reverse, stream);
}

int decode_cross_absorbed(const bf16* qp, const bf16* xn, bf16* ctx, int batch, int heads, int width, int n_kv,
                          const unsigned char* finished, const DecodeState* st, bool reverse, cudaStream_t stream) {
  EC_REQUIRE(decode_cross_absorbed_supported(width, heads, width / std::max(heads, 1), n_kv), EC_ERR_UNSUPPORTED,
             "absorbed cross-attention: width %d, heads %d unsupported", width, heads);
  EC_REQUIRE(batch > 0 && qp && xn && ctx, EC_ERR_INVALID, "absorbed cross-attention: bad arguments");
  switch (width) {
    case 256: return launch_xa<256>(qp, xn, ctx, batch, heads, n_kv, finished, st, reverse ? 1 : 0, stream);
    case 512: return launch_xa<512>(qp, xn, ctx, batch, heads, n_kv, finished, st, reverse ? 1 : 0, stream);
    default: return launch_xa<768>(qp, xn, ctx, batch, heads, n_kv, finished, st, reverse ? 1 : 0, stream);
  }
}

size_t packed_image_tokens_elems(int n, int n_kv, int width) {
  return (size_t)n * (size_t)(ceil_div(n_kv, xa::kKC) * xa::kKC) * (size_t)(width + 8);
}

int pack_image_tokens(const float* x, bf16* out, int n, int n_kv, int width, bool normalize, cudaStream_t stream) {
  EC_REQUIRE(width % 2 == 0 && n > 0 && n_kv > 0, EC_ERR_INVALID, "pack_image_tokens: bad shape");
  const long long rows = (long long)n * ceil_div(n_kv, xa::kKC) * xa::kKC;
  const int grid = (int)std::min<long long>(ceil_div<long long>(rows, 8), (long long)num_sms() * 16);
  xa::pack_tokens_kernel<<<grid, 256, 0, stream>>>(x, out, n, n_kv, width, normalize ? 1 : 0);
  EC_POST_LAUNCH();
  return EC_OK;
}

int absorb_k_weights(const bf16* wk, bf16* out, int heads, int head_dim, int width, float scale, cudaStream_t stream) {
  const long long total = (long long)heads * width * head_dim;
  const int grid = (int)std::min<long long>(ceil_div<long long>(total, 256), 4096);
  xa::absorb_k_kernel<<<grid, 256, 0, stream>>>(wk, out, heads, head_dim, width, scale);
  EC_POST_LAUNCH();
  return EC_OK;
}

int decode_cross_absorbed_check_watchdog(cudaStream_t stream) {
  EC_CHECK_CUDA(cudaStreamSynchronize(stream));
  int v = 0;
  EC_CHECK_CUDA(cudaMemcpyFromSymbol(&v, xa::g_xa_watchdog, sizeof(int)));
  if (v != 0) {
    int zero = 0;
    cudaMemcpyToSymbol(xa::g_xa_watchdog, &zero, sizeof(int));
    set_error("absorbed cross-attention: a bulk-copy wait timed out");
    return EC_ERR_TIMEOUT;
  }
  return EC_OK;
}

}  // namespace ec
