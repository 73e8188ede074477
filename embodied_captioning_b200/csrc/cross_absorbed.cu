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
// 32 x (2T + 16)-byte block and arrives through a 3-deep ring of shared-memory stages with a single cp.async.bulk per
// chunk completing on an mbarrier (measured: one copy per key ROW -- 32 small bulk copies per chunk issued by one warp
// -- made the copy issue the critical path, 127 us per launch at 1024 crops); the ring runs ahead across crop
// boundaries.  8 warps: in the score phase each warp contracts its T/8 slice of the feature dimension for all 32 keys
// (Q' fragments stay in registers for the whole crop) and the 8 partial tiles are summed through shared memory; in the
// context phase each warp owns T/8 output columns.  Two __syncthreads per chunk.
#include "kernels.cuh"

namespace ec {

namespace xa {

constexpr int kKC = 32;          // keys per chunk
constexpr int kNS = 3;           // ring depth (chunks)
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSpLd = 40;        // floats per row of a partial-score tile (32 + 8 pad: conflict-free float2 stores)
constexpr int kPLd = 40;         // bf16 per row of the probability tile (80 B: conflict-free ldmatrix)

template <int T> struct Cfg {
  static constexpr int kRowBytes = T * 2 + 16;
  static constexpr int kStageBytes = kKC * kRowBytes;
  static constexpr int kQRowBytes = T * 2;             // Q' rows are staged unpadded (one contiguous copy per crop)
  static constexpr int kQBytes = 16 * kQRowBytes;
  static constexpr int kSpBytes = kWarps * 16 * kSpLd * 4;
  static constexpr int kPBytes = 16 * kPLd * 2;
  static constexpr int kStatBytes = 2 * 16 * 4;       // corr[16], l[16]
  static constexpr int kBarBytes = 64;                // full[kNS], qfull
  static constexpr int kSmem = kNS * kStageBytes + kQBytes + kSpBytes + kPBytes + kStatBytes + kBarBytes + 128;
  static_assert(kSmem <= 227 * 1024, "shared memory budget exceeded");
  static_assert(T % (kWarps * 16) == 0, "feature width must split into 16-wide k-steps per warp");
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int T>
__global__ void __launch_bounds__(kThreads, 1)
cross_absorbed_kernel(const bf16* __restrict__ qp, const bf16* __restrict__ xn, bf16* __restrict__ ctx, int n, int heads,
                      int n_kv, const unsigned char* __restrict__ finished, const DecodeState* __restrict__ st) {
  using C = Cfg<T>;
  constexpr int DPW = T / kWarps;      // feature columns per warp
  constexpr int KS = DPW / 16;         // k-steps of the score phase
  constexpr int NT = DPW / 8;          // n-tiles of the context phase
  constexpr float kLog2e = 1.4426950408889634f;
  pdl_trigger_top();
  pdl_wait();
  if (st != nullptr && st->done) return;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage0 = base;
  const uint32_t qs = base + kNS * C::kStageBytes;
  float* sp = reinterpret_cast<float*>(gen + kNS * C::kStageBytes + C::kQBytes);
  uint8_t* pb_gen = reinterpret_cast<uint8_t*>(sp) + C::kSpBytes;
  const uint32_t pb = qs + C::kQBytes + C::kSpBytes;
  float* corr_s = reinterpret_cast<float*>(pb_gen + C::kPBytes);
  float* l_s = corr_s + 16;
  const uint32_t bar0 = pb + C::kPBytes + C::kStatBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  const uint32_t q_bar = bar0 + 8u * kNS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q4 = lane & 3;
  const int nchunks = (n_kv + kKC - 1) / kKC;

  auto next_active = [&](int b) {
    if (finished != nullptr)
      while (b < n && finished[b]) b += (int)gridDim.x;
    return b;
  };

  // ---- prologue: barriers; Q' rows past `heads` are zeroed once (0 x NaN = NaN; key rows past n_kv are zero in HBM)
  if (tid == 0) {
    for (int s = 0; s < kNS; ++s) mbar_init(full_bar(s), 1);
    mbar_init(q_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    uint4* z = reinterpret_cast<uint4*>(gen + kNS * C::kStageBytes);
    for (int i = tid; i < C::kQBytes / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic zero-fill before the async-proxy copies
  __syncthreads();

  int cb = next_active((int)blockIdx.x);      // crop being consumed
  if (cb >= n) return;
  int pcrop = cb, pchunk = 0;                 // producer cursor (runs ahead, across crop boundaries)
  int g_issue = 0, g_use = 0;                 // chunks issued / consumed by this CTA

  auto issue_chunk = [&]() {                  // uniform bookkeeping; one thread issues the chunk's single bulk copy
    if (pcrop >= n) return;
    const int stage = g_issue % kNS;
    if (tid == 0) {
      mbar_expect_tx(full_bar(stage), (uint32_t)C::kStageBytes);
      bulk_g2s(stage0 + stage * C::kStageBytes,
               reinterpret_cast<const uint8_t*>(xn) + ((size_t)pcrop * nchunks + pchunk) * C::kStageBytes,
               (uint32_t)C::kStageBytes, full_bar(stage));
    }
    ++g_issue;
    if (++pchunk == nchunks) { pchunk = 0; pcrop = next_active(pcrop + (int)gridDim.x); }
  };
  auto issue_q = [&](int b) {
    if (b >= n) return;
    if (tid == 0) {
      mbar_expect_tx(q_bar, (uint32_t)heads * (uint32_t)C::kQRowBytes);
      bulk_g2s(qs, qp + (size_t)b * heads * T, (uint32_t)heads * (uint32_t)C::kQRowBytes, q_bar);
    }
  };

  issue_q(cb);
  issue_chunk();
  issue_chunk();

  for (int it = 0; cb < n; ++it) {
    const int nb = next_active(cb + (int)gridDim.x);
    // ---- the crop's absorbed queries: A fragments of Q' (16 x T, this warp's DPW columns) stay in registers
    if (!mbar_wait(q_bar, (uint32_t)(it & 1))) return;
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldsm_x4(qa[ks], qs + (lane & 15) * C::kQRowBytes + (warp * DPW + ks * 16 + (lane >> 4) * 8) * 2);

    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    float m_run = -INFINITY, l_run = 0.f;        // softmax state of row tid >> 4 (replicated over its 16 threads)

    for (int c = 0; c < nchunks; ++c, ++g_use) {
      const int stage = g_use % kNS;
      const uint32_t sbase = stage0 + stage * C::kStageBytes;
      if (!mbar_wait(full_bar(stage), (uint32_t)((g_use / kNS) & 1))) return;

      // ---- score phase: partial S[16 x 32] over this warp's feature columns
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b[4];
          ldsm_x4(b, sbase + (np * 16 + (lane & 7) + (lane >> 4) * 8) * C::kRowBytes +
                         (warp * DPW + ks * 16 + ((lane >> 3) & 1) * 8) * 2);
          mma16816(s[2 * np], qa[ks], b[0], b[1]);
          mma16816(s[2 * np + 1], qa[ks], b[2], b[3]);
        }
      }
      {
        float* my = sp + warp * 16 * kSpLd;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          *reinterpret_cast<float2*>(my + g * kSpLd + nt * 8 + 2 * q4) = make_float2(s[nt][0], s[nt][1]);
          *reinterpret_cast<float2*>(my + (g + 8) * kSpLd + nt * 8 + 2 * q4) = make_float2(s[nt][2], s[nt][3]);
        }
      }
      __syncthreads();                                   // (A) partials visible; every warp is past context(c-1)
      issue_chunk();                                     // refills the stage chunk c-1 used
      if (c == 0) issue_q(nb);                           // every warp holds its Q' fragments: the buffer is free

      // ---- softmax: thread (r, cp) owns keys 2cp, 2cp+1 of row r
      {
        const int r = tid >> 4, cp = tid & 15;
        float2 v = make_float2(0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          const float2 t = *reinterpret_cast<const float2*>(sp + (w * 16 + r) * kSpLd + 2 * cp);
          v.x += t.x; v.y += t.y;
        }
        const int key0 = c * kKC + 2 * cp;
        const float s0 = (key0 < n_kv) ? v.x : -INFINITY;
        const float s1 = (key0 + 1 < n_kv) ? v.y : -INFINITY;
        float mx = fmaxf(s0, s1);
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float m_new = fmaxf(m_run, mx);            // finite: chunk 0 always holds a valid key
        const float corr = ex2((m_run - m_new) * kLog2e);
        const uint32_t pp = pack_bf16x2(ex2((s0 - m_new) * kLog2e), ex2((s1 - m_new) * kLog2e));
        const float2 pr = unpack_bf16x2(pp);             // the normaliser sums the ROUNDED probabilities the MMA uses
        float sum = pr.x + pr.y;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        l_run = l_run * corr + sum;
        m_run = m_new;
        *reinterpret_cast<uint32_t*>(pb_gen + (r * kPLd + 2 * cp) * 2) = pp;
        if (cp == 0) { corr_s[r] = corr; l_s[r] = l_run; }
      }
      __syncthreads();                                   // (B) P, corr visible

      // ---- context phase: C[16 x DPW] = corr * C + P[16 x 32] X^[32 x DPW]
      {
        const float c_lo = corr_s[g], c_hi = corr_s[g + 8];
#pragma unroll
        for (int i = 0; i < NT; ++i) { acc[i][0] *= c_lo; acc[i][1] *= c_lo; acc[i][2] *= c_hi; acc[i][3] *= c_hi; }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          uint32_t pa[4];
          ldsm_x4(pa, pb + ((lane & 15) * kPLd + kk * 16 + (lane >> 4) * 8) * 2);
#pragma unroll
          for (int np = 0; np < NT / 2; ++np) {
            uint32_t b[4];
            ldsm_x4_t(b, sbase + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::kRowBytes +
                             (warp * DPW + np * 16 + (lane >> 4) * 8) * 2);
            mma16816(acc[2 * np], pa, b[0], b[1]);
            mma16816(acc[2 * np + 1], pa, b[2], b[3]);
          }
        }
      }
    }

    // ---- normalise and store the crop's context rows (heads x T)
    {
      const float i_lo = 1.0f / l_s[g], i_hi = 1.0f / l_s[g + 8];
      bf16* o_lo = ctx + ((size_t)cb * heads + g) * T + warp * DPW + 2 * q4;
      bf16* o_hi = ctx + ((size_t)cb * heads + g + 8) * T + warp * DPW + 2 * q4;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (g < heads) *reinterpret_cast<uint32_t*>(o_lo + nt * 8) = pack_bf16x2(acc[nt][0] * i_lo, acc[nt][1] * i_lo);
        if (g + 8 < heads) *reinterpret_cast<uint32_t*>(o_hi + nt * 8) = pack_bf16x2(acc[nt][2] * i_hi, acc[nt][3] * i_hi);
      }
    }
    cb = nb;
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

template <int T>
static int launch_xa(const bf16* qp, const bf16* xn, bf16* ctx, int n, int heads, int n_kv, const unsigned char* finished,
                     const DecodeState* st, cudaStream_t stream) {
  using C = xa::Cfg<T>;
  static bool attr_done[64] = {};
  const int dev = current_device();
  EC_REQUIRE(dev >= 0 && dev < 64, EC_ERR_DEVICE, "no current CUDA device");
  if (!attr_done[dev]) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(xa::cross_absorbed_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    attr_done[dev] = true;
  }
  const int grid = std::min(n, num_sms());
  note_launches(1);
  EC_CHECK_CUDA(launch_kernel(xa::cross_absorbed_kernel<T>, dim3(grid), dim3(xa::kThreads), (size_t)C::kSmem, stream, qp, xn,
                              ctx, n, heads, n_kv, finished, st));
  return EC_OK;
}

int decode_cross_absorbed(const bf16* qp, const bf16* xn, bf16* ctx, int batch, int heads, int width, int n_kv,
                          const unsigned char* finished, const DecodeState* st, cudaStream_t stream) {
  EC_REQUIRE(decode_cross_absorbed_supported(width, heads, width / std::max(heads, 1), n_kv), EC_ERR_UNSUPPORTED,
             "absorbed cross-attention: width %d, heads %d unsupported", width, heads);
  EC_REQUIRE(batch > 0 && qp && xn && ctx, EC_ERR_INVALID, "absorbed cross-attention: bad arguments");
  switch (width) {
    case 256: return launch_xa<256>(qp, xn, ctx, batch, heads, n_kv, finished, st, stream);
    case 512: return launch_xa<512>(qp, xn, ctx, batch, heads, n_kv, finished, st, stream);
    default: return launch_xa<768>(qp, xn, ctx, batch, heads, n_kv, finished, st, stream);
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
