// Encoder / attentional-pooler attention (non-causal, short sequences: 257 / 50 keys).
// One CTA per (head, crop): K and V of that head are staged once in padded shared memory, each warp
// owns 16-query tiles and runs a flash-style online softmax over 32-key chunks with
// mma.sync.m16n8k16 bf16 tensor-core tiles (fp32 accumulate, fp32 softmax).
#include "kernels.cuh"

namespace ec {

constexpr int kAttnWarps = 6;
constexpr int kAttnThreads = kAttnWarps * 32;
constexpr int kKvChunk = 32;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads) attention_full_kernel(AttnArgs a, int skv_pad) {
  constexpr int LDS = D + 8;          // padded row (elements): conflict-free ldmatrix
  constexpr int KS = D / 16;          // k-steps of QK^T
  constexpr int DT = D / 8;           // n8 tiles of the output
  extern __shared__ __align__(16) unsigned char smem_attn[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_attn);
  bf16* Vs = Ks + (size_t)skv_pad * LDS;

  const int b = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tg = lane & 3;
  // per-sequence key count (padded text batches); keys beyond it are masked like the padding of the last chunk
  const int skv = (a.kv_len != nullptr) ? max(1, min(a.kv_len[b], a.skv)) : a.skv;
  const int kv_end = ((skv + kKvChunk - 1) / kKvChunk) * kKvChunk;

  // ---- stage K, V of this (crop, head); rows >= skv are zero so that masked columns contribute exactly 0
  {
    const bf16* kg = a.k + (size_t)b * a.k_bs + (size_t)h * D;
    const bf16* vg = a.v + (size_t)b * a.v_bs + (size_t)h * D;
    constexpr int CPR = D / 8;        // 16-byte chunks per row
    for (int idx = tid; idx < skv_pad * CPR; idx += kAttnThreads) {
      const int r = idx / CPR, c = idx % CPR;
      uint4 kv4 = make_uint4(0, 0, 0, 0), vv4 = make_uint4(0, 0, 0, 0);
      if (r < skv) {
        kv4 = *reinterpret_cast<const uint4*>(kg + (size_t)r * a.k_rs + c * 8);
        vv4 = *reinterpret_cast<const uint4*>(vg + (size_t)r * a.v_rs + c * 8);
      }
      *reinterpret_cast<uint4*>(Ks + (size_t)r * LDS + c * 8) = kv4;
      *reinterpret_cast<uint4*>(Vs + (size_t)r * LDS + c * 8) = vv4;
    }
  }
  __syncthreads();

  const uint32_t ks_base = smem_u32(Ks), vs_base = smem_u32(Vs);
  const float sl2 = a.scale * 1.4426950408889634f;   // softmax scale folded into exp2
  const int n_qtiles = (a.sq + 15) >> 4;
  const bf16* qg = a.q + (size_t)b * a.q_bs + (size_t)h * D;
  bf16* og = a.o + (size_t)b * a.o_bs + (size_t)h * D;

  for (int qt = warp; qt < n_qtiles; qt += kAttnWarps) {
    const int q0 = qt * 16;
    const int r0 = min(q0 + g, a.sq - 1), r1 = min(q0 + g + 8, a.sq - 1);
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int c = ks * 16 + tg * 2;
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qg + (size_t)r0 * a.q_rs + c);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qg + (size_t)r1 * a.q_rs + c);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(qg + (size_t)r0 * a.q_rs + c + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(qg + (size_t)r1 * a.q_rs + c + 8);
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }

    for (int kv0 = 0; kv0 < kv_end; kv0 += kKvChunk) {
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          const int row = kv0 + np * 16 + (lane & 7) + ((lane >> 4) ? 8 : 0);
          const int col = ks * 16 + (((lane >> 3) & 1) ? 8 : 0);
          uint32_t b0, b1, b2, b3;
          ldsm_x4(ks_base + (uint32_t)(row * LDS + col) * 2u, b0, b1, b2, b3);
          mma_bf16_16816(s[np * 2], qa[ks], b0, b1);
          mma_bf16_16816(s[np * 2 + 1], qa[ks], b2, b3);
        }
      }
      // mask keys beyond skv, running max
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = kv0 + nt * 8 + tg * 2;
        if (c >= skv) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (c + 1 >= skv) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every chunk holds >= 1 valid key
      const float c0 = exp2f((m0 - mn0) * sl2), c1 = exp2f((m1 - mn1) * sl2);
      m0 = mn0; m1 = mn1;
      const float ms0 = mn0 * sl2, ms1 = mn1 * sl2;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = exp2f(s[nt][0] * sl2 - ms0); s[nt][1] = exp2f(s[nt][1] * sl2 - ms0);
        s[nt][2] = exp2f(s[nt][2] * sl2 - ms1); s[nt][3] = exp2f(s[nt][3] * sl2 - ms1);
        ps0 += s[nt][0] + s[nt][1];
        ps1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + ps0;
      l1 = l1 * c1 + ps1;
#pragma unroll
      for (int i = 0; i < DT; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      // O += P V
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < D / 16; ++dp) {
          const int row = kv0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) ? 8 : 0);
          const int col = dp * 16 + ((lane >> 4) ? 8 : 0);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(vs_base + (uint32_t)(row * LDS + col) * 2u, b0, b1, b2, b3);
          mma_bf16_16816(o[2 * dp], pa, b0, b1);
          mma_bf16_16816(o[2 * dp + 1], pa, b2, b3);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int row0 = q0 + g, row1 = q0 + g + 8;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      const int c = i * 8 + tg * 2;
      if (row0 < a.sq) *reinterpret_cast<uint32_t*>(og + (size_t)row0 * a.o_rs + c) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
      if (row1 < a.sq) *reinterpret_cast<uint32_t*>(og + (size_t)row1 * a.o_rs + c) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
    }
  }
}

template <int D>
static int launch_attn(const AttnArgs& a, cudaStream_t stream) {
  const int skv_pad = ceil_div(a.skv, kKvChunk) * kKvChunk;
  const size_t smem = (size_t)2 * skv_pad * (D + 8) * sizeof(bf16);
  EC_REQUIRE(smem <= 220 * 1024, EC_ERR_UNSUPPORTED, "attention: %d keys x head_dim %d exceed shared memory", a.skv, D);
  static size_t configured_dev[64] = {};
  const int dev_ = current_device();
  EC_REQUIRE(dev_ >= 0, EC_ERR_DEVICE, "no current CUDA device");
  size_t& configured = configured_dev[dev_];
  if (smem > configured) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(attention_full_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(a.batch, a.heads);
  attention_full_kernel<D><<<grid, kAttnThreads, smem, stream>>>(a, skv_pad);
  EC_POST_LAUNCH();
  return EC_OK;
}

int attention_full(const AttnArgs& a, cudaStream_t stream) {
  EC_REQUIRE(a.batch > 0 && a.heads > 0 && a.sq > 0 && a.skv > 0, EC_ERR_INVALID, "attention: empty problem");
  EC_REQUIRE(a.heads <= 65535, EC_ERR_UNSUPPORTED, "attention: %d heads > 65535", a.heads);
  auto ok8 = [](long long v) { return v % 8 == 0; };
  EC_REQUIRE(ok8(a.q_rs) && ok8(a.k_rs) && ok8(a.v_rs) && ok8(a.o_rs) && ok8(a.q_bs) && ok8(a.k_bs) && ok8(a.v_bs) &&
                 ok8(a.o_bs),
             EC_ERR_INVALID, "attention: strides must be multiples of 8 elements");
  switch (a.head_dim) {
    case 32: return launch_attn<32>(a, stream);
    case 64: return launch_attn<64>(a, stream);
    case 96: return launch_attn<96>(a, stream);
    default:
      set_error("attention: head_dim %d unsupported (32, 64, 96)", a.head_dim);
      return EC_ERR_UNSUPPORTED;
  }
}

}  // namespace ec
