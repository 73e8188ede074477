// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
// A (activations) and W (nn.Linear weight layout, open_clip `in_proj_weight` / `c_fc.weight` ...) are
// both K-major bf16, staged by TMA into 128B-swizzled shared-memory tiles and consumed by tcgen05.mma with fp32
// accumulators in TMEM: CTA pairs with cta_group::2 (UMMA 256 x 256 x 16) when the problem has at least one
// 256 x 256 tile per SM pair, otherwise cta_group::1 (UMMA 128 x BN x 16).  Persistent CTAs (one per SM),
// warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4.. = epilogue
// (TMEM -> registers -> global).  Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of
// tile i+1.
#pragma once
#include "common.cuh"

namespace ec {

enum GemmEpilogue : int {
  EPI_BF16 = 0,        // out bf16 = acc + bias
  EPI_BF16_GELU = 1,   // out bf16 = gelu(acc + bias)
  EPI_F32 = 2,         // out f32  = acc + bias
  EPI_F32_RESID = 3,   // out f32  = residual + acc + bias   (residual may alias out)
  EPI_ARGMAX = 4       // vocab head: optional f32 logits + per-(row, column-slab) arg-max partials
};

struct GemmArgs {
  const bf16* A;   int lda;       // [M, K], row stride lda elements (multiple of 8)
  const bf16* W;   int ldw;       // [N, K], row stride ldw elements (multiple of 8)
  int M, N, K;
  const float* bias;              // [N] or nullptr
  const float* residual; int ldr; // EPI_F32_RESID
  void* out; int ldo;             // bf16 or f32 [M, ldo]
  // EPI_ARGMAX only
  float* logits; long long ld_logits;   // optional [M, N] f32 (may be nullptr)
  long long logits_step_stride;         // logits base advances by (*cur_len - 1) * stride elements (0 if unused)
  float* part_val; int* part_idx; int n_parts;  // [M, n_parts]; n_parts = gemm_argmax_parts(N)
  int eos_col;                    // column forced to -inf when *cur_len < min_len
  const int* cur_len; int min_len;
  const int* skip_flag;           // device int; kernel returns immediately when *skip_flag != 0 (nullable)
  // EPI_F32_RESID only: fused LayerNorm tail.  When ln_out != nullptr the kernel also writes
  // ln_out[M, N] (bf16) = LayerNorm(out rows; ln_gamma, ln_beta, eps 1e-5) once a 128-row block of `out` is complete.
  // ln_counters: >= ceil(M / 128) ints, zero before the first launch (the kernel re-arms them); N <= 1024, N % 4 == 0.
  const float* ln_gamma; const float* ln_beta; bf16* ln_out; int ld_ln; int* ln_counters;
  // optional: bring [prefetch_ptr, +prefetch_bytes) into L2 while this GEMM runs (the next GEMM's weights; 1-CTA kernel)
  const void* prefetch_ptr; size_t prefetch_bytes;
  // Grouped (block-diagonal) form, group_cols > 0:  out[m, n] = sum_k A[m, (n / group_cols) * K + k] * W[n, k]  -- every
  // group of `group_cols` output columns reads its own K-wide column slab of A (A is [M, (N / group_cols) * K]).  Used
  // by the absorbed cross-attention of the decoder (per-head projections).  K % 64 == 0; 1-CTA kernel, tile width
  // `bn_hint` (0 = the widest of 256 / 128 / 64 that divides group_cols).
  int group_cols; int bn_hint;
  // Deterministic split-K for the residual epilogue of small-M GEMMs with a long K (decoder FC2: 48 tiles of 128 x 128
  // with K = 3072 leave 100 SMs idle while every CTA streams 1.5 MB through its own L2 port): split_k in {2, 3, 4}
  // slices per output tile on adjacent CTAs, partial accumulators through sk_ws ((split_k - 1) * M * N floats), summed
  // in slice order by slice 0 (bit-reproducible).  sk_cnt: >= tiles ints, zero before the first launch (re-armed by the
  // kernel).  Needs tiles(128 x 128) * split_k <= SM count and (K / 64) % split_k == 0; 0 / 1 = off.
  int split_k; float* sk_ws; int* sk_cnt;
};

// largest usable split_k (0 = none) for a residual GEMM of this shape
int gemm_choose_split_k(int M, int N, int K);

// true when gemm_bf16 runs this shape on the CTA-pair kernel (whose fused LayerNorm tail needs no inter-CTA sync)
bool gemm_uses_pair_kernel(int M, int N);

// number of arg-max partial slots per row produced by EPI_ARGMAX for vocabulary size N
int gemm_argmax_parts(int N);

// Launch. Returns EC_OK or an error code (set_error() filled in).
int gemm_bf16(const GemmArgs& a, GemmEpilogue epi, cudaStream_t stream);

// ---- GEMM sequence (gemm_seq.cuh): up to kGemmSeqMaxStages small dependent GEMMs, each reading the rows the one
// before it (or its LayerNorm tail: a.ln_out != nullptr, EPI_F32_RESID) produced, as ONE persistent launch.  Stage
// shapes are those of the 1-CTA kernel; results are bit-identical to launching the stages one after the other.
// `counters`: kGemmSeqCounterInts ints, ZERO at launch.  gemm_seq_supported() says whether a list can run fused.
struct GemmSeqStage { GemmArgs a; GemmEpilogue epi; };
constexpr int kGemmSeqMaxStages = 6;
constexpr int kGemmSeqCounterInts = kGemmSeqMaxStages * 2 * 16 * 32;      // one 128-byte line per counter
bool gemm_seq_supported(const GemmSeqStage* st, int n);
int gemm_seq(const GemmSeqStage* st, int n, int* counters, const int* skip_flag, const void* prefetch_ptr,
             size_t prefetch_bytes, cudaStream_t stream, int timing_tag = -1);
// stage clocks of CTA 0 recorded by -DEC_SEQ_TIMING builds: [64 tags][6 stages][16 events] cycles (zeros otherwise)
int gemm_seq_timing_read(long long* host, int count);
// -DEC_GEMM_TIMING builds: ring of 8192 records x 10 of CTA 0 of the 1-CTA kernel (gemm_kernel.cuh)
int gemm_timing_read(long long* host, int count);

// Test / debug helper: plain SIMT reference GEMM with the same epilogues (used by the unit tests to
// cross-check the tensor-core kernel on device; never called by the model path).
int gemm_reference_simt(const GemmArgs& a, GemmEpilogue epi, cudaStream_t stream);

// per-launch CUDA-event timing of every non-captured gemm_bf16() call (bench.py roofline line)
void gemm_profile_enable(bool on);
int gemm_profile_read(double* ms, double* flops, long long* launches);

// device-side watchdog flag (0 = fine). Reading it synchronises the stream.
int gemm_check_watchdog(cudaStream_t stream);

}  // namespace ec
