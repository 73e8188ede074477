// Host side of the tcgen05 GEMM: tensor maps, tile-shape heuristic, launch, profiling hooks, SIMT cross-check.
#include "gemm_kernel.cuh"
#include "gemm_seq.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace ec {

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

static bool pair_shape(int M, int N) { return ceil_div(M, 2 * kBM) * ceil_div(N, 256) >= num_sms() / 2; }

int gemm_choose_split_k(int M, int N, int K) {
  // Measured on B200 inside the decoder step (profiles/r02_decoder_experiments.txt): correct and bit-reproducible, but
  // SLOWER for the decoder's FC2 (decode 90.4 -> 96.4 ms at 1024 crops): slice 0 sits behind the other slices'
  // epilogues, then re-reads their partial tiles, and 144 CTAs poll / fence through L2.  Opt-in (EC_SPLITK=1).
  static const int mode = [] { const char* e = getenv("EC_SPLITK"); return e ? atoi(e) : 0; }();
  if (mode == 0 || K < 2048 || K % kBK != 0 || pair_shape(M, N)) return 0;
  const int tiles = ceil_div(M, kBM) * ceil_div(N, 128), kb = K / kBK;
  for (int s = 4; s >= 2; --s)
    if (kb % s == 0 && kb / s >= 2 && tiles * s <= num_sms()) return s;
  return 0;
}

bool gemm_uses_pair_kernel(int M, int N) {
  static const int mode = [] { const char* e = getenv("EC_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  return mode != 0 && pair_shape(M, N);
}

template <int BN, int EPI>
static int launch_tile(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  CUtensorMap ta, tw;
  const int groups = a.group_cols > 0 ? a.N / a.group_cols : 1;
  EC_TRY(make_tmap(&ta, a.A, a.M, a.K * groups, a.lda, kBM));
  EC_TRY(make_tmap(&tw, a.W, a.N, a.K, a.ldw, BN));
  EpiParams p;
  p.M = a.M; p.N = a.N; p.bias = a.bias; p.residual = a.residual; p.ldr = a.ldr; p.out = a.out; p.ldo = a.ldo;
  p.logits = a.logits; p.ld_logits = a.ld_logits; p.logits_step_stride = a.logits_step_stride; p.part_val = a.part_val; p.part_idx = a.part_idx;
  p.n_parts = a.n_parts; p.eos_col = a.eos_col; p.cur_len = a.cur_len; p.min_len = a.min_len;
  p.skip_flag = a.skip_flag;
  p.ln_g = a.ln_gamma; p.ln_b = a.ln_beta; p.ln_out = a.ln_out; p.ld_ln = a.ld_ln; p.ln_cnt = a.ln_counters;
  p.ln_tiles_n = ceil_div(a.N, BN);
  p.pf_ptr = reinterpret_cast<const unsigned char*>(a.prefetch_ptr); p.pf_bytes = a.prefetch_bytes;
  p.grp_cols = a.group_cols;
  const int S = (EPI == EPI_F32_RESID && a.split_k > 1) ? a.split_k : 1;
  p.split_k = S; p.sk_ws = a.sk_ws; p.sk_cnt = a.sk_cnt;
  const int tiles_m = ceil_div(a.M, kBM), tiles_n = ceil_div(a.N, BN);
  // measured (profiles/r02_l2_policies.txt): +0.4 ms per decode -- the eight row-block CTAs that share a W tile do not
  // arrive close enough together -- so the hint is opt-in (EC_W_EVICT_FIRST=1)
  static const int env_wef = [] { const char* e = getenv("EC_W_EVICT_FIRST"); return e ? atoi(e) : 0; }();
  p.w_evict_first = (env_wef != 0 && tiles_m <= 16) ? 1 : 0;
  EC_REQUIRE(S == 1 || tiles_m * tiles_n * S <= num_sms(), EC_ERR_UNSUPPORTED, "split-K needs one tile slice per SM");
  const int grid = std::min(tiles_m * tiles_n * S, num_sms());
  static bool attr_done[64] = {};   // per template instantiation and device
  const int dev_ = current_device();
  EC_REQUIRE(dev_ >= 0, EC_ERR_DEVICE, "no current CUDA device");
  bool& attr_set = attr_done[dev_];
  if (!attr_set) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_set = true;
  }
  note_launches(1);
  EC_CHECK_CUDA(launch_kernel(gemm_tcgen05_kernel<BN, EPI>, dim3(grid), dim3(kThreads1), (size_t)Cfg::kSmemBytes, stream,
                              ta, tw, p, a.K, tiles_m, tiles_n));
  return EC_OK;
}

template <int BN, int EPI, int EW>
static int launch_pair_ew(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = PairCfg<BN>;
  constexpr int kSmem = Cfg::template smem_bytes<EPI, EW>();
  CUtensorMap ta, tw;
  EC_TRY(make_tmap(&ta, a.A, a.M, a.K, a.lda, kBM));
  EC_TRY(make_tmap(&tw, a.W, a.N, a.K, a.ldw, BN / 2));
  EpiParams p;
  p.M = a.M; p.N = a.N; p.bias = a.bias; p.residual = a.residual; p.ldr = a.ldr; p.out = a.out; p.ldo = a.ldo;
  p.logits = a.logits; p.ld_logits = a.ld_logits; p.logits_step_stride = a.logits_step_stride; p.part_val = a.part_val; p.part_idx = a.part_idx;
  p.n_parts = a.n_parts; p.eos_col = a.eos_col; p.cur_len = a.cur_len; p.min_len = a.min_len;
  p.skip_flag = a.skip_flag;
  p.ln_g = a.ln_gamma; p.ln_b = a.ln_beta; p.ln_out = a.ln_out; p.ld_ln = a.ld_ln; p.ln_cnt = a.ln_counters;
  p.ln_tiles_n = ceil_div(a.N, BN);
  p.pf_ptr = nullptr; p.pf_bytes = 0; p.grp_cols = 0; p.split_k = 1; p.sk_ws = nullptr; p.sk_cnt = nullptr;
  p.w_evict_first = 0;
  const int tiles_m = ceil_div(a.M, 2 * kBM), tiles_n = ceil_div(a.N, BN);
  const int clusters = std::min(tiles_m * tiles_n, num_sms() / 2);
  static bool attr_done[64] = {};   // per template instantiation and device
  const int dev_ = current_device();
  EC_REQUIRE(dev_ >= 0, EC_ERR_DEVICE, "no current CUDA device");
  bool& attr_set = attr_done[dev_];
  if (!attr_set) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<BN, EPI, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmem));
    attr_set = true;
  }
  note_launches(1);
  EC_CHECK_CUDA(launch_kernel(gemm_tcgen05_pair_kernel<BN, EPI, EW>, dim3(2 * clusters), dim3((kFirstEpiWarp + EW) * 32),
                              (size_t)kSmem, stream, ta, tw, p, a.K, tiles_m, tiles_n));
  return EC_OK;
}

template <int BN, int EPI>
static int launch_pair(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = PairCfg<BN>;
  // 16 epilogue warps for the bf16-output epilogues (bias / GELU), 8 for fp32 residual and arg-max
  static const int ew_env = [] { const char* e = getenv("EC_GEMM_EW"); return e ? atoi(e) : 0; }();
  constexpr bool kCan16 = (EPI == EPI_BF16 || EPI == EPI_BF16_GELU);
  constexpr int EW = kCan16 ? 16 : 8;
  if (kCan16 && ew_env == 8) return launch_pair_ew<BN, EPI, 8>(a, stream);
  return launch_pair_ew<BN, EPI, EW>(a, stream);
}

// CTA-pair (cta_group::2) kernel for GEMMs with enough 256 x 256 tiles to occupy every SM pair
static bool use_pair_kernel(const GemmArgs& a) {
  static const int mode = [] { const char* e = getenv("EC_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (mode == 0) return false;
  return pair_shape(a.M, a.N);
}

template <int EPI>
static int launch_epi(const GemmArgs& a, cudaStream_t stream) {
  if constexpr (EPI == EPI_BF16 || EPI == EPI_F32) {
    if (a.group_cols > 0) {        // grouped GEMM: 1-CTA kernel, tiles must not straddle two groups
      int bn = a.bn_hint;
      if (bn == 0) bn = (a.group_cols % 256 == 0) ? 256 : (a.group_cols % 128 == 0 ? 128 : 64);
      EC_REQUIRE(a.group_cols % bn == 0, EC_ERR_UNSUPPORTED, "grouped GEMM: tile width %d does not divide group_cols %d",
                 bn, a.group_cols);
      if (bn == 256) return launch_tile<256, EPI>(a, stream);
      if (bn == 128) return launch_tile<128, EPI>(a, stream);
      if (bn == 64) return launch_tile<64, EPI>(a, stream);
      set_error("grouped GEMM: unsupported tile width %d", bn);
      return EC_ERR_UNSUPPORTED;
    }
  }
  if constexpr (EPI == EPI_F32_RESID) {
    if (a.split_k > 1) return launch_tile<128, EPI>(a, stream);
  }
  if (use_pair_kernel(a)) return launch_pair<256, EPI>(a, stream);
  if constexpr (EPI == EPI_ARGMAX) {
    return launch_tile<256, EPI>(a, stream);
  } else {
  static const int forced = [] { const char* e = getenv("EC_GEMM_BN"); return e ? atoi(e) : 0; }();
  if (forced == 256) return launch_tile<256, EPI>(a, stream);
  if constexpr (EPI == EPI_BF16 || EPI == EPI_BF16_GELU) {
    // 128 x 192 tiles when they turn a ragged second wave of 128 x 128 tiles into one wave (decoder FC1:
    // 1024 x 3072 is 192 tiles of 128 x 128 on 148 SMs, but 128 tiles of 128 x 192)
    const int t128 = ceil_div(a.M, kBM) * ceil_div(a.N, 128), t192 = ceil_div(a.M, kBM) * ceil_div(a.N, 192);
    if (forced == 192 || (forced == 0 && a.N % 192 == 0 && t128 > num_sms() && t192 <= num_sms()))
      return launch_tile<192, EPI>(a, stream);
  }
  if (forced == 128) return launch_tile<128, EPI>(a, stream);
  if (forced == 64) return launch_tile<64, EPI>(a, stream);
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
  if (a.group_cols > 0) {
    EC_REQUIRE(epi == EPI_BF16 || epi == EPI_F32, EC_ERR_UNSUPPORTED, "grouped GEMM supports the plain epilogues only");
    EC_REQUIRE(a.N % a.group_cols == 0 && a.K % kBK == 0 && a.group_cols % 64 == 0, EC_ERR_UNSUPPORTED,
               "grouped GEMM: N %d, group_cols %d, K %d", a.N, a.group_cols, a.K);
    EC_REQUIRE((long long)a.lda >= (long long)a.K * (a.N / a.group_cols), EC_ERR_INVALID,
               "grouped GEMM: A row stride %d shorter than groups * K", a.lda);
    EC_REQUIRE(a.ln_out == nullptr, EC_ERR_UNSUPPORTED, "grouped GEMM has no fused LayerNorm tail");
  }
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
  if (a.ln_out != nullptr) {
    EC_REQUIRE(epi == EPI_F32_RESID, EC_ERR_INVALID, "the fused LayerNorm tail belongs to the residual epilogue");
    EC_REQUIRE(a.ln_gamma && a.ln_beta && a.ln_counters, EC_ERR_INVALID, "fused LayerNorm tail: null parameter");
    EC_REQUIRE(a.N % 4 == 0 && a.N <= kLnTailMaxVec * 128 && a.ld_ln % 4 == 0, EC_ERR_UNSUPPORTED,
               "fused LayerNorm tail: row width %d unsupported", a.N);
  }
  if (a.split_k > 1) {
    EC_REQUIRE(epi == EPI_F32_RESID && a.ln_out == nullptr && a.group_cols == 0, EC_ERR_UNSUPPORTED,
               "split-K belongs to the plain residual epilogue");
    EC_REQUIRE(a.sk_ws && a.sk_cnt, EC_ERR_INVALID, "split-K needs a partial workspace and tile counters");
    EC_REQUIRE(a.K % kBK == 0 && (a.K / kBK) % a.split_k == 0 && a.split_k <= 4, EC_ERR_UNSUPPORTED,
               "split-K: K %d not divisible into %d slices of whole k-blocks", a.K, a.split_k);
    EC_REQUIRE((reinterpret_cast<uintptr_t>(a.sk_ws) & 15u) == 0 && a.N % 4 == 0, EC_ERR_INVALID, "split-K workspace alignment");
  }
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

// ============================================================================================
// GEMM sequence (gemm_seq.cuh)
// ============================================================================================
static_assert(kGemmSeqMaxStages == kSeqMaxStages && kGemmSeqCounterInts == kSeqCounterInts, "gemm_seq constants");

// tile width the stand-alone dispatch (launch_epi) would give this stage on the 1-CTA kernel; -1: not a sequence stage
static int seq_stage_bn(const GemmArgs& a, GemmEpilogue epi) {
  if (epi != EPI_BF16 && epi != EPI_BF16_GELU && epi != EPI_F32_RESID) return -1;
  if (a.split_k > 1 || a.M <= 0 || a.N <= 0 || a.K <= 0 || a.N % 8 != 0) return -1;
  if (a.group_cols > 0) {
    if (epi != EPI_BF16 || a.N % a.group_cols != 0 || a.K % kBK != 0 || a.group_cols % 64 != 0) return -1;
    int bn = a.bn_hint;
    if (bn == 0) bn = (a.group_cols % 256 == 0) ? 256 : (a.group_cols % 128 == 0 ? 128 : 64);
    return (a.group_cols % bn == 0 && (bn == 64 || bn == 128 || bn == 256)) ? bn : -1;
  }
  if (pair_shape(a.M, a.N)) return -1;
  const int sms = num_sms(), tm = ceil_div(a.M, kBM);
  if (epi != EPI_F32_RESID) {
    const int t128 = tm * ceil_div(a.N, 128), t192 = tm * ceil_div(a.N, 192);
    if (a.N % 192 == 0 && t128 > sms && t192 <= sms) return 192;
  }
  if (tm * ceil_div(a.N, 256) >= sms) return epi == EPI_F32_RESID ? -1 : 256;
  if (tm * ceil_div(a.N, 128) >= sms / 2 || a.N >= 1024) return 128;
  return 64;
}

bool gemm_seq_supported(const GemmSeqStage* st, int n) {
  if (st == nullptr || n < 1 || n > kSeqMaxStages) return false;
  const int M = st[0].a.M;
  if (M < 1 || ceil_div(M, kBM) > kSeqMaxRowBlocks) return false;
  for (int i = 0; i < n; ++i) {
    const GemmArgs& a = st[i].a;
    const int bn = seq_stage_bn(a, st[i].epi);
    if (bn < 0 || a.M != M || !a.A || !a.W || !a.out) return false;
    if (a.lda % 8 != 0 || a.ldw % 8 != 0 || a.ldo % 8 != 0) return false;
    if (st[i].epi == EPI_F32_RESID && (a.residual == nullptr || a.ldr % 4 != 0)) return false;
    if (a.ln_out != nullptr) {
      if (st[i].epi != EPI_F32_RESID || !a.ln_gamma || !a.ln_beta || a.N % 4 != 0 || a.N > kLnTailMaxVec * 128 || a.ld_ln % 4 != 0)
        return false;
      // the tail waits inside the epilogue for the block's other n-tiles: they must all be resident at once
      if (ceil_div(M, kBM) * ceil_div(a.N, bn) > num_sms()) return false;
    }
  }
  return true;
}

int gemm_timing_read(long long* host, int count) {
  EC_REQUIRE(host != nullptr && count > 0 && count <= kGemmTimingSlots * 12, EC_ERR_INVALID, "gemm_timing_read: bad arguments");
  EC_CHECK_CUDA(cudaDeviceSynchronize());
  EC_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_gemm_timing, (size_t)count * sizeof(long long)));
  return EC_OK;
}

int gemm_seq_timing_read(long long* host, int count) {
  const size_t total = (size_t)kSeqTimingTags * kSeqMaxStages * kSeqTimingEvents;
  EC_REQUIRE(host != nullptr && count > 0 && (size_t)count <= total, EC_ERR_INVALID, "gemm_seq_timing_read: bad arguments");
  EC_CHECK_CUDA(cudaDeviceSynchronize());
  EC_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_seq_timing, (size_t)count * sizeof(long long)));
  return EC_OK;
}

int gemm_seq(const GemmSeqStage* st, int n, int* counters, const int* skip_flag, const void* prefetch_ptr,
             size_t prefetch_bytes, cudaStream_t stream, int timing_tag) {
  EC_REQUIRE(gemm_seq_supported(st, n), EC_ERR_UNSUPPORTED, "GEMM sequence: unsupported stage list");
  EC_REQUIRE(counters != nullptr, EC_ERR_INVALID, "GEMM sequence: null counters");
  SeqMaps maps;
  SeqParams cp;
  memset(&maps, 0, sizeof(maps));
  memset(&cp, 0, sizeof(cp));
  const int M = st[0].a.M;
  cp.n_stages = n; cp.tiles_m = ceil_div(M, kBM); cp.counters = counters; cp.skip_flag = skip_flag;
  cp.pf_ptr = reinterpret_cast<const unsigned char*>(prefetch_ptr); cp.pf_bytes = prefetch_bytes;
  cp.tag = timing_tag;
  int max_tiles = 1;
  for (int i = 0; i < n; ++i) {
    const GemmArgs& a = st[i].a;
    const int bn = seq_stage_bn(a, st[i].epi);
    const int groups = a.group_cols > 0 ? a.N / a.group_cols : 1;
    EC_REQUIRE(a.lda >= a.K * groups && a.ldw >= a.K, EC_ERR_INVALID, "GEMM sequence: row strides shorter than K");
    EC_TRY(make_tmap(&maps.a[i], a.A, a.M, a.K * groups, a.lda, kBM));
    EC_TRY(make_tmap(&maps.w[i], a.W, a.N, a.K, a.ldw, bn));
    SeqStageDev& d = cp.st[i];
    EpiParams& p = d.p;
    p.M = a.M; p.N = a.N; p.bias = a.bias; p.residual = a.residual; p.ldr = a.ldr; p.out = a.out; p.ldo = a.ldo;
    p.eos_col = -1; p.split_k = 1; p.grp_cols = a.group_cols;
    p.ln_g = a.ln_gamma; p.ln_b = a.ln_beta; p.ln_out = a.ln_out; p.ld_ln = a.ld_ln;
    p.pf_ptr = reinterpret_cast<const unsigned char*>(a.W);            // warp 2 pulls the later stages' weights into L2
    p.pf_bytes = (unsigned long long)a.N * a.ldw * sizeof(bf16);
    d.K = a.K; d.bn = bn; d.epi = (int)st[i].epi; d.tiles_n = ceil_div(a.N, bn);
    d.ln_tail = a.ln_out != nullptr ? 1 : 0;
    d.dep = 0;
    if (i > 0) {
      const GemmArgs& prev = st[i - 1].a;
      if (prev.ln_out != nullptr && a.A == prev.ln_out) d.dep = 2;
      else if (a.A == reinterpret_cast<const bf16*>(prev.out)) d.dep = 1;
      // the residual stream of a later stage is complete only when the stage before it has stored the whole row block
      else if (st[i].epi == EPI_F32_RESID) d.dep = 1;
    }
    max_tiles = std::max(max_tiles, cp.tiles_m * d.tiles_n);
  }
  const int grid = std::min(max_tiles, num_sms());
  static bool attr_done[64] = {};
  const int dev_ = current_device();
  EC_REQUIRE(dev_ >= 0 && dev_ < 64, EC_ERR_DEVICE, "no current CUDA device");
  if (!attr_done[dev_]) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(gemm_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqSmemBytes));
    attr_done[dev_] = true;
  }
  note_launches(1);
  // no programmatic dependent launch: the kernel has no griddepcontrol.wait
  EC_CHECK_CUDA(launch_kernel_pdl(false, gemm_seq_kernel, dim3(grid), dim3(kThreads1), (size_t)kSeqSmemBytes, stream, maps, cp));
  return EC_OK;
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
  const size_t a0 = (size_t)row * lda + (p.grp_cols > 0 ? (size_t)(col / p.grp_cols) * K : 0);
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[a0 + k]) * __bfloat162float(W[(size_t)col * ldw + k]);
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
  p.ln_g = nullptr; p.ln_b = nullptr; p.ln_out = nullptr; p.ld_ln = 0; p.ln_cnt = nullptr; p.ln_tiles_n = 0;
  p.pf_ptr = nullptr; p.pf_bytes = 0; p.grp_cols = a.group_cols; p.split_k = 1; p.sk_ws = nullptr; p.sk_cnt = nullptr;
  p.w_evict_first = 0;
  dim3 grid(ceil_div(a.N, 128), a.M);
  gemm_simt_kernel<<<grid, 128, 0, stream>>>(a.A, a.lda, a.W, a.ldw, a.M, a.N, a.K, p, (int)epi);
  EC_POST_LAUNCH();
  return EC_OK;
}

}  // namespace ec
