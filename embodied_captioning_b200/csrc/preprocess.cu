// Crop -> Pillow-exact antialiased bicubic resize -> centre crop -> /255 -> normalise, one gather kernel.
//
// Restates (bit-exactly) what `self.preprocess(PIL.Image)` does at coca.py:28 of the reference: torchvision
// Resize(shortest edge -> S, BICUBIC) + CenterCrop(S) + ToTensor + Normalize(OPENAI mean/std), whose resize is
// Pillow's two-pass (horizontal, then vertical) 8-bit resampler with 22-bit fixed-point coefficients
// (SURVEY.md Appendix B).  Only the SxS centre window of the resized image is ever computed.
//
// grid = (S/8 output-row bands, N crops).  Each CTA builds the fixed-point tap tables in fp64 with
// round-to-nearest intrinsics (no FMA contraction, so the taps equal Pillow's), runs the horizontal pass for
// the source rows its band needs into a shared-memory uint8 strip, then the vertical pass, and writes
// normalised bf16 / fp32 planes with coalesced stores.  HBM-bound: reads w*h*3 B, writes 3*S*S*2 B per crop.
#include "kernels.cuh"

namespace ec {

constexpr int kBandRows = 8;
constexpr int kMaxTaps = 48;
constexpr int kPrecBits = 22;
constexpr int kPreThreads = 256;

__constant__ float c_mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
__constant__ float c_std[3] = {0.26862954f, 0.26130258f, 0.27577711f};

__device__ __forceinline__ double bicubic_w(double x) {
  // Pillow bicubic_filter with a = -0.5, evaluated without contraction
  if (x < 0.0) x = -x;
  if (x < 1.0) {
    // ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    double t = __dsub_rn(__dmul_rn(1.5, x), 2.5);
    t = __dmul_rn(__dmul_rn(t, x), x);
    return __dadd_rn(t, 1.0);
  }
  if (x < 2.0) {
    // (((x - 5) * x + 8) * x - 4) * a
    double t = __dmul_rn(__dsub_rn(x, 5.0), x);
    t = __dmul_rn(__dadd_rn(t, 8.0), x);
    return __dmul_rn(__dsub_rn(t, 4.0), -0.5);
  }
  return 0.0;
}

// taps of output index `xx` when resampling in_size -> out_size; returns tap count, writes xmin and int taps
__device__ int resample_taps(int xx, int in_size, int out_size, int* __restrict__ taps, int* xmin_out) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, fs);
  const double ss = __ddiv_rn(1.0, fs);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  int n = xmax - xmin;
  if (n > kMaxTaps) n = kMaxTaps;    // guarded by the host-side scale bound; never taken for valid input
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    ww = __dadd_rn(ww, bicubic_w(arg));
  }
  for (int x = 0; x < n; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    double k = bicubic_w(arg);
    if (ww != 0.0) k = __ddiv_rn(k, ww);
    const double scaled = __dmul_rn(k, (double)(1 << kPrecBits));
    taps[x] = (k < 0.0) ? (int)__dadd_rn(-0.5, scaled) : (int)__dadd_rn(0.5, scaled);
  }
  *xmin_out = xmin;
  return n;
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> kPrecBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

struct PreArgs {
  const unsigned char* frames; int n_frames, fh, fw, fc;
  const int* rects; const int* frame_idx; const int* flags;
  int n, size;
  bf16* out_bf16; float* out_f32; int* status;
  int max_strip_rows;
};

__global__ void __launch_bounds__(kPreThreads) preprocess_kernel(PreArgs a) {
  extern __shared__ __align__(16) unsigned char smem_pre[];
  const int S = a.size;
  // shared layout: h_taps[S][kMaxTaps] int | h_xmin[S] | h_n[S] | v_taps[8][kMaxTaps] | v_ymin[8] | v_n[8] | strip u8
  int* h_taps = reinterpret_cast<int*>(smem_pre);
  int* h_xmin = h_taps + S * kMaxTaps;
  int* h_n = h_xmin + S;
  int* v_taps = h_n + S;
  int* v_ymin = v_taps + kBandRows * kMaxTaps;
  int* v_n = v_ymin + kBandRows;
  unsigned char* strip = reinterpret_cast<unsigned char*>(v_n + kBandRows);
  __shared__ int s_info[4];   // y_first, n_strip_rows, error
  // ToTensor + Normalize depend only on (channel, 8-bit value): 768 exactly-rounded results per CTA instead of two
  // IEEE divisions per output element (they were most of this kernel's instructions)
  __shared__ float s_lut[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += kPreThreads)
    s_lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), c_mean[i >> 8]), c_std[i >> 8]);

  const int crop = blockIdx.y, band = blockIdx.x, tid = threadIdx.x;
  const int x1 = a.rects[crop * 4 + 0], y1 = a.rects[crop * 4 + 1];
  const int x2 = a.rects[crop * 4 + 2], y2 = a.rects[crop * 4 + 3];
  const int fi = a.frame_idx[crop];
  const bool swap_rb = (a.flags != nullptr) && (a.flags[crop] & 1);
  const int w = x2 - x1, h = y2 - y1;
  const size_t plane = (size_t)S * S;
  const int oy0 = band * kBandRows;

  int err = 0;
  if (w <= 0 || h <= 0 || fi < 0 || fi >= a.n_frames) err = 1;           // empty crop / bad frame index
  int new_w = S, new_h = S;
  if (!err) {
    const int shorte = min(w, h), longe = max(w, h);
    const int new_long = (int)__ddiv_rn((double)((long long)S * longe), (double)shorte);   // int(S*long/short)
    if (w <= h) { new_w = S; new_h = new_long; } else { new_w = new_long; new_h = S; }
    const double sch = (double)w / (double)new_w, scv = (double)h / (double)new_h;
    const double sc = sch > scv ? sch : scv;
    const int need_taps = (int)ceil(2.0 * (sc < 1.0 ? 1.0 : sc)) * 2 + 1;
    if (need_taps > kMaxTaps) err = 2;                                     // down-scale factor too large
  }
  if (err) {
    if (band == 0 && tid == 0 && a.status) a.status[crop] = err;
    for (int i = tid; i < 3 * kBandRows * S; i += kPreThreads) {
      const int c = i / (kBandRows * S), r = (i / S) % kBandRows, ox = i % S;
      const size_t o = (size_t)crop * 3 * plane + (size_t)c * plane + (size_t)(oy0 + r) * S + ox;
      if (a.out_bf16) a.out_bf16[o] = __float2bfloat16_rn(0.f);
      if (a.out_f32) a.out_f32[o] = 0.f;
    }
    return;
  }
  if (band == 0 && tid == 0 && a.status) a.status[crop] = 0;
  // centre-crop offsets: int(round((dim - S) / 2.0)), Python round-half-even
  auto half_even = [](int d) { int q = d >> 1; if (d & 1) q += (q & 1); return q; };
  const int left = half_even(new_w - S), top = half_even(new_h - S);
  const bool need_h = (w != new_w), need_v = (h != new_h);

  if (!need_h && !need_v) {
    // ---- identity fast path (the crop already is S x S: every synthetic 224 x 224 crop of the headline config):
    // resampling with a single unit tap is the identity on 8-bit pixels, so gather -> table -> store, no strip
    __syncthreads();                                   // s_lut
    const unsigned char* frame = a.frames + (size_t)fi * a.fh * a.fw * a.fc;
    const int c0 = swap_rb ? 2 : 0, c2 = swap_rb ? 0 : 2;
    for (int i = tid; i < kBandRows * S; i += kPreThreads) {
      const int r = i / S, ox = i % S;
      const int sy = y1 + top + oy0 + r, sx = x1 + left + ox;
      int u0 = 0, u1 = 0, u2 = 0;                      // outside the frame: PIL crop pads with zeros
      if (sy >= 0 && sy < a.fh && sx >= 0 && sx < a.fw) {
        const unsigned char* px = frame + ((size_t)sy * a.fw + sx) * a.fc;
        u0 = px[c0]; u1 = px[1]; u2 = px[c2];
      }
      const size_t o = (size_t)crop * 3 * plane + (size_t)(oy0 + r) * S + ox;
      const float f0 = s_lut[u0], f1 = s_lut[256 + u1], f2 = s_lut[512 + u2];
      if (a.out_bf16) {
        a.out_bf16[o] = __float2bfloat16_rn(f0); a.out_bf16[o + plane] = __float2bfloat16_rn(f1);
        a.out_bf16[o + 2 * plane] = __float2bfloat16_rn(f2);
      }
      if (a.out_f32) { a.out_f32[o] = f0; a.out_f32[o + plane] = f1; a.out_f32[o + 2 * plane] = f2; }
    }
    return;
  }

  // ---- tap tables
  if (tid < S) {
    if (need_h) {
      h_n[tid] = resample_taps(left + tid, w, new_w, h_taps + tid * kMaxTaps, &h_xmin[tid]);
    } else {
      h_n[tid] = 1; h_xmin[tid] = left + tid; h_taps[tid * kMaxTaps] = 1 << kPrecBits;
    }
  } else if (tid < S + kBandRows) {
    const int r = tid - S;
    if (need_v) {
      v_n[r] = resample_taps(top + oy0 + r, h, new_h, v_taps + r * kMaxTaps, &v_ymin[r]);
    } else {
      v_n[r] = 1; v_ymin[r] = top + oy0 + r; v_taps[r * kMaxTaps] = 1 << kPrecBits;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int y_first = v_ymin[0];
    const int y_last = v_ymin[kBandRows - 1] + v_n[kBandRows - 1];
    s_info[0] = y_first;
    s_info[1] = y_last - y_first;
  }
  __syncthreads();
  const int y_first = s_info[0];
  int n_rows = s_info[1];
  if (n_rows > a.max_strip_rows) {          // cannot happen when the host bound is honoured; fail safe, no OOB
    if (band == 0 && tid == 0 && a.status) a.status[crop] = 2;
    n_rows = a.max_strip_rows;
  }

  // ---- horizontal pass into the strip (u8, [row][S][3])
  const unsigned char* frame = a.frames + (size_t)fi * a.fh * a.fw * a.fc;
  const int c0 = swap_rb ? 2 : 0, c2 = swap_rb ? 0 : 2;
  for (int i = tid; i < n_rows * S; i += kPreThreads) {
    const int r = i / S, ox = i % S;
    const int sy = y1 + y_first + r;                    // frame row
    const int n = h_n[ox], xm = h_xmin[ox];
    const int* tp = h_taps + ox * kMaxTaps;
    int acc0 = 1 << (kPrecBits - 1), acc1 = acc0, acc2 = acc0;
    if (sy >= 0 && sy < a.fh) {
      const unsigned char* rowp = frame + (size_t)sy * a.fw * a.fc;
      for (int t = 0; t < n; ++t) {
        const int sx = x1 + xm + t;
        if (sx >= 0 && sx < a.fw) {
          const unsigned char* px = rowp + (size_t)sx * a.fc;
          const int k = tp[t];
          acc0 += (int)px[c0] * k; acc1 += (int)px[1] * k; acc2 += (int)px[c2] * k;
        }
      }
    }
    unsigned char* d = strip + ((size_t)r * S + ox) * 3;
    d[0] = (unsigned char)clip8(acc0); d[1] = (unsigned char)clip8(acc1); d[2] = (unsigned char)clip8(acc2);
  }
  __syncthreads();

  // ---- vertical pass + normalise + store (ox fastest -> coalesced)
  for (int i = tid; i < 3 * kBandRows * S; i += kPreThreads) {
    const int c = i / (kBandRows * S), r = (i / S) % kBandRows, ox = i % S;
    const int n = v_n[r], ym = v_ymin[r] - y_first;
    const int* tp = v_taps + r * kMaxTaps;
    int acc = 1 << (kPrecBits - 1);
    for (int t = 0; t < n; ++t) {
      const int rr = ym + t;
      if (rr < n_rows) acc += (int)strip[((size_t)rr * S + ox) * 3 + c] * tp[t];
    }
    const int u = clip8(acc);
    // ToTensor (/255) then Normalize ((x - mean) / std), each rounded to fp32 like torch (tabulated above)
    const float f = s_lut[c * 256 + u];
    const size_t o = (size_t)crop * 3 * plane + (size_t)c * plane + (size_t)(oy0 + r) * S + ox;
    if (a.out_bf16) a.out_bf16[o] = __float2bfloat16_rn(f);
    if (a.out_f32) a.out_f32[o] = f;
  }
}

int preprocess_crops(const unsigned char* frames, int n_frames, int frame_h, int frame_w, int frame_c,
                     const int* rects, const int* frame_idx, const int* flags, int n, int size, bf16* out_bf16,
                     float* out_f32, int* status, cudaStream_t stream) {
  EC_REQUIRE(size % kBandRows == 0 && size > 0 && size + kBandRows <= kPreThreads, EC_ERR_UNSUPPORTED,
             "preprocess: output size %d unsupported", size);
  EC_REQUIRE(frame_c == 3 || frame_c == 4, EC_ERR_INVALID, "preprocess: frames must have 3 or 4 channels");
  EC_REQUIRE(n <= 65535, EC_ERR_UNSUPPORTED, "preprocess: more than 65535 crops per call (chunk the call)");
  if (n <= 0) return EC_OK;
  // strip rows needed by an 8-row band at the largest supported down-scale (kMaxTaps taps): 8*s + 2*support + 2
  const double s_max = (double)(kMaxTaps - 1) / 4.0;
  const int max_strip_rows = (int)(kBandRows * s_max + 4.0 * s_max) + 4;
  const size_t smem = (size_t)(size * kMaxTaps + 2 * size + kBandRows * kMaxTaps + 2 * kBandRows) * sizeof(int) +
                      (size_t)max_strip_rows * size * 3;
  static size_t configured_dev[64] = {};
  const int dev_ = current_device();
  EC_REQUIRE(dev_ >= 0, EC_ERR_DEVICE, "no current CUDA device");
  size_t& configured = configured_dev[dev_];
  if (smem > configured) {
    EC_CHECK_CUDA(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  PreArgs a;
  a.frames = frames; a.n_frames = n_frames; a.fh = frame_h; a.fw = frame_w; a.fc = frame_c;
  a.rects = rects; a.frame_idx = frame_idx; a.flags = flags; a.n = n; a.size = size;
  a.out_bf16 = out_bf16; a.out_f32 = out_f32; a.status = status; a.max_strip_rows = max_strip_rows;
  dim3 grid(size / kBandRows, n);
  preprocess_kernel<<<grid, kPreThreads, smem, stream>>>(a);
  EC_POST_LAUNCH();
  return EC_OK;
}

}  // namespace ec
