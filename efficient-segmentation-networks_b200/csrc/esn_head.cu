// Segmentation heads (logits in the caller's NCHW layout and/or the uint8 argmax mask) and the
// weighted cross-entropy loss.  All are HBM-bound: one thread per (low-res or output) pixel,
// stores coalesced along W inside each class plane.
#include "esn_common.cuh"

namespace {

constexpr int kMaxClasses = 32;

template <typename T> __device__ __forceinline__ void st2(T* p, float a, float b);
template <> __device__ __forceinline__ void st2<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void st2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// ---------------------------------------------------------------- ConvTranspose2d(Cin, classes, 2, stride 2)
struct ConvtHeadArgs {
  const void* x;
  const float* w;     // [2][2][Cin][32]
  const float* bias;  // [classes]
  void* logits;       // NCHW or null
  uint8_t* mask;      // or null
  int N, Hi, Wi, Cin, x_cs, classes;
};

template <typename TI, typename TL, int CIN>
__global__ void __launch_bounds__(128) convt2x2_head_kernel(const ConvtHeadArgs a) {
  extern __shared__ float sw[];  // [4][CIN][32] + bias[32]
  float* sb = sw + 4 * CIN * 32;
  for (int i = threadIdx.x; i < 4 * CIN * 32; i += blockDim.x) sw[i] = a.w[i];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) sb[i] = i < a.classes ? a.bias[i] : 0.f;
  __syncthreads();
  const long long total = (long long)a.N * a.Hi * a.Wi;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % a.Wi);
  const int i = (int)((idx / a.Wi) % a.Hi);
  const int n = (int)(idx / ((long long)a.Wi * a.Hi));
  float f[CIN];
  const TI* xp = reinterpret_cast<const TI*>(a.x) + (size_t)idx * a.x_cs;
#pragma unroll
  for (int c = 0; c < CIN; c += 4) {
    const float4 t = ld4<TI>(xp + c);
    f[c] = t.x;
    f[c + 1] = t.y;
    f[c + 2] = t.z;
    f[c + 3] = t.w;
  }
  const int Ho = 2 * a.Hi, Wo = 2 * a.Wi;
  const int ncls4 = (a.classes + 3) / 4;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float acc[2][kMaxClasses];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
#pragma unroll
      for (int q = 0; q < kMaxClasses / 4; ++q) {
        if (q < ncls4) {
          float4 v = *reinterpret_cast<const float4*>(sb + 4 * q);
          const float* wp = sw + (size_t)((r * 2 + s) * CIN) * 32 + 4 * q;
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + c * 32);
            v.x += f[c] * wv.x;
            v.y += f[c] * wv.y;
            v.z += f[c] * wv.z;
            v.w += f[c] * wv.w;
          }
          acc[s][4 * q] = v.x;
          acc[s][4 * q + 1] = v.y;
          acc[s][4 * q + 2] = v.z;
          acc[s][4 * q + 3] = v.w;
        }
      }
    }
    const int ho = 2 * i + r;
    if (a.logits) {
      TL* lp = reinterpret_cast<TL*>(a.logits);
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < a.classes)
          st2<TL>(lp + ((size_t)((size_t)n * a.classes + k) * Ho + ho) * Wo + 2 * j, acc[0][k], acc[1][k]);
    }
    if (a.mask) {
      int b0 = 0, b1 = 0;
      float m0 = acc[0][0], m1 = acc[1][0];
#pragma unroll
      for (int k = 1; k < kMaxClasses; ++k)
        if (k < a.classes) {
          if (acc[0][k] > m0) { m0 = acc[0][k]; b0 = k; }
          if (acc[1][k] > m1) { m1 = acc[1][k]; b1 = k; }
        }
      *reinterpret_cast<uchar2*>(a.mask + ((size_t)n * Ho + ho) * Wo + 2 * j) = make_uchar2((uint8_t)b0, (uint8_t)b1);
    }
  }
}

// Vector variant: one thread = 4 consecutive input pixels -> 8 consecutive outputs in each of the two output
// rows, so every class plane gets 16-byte (bf16) / 2x16-byte (fp32) stores and each shared-memory weight
// vector feeds 16 FMAs instead of 4.
template <typename TL> __device__ __forceinline__ void store8(TL* p, const float* v);
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  *reinterpret_cast<uint4*>(p) = float_to_bf16x8(v);
}
template <> __device__ __forceinline__ void store8<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

template <typename TI, typename TL, int CIN>
__global__ void __launch_bounds__(128) convt2x2_head_vec_kernel(const ConvtHeadArgs a) {
  extern __shared__ float sw[];  // [4][CIN][32] + bias[32]
  float* sb = sw + 4 * CIN * 32;
  for (int i = threadIdx.x; i < 4 * CIN * 32; i += blockDim.x) sw[i] = a.w[i];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) sb[i] = i < a.classes ? a.bias[i] : 0.f;
  __syncthreads();
  const int wq = a.Wi / 4;
  const long long total = (long long)a.N * a.Hi * wq;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j0 = (int)(idx % wq) * 4;
  const int i = (int)((idx / wq) % a.Hi);
  const int n = (int)(idx / ((long long)wq * a.Hi));
  float f[4][CIN];
  const TI* xp = reinterpret_cast<const TI*>(a.x) + ((size_t)((size_t)n * a.Hi + i) * a.Wi + j0) * a.x_cs;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < CIN; c += 4) {
      const float4 t = ld4<TI>(xp + (size_t)p * a.x_cs + c);
      f[p][c] = t.x; f[p][c + 1] = t.y; f[p][c + 2] = t.z; f[p][c + 3] = t.w;
    }
  const int Ho = 2 * a.Hi, Wo = 2 * a.Wi;
  const int ncls4 = (a.classes + 3) / 4;
#pragma unroll 1
  for (int r = 0; r < 2; ++r) {
    const int ho = 2 * i + r;
    float best[8];
    int bi[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { best[o] = -INFINITY; bi[o] = 0; }
#pragma unroll 1
    for (int q = 0; q < ncls4; ++q) {
      float2 acc2[2][8];  // [class pair in quad][output pixel = 2*p + s]: packed FFMA2, two classes per instruction
      const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);
#pragma unroll
      for (int o = 0; o < 8; ++o) { acc2[0][o] = make_float2(b.x, b.y); acc2[1][o] = make_float2(b.z, b.w); }
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + (size_t)((r * 2 + 0) * CIN + c) * 32 + 4 * q);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + (size_t)((r * 2 + 1) * CIN + c) * 32 + 4 * q);
        const float2 w0a = make_float2(w0.x, w0.y), w0b = make_float2(w0.z, w0.w);
        const float2 w1a = make_float2(w1.x, w1.y), w1b = make_float2(w1.z, w1.w);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 vv = make_float2(f[p][c], f[p][c]);
          acc2[0][2 * p] = ffma2(vv, w0a, acc2[0][2 * p]);
          acc2[1][2 * p] = ffma2(vv, w0b, acc2[1][2 * p]);
          acc2[0][2 * p + 1] = ffma2(vv, w1a, acc2[0][2 * p + 1]);
          acc2[1][2 * p + 1] = ffma2(vv, w1b, acc2[1][2 * p + 1]);
        }
      }
      float acc[4][8];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        acc[0][o] = acc2[0][o].x; acc[1][o] = acc2[0][o].y; acc[2][o] = acc2[1][o].x; acc[3][o] = acc2[1][o].y;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = 4 * q + kk;
        if (k < a.classes) {
          if (a.logits)
            store8<TL>(reinterpret_cast<TL*>(a.logits) + ((size_t)((size_t)n * a.classes + k) * Ho + ho) * Wo + 2 * j0, acc[kk]);
#pragma unroll
          for (int o = 0; o < 8; ++o)
            if (acc[kk][o] > best[o]) { best[o] = acc[kk][o]; bi[o] = k; }
        }
      }
    }
    if (a.mask) {
      uint8_t m[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) m[o] = (uint8_t)bi[o];
      *reinterpret_cast<uint2*>(a.mask + ((size_t)n * Ho + ho) * Wo + 2 * j0) = *reinterpret_cast<const uint2*>(m);
    }
  }
}

// ---------------------------------------------------------------- bilinear (align_corners=False) head
struct BilinearHeadArgs {
  const void* x;  // NHWC low-res scores
  void* logits;
  uint8_t* mask;
  int N, Hi, Wi, x_cs, classes, Ho, Wo;
  float sh, sw;   // in/out  (align_corners: (in-1)/(out-1))
  int align;
};

template <typename TI, typename TL>
__global__ void __launch_bounds__(256) bilinear_head_kernel(const BilinearHeadArgs a) {
  const long long total = (long long)a.N * a.Ho * a.Wo;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % a.Wo);
  const int ho = (int)((idx / a.Wo) % a.Ho);
  const int n = (int)(idx / ((long long)a.Wo * a.Ho));
  // torch area_pixel_compute_source_index, align_corners=False, cubic=False
  float fh, fw;
  if (a.align) {
    fh = a.sh * ho;
    fw = a.sw * wo;
  } else {
    fh = a.sh * (ho + 0.5f) - 0.5f;
    fh = fh < 0.f ? 0.f : fh;
    fw = a.sw * (wo + 0.5f) - 0.5f;
    fw = fw < 0.f ? 0.f : fw;
  }
  const int h0 = min((int)fh, a.Hi - 1), w0 = min((int)fw, a.Wi - 1);
  const int hp = (h0 < a.Hi - 1) ? 1 : 0, wp = (w0 < a.Wi - 1) ? 1 : 0;
  const float lh1 = fh - h0, lh0 = 1.f - lh1, lw1 = fw - w0, lw0 = 1.f - lw1;
  const TI* x = reinterpret_cast<const TI*>(a.x);
  const TI* p00 = x + ((size_t)((size_t)n * a.Hi + h0) * a.Wi + w0) * a.x_cs;
  const TI* p01 = p00 + (size_t)wp * a.x_cs;
  const TI* p10 = p00 + (size_t)hp * a.Wi * a.x_cs;
  const TI* p11 = p10 + (size_t)wp * a.x_cs;
  float best = 0.f;
  int bi = 0;
  TL* lp = reinterpret_cast<TL*>(a.logits);
  for (int k = 0; k < a.classes; ++k) {
    const float v = lh0 * (lw0 * ld1<TI>(p00 + k) + lw1 * ld1<TI>(p01 + k)) +
                    lh1 * (lw0 * ld1<TI>(p10 + k) + lw1 * ld1<TI>(p11 + k));
    if (lp) st1<TL>(lp + ((size_t)((size_t)n * a.classes + k) * a.Ho + ho) * a.Wo + wo, v);
    if (k == 0 || v > best) { best = v; bi = k; }
  }
  if (a.mask) a.mask[idx] = (uint8_t)bi;
}

// Vector variant: one thread = PX consecutive output pixels of one row (16-byte stores into every class
// plane, PX bytes of mask).  The two source rows are blended once per source column and the column pair is
// rolled along the row, so an 8x upsample reads 3 class vectors per 8 outputs instead of 32.
template <typename TI, int NC> __device__ __forceinline__ void load_classes(const TI* p, float* f);
template <int NC> __device__ __forceinline__ void load_classes_f32(const float* p, float* f) {
#pragma unroll
  for (int k = 0; k < NC; k += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + k));
    f[k] = t.x; f[k + 1] = t.y; f[k + 2] = t.z; f[k + 3] = t.w;
  }
}
template <int NC> __device__ __forceinline__ void load_classes_bf16(const __nv_bfloat16* p, float* f) {
#pragma unroll
  for (int k = 0; k < NC; k += 4) {
    const float4 t = ld4<__nv_bfloat16>(p + k);
    f[k] = t.x; f[k + 1] = t.y; f[k + 2] = t.z; f[k + 3] = t.w;
  }
}
template <int NC> struct ClsLoad {
  static __device__ __forceinline__ void ld(const float* p, float* f) { load_classes_f32<NC>(p, f); }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) { load_classes_bf16<NC>(p, f); }
};

template <typename TL, int PX> struct RowStore;
template <> struct RowStore<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) { *reinterpret_cast<uint4*>(p) = float_to_bf16x8(v); }
};
template <> struct RowStore<float, 4> {
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <typename TI, typename TL, int NC, int PX>
__global__ void __launch_bounds__(128) bilinear_head_vec_kernel(const BilinearHeadArgs a) {
  const int wo0 = (blockIdx.x * 128 + threadIdx.x) * PX;
  if (wo0 >= a.Wo) return;
  const int ho = blockIdx.y, n = blockIdx.z;
  float fh;
  if (a.align) {
    fh = a.sh * ho;
  } else {
    fh = a.sh * (ho + 0.5f) - 0.5f;
    fh = fh < 0.f ? 0.f : fh;
  }
  const int h0 = min((int)fh, a.Hi - 1);
  const int hp = (h0 < a.Hi - 1) ? 1 : 0;
  const float lh1 = fh - h0, lh0 = 1.f - lh1;
  const TI* r0 = reinterpret_cast<const TI*>(a.x) + ((size_t)((size_t)n * a.Hi + h0) * a.Wi) * a.x_cs;
  const TI* r1 = r0 + (size_t)hp * a.Wi * a.x_cs;
  float c0[NC], c1[NC];
  float out[NC][PX];
  int curw = -2;
  auto blend = [&](int w, float* c) {
    float t[NC], b[NC];
    ClsLoad<NC>::ld(r0 + (size_t)w * a.x_cs, t);
    ClsLoad<NC>::ld(r1 + (size_t)w * a.x_cs, b);
#pragma unroll
    for (int k = 0; k < NC; ++k) c[k] = lh0 * t[k] + lh1 * b[k];
  };
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    const int wo = wo0 + j;
    float fw;
    if (a.align) {
      fw = a.sw * wo;
    } else {
      fw = a.sw * (wo + 0.5f) - 0.5f;
      fw = fw < 0.f ? 0.f : fw;
    }
    const int w0 = min((int)fw, a.Wi - 1);
    const int wp = (w0 < a.Wi - 1) ? 1 : 0;
    const float lw1 = fw - w0, lw0 = 1.f - lw1;
    if (w0 != curw) {
      if (w0 == curw + 1) {
#pragma unroll
        for (int k = 0; k < NC; ++k) c0[k] = c1[k];
      } else {
        blend(w0, c0);
      }
      if (wp) {
        blend(w0 + 1, c1);
      } else {
#pragma unroll
        for (int k = 0; k < NC; ++k) c1[k] = c0[k];
      }
      curw = w0;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) out[k][j] = lw0 * c0[k] + lw1 * c1[k];
  }
  if (a.logits) {
    TL* lp = reinterpret_cast<TL*>(a.logits) + ((size_t)n * a.classes * a.Ho + ho) * a.Wo + wo0;
#pragma unroll
    for (int k = 0; k < NC; ++k)
      if (k < a.classes) RowStore<TL, PX>::st(lp + (size_t)k * a.Ho * a.Wo, out[k]);
  }
  if (a.mask) {
    uint8_t m[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      float best = out[0][j];
      int bi = 0;
#pragma unroll
      for (int k = 1; k < NC; ++k)
        if (k < a.classes && out[k][j] > best) { best = out[k][j]; bi = k; }
      m[j] = (uint8_t)bi;
    }
    uint8_t* mp = a.mask + ((size_t)n * a.Ho + ho) * a.Wo + wo0;
    if (PX == 8) *reinterpret_cast<uint2*>(mp) = *reinterpret_cast<const uint2*>(m);
    else *reinterpret_cast<uint32_t*>(mp) = *reinterpret_cast<const uint32_t*>(m);
  }
}

template <typename TI, typename TL, int PX>
void launch_bilinear_vec(const BilinearHeadArgs& a, cudaStream_t st) {
  dim3 grid((unsigned)((a.Wo / PX + 127) / 128), (unsigned)a.Ho, (unsigned)a.N);
  if (a.classes <= 20) bilinear_head_vec_kernel<TI, TL, 20, PX><<<grid, 128, 0, st>>>(a);
  else bilinear_head_vec_kernel<TI, TL, 32, PX><<<grid, 128, 0, st>>>(a);
}

// ---------------------------------------------------------------- weighted cross-entropy on NCHW logits
struct CEArgs {
  const void* logits;
  const long long* target;
  const float* weight;
  float* sums;
  void* dlogits;
  int N, C, H, W, ignore;
  const float* gnorm;
  const float* gout;
  float* prob_out;           // optional: softmax probability of the target class per pixel (1 where ignored)
  const float* keep_thresh;  // optional device scalar: pixels whose target probability exceeds it count as ignored (OHEM)
};

template <typename T>
__global__ void __launch_bounds__(256) weighted_ce_kernel(const CEArgs a) {
  const long long hw = (long long)a.H * a.W;
  const long long total = (long long)a.N * hw;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float wl = 0.f, wsum = 0.f;
  if (idx < total) {
    const int n = (int)(idx / hw);
    const long long px = idx % hw;
    const T* lp = reinterpret_cast<const T*>(a.logits) + (size_t)n * a.C * hw + px;
    const long long y = a.target[idx];
    bool valid = (y != a.ignore) && y >= 0 && y < a.C;
    float v[kMaxClasses];
    float m = -INFINITY, xy = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < a.C) {
        v[k] = ld1<T>(lp + (size_t)k * hw);
        m = fmaxf(m, v[k]);
        if (k == (int)y) xy = v[k];
      }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < a.C) {
        v[k] = __expf(v[k] - m);
        s += v[k];
      }
    if (a.prob_out || a.keep_thresh) {
      // OHEM (utils/losses/loss.py:189-206): probability of the labelled class, 1 for ignored pixels; the same expression in
      // the recording pass and in the thresholded pass, so both see identical bits
      const float py = valid ? __expf(xy - m) / s : 1.f;
      if (a.prob_out) a.prob_out[idx] = py;
      if (a.keep_thresh && !(py <= __ldg(a.keep_thresh))) valid = false;
    }
    const float wy = valid ? (a.weight ? __ldg(a.weight + y) : 1.f) : 0.f;
    if (valid) {
      wl = wy * (m + logf(s) - xy);  // w * (lse - x_y)
      wsum = wy;
    }
    if (a.dlogits) {
      T* gp = reinterpret_cast<T*>(a.dlogits) + (size_t)n * a.C * hw + px;
      const float gs = (a.gout ? __ldg(a.gout) : 1.f) / (a.gnorm ? __ldg(a.gnorm) : 1.f);
      const float inv = wy / s * gs;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < a.C) st1<T>(gp + (size_t)k * hw, v[k] * inv - ((valid && k == (int)y) ? wy * gs : 0.f));
    }
  }
  // block reduction: warp shuffle -> smem -> one atomic pair per CTA
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wl += __shfl_xor_sync(0xffffffffu, wl, o);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = wl;
    red[1][warp] = wsum;
  }
  __syncthreads();
  if (warp == 0) {
    wl = lane < 8 ? red[0][lane] : 0.f;
    wsum = lane < 8 ? red[1][lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      wl += __shfl_xor_sync(0xffffffffu, wl, o);
      wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if (lane == 0) {
      atomicAdd(a.sums, wl);
      atomicAdd(a.sums + 1, wsum);
    }
  }
}

}  // namespace

extern "C" int esn_head_convt2x2(const EsnHead* p, void* stream) {
  if (!p || !p->w || !p->bias || !esn_valid_nhwc(p->x)) return ESN_ERR_BAD_ARG;
  if (!p->logits.ptr && !p->mask) return ESN_ERR_BAD_ARG;
  if (p->classes < 1 || p->classes > kMaxClasses) return ESN_ERR_UNSUPPORTED;
  const EsnTensor& x = p->x;
  if (p->out_h != 2 * x.h || p->out_w != 2 * x.w) return ESN_ERR_BAD_SHAPE;
  if ((x.c != 16 && x.c != 20) || x.c_stride % 4) return ESN_ERR_UNSUPPORTED;   // 20 = 19 classes + one zero channel
  if ((uintptr_t)x.ptr % 16) return ESN_ERR_ALIGN;
  int ldt = ESN_F32;
  if (p->logits.ptr) {
    const EsnTensor& l = p->logits;
    if (l.layout != ESN_NCHW || (l.dtype != ESN_F32 && l.dtype != ESN_BF16)) return ESN_ERR_BAD_ARG;
    if (l.n != x.n || l.c != p->classes || l.h != p->out_h || l.w != p->out_w) return ESN_ERR_BAD_SHAPE;
    ldt = l.dtype;
  }
  ConvtHeadArgs a;
  a.x = x.ptr;
  a.w = p->w;
  a.bias = p->bias;
  a.logits = p->logits.ptr;
  a.mask = p->mask;
  a.N = x.n;
  a.Hi = x.h;
  a.Wi = x.w;
  a.Cin = x.c;
  a.x_cs = x.c_stride;
  a.classes = p->classes;
  const long long total = (long long)x.n * x.h * x.w;
  const int block = 128, grid = esn_cdiv(total, block);
  const size_t smem = (4 * (size_t)x.c * 32 + 32) * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ESN_HEAD_LAUNCH(CIN)                                                                   \
  do {                                                                                         \
    if (x.dtype == ESN_F32 && ldt == ESN_F32)                                                  \
      convt2x2_head_kernel<float, float, CIN><<<grid, block, smem, st>>>(a);                   \
    else if (x.dtype == ESN_F32)                                                               \
      convt2x2_head_kernel<float, __nv_bfloat16, CIN><<<grid, block, smem, st>>>(a);           \
    else if (ldt == ESN_F32)                                                                   \
      convt2x2_head_kernel<__nv_bfloat16, float, CIN><<<grid, block, smem, st>>>(a);           \
    else                                                                                       \
      convt2x2_head_kernel<__nv_bfloat16, __nv_bfloat16, CIN><<<grid, block, smem, st>>>(a);   \
  } while (0)
  const bool vec = x.w % 4 == 0 && (!p->logits.ptr || (uintptr_t)p->logits.ptr % 16 == 0) && (!p->mask || (uintptr_t)p->mask % 8 == 0);
  if (vec) {
    const int vgrid = esn_cdiv(total / 4, block);
#define ESN_HEADV_LAUNCH(CIN)                                                                       \
  do {                                                                                              \
    if (x.dtype == ESN_F32 && ldt == ESN_F32)                                                       \
      convt2x2_head_vec_kernel<float, float, CIN><<<vgrid, block, smem, st>>>(a);                   \
    else if (x.dtype == ESN_F32)                                                                    \
      convt2x2_head_vec_kernel<float, __nv_bfloat16, CIN><<<vgrid, block, smem, st>>>(a);           \
    else if (ldt == ESN_F32)                                                                        \
      convt2x2_head_vec_kernel<__nv_bfloat16, float, CIN><<<vgrid, block, smem, st>>>(a);           \
    else                                                                                            \
      convt2x2_head_vec_kernel<__nv_bfloat16, __nv_bfloat16, CIN><<<vgrid, block, smem, st>>>(a);   \
  } while (0)
    if (x.c == 16) ESN_HEADV_LAUNCH(16);
    else ESN_HEADV_LAUNCH(20);
#undef ESN_HEADV_LAUNCH
  } else if (x.c == 16) {
    ESN_HEAD_LAUNCH(16);
  } else {
    ESN_HEAD_LAUNCH(20);
  }
#undef ESN_HEAD_LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_head_bilinear(const EsnHead* p, void* stream) {
  if (!p || !esn_valid_nhwc(p->x)) return ESN_ERR_BAD_ARG;
  if (!p->logits.ptr && !p->mask) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  if (p->classes != x.c || p->out_h < 1 || p->out_w < 1) return ESN_ERR_BAD_SHAPE;
  int ldt = ESN_F32;
  if (p->logits.ptr) {
    const EsnTensor& l = p->logits;
    if (l.layout != ESN_NCHW || (l.dtype != ESN_F32 && l.dtype != ESN_BF16)) return ESN_ERR_BAD_ARG;
    if (l.n != x.n || l.c != p->classes || l.h != p->out_h || l.w != p->out_w) return ESN_ERR_BAD_SHAPE;
    ldt = l.dtype;
  }
  BilinearHeadArgs a;
  a.x = x.ptr;
  a.logits = p->logits.ptr;
  a.mask = p->mask;
  a.N = x.n;
  a.Hi = x.h;
  a.Wi = x.w;
  a.x_cs = x.c_stride;
  a.classes = p->classes;
  a.Ho = p->out_h;
  a.Wo = p->out_w;
  a.align = p->align_corners;
  if (a.align) {
    a.sh = p->out_h > 1 ? (float)(x.h - 1) / (float)(p->out_h - 1) : 0.f;
    a.sw = p->out_w > 1 ? (float)(x.w - 1) / (float)(p->out_w - 1) : 0.f;
  } else {
    a.sh = (float)x.h / (float)p->out_h;
    a.sw = (float)x.w / (float)p->out_w;
  }
  const long long total = (long long)x.n * p->out_h * p->out_w;
  const int block = 256, grid = esn_cdiv(total, block);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {  // vector path: class vectors readable as whole 4-channel groups, rows storable as 16-byte pieces
    const int nc = p->classes <= 20 ? 20 : 32;
    const int px = ldt == ESN_BF16 ? 8 : 4;
    const size_t xsz = x.dtype == ESN_F32 ? 4 : 2;
    const bool ok = p->classes <= 32 && x.c_stride >= nc && x.c_stride % 4 == 0 && (uintptr_t)x.ptr % (4 * xsz) == 0 &&
                    p->out_w % px == 0 && p->out_h <= 65535 && x.n <= 65535 &&
                    (!p->logits.ptr || (uintptr_t)p->logits.ptr % 16 == 0) && (!p->mask || (uintptr_t)p->mask % 8 == 0);
    if (ok) {
      if (x.dtype == ESN_F32 && ldt == ESN_F32) launch_bilinear_vec<float, float, 4>(a, st);
      else if (x.dtype == ESN_F32) launch_bilinear_vec<float, __nv_bfloat16, 8>(a, st);
      else if (ldt == ESN_F32) launch_bilinear_vec<__nv_bfloat16, float, 4>(a, st);
      else launch_bilinear_vec<__nv_bfloat16, __nv_bfloat16, 8>(a, st);
      ESN_CHECK_LAUNCH();
      return ESN_OK;
    }
  }
  if (x.dtype == ESN_F32 && ldt == ESN_F32)
    bilinear_head_kernel<float, float><<<grid, block, 0, st>>>(a);
  else if (x.dtype == ESN_F32)
    bilinear_head_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>(a);
  else if (ldt == ESN_F32)
    bilinear_head_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>(a);
  else
    bilinear_head_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_weighted_ce(const EsnCE* p, void* stream) {
  if (!p || !p->logits.ptr || !p->target || !p->sums) return ESN_ERR_BAD_ARG;
  const EsnTensor& l = p->logits;
  if (l.layout != ESN_NCHW || (l.dtype != ESN_F32 && l.dtype != ESN_BF16)) return ESN_ERR_BAD_ARG;
  if (l.c < 1 || l.c > kMaxClasses) return ESN_ERR_UNSUPPORTED;
  if (p->dlogits.ptr) {
    const EsnTensor& g = p->dlogits;
    if (g.layout != ESN_NCHW || g.dtype != l.dtype) return ESN_ERR_BAD_ARG;
    if (g.n != l.n || g.c != l.c || g.h != l.h || g.w != l.w) return ESN_ERR_BAD_SHAPE;
  }
  CEArgs a;
  a.logits = l.ptr;
  a.target = reinterpret_cast<const long long*>(p->target);
  a.weight = p->weight;
  a.sums = p->sums;
  a.dlogits = p->dlogits.ptr;
  a.N = l.n;
  a.C = l.c;
  a.H = l.h;
  a.W = l.w;
  a.ignore = p->ignore_label;
  a.gnorm = p->gnorm;
  a.gout = p->gout;
  a.prob_out = p->prob_out;
  a.keep_thresh = p->keep_thresh;
  const long long total = (long long)l.n * l.h * l.w;
  const int block = 256, grid = esn_cdiv(total, block);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (l.dtype == ESN_F32)
    weighted_ce_kernel<float><<<grid, block, 0, st>>>(a);
  else
    weighted_ce_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
