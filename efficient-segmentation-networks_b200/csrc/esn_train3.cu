// Backward of ENet's pooling pair (training path; reference: autograd of nn.MaxPool2d(3, 2, 1, return_indices=True) and
// nn.MaxUnpool2d(2), model/ENet.py:126-130, 225, 262 reached from loss.backward() at train.py:353).
//   esn_maxpool3x3s2_idx_bwd : dx[n,h,w,c] (+)= sum of dy over the (at most four) windows whose recorded arg-max is (h,w) --
//                              written as a GATHER over the input pixels, so it is deterministic and needs no atomics
//                              although the 3x3 / stride-2 windows overlap
//   esn_max_unpool2x2_bwd    : dv[n,i,j,c] = dy[n, idx / W2, idx % W2, c] -- every pooled cell reads the gradient at the
//                              position it was scattered to (cells that lost a collision in the forward still receive it:
//                              that is what torch's max_unpool2d backward does)
// HBM-bound elementwise passes, 8 channels (16 bytes of bf16) per thread where the views allow.
#include "esn_common.cuh"

namespace {

template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool3x3s2_idx_bwd_kernel(const T* __restrict__ dy, const int32_t* __restrict__ idx,
                                                                   T* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo,
                                                                   int dy_cs, int dx_cs, int accumulate) {
  const int cg = C / V;
  const long long total = (long long)N * H * W * cg;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % cg) * V;
  const long long pix = t / cg;
  const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
  const int target = h * W + w;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = accumulate ? ld1<T>(dx + pix * dx_cs + c + k) : 0.f;
  // windows (i, j) cover input rows 2i-1 .. 2i+1: i in {h/2, (h+1)/2} (equal for even h), same for columns
  const int i0 = h >> 1, i1 = (h + 1) >> 1, j0 = w >> 1, j1 = (w + 1) >> 1;
  for (int ii = 0; ii < 2; ++ii) {
    const int i = ii ? i1 : i0;
    if ((ii && i1 == i0) || i >= Ho) continue;
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jj ? j1 : j0;
      if ((jj && j1 == j0) || j >= Wo) continue;
      const long long q = ((long long)n * Ho + i) * Wo + j;
#pragma unroll
      for (int k = 0; k < V; ++k)
        if (__ldg(idx + q * C + c + k) == target) acc[k] += ld1<T>(dy + q * dy_cs + c + k);
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) st1<T>(dx + pix * dx_cs + c + k, acc[k]);
}

template <typename T, int V>
__global__ void __launch_bounds__(256) max_unpool2x2_bwd_kernel(const T* __restrict__ dy, const int32_t* __restrict__ idx,
                                                                T* __restrict__ dv, long long npix_lo, int C, int Hp, int Wp,
                                                                int dy_cs, int dv_cs) {
  const int cg = C / V;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= npix_lo * cg) return;
  const int c = (int)(t % cg) * V;
  const long long q = t / cg;
  const long long n = q / ((long long)Hp * Wp);
  const long long plane = n * (4LL * Hp * Wp);          // output plane of this image: (2Hp) x (2Wp) pixels
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int p = __ldg(idx + q * C + c + k);
    st1<T>(dv + q * dv_cs + c + k, ld1<T>(dy + (plane + p) * dy_cs + c + k));
  }
}

// ---------------------------------------------------------------- bilinear backward, separable and coalesced
// d low[n,h,w,c] = sum over output pixels (ho, wo) of d high[n,c,ho,wo] * weight.  The weight of (ho, wo) on source pixel
// (h, w) is a product of two hat functions of the clamped source coordinates, max(0, 1 - |fh - h|) * max(0, 1 - |fw - w|), with
// f(o) = clamp(s * o + t, 0, size - 1): s = in / out, t = s / 2 - 1/2 for align_corners = False; s = (in - 1) / (out - 1), t = 0
// for align_corners = True -- the same numbers as (1 - frac, frac) on (floor, floor + 1), both border clamps included.
// A CTA owns (n, c, source row h, kTWL source columns).  Pass 1: every thread takes one output column of the CTA's window
// and sums it over the ~2/s output rows under the vertical hat -- consecutive threads read consecutive addresses of the NCHW
// gradient, all loads of a thread issued before the first use (the per-element kernels have neighbouring lanes 1/s floats
// apart and one load in flight: 1 TB/s on DABNet's 318 MB of d logits, 4.4 ms of Fast-SCNN's step).  Pass 2: kTWL threads
// finish the horizontal hat from shared memory.  Replaces upsample_bilinear2d_backward of DABNet.py:181 / FastSCNN.py:233.
constexpr int kTWL = 32;
constexpr int kRB = 24;
constexpr int kMaxRows = 96;      // output rows under one vertical hat (2 / s + 3): up-sampling factors up to ~46
template <typename TL, typename TO, int HG>
__global__ void __launch_bounds__(288) bilinear_bwd_rows_kernel(const TL* __restrict__ dl, TO* __restrict__ dlow, int C, int Hi, int Wi,
                                                                int Ho, int Wo, int low_cs, float sh, float th, float sw, float tw,
                                                                float gscale, int wtiles, int hgroups, int window, int accumulate) {
  // HG consecutive source rows per CTA: their vertical hats overlap by half, so the union of output rows is (HG + 1) / s
  // instead of 2 HG / s -- each output row is loaded once and feeds (at most two of) the HG column sums
  extern __shared__ float colsum[];           // [HG][window]
  const int wt = blockIdx.x % wtiles;
  const int hg = (blockIdx.x / wtiles) % hgroups;
  const int c = (blockIdx.x / (wtiles * hgroups)) % C;
  const int n = blockIdx.x / (wtiles * hgroups * C);
  const int h0 = hg * HG, h1 = min(h0 + HG, Hi) - 1;
  const int w0 = wt * kTWL, w1 = min(w0 + kTWL, Wi) - 1;
  const float rh = 1.f / sh, rw = 1.f / sw;
  // outputs o with |s*o + t - h| < 1 for some h of the group, one extra on each side for rounding; the border clamps only
  // add outputs that the image bounds cut off anyway
  const int ho0 = max((int)ceilf(((float)h0 - 1.f - th) * rh) - 1, 0), ho1 = min((int)floorf(((float)h1 + 1.f - th) * rh) + 1, Ho - 1);
  const int wlo = max((int)ceilf(((float)w0 - 1.f - tw) * rw) - 1, 0), whi = min((int)floorf(((float)w1 + 1.f - tw) * rw) + 1, Wo - 1);
  const TL* plane = dl + ((size_t)n * C + c) * Ho * Wo;
  const float hmax = (float)(Hi - 1), wmax = (float)(Wi - 1);
  // the vertical weights are the same for every thread of the CTA: computed once (ncu: the first version of this kernel
  // rebuilt them per element and was instruction-bound -- 295 M warp instructions, SM pipes 85 % busy at 1 TB/s)
  __shared__ float wh_s[HG][kMaxRows];
  const int nrows = ho1 - ho0 + 1;
  for (int i = threadIdx.x; i < HG * nrows; i += blockDim.x) {
    const int j = i / nrows, r = i - j * nrows;
    const float fh = fminf(fmaxf(fmaf(sh, (float)(ho0 + r), th), 0.f), hmax);
    wh_s[j][r] = (h0 + j <= h1) ? fmaxf(1.f - fabsf(fh - (float)(h0 + j)), 0.f) : 0.f;
  }
  __syncthreads();
  for (int wo = wlo + threadIdx.x; wo <= whi; wo += blockDim.x) {
    float acc[HG];
#pragma unroll
    for (int j = 0; j < HG; ++j) acc[j] = 0.f;
    const TL* col = plane + (size_t)ho0 * Wo + wo;
    for (int r0 = 0; r0 < nrows; r0 += kRB) {
      float v[kRB];
#pragma unroll
      for (int k = 0; k < kRB; ++k) v[k] = (r0 + k < nrows) ? ld1<TL>(col + (size_t)(r0 + k) * Wo) : 0.f;
#pragma unroll
      for (int k = 0; k < kRB; ++k)
        if (r0 + k < nrows) {
#pragma unroll
          for (int j = 0; j < HG; ++j) acc[j] = fmaf(wh_s[j][r0 + k], v[k], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < HG; ++j) colsum[j * window + wo - wlo] = acc[j];
  }
  __syncthreads();
  // HG x kTWL results: thread -> (row of the group, source column)
  const int j = threadIdx.x / kTWL, w = w0 + (int)threadIdx.x % kTWL;
  if (j < HG && h0 + j <= h1 && w <= w1) {
    const int a0 = max((int)ceilf(((float)w - 1.f - tw) * rw) - 1, wlo), a1 = min((int)floorf(((float)w + 1.f - tw) * rw) + 1, whi);
    float acc = 0.f;
    const float* cs = colsum + j * window - wlo;
    for (int wo = a0; wo <= a1; ++wo) {
      const float fw = fminf(fmaxf(fmaf(sw, (float)wo, tw), 0.f), wmax);
      acc = fmaf(fmaxf(1.f - fabsf(fw - (float)w), 0.f), cs[wo], acc);
    }
    TO* o = dlow + ((size_t)((size_t)n * Hi + h0 + j) * Wi + w) * low_cs + c;
    acc *= gscale;
    st1<TO>(o, accumulate ? ld1<TO>(o) + acc : acc);
  }
}

}  // namespace

// dy: NCHW gradient of the up-sampled tensor, dx: NHWC gradient of the source; false when the problem is not this kernel's
bool esn_bilinear_bwd_rows_try(const EsnTensor* dy, const EsnTensor* dx, int align_corners, int accumulate, float gscale,
                               void* stream) {
  // up-sampling factors >= 4 only: a CTA covers 32 source columns, i.e. 32 x factor output columns with its 288 threads (at a
  // factor of 2 -- ESPNetv2's final up-sampling -- three quarters of the CTA idle: 17 ms against 11 for the per-element kernel)
  if (dy->layout != ESN_NCHW || dy->w < 4 * dx->w || dy->h < 2 || dy->w < 2 || dx->h < 1) return false;
  float sh, th, sw, tw;
  if (align_corners) {
    if (dx->h < 2 || dx->w < 2) return false;
    sh = (float)(dx->h - 1) / (float)(dy->h - 1), th = 0.f;
    sw = (float)(dx->w - 1) / (float)(dy->w - 1), tw = 0.f;
  } else {
    sh = (float)dx->h / (float)dy->h, th = 0.5f * sh - 0.5f;
    sw = (float)dx->w / (float)dy->w, tw = 0.5f * sw - 0.5f;
  }
  const int wtiles = esn_cdiv(dx->w, kTWL);
  const int window = (int)((kTWL + 2) / sw) + 8;          // output columns under the hats of kTWL source columns
  // four source rows per CTA when the union of their output rows fits the weight table, else one
  const int hgsize = ((int)(5.f / sh) + 4 <= kMaxRows && dx->h >= 4) ? 4 : 1;
  if ((int)((hgsize + 1) / sh) + 4 > kMaxRows) return false;
  const int hgroups = esn_cdiv(dx->h, hgsize);
  const long long ctas = (long long)dx->n * dx->c * hgroups * wtiles;
  if (window > 8192 || ctas >= (1LL << 31)) return false;
  const int smem = hgsize * window * (int)sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool lf = dy->dtype == ESN_F32, of = dx->dtype == ESN_F32;
#define ESN_BLR2(TL, TO, HG)                                                                                                  \
  bilinear_bwd_rows_kernel<TL, TO, HG><<<(unsigned)ctas, 288, smem, st>>>((const TL*)dy->ptr, (TO*)dx->ptr, dx->c, dx->h, dx->w, \
                                                                          dy->h, dy->w, dx->c_stride, sh, th, sw, tw, gscale,  \
                                                                          wtiles, hgroups, window, accumulate)
#define ESN_BLR(TL, TO)                    \
  do {                                     \
    if (hgsize == 4) ESN_BLR2(TL, TO, 4);  \
    else ESN_BLR2(TL, TO, 1);              \
  } while (0)
  if (lf && of) ESN_BLR(float, float);
  else if (lf) ESN_BLR(float, __nv_bfloat16);
  else if (of) ESN_BLR(__nv_bfloat16, float);
  else ESN_BLR(__nv_bfloat16, __nv_bfloat16);
#undef ESN_BLR
#undef ESN_BLR2
  return true;
}


extern "C" int esn_maxpool3x3s2_idx_bwd(const EsnTensor* dy, const int32_t* idx, const EsnTensor* dx, int32_t accumulate,
                                        void* stream) {
  if (!dy || !dx || !idx || !esn_valid_nhwc(*dy) || !esn_valid_nhwc(*dx) || dy->dtype != dx->dtype) return ESN_ERR_BAD_ARG;
  if (dy->n != dx->n || dy->c != dx->c || dy->h != (dx->h - 1) / 2 + 1 || dy->w != (dx->w - 1) / 2 + 1) return ESN_ERR_BAD_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C = dx->c;
  const int V = (C % 8 == 0) ? 8 : (C % 4 == 0 ? 4 : 1);
  const long long total = (long long)dx->n * dx->h * dx->w * (C / V);
  const int grid = esn_cdiv(total, 256);
#define LAUNCH(T, VV)                                                                                                      \
  maxpool3x3s2_idx_bwd_kernel<T, VV><<<grid, 256, 0, st>>>((const T*)dy->ptr, idx, (T*)dx->ptr, dx->n, dx->h, dx->w, C, dy->h, \
                                                           dy->w, dy->c_stride, dx->c_stride, accumulate)
  if (dx->dtype == ESN_BF16) {
    if (V == 8) LAUNCH(__nv_bfloat16, 8); else if (V == 4) LAUNCH(__nv_bfloat16, 4); else LAUNCH(__nv_bfloat16, 1);
  } else {
    if (V == 8) LAUNCH(float, 8); else if (V == 4) LAUNCH(float, 4); else LAUNCH(float, 1);
  }
#undef LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_max_unpool2x2_bwd(const EsnTensor* dy, const int32_t* idx, const EsnTensor* dv, void* stream) {
  if (!dy || !dv || !idx || !esn_valid_nhwc(*dy) || !esn_valid_nhwc(*dv) || dy->dtype != dv->dtype) return ESN_ERR_BAD_ARG;
  if (dy->n != dv->n || dy->c != dv->c || dy->h != 2 * dv->h || dy->w != 2 * dv->w) return ESN_ERR_BAD_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C = dv->c;
  const int V = (C % 8 == 0) ? 8 : (C % 4 == 0 ? 4 : 1);
  const long long npix = (long long)dv->n * dv->h * dv->w;
  const int grid = esn_cdiv(npix * (C / V), 256);
#define LAUNCH(T, VV)                                                                                                       \
  max_unpool2x2_bwd_kernel<T, VV><<<grid, 256, 0, st>>>((const T*)dy->ptr, idx, (T*)dv->ptr, npix, C, dv->h, dv->w, dy->c_stride, \
                                                        dv->c_stride)
  if (dv->dtype == ESN_BF16) {
    if (V == 8) LAUNCH(__nv_bfloat16, 8); else if (V == 4) LAUNCH(__nv_bfloat16, 4); else LAUNCH(__nv_bfloat16, 1);
  } else {
    if (V == 8) LAUNCH(float, 8); else if (V == 4) LAUNCH(float, 4); else LAUNCH(float, 1);
  }
#undef LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of CGNet's global-context gate FGlo (CGNet.py:173-191, y = x * g[n][c] with g = sigmoid(MLP(mean_hw x))):
//   esn_dot_nc        : out[n][c] += sum_hw a[n,h,w,c] * b[n,h,w,c]   (dg = sum dy * x; fp32 atomics, out zeroed by the caller)
//   esn_scale_add_nc  : y = a * s[n][c] + t[n][c] (+ extra)            (dx = dy * g + dpooled / HW, plus an earlier gradient)
// Both HBM-bound passes over the activation; the two-layer MLP on the (N, C) vectors is host-side torch (a few kFLOP).
namespace {

template <typename T, int V>
__global__ void __launch_bounds__(256) dot_nc_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out,
                                                     int HW, int C, int a_cs, int b_cs, int rows_per_cta) {
  __shared__ float red[256 * V];
  const int cg = C / V;                       // channel groups; 256 % cg == 0 is checked on the host
  const int lane_c = threadIdx.x % cg, lane_p = threadIdx.x / cg, np = 256 / cg;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * rows_per_cta;
  const int p1 = min(HW, p0 + rows_per_cta);
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (int p = p0 + lane_p; p < p1; p += np) {
    const size_t pix = (size_t)n * HW + p;
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += ld1<T>(a + pix * a_cs + lane_c * V + k) * ld1<T>(b + pix * b_cs + lane_c * V + k);
  }
#pragma unroll
  for (int k = 0; k < V; ++k) red[threadIdx.x * V + k] = acc[k];
  __syncthreads();
  if (lane_p == 0) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float s = 0.f;
      for (int q = 0; q < np; ++q) s += red[(q * cg + lane_c) * V + k];
      atomicAdd(out + (size_t)n * C + lane_c * V + k, s);
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) scale_add_nc_kernel(const T* __restrict__ a, const float* __restrict__ s, const float* __restrict__ t,
                                                           const T* __restrict__ extra, T* __restrict__ y, long long npix, int HW,
                                                           int C, int a_cs, int e_cs, int y_cs) {
  const int cg = C / V;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= npix * cg) return;
  const int c = (int)(i % cg) * V;
  const long long pix = i / cg;
  const long long n = pix / HW;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    float v = ld1<T>(a + pix * a_cs + c + k) * __ldg(s + n * C + c + k) + (t ? __ldg(t + n * C + c + k) : 0.f);
    if (extra) v += ld1<T>(extra + pix * e_cs + c + k);
    st1<T>(y + pix * y_cs + c + k, v);
  }
}

}  // namespace

extern "C" int esn_dot_nc(const EsnTensor* a, const EsnTensor* b, float* out, void* stream) {
  if (!a || !b || !out || !esn_valid_nhwc(*a) || !esn_valid_nhwc(*b) || a->dtype != b->dtype) return ESN_ERR_BAD_ARG;
  if (a->n != b->n || a->h != b->h || a->w != b->w || a->c != b->c) return ESN_ERR_BAD_SHAPE;
  const int C = a->c;
  const int V = (C % 8 == 0 && 256 % (C / 8) == 0) ? 8 : ((C % 4 == 0 && 256 % (C / 4) == 0) ? 4 : 0);
  if (!V) return ESN_ERR_UNSUPPORTED;
  const int HW = a->h * a->w;
  const int rows = 2048;
  dim3 grid((unsigned)esn_cdiv(HW, rows), (unsigned)a->n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == ESN_BF16) {
    if (V == 8) dot_nc_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a->ptr, (const __nv_bfloat16*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
    else dot_nc_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a->ptr, (const __nv_bfloat16*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
  } else {
    if (V == 8) dot_nc_kernel<float, 8><<<grid, 256, 0, st>>>((const float*)a->ptr, (const float*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
    else dot_nc_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)a->ptr, (const float*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
  }
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_scale_add_nc(const EsnTensor* a, const float* s, const float* t, const EsnTensor* extra, const EsnTensor* y,
                                void* stream) {
  if (!a || !s || !y || !esn_valid_nhwc(*a) || !esn_valid_nhwc(*y) || a->dtype != y->dtype) return ESN_ERR_BAD_ARG;
  if (a->n != y->n || a->h != y->h || a->w != y->w || a->c != y->c) return ESN_ERR_BAD_SHAPE;
  const bool has_e = extra && extra->ptr;
  if (has_e && (!esn_valid_nhwc(*extra) || extra->dtype != a->dtype || extra->n != a->n || extra->h != a->h || extra->w != a->w ||
                extra->c != a->c))
    return ESN_ERR_BAD_ARG;
  const int C = a->c;
  const int V = (C % 8 == 0) ? 8 : (C % 4 == 0 ? 4 : 1);
  const long long npix = (long long)a->n * a->h * a->w;
  const int grid = esn_cdiv(npix * (C / V), 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define LAUNCH(T, VV)                                                                                                          \
  scale_add_nc_kernel<T, VV><<<grid, 256, 0, st>>>((const T*)a->ptr, s, t, has_e ? (const T*)extra->ptr : nullptr, (T*)y->ptr, npix, \
                                                   a->h * a->w, C, a->c_stride, has_e ? extra->c_stride : 0, y->c_stride)
  if (a->dtype == ESN_BF16) {
    if (V == 8) LAUNCH(__nv_bfloat16, 8); else if (V == 4) LAUNCH(__nv_bfloat16, 4); else LAUNCH(__nv_bfloat16, 1);
  } else {
    if (V == 8) LAUNCH(float, 8); else if (V == 4) LAUNCH(float, 4); else LAUNCH(float, 1);
  }
#undef LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
