// Backward kernels of the resize / pooling / dropout ops of the pointwise-heavy nets (Fast-SCNN, ESPNetv2):
// gather-form (deterministic, no atomics), NHWC, one thread per (pixel, channel) of the gradient it produces.
// All are HBM-bound and small (they run at 1/8 ... 1/32 resolution, or on the 19-class logits).
#include "esn_common.cuh"

namespace {

// ---- bilinear backward, align_corners = False / True, upstream gradient in NCHW (logits) or NHWC
template <typename TG, typename TO>
__global__ void __launch_bounds__(128) bilinear_bwd2_kernel(const TG* __restrict__ dy, TO* __restrict__ dx, int N, int C, int Hi,
                                                            int Wi, int Ho, int Wo, int dy_nchw, int dy_cs, int dx_cs, float sh,
                                                            float sw, int align, int accumulate) {
  const long long total = (long long)N * Hi * Wi * C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int c, w, h, n;
  if (dy_nchw) {      // upstream gradient in class planes: consecutive threads walk along w (coalesced plane reads)
    w = (int)(idx % Wi);
    h = (int)((idx / Wi) % Hi);
    c = (int)((idx / ((long long)Wi * Hi)) % C);
    n = (int)(idx / ((long long)Wi * Hi * C));
  } else {
    c = (int)(idx % C);
    w = (int)((idx / C) % Wi);
    h = (int)((idx / ((long long)C * Wi)) % Hi);
    n = (int)(idx / ((long long)C * Wi * Hi));
  }
  // candidate outputs: source index within (h-1, h+1)
  const float rh = sh > 0.f ? 1.f / sh : 0.f, rw = sw > 0.f ? 1.f / sw : 0.f;
  const float off = align ? 0.f : 0.5f;
  int ho0 = sh > 0.f ? (int)floorf(((float)h - 1.f + off) * rh - off) - 1 : 0;
  int ho1 = sh > 0.f ? (int)ceilf(((float)h + 1.f + off) * rh - off) + 1 : Ho - 1;
  int wo0 = sw > 0.f ? (int)floorf(((float)w - 1.f + off) * rw - off) - 1 : 0;
  int wo1 = sw > 0.f ? (int)ceilf(((float)w + 1.f + off) * rw - off) + 1 : Wo - 1;
  ho0 = max(ho0, 0); ho1 = min(ho1, Ho - 1);
  wo0 = max(wo0, 0); wo1 = min(wo1, Wo - 1);
  float acc = 0.f;
  for (int ho = ho0; ho <= ho1; ++ho) {
    float fh = align ? sh * ho : sh * (ho + 0.5f) - 0.5f;
    fh = fh < 0.f ? 0.f : fh;
    const int h0 = min((int)fh, Hi - 1);
    const int h1 = h0 + ((h0 < Hi - 1) ? 1 : 0);
    const float l1 = fh - h0, l0 = 1.f - l1;
    const float wh = (h0 == h ? l0 : 0.f) + (h1 == h ? l1 : 0.f);
    if (wh == 0.f) continue;
    float rowacc = 0.f;
    for (int wo = wo0; wo <= wo1; ++wo) {
      float fw = align ? sw * wo : sw * (wo + 0.5f) - 0.5f;
      fw = fw < 0.f ? 0.f : fw;
      const int w0 = min((int)fw, Wi - 1);
      const int w1 = w0 + ((w0 < Wi - 1) ? 1 : 0);
      const float m1 = fw - w0, m0 = 1.f - m1;
      const float ww = (w0 == w ? m0 : 0.f) + (w1 == w ? m1 : 0.f);
      if (ww != 0.f) {
        const size_t gi = dy_nchw ? (((size_t)n * C + c) * Ho + ho) * Wo + wo : (((size_t)n * Ho + ho) * Wo + wo) * dy_cs + c;
        rowacc += ww * ld1<TG>(dy + gi);
      }
    }
    acc += wh * rowacc;
  }
  TO* p = dx + ((size_t)((size_t)n * Hi + h) * Wi + w) * dx_cs + c;
  if (accumulate) acc += ld1<TO>(p);
  st1<TO>(p, acc);
}

// the NHWC bf16 case on 16-byte channel vectors: a thread owns eight channels of one source pixel (same candidate windows and
// weights as above).  The scalar kernel ran ESPNetv2's seven decoder / pyramid up-sampling gradients at 0.25 TB/s (9.1 ms).
__global__ void __launch_bounds__(256) bilinear_bwd2_v8_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                               long long total, int C8, int Hi, int Wi, int Ho, int Wo, int dy_cs,
                                                               int dx_cs, float sh, float sw, int align, int accumulate) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % C8);
  const long long pix = idx / C8;
  const int w = (int)(pix % Wi);
  const int h = (int)((pix / Wi) % Hi);
  const long long n = pix / ((long long)Wi * Hi);
  const float rh = sh > 0.f ? 1.f / sh : 0.f, rw = sw > 0.f ? 1.f / sw : 0.f;
  const float off = align ? 0.f : 0.5f;
  int ho0 = sh > 0.f ? (int)floorf(((float)h - 1.f + off) * rh - off) - 1 : 0;
  int ho1 = sh > 0.f ? (int)ceilf(((float)h + 1.f + off) * rh - off) + 1 : Ho - 1;
  int wo0 = sw > 0.f ? (int)floorf(((float)w - 1.f + off) * rw - off) - 1 : 0;
  int wo1 = sw > 0.f ? (int)ceilf(((float)w + 1.f + off) * rw - off) + 1 : Wo - 1;
  ho0 = max(ho0, 0); ho1 = min(ho1, Ho - 1);
  wo0 = max(wo0, 0); wo1 = min(wo1, Wo - 1);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int ho = ho0; ho <= ho1; ++ho) {
    float fh = align ? sh * ho : sh * (ho + 0.5f) - 0.5f;
    fh = fh < 0.f ? 0.f : fh;
    const int h0 = min((int)fh, Hi - 1);
    const int h1 = h0 + ((h0 < Hi - 1) ? 1 : 0);
    const float l1 = fh - h0, l0 = 1.f - l1;
    const float wh = (h0 == h ? l0 : 0.f) + (h1 == h ? l1 : 0.f);
    if (wh == 0.f) continue;
    const __nv_bfloat16* row = dy + ((size_t)(n * Ho + ho) * Wo) * dy_cs;
    for (int wo = wo0; wo <= wo1; ++wo) {
      float fw = align ? sw * wo : sw * (wo + 0.5f) - 0.5f;
      fw = fw < 0.f ? 0.f : fw;
      const int w0 = min((int)fw, Wi - 1);
      const int w1 = w0 + ((w0 < Wi - 1) ? 1 : 0);
      const float m1 = fw - w0, m0 = 1.f - m1;
      const float ww = ((w0 == w ? m0 : 0.f) + (w1 == w ? m1 : 0.f)) * wh;
      if (ww != 0.f) {
        float t[8];
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(row + (size_t)wo * dy_cs) + cv), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(ww, t[j], acc[j]);
      }
    }
  }
  uint4* p = reinterpret_cast<uint4*>(dx + (size_t)pix * dx_cs) + cv;
  if (accumulate) {
    float prev[8];
    bf16x8_to_float(*p, prev);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += prev[j];
  }
  *p = float_to_bf16x8(acc);
}

// ---- adaptive average pool backward: dx[h,w] = sum over windows containing (h,w) of dy / |window|
template <typename T>
__global__ void __launch_bounds__(128) adaptive_avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int C, int H,
                                                                   int W, int S, int dy_cs, int dx_cs, int accumulate) {
  const long long total = (long long)N * H * W * C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int w = (int)((idx / C) % W);
  const int h = (int)((idx / ((long long)C * W)) % H);
  const int n = (int)(idx / ((long long)C * W * H));
  float acc = 0.f;
  for (int i = 0; i < S; ++i) {
    const int hs = (i * H) / S, he = ((i + 1) * H + S - 1) / S;
    if (h < hs || h >= he) continue;
    for (int j = 0; j < S; ++j) {
      const int ws = (j * W) / S, we = ((j + 1) * W + S - 1) / S;
      if (w < ws || w >= we) continue;
      acc += ld1<T>(dy + (((size_t)n * S + i) * S + j) * dy_cs + c) / (float)((he - hs) * (we - ws));
    }
  }
  T* p = dx + ((size_t)((size_t)n * H + h) * W + w) * dx_cs + c;
  if (accumulate) acc += ld1<T>(p);
  st1<T>(p, acc);
}

// ---- AvgPool2d(3, stride 2, pad 1, count_include_pad) backward: every window divides by 9
template <typename T>
__global__ void __launch_bounds__(128) avgpool3x3s2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int C, int H, int W,
                                                               int Ho, int Wo, int dy_cs, int dx_cs, int accumulate) {
  const long long total = (long long)N * H * W * C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int w = (int)((idx / C) % W);
  const int h = (int)((idx / ((long long)C * W)) % H);
  const int n = (int)(idx / ((long long)C * W * H));
  float acc = 0.f;
  for (int ho = h / 2; ho <= (h + 1) / 2; ++ho) {          // windows [2ho-1, 2ho+1] containing h
    if (ho >= Ho) continue;
    for (int wo = w / 2; wo <= (w + 1) / 2; ++wo) {
      if (wo >= Wo) continue;
      acc += ld1<T>(dy + (((size_t)n * Ho + ho) * Wo + wo) * dy_cs + c);
    }
  }
  acc *= (1.f / 9.f);
  T* p = dx + ((size_t)((size_t)n * H + h) * W + w) * dx_cs + c;
  if (accumulate) acc += ld1<T>(p);
  st1<T>(p, acc);
}

// the same for bf16 tensors on 16-byte channel vectors: a thread owns eight channels of one input pixel and gathers the (up to
// four) windows that contain it with 16-byte loads.  The scalar kernel above ran ESPNetv2's seven down-sampler gradients at
// 0.2 TB/s (9.9 ms of its training step).
__global__ void __launch_bounds__(256) avgpool3x3s2_bwd_v8_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                                  long long total, int C8, int H, int W, int Ho, int Wo, int dy_cs,
                                                                  int dx_cs, int accumulate) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = (int)(idx % C8);
  const long long pix = idx / C8;
  const int w = (int)(pix % W);
  const int h = (int)((pix / W) % H);
  const long long n = pix / ((long long)W * H);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int ho = h / 2; ho <= (h + 1) / 2; ++ho) {          // windows [2ho-1, 2ho+1] containing h
    if (ho >= Ho) continue;
    for (int wo = w / 2; wo <= (w + 1) / 2; ++wo) {
      if (wo >= Wo) continue;
      float t[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(dy + ((size_t)(n * Ho + ho) * Wo + wo) * dy_cs) + cv), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
  }
  uint4* p = reinterpret_cast<uint4*>(dx + (size_t)pix * dx_cs) + cv;
  float prev[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (accumulate) bf16x8_to_float(*p, prev);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = acc[j] * (1.f / 9.f) + prev[j];
  *p = float_to_bf16x8(acc);
}

// ---- dropout (element-wise nn.Dropout or per-(n, c) nn.Dropout2d): counter-based hash, so the backward
// pass regenerates the mask from the seed instead of storing it
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int C, int H, int W, int x_cs,
                                                      int y_cs, uint64_t seed, const unsigned long long* __restrict__ step,
                                                      uint32_t thresh, float scale, int per_channel) {
  if (step) seed += 0xd1b54a32d192ed03ULL * (uint64_t)(*step);   // device-side iteration counter (CUDA-graph replays)
  const long long total = (long long)N * H * W * C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long pix = idx / C;
  const int n = (int)(pix / ((long long)H * W));
  const uint64_t key = per_channel ? ((uint64_t)n * C + c) : (uint64_t)idx;
  const bool keep = mix32(seed + 0x9e3779b97f4a7c15ULL * (key + 1)) >= thresh;
  st1<T>(y + (size_t)pix * y_cs + c, keep ? ld1<T>(x + (size_t)pix * x_cs + c) * scale : 0.f);
}

// Dropout2d: the N x C scale factors (0 or 1 / (1 - p)), same keys and hash as the per-element kernel's per_channel mode
__global__ void dropout_mask_nc_kernel(float* __restrict__ mask, long long count, uint64_t seed,
                                       const unsigned long long* __restrict__ step, uint32_t thresh, float scale) {
  if (step) seed += 0xd1b54a32d192ed03ULL * (uint64_t)(*step);
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= count) return;
  mask[i] = mix32(seed + 0x9e3779b97f4a7c15ULL * ((uint64_t)i + 1)) >= thresh ? scale : 0.f;
}

}  // namespace

bool esn_bilinear_bwd_rows_try(const EsnTensor* dy, const EsnTensor* dx, int align_corners, int accumulate, float gscale,
                               void* stream);      // esn_train3.cu

extern "C" int esn_dropout_mask_nc(float* mask, int64_t count, uint64_t seed, const uint64_t* step, float p, void* stream) {
  if (!mask || count < 1 || !(p >= 0.f) || p >= 1.f) return ESN_ERR_BAD_ARG;
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  dropout_mask_nc_kernel<<<esn_cdiv(count, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      mask, count, seed, reinterpret_cast<const unsigned long long*>(step), thresh, 1.f / (1.f - p));
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_bilinear_bwd_nhwc(const EsnTensor* dy, const EsnTensor* dx, int32_t align_corners, int32_t accumulate,
                                     void* stream) {
  if (!dy || !dx || !dy->ptr || !esn_valid_nhwc(*dx)) return ESN_ERR_BAD_ARG;
  const bool nchw = dy->layout == ESN_NCHW;
  if (!nchw && !esn_valid_nhwc(*dy)) return ESN_ERR_BAD_ARG;
  if (dy->n != dx->n || dy->c != dx->c) return ESN_ERR_BAD_SHAPE;
  float sh, sw;
  if (align_corners) {
    sh = dy->h > 1 ? (float)(dx->h - 1) / (float)(dy->h - 1) : 0.f;
    sw = dy->w > 1 ? (float)(dx->w - 1) / (float)(dy->w - 1) : 0.f;
  } else {
    sh = (float)dx->h / (float)dy->h;
    sw = (float)dx->w / (float)dy->w;
  }
  if (nchw && esn_bilinear_bwd_rows_try(dy, dx, align_corners ? 1 : 0, accumulate ? 1 : 0, 1.f, stream)) {
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  const long long total = (long long)dx->n * dx->h * dx->w * dx->c;
  const int grid = esn_cdiv(total, 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool gf = dy->dtype == ESN_F32, of = dx->dtype == ESN_F32;
  // (not for up-sampling factors above 4: a source pixel of Fast-SCNN's 1x1 ... 6x6 pyramid levels gathers hundreds of outputs,
  // and one thread per channel is then the better split -- measured, 19.7 against 20.6 ms per Fast-SCNN step)
  if (!nchw && !gf && !of && dx->c % 8 == 0 && dx->c_stride % 8 == 0 && dy->c_stride % 8 == 0 && ((uintptr_t)dx->ptr % 16) == 0 &&
      ((uintptr_t)dy->ptr % 16) == 0 && total / 8 < 0x7fffffffLL * 256 && (long long)dy->h * dy->w <= 16LL * dx->h * dx->w) {
    const long long tv = total / 8;
    bilinear_bwd2_v8_kernel<<<esn_cdiv(tv, 256), 256, 0, st>>>((const __nv_bfloat16*)dy->ptr, (__nv_bfloat16*)dx->ptr, tv, dx->c / 8,
                                                               dx->h, dx->w, dy->h, dy->w, dy->c_stride, dx->c_stride, sh, sw,
                                                               align_corners ? 1 : 0, accumulate ? 1 : 0);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
#define ESN_BB2(TG, TO)                                                                                                  \
  bilinear_bwd2_kernel<TG, TO><<<grid, 128, 0, st>>>((const TG*)dy->ptr, (TO*)dx->ptr, dx->n, dx->c, dx->h, dx->w, dy->h, \
                                                     dy->w, nchw ? 1 : 0, dy->c_stride, dx->c_stride, sh, sw,             \
                                                     align_corners ? 1 : 0, accumulate)
  if (gf && of) ESN_BB2(float, float);
  else if (gf) ESN_BB2(float, __nv_bfloat16);
  else if (of) ESN_BB2(__nv_bfloat16, float);
  else ESN_BB2(__nv_bfloat16, __nv_bfloat16);
#undef ESN_BB2
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_adaptive_avgpool_bwd(const EsnTensor* dy, const EsnTensor* dx, int32_t accumulate, void* stream) {
  if (!dy || !dx || !esn_valid_nhwc(*dy) || !esn_valid_nhwc(*dx)) return ESN_ERR_BAD_ARG;
  if (dy->n != dx->n || dy->c != dx->c || dy->h != dy->w || dy->dtype != dx->dtype) return ESN_ERR_BAD_SHAPE;
  const long long total = (long long)dx->n * dx->h * dx->w * dx->c;
  const int grid = esn_cdiv(total, 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dx->dtype == ESN_F32)
    adaptive_avgpool_bwd_kernel<float><<<grid, 128, 0, st>>>((const float*)dy->ptr, (float*)dx->ptr, dx->n, dx->c, dx->h, dx->w,
                                                             dy->h, dy->c_stride, dx->c_stride, accumulate);
  else
    adaptive_avgpool_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)dy->ptr, (__nv_bfloat16*)dx->ptr, dx->n,
                                                                     dx->c, dx->h, dx->w, dy->h, dy->c_stride, dx->c_stride,
                                                                     accumulate);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_avgpool3x3s2_bwd(const EsnTensor* dy, const EsnTensor* dx, int32_t accumulate, void* stream) {
  if (!dy || !dx || !esn_valid_nhwc(*dy) || !esn_valid_nhwc(*dx)) return ESN_ERR_BAD_ARG;
  if (dy->n != dx->n || dy->c != dx->c || dy->h != (dx->h - 1) / 2 + 1 || dy->w != (dx->w - 1) / 2 + 1 || dy->dtype != dx->dtype)
    return ESN_ERR_BAD_SHAPE;
  const long long total = (long long)dx->n * dx->h * dx->w * dx->c;
  const int grid = esn_cdiv(total, 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dx->dtype == ESN_BF16 && dx->c % 8 == 0 && dx->c_stride % 8 == 0 && dy->c_stride % 8 == 0 &&
      ((uintptr_t)dx->ptr % 16) == 0 && ((uintptr_t)dy->ptr % 16) == 0 && total / 8 < 0x7fffffffLL * 256) {
    const long long tv = total / 8;
    avgpool3x3s2_bwd_v8_kernel<<<esn_cdiv(tv, 256), 256, 0, st>>>((const __nv_bfloat16*)dy->ptr, (__nv_bfloat16*)dx->ptr, tv,
                                                                  dx->c / 8, dx->h, dx->w, dy->h, dy->w, dy->c_stride, dx->c_stride,
                                                                  accumulate);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  if (dx->dtype == ESN_F32)
    avgpool3x3s2_bwd_kernel<float><<<grid, 128, 0, st>>>((const float*)dy->ptr, (float*)dx->ptr, dx->n, dx->c, dx->h, dx->w, dy->h,
                                                         dy->w, dy->c_stride, dx->c_stride, accumulate);
  else
    avgpool3x3s2_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)dy->ptr, (__nv_bfloat16*)dx->ptr, dx->n, dx->c,
                                                                 dx->h, dx->w, dy->h, dy->w, dy->c_stride, dx->c_stride, accumulate);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_dropout(const EsnTensor* x, const EsnTensor* y, uint64_t seed, float p, int32_t per_channel, void* stream) {
  return esn_dropout_step(x, y, seed, nullptr, p, per_channel, stream);
}

extern "C" int esn_dropout_step(const EsnTensor* x, const EsnTensor* y, uint64_t seed, const uint64_t* step, float p,
                                int32_t per_channel, void* stream) {
  if (!x || !y || !esn_valid_nhwc(*x) || !esn_valid_nhwc(*y)) return ESN_ERR_BAD_ARG;
  if (x->n != y->n || x->c != y->c || x->h != y->h || x->w != y->w || x->dtype != y->dtype) return ESN_ERR_BAD_SHAPE;
  if (!(p >= 0.f) || p >= 1.f) return ESN_ERR_BAD_ARG;
  const long long total = (long long)x->n * x->h * x->w * x->c;
  const int grid = esn_cdiv(total, 256);
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  const float scale = 1.f / (1.f - p);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x->dtype == ESN_F32)
    dropout_kernel<float><<<grid, 256, 0, st>>>((const float*)x->ptr, (float*)y->ptr, x->n, x->c, x->h, x->w, x->c_stride,
                                                y->c_stride, seed, reinterpret_cast<const unsigned long long*>(step), thresh, scale, per_channel);
  else
    dropout_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x->ptr, (__nv_bfloat16*)y->ptr, x->n, x->c, x->h, x->w,
                                                        x->c_stride, y->c_stride, seed, reinterpret_cast<const unsigned long long*>(step), thresh, scale, per_channel);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
