// Resampling kernels for the pyramid-pooling / feature-fusion paths (FastSCNN.py:85-112,157-182;
// ESPNetv2 PSP): adaptive average pooling and NHWC->NHWC bilinear interpolation (align_corners False or True),
// written into a channel slice of a concat buffer.  Small, HBM/L2-bound.
#include "esn_common.cuh"

namespace {

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) adaptive_avgpool_kernel(const TI* __restrict__ x, TO* __restrict__ y, int N, int Hi,
                                                               int Wi, int C, int x_cs, int Ho, int Wo, int y_cs) {
  const long long total = (long long)N * Ho * Wo * C;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
  // torch adaptive pooling windows: [floor(o*I/O), ceil((o+1)*I/O))
  const int h0 = (oy * Hi) / Ho, h1 = ((oy + 1) * Hi + Ho - 1) / Ho;
  const int w0 = (ox * Wi) / Wo, w1 = ((ox + 1) * Wi + Wo - 1) / Wo;
  float s = 0.f;
  for (int h = h0; h < h1; ++h)
    for (int w = w0; w < w1; ++w) s += ld1<TI>(x + ((size_t)((size_t)n * Hi + h) * Wi + w) * x_cs + c);
  st1<TO>(y + (size_t)p * y_cs + c, s / (float)((h1 - h0) * (w1 - w0)));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bilinear_nhwc_kernel(const TI* __restrict__ x, TO* __restrict__ y, int N, int Hi,
                                                            int Wi, int C, int x_cs, int Ho, int Wo, int y_cs, float sh,
                                                            float sw, int align) {
  const int ncg = (C + 3) / 4;
  const long long total = (long long)N * Ho * Wo * ncg;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % ncg) * 4;
  const long long p = i / ncg;
  const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
  float fh, fw;
  if (align) {          // torch area_pixel_compute_source_index, align_corners=True: dst * (in-1)/(out-1)
    fh = sh * ho;
    fw = sw * wo;
  } else {
    fh = sh * (ho + 0.5f) - 0.5f;
    fh = fh < 0.f ? 0.f : fh;
    fw = sw * (wo + 0.5f) - 0.5f;
    fw = fw < 0.f ? 0.f : fw;
  }
  const int h0 = min((int)fh, Hi - 1), w0 = min((int)fw, Wi - 1);
  const int hp = (h0 < Hi - 1) ? 1 : 0, wp = (w0 < Wi - 1) ? 1 : 0;
  const float lh1 = fh - h0, lh0 = 1.f - lh1, lw1 = fw - w0, lw0 = 1.f - lw1;
  const TI* p00 = x + ((size_t)((size_t)n * Hi + h0) * Wi + w0) * x_cs + c;
  const TI* p01 = p00 + (size_t)wp * x_cs;
  const TI* p10 = p00 + (size_t)hp * Wi * x_cs;
  const TI* p11 = p10 + (size_t)wp * x_cs;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (c + j < C) {
      const float v = lh0 * (lw0 * ld1<TI>(p00 + j) + lw1 * ld1<TI>(p01 + j)) + lh1 * (lw0 * ld1<TI>(p10 + j) + lw1 * ld1<TI>(p11 + j));
      st1<TO>(y + (size_t)p * y_cs + c + j, v);
    }
  }
}

}  // namespace

extern "C" int esn_adaptive_avgpool(const EsnTensor* x, const EsnTensor* y, void* stream) {
  if (!x || !y || !esn_valid_nhwc(*x) || !esn_valid_nhwc(*y)) return ESN_ERR_BAD_ARG;
  if (x->n != y->n || x->c != y->c) return ESN_ERR_BAD_SHAPE;   // output may exceed the input (windows then repeat)
  const long long total = (long long)y->n * y->h * y->w * y->c;
  const int grid = esn_cdiv(total, 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ESN_AAP(TI, TO) adaptive_avgpool_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)x->ptr, (TO*)y->ptr, x->n, x->h, x->w, x->c, x->c_stride, y->h, y->w, y->c_stride)
  const bool xf = x->dtype == ESN_F32, yf = y->dtype == ESN_F32;
  if (xf && yf) ESN_AAP(float, float);
  else if (xf) ESN_AAP(float, __nv_bfloat16);
  else if (yf) ESN_AAP(__nv_bfloat16, float);
  else ESN_AAP(__nv_bfloat16, __nv_bfloat16);
#undef ESN_AAP
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_bilinear_nhwc(const EsnTensor* x, const EsnTensor* y, int32_t align_corners, void* stream) {
  if (!x || !y || !esn_valid_nhwc(*x) || !esn_valid_nhwc(*y)) return ESN_ERR_BAD_ARG;
  if (x->n != y->n || x->c != y->c) return ESN_ERR_BAD_SHAPE;
  float sh, sw;
  if (align_corners) {
    sh = y->h > 1 ? (float)(x->h - 1) / (float)(y->h - 1) : 0.f;
    sw = y->w > 1 ? (float)(x->w - 1) / (float)(y->w - 1) : 0.f;
  } else {
    sh = (float)x->h / (float)y->h;
    sw = (float)x->w / (float)y->w;
  }
  const long long total = (long long)y->n * y->h * y->w * ((y->c + 3) / 4);
  const int grid = esn_cdiv(total, 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ESN_BL(TI, TO) bilinear_nhwc_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)x->ptr, (TO*)y->ptr, x->n, x->h, x->w, x->c, x->c_stride, y->h, y->w, y->c_stride, sh, sw, align_corners)
  const bool xf = x->dtype == ESN_F32, yf = y->dtype == ESN_F32;
  if (xf && yf) ESN_BL(float, float);
  else if (xf) ESN_BL(float, __nv_bfloat16);
  else if (yf) ESN_BL(__nv_bfloat16, float);
  else ESN_BL(__nv_bfloat16, __nv_bfloat16);
#undef ESN_BL
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
