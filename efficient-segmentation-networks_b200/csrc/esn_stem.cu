// Network stem: 3x3 stride-2 pad-1 convolution read straight from the caller's NCHW fp32 image
// (Cin = 3), optionally concatenated with MaxPool2d(2,2) of the same image, then the per-channel
// affine (bias + folded BatchNorm) and activation, written as NHWC (bf16 or fp32).
//   ERFNet.py:16-27  DownsamplerBlock(3,16): 13 conv + 3 pool channels, BN, ReLU
//   DABNet.py:132    Conv(3,32,3,2) + BNPReLU
// HBM-bound: reads 12 B/input pixel once (neighbour re-reads hit L1), writes Ctot*2 B per output pixel.
#include "esn_common.cuh"

namespace {

struct StemArgs {
  const float* x;
  void* y;
  const float* w;  // [9][3][cconv]
  int N, H, W, Ho, Wo, cconv, ctot, y_cs, with_pool, pad;
  EpiArgs ep;
};

template <typename TO, int CPAD>
__global__ void __launch_bounds__(128) stem_kernel(const StemArgs a) {
  __shared__ float sw[27 * CPAD];
  __shared__ float sp[3 * CPAD];
  for (int i = threadIdx.x; i < 27 * CPAD; i += blockDim.x) {
    const int t = i / CPAD, c = i % CPAD;
    sw[i] = c < a.cconv ? a.w[t * a.cconv + c] : 0.f;
  }
  for (int i = threadIdx.x; i < CPAD; i += blockDim.x) {
    const bool in = i < a.ctot;
    sp[i] = (in && a.ep.scale) ? a.ep.scale[i] : 1.f;
    sp[CPAD + i] = (in && a.ep.shift) ? a.ep.shift[i] : 0.f;
    sp[2 * CPAD + i] = (in && a.ep.act == ESN_ACT_PRELU) ? a.ep.alpha[i] : 0.f;
  }
  __syncthreads();
  const long long total = (long long)a.N * a.Ho * a.Wo;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % a.Wo);
  const int ho = (int)((idx / a.Wo) % a.Ho);
  const int n = (int)(idx / ((long long)a.Wo * a.Ho));
  const size_t plane = (size_t)a.H * a.W;
  const float* xb = a.x + (size_t)n * 3 * plane;

  float v[27];
  unsigned okmask = 0;   // bit (r*3+s): tap inside the image (needed by the 3x3 pool: padding is -inf there)
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = 2 * ho - a.pad + r;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int wi = 2 * wo - a.pad + s;
      const bool ok = hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
      okmask |= (ok ? 1u : 0u) << (r * 3 + s);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = ok ? __ldg(xb + c * plane + (size_t)hi * a.W + wi) : 0.f;
    }
  }
  float acc[CPAD];
#pragma unroll
  for (int c = 0; c < CPAD; ++c) acc[c] = 0.f;
#pragma unroll
  for (int t = 0; t < 27; ++t) {
#pragma unroll
    for (int c4 = 0; c4 < CPAD; c4 += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(sw + t * CPAD + c4);
      acc[c4] = fmaf(v[t], wv.x, acc[c4]);
      acc[c4 + 1] = fmaf(v[t], wv.y, acc[c4 + 1]);
      acc[c4 + 2] = fmaf(v[t], wv.z, acc[c4 + 2]);
      acc[c4 + 3] = fmaf(v[t], wv.w, acc[c4 + 3]);
    }
  }
  if (a.with_pool) {  // 1: taps (1,1),(1,2),(2,1),(2,2) = the 2x2 window; 2: all valid taps = MaxPool2d(3,2,1)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float m;
      if (a.with_pool == 2) {
        m = -INFINITY;
#pragma unroll
        for (int t = 0; t < 9; ++t)
          if (okmask & (1u << t)) m = fmaxf(m, v[t * 3 + c]);
      } else {
        m = fmaxf(fmaxf(v[(1 * 3 + 1) * 3 + c], v[(1 * 3 + 2) * 3 + c]),
                  fmaxf(v[(2 * 3 + 1) * 3 + c], v[(2 * 3 + 2) * 3 + c]));
      }
#pragma unroll
      for (int k = 0; k < CPAD; ++k)
        if (k == a.cconv + c) acc[k] = m;
    }
  }
#pragma unroll
  for (int c = 0; c < CPAD; ++c) {
    const float t = fmaf(acc[c], sp[c], sp[CPAD + c]);
    acc[c] = apply_act(t, a.ep.act, sp[2 * CPAD + c]);
  }
  TO* yp = reinterpret_cast<TO*>(a.y) + (size_t)idx * a.y_cs;
#pragma unroll
  for (int c4 = 0; c4 < CPAD; c4 += 4)
    if (c4 < a.ctot) st4<TO>(yp + c4, make_float4(acc[c4], acc[c4 + 1], acc[c4 + 2], acc[c4 + 3]));
}

}  // namespace

extern "C" int esn_stem_conv3x3s2(const EsnStem* p, void* stream) {
  if (!p || !p->w || !p->x.ptr || !esn_valid_nhwc(p->y)) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (x.layout != ESN_NCHW || x.dtype != ESN_F32 || x.c != 3) return ESN_ERR_UNSUPPORTED;
  const int pad = (p->with_pool & ESN_STEM_PAD0) ? 0 : 1;
  const int pool = p->with_pool & 3;
  if (x.n != y.n || y.h != (x.h + 2 * pad - 3) / 2 + 1 || y.w != (x.w + 2 * pad - 3) / 2 + 1) return ESN_ERR_BAD_SHAPE;
  if (pool == 1 && ((x.h | x.w) & 1)) return ESN_ERR_UNSUPPORTED;
  if (pool && !pad) return ESN_ERR_UNSUPPORTED;
  const int ctot = p->cconv + (pool ? 3 : 0);
  if (y.c != ctot || ctot > 32 || ctot % 4 || y.c_stride % 4) return ESN_ERR_UNSUPPORTED;
  const size_t ysz = y.dtype == ESN_F32 ? 4 : 2;
  if ((uintptr_t)y.ptr % (4 * ysz)) return ESN_ERR_ALIGN;
  int rc = esn_check_epilogue(p->ep, y);
  if (rc) return rc;
  if (p->ep.residual.ptr) return ESN_ERR_UNSUPPORTED;
  StemArgs a;
  a.x = reinterpret_cast<const float*>(x.ptr);
  a.y = y.ptr;
  a.w = p->w;
  a.N = x.n;
  a.H = x.h;
  a.W = x.w;
  a.Ho = y.h;
  a.Wo = y.w;
  a.cconv = p->cconv;
  a.ctot = ctot;
  a.y_cs = y.c_stride;
  a.with_pool = pool;
  a.pad = pad;
  a.ep = make_epi(p->ep);
  const long long total = (long long)y.n * y.h * y.w;
  const int block = 128, grid = esn_cdiv(total, block);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool c16 = ctot <= 16;
  if (y.dtype == ESN_BF16) {
    if (c16) stem_kernel<__nv_bfloat16, 16><<<grid, block, 0, st>>>(a);
    else stem_kernel<__nv_bfloat16, 32><<<grid, block, 0, st>>>(a);
  } else {
    if (c16) stem_kernel<float, 16><<<grid, block, 0, st>>>(a);
    else stem_kernel<float, 32><<<grid, block, 0, st>>>(a);
  }
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
