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

// One thread = PX (2) horizontally adjacent output pixels: the 27 weight vectors are read from shared memory once for
// both (the single-pixel form issued 2 LDS.128 per 4 FFMA2 and was shared-memory-issue-bound), FMAs are packed FFMA2
// over channel pairs.
template <typename TO, int CPAD>
__global__ void __launch_bounds__(128) stem_kernel(const StemArgs a) {
  constexpr int PX = 2;
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
  const int Wq = (a.Wo + PX - 1) / PX;
  const long long total = (long long)a.N * a.Ho * Wq;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo0 = (int)(idx % Wq) * PX;
  const int ho = (int)((idx / Wq) % a.Ho);
  const int n = (int)(idx / ((long long)Wq * a.Ho));
  const size_t plane = (size_t)a.H * a.W;
  const float* xb = a.x + (size_t)n * 3 * plane;

  float v[PX][27];
  unsigned okmask[PX];   // bit (r*3+s): tap inside the image (needed by the 3x3 pool: padding is -inf there)
#pragma unroll
  for (int q = 0; q < PX; ++q) {
    okmask[q] = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hi = 2 * ho - a.pad + r;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int wi = 2 * (wo0 + q) - a.pad + s;
        const bool ok = hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
        okmask[q] |= (ok ? 1u : 0u) << (r * 3 + s);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[q][(r * 3 + s) * 3 + c] = ok ? __ldg(xb + c * plane + (size_t)hi * a.W + wi) : 0.f;
      }
    }
  }
  float2 acc2[PX][CPAD / 2];
#pragma unroll
  for (int q = 0; q < PX; ++q)
#pragma unroll
    for (int c = 0; c < CPAD / 2; ++c) acc2[q][c] = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 27; ++t) {
#pragma unroll
    for (int c4 = 0; c4 < CPAD; c4 += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(sw + t * CPAD + c4);
      const float2 wa = make_float2(wv.x, wv.y), wb = make_float2(wv.z, wv.w);
#pragma unroll
      for (int q = 0; q < PX; ++q) {
        const float2 vv = make_float2(v[q][t], v[q][t]);
        acc2[q][c4 / 2] = ffma2(vv, wa, acc2[q][c4 / 2]);
        acc2[q][c4 / 2 + 1] = ffma2(vv, wb, acc2[q][c4 / 2 + 1]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PX; ++q) {
    if (wo0 + q < a.Wo) {
    float acc[CPAD];
#pragma unroll
    for (int c = 0; c < CPAD / 2; ++c) { acc[2 * c] = acc2[q][c].x; acc[2 * c + 1] = acc2[q][c].y; }
    float m[3] = {0.f, 0.f, 0.f};
    if (a.with_pool) {  // 1: taps (1,1),(1,2),(2,1),(2,2) = the 2x2 window; 2: all valid taps = MaxPool2d(3,2,1)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.with_pool == 2) {
          m[c] = -INFINITY;
#pragma unroll
          for (int t = 0; t < 9; ++t)
            if (okmask[q] & (1u << t)) m[c] = fmaxf(m[c], v[q][t * 3 + c]);
        } else {
          m[c] = fmaxf(fmaxf(v[q][(1 * 3 + 1) * 3 + c], v[q][(1 * 3 + 2) * 3 + c]),
                       fmaxf(v[q][(2 * 3 + 1) * 3 + c], v[q][(2 * 3 + 2) * 3 + c]));
        }
      }
    }
    const int pc = a.with_pool ? a.cconv : CPAD;     // first pooled channel (selects, no dynamic register indexing)
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      float t = acc[c];
      t = (c == pc) ? m[0] : ((c == pc + 1) ? m[1] : ((c == pc + 2) ? m[2] : t));
      t = fmaf(t, sp[c], sp[CPAD + c]);
      acc[c] = apply_act(t, a.ep.act, sp[2 * CPAD + c]);
    }
    TO* yp = reinterpret_cast<TO*>(a.y) + ((size_t)((size_t)n * a.Ho + ho) * a.Wo + wo0 + q) * a.y_cs;
#pragma unroll
    for (int c4 = 0; c4 < CPAD; c4 += 4)
      if (c4 < a.ctot) st4<TO>(yp + c4, make_float4(acc[c4], acc[c4 + 1], acc[c4 + 2], acc[c4 + 3]));
    }
  }
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
  const long long total = (long long)y.n * y.h * ((y.w + 1) / 2);   // one thread per pair of output pixels
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
