// ENet's RegularBottleneck with FOUR internal channels as ONE launch (ENet.py:46-100 for channels = 16: regular5_1, at half the
// input resolution): 1x1 (16 -> 4) + BN + act -> 3x3 (4 -> 4, padding = dilation) + BN + act -> 1x1 (4 -> 16) + BN + act ->
// act(x + .).  As three launches the 4-channel convs run one pixel x 4 channels per thread on the direct kernel and the 4-channel
// intermediates cross HBM twice: 1.27 + 0.52 + 0.37 ms of ENet's 19 ms step on 32 x 512 x 1024 pixels for 272 MACs per pixel.
// Here a CTA owns an 8 x 32 pixel tile: phase 1 evaluates the first 1x1 for the tile and its halo into shared memory (fp32, zero
// outside the image = the 3x3 conv's padding), phase 2 / 3 run the 3x3 and the second 1x1 per pixel from there with the weights
// read as shared-memory broadcasts; x is read once (+ halo) and y written once.
#include "esn_common.cuh"

namespace {

constexpr int kTH = 8, kTW = 32;

struct Bneck4Args {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float *w1, *w2, *w3;        // [16][4], [9][4][4], [4][16]  (tap, cin, cout)
  const float *s1, *b1, *s2, *b2, *s3, *b3, *a1, *a2, *a3;
  int N, H, W, x_cs, y_cs, dil, act;
};

__global__ void __launch_bounds__(kTH * kTW) bneck4_kernel(const Bneck4Args a) {
  extern __shared__ __align__(16) float sm[];
  const int d = a.dil;
  const int HT = kTH + 2 * d, WT = kTW + 2 * d;
  float* sw1 = sm;                  // 64
  float* sw2 = sw1 + 64;            // 144
  float* sw3 = sw2 + 144;           // 64
  float* sp = sw3 + 64;             // s1 b1 a1 s2 b2 a2 (4 each) | s3 b3 a3 (16 each) = 24 + 48
  float4* t1 = reinterpret_cast<float4*>(sp + 72);      // [HT][WT]
  for (int i = threadIdx.x; i < 64; i += blockDim.x) { sw1[i] = a.w1[i]; sw3[i] = a.w3[i]; }
  for (int i = threadIdx.x; i < 144; i += blockDim.x) sw2[i] = a.w2[i];
  if (threadIdx.x < 4) {
    const int i = threadIdx.x;
    sp[i] = a.s1[i]; sp[4 + i] = a.b1[i]; sp[8 + i] = a.a1 ? a.a1[i] : 0.f;
    sp[12 + i] = a.s2[i]; sp[16 + i] = a.b2[i]; sp[20 + i] = a.a2 ? a.a2[i] : 0.f;
  }
  if (threadIdx.x < 16) {
    const int i = threadIdx.x;
    sp[24 + i] = a.s3[i]; sp[40 + i] = a.b3[i]; sp[56 + i] = a.a3 ? a.a3[i] : 0.f;
  }
  const int tiles_w = (a.W + kTW - 1) / kTW, tiles_h = (a.H + kTH - 1) / kTH;
  const int tw = blockIdx.x % tiles_w, th = (blockIdx.x / tiles_w) % tiles_h, n = blockIdx.x / (tiles_w * tiles_h);
  const int h0 = th * kTH, w0 = tw * kTW;
  __syncthreads();
  // ---- phase 1: t1 = act(BN(W1 x)) on the tile + halo
  for (int p = threadIdx.x; p < HT * WT; p += blockDim.x) {
    const int r = p / WT, c = p - r * WT;
    const int h = h0 - d + r, w = w0 - d + c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h >= 0 && h < a.H && w >= 0 && w < a.W) {
      const __nv_bfloat16* xp = a.x + ((size_t)((size_t)n * a.H + h) * a.W + w) * a.x_cs;
      float f[16];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(xp)), f);
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(xp) + 1), f + 8);
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
        const float4 wv = *reinterpret_cast<const float4*>(sw1 + ci * 4);
        o[0] = fmaf(f[ci], wv.x, o[0]); o[1] = fmaf(f[ci], wv.y, o[1]);
        o[2] = fmaf(f[ci], wv.z, o[2]); o[3] = fmaf(f[ci], wv.w, o[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = apply_act(fmaf(o[j], sp[j], sp[4 + j]), a.act, sp[8 + j]);
      v = make_float4(o[0], o[1], o[2], o[3]);
    }
    t1[p] = v;
  }
  __syncthreads();
  // ---- phase 2 + 3: own pixel
  const int r = threadIdx.x / kTW, c = threadIdx.x % kTW;
  const int h = h0 + r, w = w0 + c;
  if (h >= a.H || w >= a.W) return;
  float t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kr = 0; kr < 3; ++kr)
#pragma unroll
    for (int kc = 0; kc < 3; ++kc) {
      const float4 v = t1[(r + kr * d) * WT + c + kc * d];
      const float* wt = sw2 + (kr * 3 + kc) * 16;
      const float4 wa = *reinterpret_cast<const float4*>(wt), wb = *reinterpret_cast<const float4*>(wt + 4);
      const float4 wc = *reinterpret_cast<const float4*>(wt + 8), wd = *reinterpret_cast<const float4*>(wt + 12);
      t2[0] = fmaf(v.x, wa.x, fmaf(v.y, wb.x, fmaf(v.z, wc.x, fmaf(v.w, wd.x, t2[0]))));
      t2[1] = fmaf(v.x, wa.y, fmaf(v.y, wb.y, fmaf(v.z, wc.y, fmaf(v.w, wd.y, t2[1]))));
      t2[2] = fmaf(v.x, wa.z, fmaf(v.y, wb.z, fmaf(v.z, wc.z, fmaf(v.w, wd.z, t2[2]))));
      t2[3] = fmaf(v.x, wa.w, fmaf(v.y, wb.w, fmaf(v.z, wc.w, fmaf(v.w, wd.w, t2[3]))));
    }
#pragma unroll
  for (int j = 0; j < 4; ++j) t2[j] = apply_act(fmaf(t2[j], sp[12 + j], sp[16 + j]), a.act, sp[20 + j]);
  const size_t pix = (size_t)((size_t)n * a.H + h) * a.W + w;
  float xin[16], out[16];
  bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(a.x + pix * a.x_cs)), xin);
  bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(a.x + pix * a.x_cs) + 1), xin + 8);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const float4 wv = *reinterpret_cast<const float4*>(sw3 + ci * 16 + q * 4);
      o.x = fmaf(t2[ci], wv.x, o.x); o.y = fmaf(t2[ci], wv.y, o.y);
      o.z = fmaf(t2[ci], wv.z, o.z); o.w = fmaf(t2[ci], wv.w, o.w);
    }
    out[q * 4] = o.x; out[q * 4 + 1] = o.y; out[q * 4 + 2] = o.z; out[q * 4 + 3] = o.w;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float e = apply_act(fmaf(out[j], sp[24 + j], sp[40 + j]), a.act, sp[56 + j]);     // ext = act(BN(conv3))
    out[j] = apply_act(xin[j] + e, a.act, sp[56 + j]);                                          // act(main + ext)
  }
  uint4* yp = reinterpret_cast<uint4*>(a.y + pix * a.y_cs);
  yp[0] = float_to_bf16x8(out);
  yp[1] = float_to_bf16x8(out + 8);
}

}  // namespace

extern "C" int esn_bottleneck4(const EsnBneck4* p, void* stream) {
  if (!p || !esn_valid_nhwc(p->x) || !esn_valid_nhwc(p->y) || !p->w1 || !p->w2 || !p->w3 || !p->scale1 || !p->shift1 ||
      !p->scale2 || !p->shift2 || !p->scale3 || !p->shift3)
    return ESN_ERR_BAD_ARG;
  const EsnTensor &x = p->x, &y = p->y;
  if (x.n != y.n || x.h != y.h || x.w != y.w || x.c != y.c) return ESN_ERR_BAD_SHAPE;
  if (p->act == ESN_ACT_PRELU && (!p->alpha1 || !p->alpha2 || !p->alpha3)) return ESN_ERR_BAD_ARG;
  if (x.c != 16 || x.dtype != ESN_BF16 || y.dtype != ESN_BF16 || x.c_stride % 8 || y.c_stride % 8 || ((uintptr_t)x.ptr & 15) ||
      ((uintptr_t)y.ptr & 15) || p->dilation < 1 || p->dilation > 4 || x.ptr == y.ptr)
    return ESN_ERR_UNSUPPORTED;
  Bneck4Args a;
  a.x = (const __nv_bfloat16*)x.ptr;
  a.y = (__nv_bfloat16*)y.ptr;
  a.w1 = p->w1; a.w2 = p->w2; a.w3 = p->w3;
  a.s1 = p->scale1; a.b1 = p->shift1; a.s2 = p->scale2; a.b2 = p->shift2; a.s3 = p->scale3; a.b3 = p->shift3;
  a.a1 = p->alpha1; a.a2 = p->alpha2; a.a3 = p->alpha3;
  a.N = x.n; a.H = x.h; a.W = x.w; a.x_cs = x.c_stride; a.y_cs = y.c_stride; a.dil = p->dilation; a.act = p->act;
  const int d = p->dilation;
  const size_t smem = (64 + 144 + 64 + 72) * sizeof(float) + (size_t)(kTH + 2 * d) * (kTW + 2 * d) * sizeof(float4);
  const long long ctas = (long long)x.n * esn_cdiv(x.h, kTH) * esn_cdiv(x.w, kTW);
  if (ctas >= (1LL << 31)) return ESN_ERR_UNSUPPORTED;
  bneck4_kernel<<<(unsigned)ctas, kTH * kTW, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
