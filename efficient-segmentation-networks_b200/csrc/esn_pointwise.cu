// Memory-bound NHWC kernels: pools, per-channel affine + activation, layout conversion.
// One thread handles one pixel x 4 channels (128-bit fp32 / 64-bit bf16 accesses) when the view
// is 4-aligned, else one pixel x 1 channel.
#include "esn_common.cuh"

namespace {

struct PwArgs {
  const void* x;
  void* y;
  int N, Hi, Wi, C, x_cs, x_nchw;
  int Ho, Wo, y_cs;
  EpiArgs ep;
};

enum { OP_MAXPOOL2 = 0, OP_AVGPOOL3S2 = 1, OP_AFFINE = 2 };

template <typename TI, int V>
__device__ __forceinline__ void load_px(const PwArgs& a, const TI* x, int n, int h, int w, int c, float* v) {
  if (a.x_nchw) {
#pragma unroll
    for (int j = 0; j < V; ++j)
      v[j] = (c + j < a.C) ? ld1<TI>(x + ((size_t)((size_t)n * a.C + c + j) * a.Hi + h) * a.Wi + w) : 0.f;
  } else {
    const TI* p = x + ((size_t)((size_t)n * a.Hi + h) * a.Wi + w) * a.x_cs + c;
    if (V == 4) {
      const float4 t = ld4<TI>(p);
      v[0] = t.x;
      v[1 % V] = t.y;
      v[2 % V] = t.z;
      v[3 % V] = t.w;
    } else {
      v[0] = ld1<TI>(p);
    }
  }
}

template <typename TI, typename TO, int V, int OP>
__global__ void __launch_bounds__(256) pw_kernel(const PwArgs a) {
  const int ncg = (a.C + V - 1) / V;
  const long long total = (long long)a.N * a.Ho * a.Wo * ncg;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = (int)(idx % ncg);
  const long long pix = idx / ncg;
  const int wo = (int)(pix % a.Wo);
  const int ho = (int)((pix / a.Wo) % a.Ho);
  const int n = (int)(pix / ((long long)a.Wo * a.Ho));
  const int c = cg * V;
  const TI* __restrict__ x = reinterpret_cast<const TI*>(a.x);
  float v[V];
  if (OP == OP_MAXPOOL2 && (2 * ho + 1 >= a.Hi || 2 * wo + 1 >= a.Wi)) {
    // odd input: MaxPool2d(2, 2) has floor(H / 2) rows; the row / column behind them is the zero padding that
    // ESNet's DownsamplerBlock adds before the concat (ESNet.py:25-29), still followed by the BatchNorm slice + ReLU
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = 0.f;
  } else if (OP == OP_MAXPOOL2) {
    float t[V];
    load_px<TI, V>(a, x, n, 2 * ho, 2 * wo, c, v);
    load_px<TI, V>(a, x, n, 2 * ho, 2 * wo + 1, c, t);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], t[j]);
    load_px<TI, V>(a, x, n, 2 * ho + 1, 2 * wo, c, t);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], t[j]);
    load_px<TI, V>(a, x, n, 2 * ho + 1, 2 * wo + 1, c, t);
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], t[j]);
  } else if (OP == OP_AVGPOOL3S2) {
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = 0.f;
    for (int r = -1; r <= 1; ++r) {
      const int hi = 2 * ho + r;
      if (hi < 0 || hi >= a.Hi) continue;
      for (int s = -1; s <= 1; ++s) {
        const int wi = 2 * wo + s;
        if (wi < 0 || wi >= a.Wi) continue;
        float t[V];
        load_px<TI, V>(a, x, n, hi, wi, c, t);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] += t[j];
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] *= (1.f / 9.f);  // count_include_pad=True: always /9
  } else {
    load_px<TI, V>(a, x, n, ho, wo, c, v);
  }
  const size_t opix = ((size_t)((size_t)n * a.Ho + ho) * a.Wo + wo);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int cc = c + j;
    if (cc < a.C) {
      const float sc = a.ep.scale ? __ldg(a.ep.scale + cc) : 1.f;
      const float sh = a.ep.shift ? __ldg(a.ep.shift + cc) : 0.f;
      if (a.ep.res && a.ep.res_first) {
        const size_t ri = opix * a.ep.res_cstride + cc;
        v[j] += (a.ep.res_dtype == ESN_BF16) ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.ep.res)[ri])
                                              : reinterpret_cast<const float*>(a.ep.res)[ri];
      }
      float t = v[j] * sc + sh;
      if (a.ep.res && !a.ep.res_first) {
        if (a.ep.pre_act) t = apply_act(t, a.ep.act, (a.ep.act == ESN_ACT_PRELU) ? __ldg(a.ep.alpha + cc) : 0.f);
        const size_t ri = opix * a.ep.res_cstride + cc;
        t += (a.ep.res_dtype == ESN_BF16) ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.ep.res)[ri])
                                           : reinterpret_cast<const float*>(a.ep.res)[ri];
      }
      const float al = (a.ep.act == ESN_ACT_PRELU) ? __ldg(a.ep.alpha + cc) : 0.f;
      v[j] = apply_act(t, a.ep.act, al);
    }
  }
  TO* yp = reinterpret_cast<TO*>(a.y) + opix * a.y_cs + c;
  if (V == 4)
    st4<TO>(yp, make_float4(v[0], v[1 % V], v[2 % V], v[3 % V]));
  else
    st1<TO>(yp, v[0]);
}

// ---- vector path: NHWC in, NHWC out, 16-byte accesses.  A thread owns one group of V channels (its
// scale / shift / slope live in registers) and walks pixels of one output row; the threads of a CTA cover
// ncg channel groups x ppb pixels per step, channel groups fastest, so a warp touches whole pixels.
// Channel counts that are not a multiple of V (35, 131, 259 ... concat slices) load the full vector -- it
// stays inside the pixel stride -- and store the valid lanes one by one.
template <typename T, int V> struct PwVec;
template <> struct PwVec<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float* f) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct PwVec<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) {
    const float4 t = ld4<__nv_bfloat16>(p);
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* f) {
    st4<__nv_bfloat16>(p, make_float4(f[0], f[1], f[2], f[3]));
  }
};
template <> struct PwVec<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    bf16x8_to_float(r, f);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* f) {
    *reinterpret_cast<uint4*>(p) = float_to_bf16x8(f);
  }
};

template <typename TI, typename TO, int V, int OP>
__global__ void __launch_bounds__(256) pw_vec_kernel(const PwArgs a, const int ncg, const int ppb, const int wpb) {
  const int pl = threadIdx.x / ncg;
  if (pl >= ppb) return;
  const int c = (threadIdx.x - pl * ncg) * V;
  const int row = blockIdx.x;
  const int n = row / a.Ho, ho = row - n * a.Ho;
  float sc[V], sh[V], al[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const bool ok = c + j < a.C;
    sc[j] = (ok && a.ep.scale) ? __ldg(a.ep.scale + c + j) : 1.f;
    sh[j] = (ok && a.ep.shift) ? __ldg(a.ep.shift + c + j) : 0.f;
    al[j] = (ok && a.ep.act == ESN_ACT_PRELU) ? __ldg(a.ep.alpha + c + j) : 0.f;
  }
  const bool full = c + V <= a.C;
  const int act = a.ep.act;
  const TI* __restrict__ x = reinterpret_cast<const TI*>(a.x) + (size_t)n * a.Hi * a.Wi * a.x_cs + c;
  TO* __restrict__ y = reinterpret_cast<TO*>(a.y) + (size_t)row * a.Wo * a.y_cs + c;
  const int w_end = min(a.Wo, (int)(blockIdx.y + 1) * wpb);
  for (int wo = blockIdx.y * wpb + pl; wo < w_end; wo += ppb) {
    float v[V];
    if (OP == OP_MAXPOOL2) {
      float t[V];
      const TI* p0 = x + ((size_t)(2 * ho) * a.Wi + 2 * wo) * a.x_cs;
      PwVec<TI, V>::ld(p0, v);
      PwVec<TI, V>::ld(p0 + a.x_cs, t);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], t[j]);
      PwVec<TI, V>::ld(p0 + (size_t)a.Wi * a.x_cs, t);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], t[j]);
      PwVec<TI, V>::ld(p0 + (size_t)a.Wi * a.x_cs + a.x_cs, t);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = fmaxf(v[j], t[j]);
    } else if (OP == OP_AVGPOOL3S2) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = 0.f;
#pragma unroll
      for (int r = -1; r <= 1; ++r) {
        const int hi = 2 * ho + r;
        if (hi < 0 || hi >= a.Hi) continue;
#pragma unroll
        for (int q = -1; q <= 1; ++q) {
          const int wi = 2 * wo + q;
          if (wi < 0 || wi >= a.Wi) continue;
          float t[V];
          PwVec<TI, V>::ld(x + ((size_t)hi * a.Wi + wi) * a.x_cs, t);
#pragma unroll
          for (int j = 0; j < V; ++j) v[j] += t[j];
        }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] *= (1.f / 9.f);
    } else {
      PwVec<TI, V>::ld(x + ((size_t)ho * a.Wi + wo) * a.x_cs, v);
    }
    if (a.ep.res) {
      float r[V];
      const size_t ri = ((size_t)row * a.Wo + wo) * a.ep.res_cstride + c;
      if (a.ep.res_dtype == ESN_BF16) PwVec<__nv_bfloat16, V>::ld(reinterpret_cast<const __nv_bfloat16*>(a.ep.res) + ri, r);
      else if (V == 4) PwVec<float, 4>::ld(reinterpret_cast<const float*>(a.ep.res) + ri, r);
      if (a.ep.res_first) {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = apply_act(fmaf(v[j] + r[j], sc[j], sh[j]), act, al[j]);
      } else if (a.ep.pre_act) {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = apply_act(apply_act(fmaf(v[j], sc[j], sh[j]), act, al[j]) + r[j], act, al[j]);
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = apply_act(fmaf(v[j], sc[j], sh[j]) + r[j], act, al[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = apply_act(fmaf(v[j], sc[j], sh[j]), act, al[j]);
    }
    TO* yp = y + (size_t)wo * a.y_cs;
    if (full) {
      PwVec<TO, V>::st(yp, v);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (c + j < a.C) st1<TO>(yp + j, v[j]);
    }
  }
}

template <typename TI, typename TO, int V, int OP>
void launch_pw_vec(const PwArgs& a, cudaStream_t st) {
  const int ncg = (a.C + V - 1) / V;
  const int ppb = 256 / ncg;
  int iters = 4;
  while (iters > 1 && (long long)a.N * a.Ho * ((a.Wo + ppb * iters - 1) / (ppb * iters)) < 148 * 8) iters >>= 1;
  const int wpb = ppb * iters;
  dim3 grid((unsigned)(a.N * a.Ho), (unsigned)((a.Wo + wpb - 1) / wpb));
  pw_vec_kernel<TI, TO, V, OP><<<grid, 256, 0, st>>>(a, ncg, ppb, wpb);
}

// MaxPool2d(2,2) + affine + activation into a channel slice whose start is NOT 16-byte aligned (DABNet's first
// DownSamplingBlock writes the 35 pooled channels at channel 29 of its 64-channel output, DABNet.py:104-108).  The
// scalar path did 4 x 35 two-byte loads per output pixel; here a thread owns one ALIGNED 16-byte vector of the output
// row (slice channels 8j-OFF .. 8j-OFF+7), reads the two aligned input vectors that straddle it for each of the four
// window pixels, and stores the vector whole when all 8 channels belong to the slice (per-channel stores at the two ends).
template <int OFF>
__global__ void __launch_bounds__(256) maxpool2x2_shift_kernel(const PwArgs a, const int nv) {
  const long long total = (long long)a.N * a.Ho * a.Wo * nv;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % nv);
  const long long pix = idx / nv;
  const int wo = (int)(pix % a.Wo);
  const int ho = (int)((pix / a.Wo) % a.Ho);
  const int n = (int)(pix / ((long long)a.Wo * a.Ho));
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a.x);
  const int k0 = 8 * j - OFF;                      // first slice channel of this output vector
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat16* px = x + ((size_t)((size_t)n * a.Hi + 2 * ho + (q >> 1)) * a.Wi + 2 * wo + (q & 1)) * a.x_cs;
    float fa[8], fb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) fa[i] = fb[i] = -INFINITY;
    if (OFF > 0 && j >= 1) bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(px + 8 * (j - 1))), fa);
    if (8 * j < a.x_cs) bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(px + 8 * j)), fb);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = (i < OFF) ? fa[8 - OFF + i] : fb[i - OFF];     // input channel k0 + i
      m[i] = fmaxf(m[i], v);
    }
  }
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)pix * a.y_cs + k0;   // 16-byte aligned
  float o[8];
  bool all = true;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + i;
    const bool ok = k >= 0 && k < a.C;
    all = all && ok;
    float t = 0.f;
    if (ok) {
      const float sc = a.ep.scale ? __ldg(a.ep.scale + k) : 1.f, sh = a.ep.shift ? __ldg(a.ep.shift + k) : 0.f;
      const float al = a.ep.act == ESN_ACT_PRELU ? __ldg(a.ep.alpha + k) : 0.f;
      t = apply_act(fmaf(m[i], sc, sh), a.ep.act, al);
    }
    o[i] = t;
  }
  if (all) {
    *reinterpret_cast<uint4*>(yp) = float_to_bf16x8(o);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (k0 + i >= 0 && k0 + i < a.C) yp[i] = __float2bfloat16_rn(o[i]);
  }
}

template <int OP>
int run_pw(const EsnPool* p, void* stream) {
  if (!p) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (!esn_valid_nhwc(y)) return ESN_ERR_BAD_ARG;
  const bool nchw = x.layout == ESN_NCHW;
  if (nchw ? (!x.ptr || x.dtype != ESN_F32) : !esn_valid_nhwc(x)) return ESN_ERR_BAD_ARG;
  if (x.n != y.n || x.c != y.c) return ESN_ERR_BAD_SHAPE;
  // max-pool: floor(H / 2) x floor(W / 2), or -- odd sizes -- ceil x ceil with a zero last row / column (scalar kernel only)
  const bool padded = OP == OP_MAXPOOL2 && (y.h != x.h / 2 || y.w != x.w / 2);
  if (padded && (y.h != (x.h + 1) / 2 || y.w != (x.w + 1) / 2)) return ESN_ERR_BAD_SHAPE;
  if (OP == OP_AVGPOOL3S2 && (y.h != (x.h - 1) / 2 + 1 || y.w != (x.w - 1) / 2 + 1)) return ESN_ERR_BAD_SHAPE;
  if (OP == OP_AFFINE && (y.h != x.h || y.w != x.w)) return ESN_ERR_BAD_SHAPE;
  int rc = esn_check_epilogue(p->ep, y, OP == OP_AFFINE);
  if (rc) return rc;
  PwArgs a;
  a.x = x.ptr;
  a.y = y.ptr;
  a.N = x.n;
  a.Hi = x.h;
  a.Wi = x.w;
  a.C = x.c;
  a.x_cs = nchw ? 0 : x.c_stride;
  a.x_nchw = nchw;
  a.Ho = y.h;
  a.Wo = y.w;
  a.y_cs = y.c_stride;
  a.ep = make_epi(p->ep);
  const size_t ysz = y.dtype == ESN_F32 ? 4 : 2, xsz = x.dtype == ESN_F32 ? 4 : 2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!nchw && !padded) {
    const bool all16 = x.dtype == ESN_BF16 && y.dtype == ESN_BF16 && (!p->ep.residual.ptr || p->ep.residual.dtype == ESN_BF16);
    const int V = all16 ? 8 : 4;
    const EsnTensor& r = p->ep.residual;
    const bool res_ok = !r.ptr || (r.c_stride % V == 0 && (uintptr_t)r.ptr % (V * (r.dtype == ESN_F32 ? 4 : 2)) == 0);
    if (x.c_stride % V == 0 && y.c_stride % V == 0 && (uintptr_t)x.ptr % (V * xsz) == 0 && (uintptr_t)y.ptr % (V * ysz) == 0 &&
        res_ok && (y.c + V - 1) / V <= 256) {
      if (all16) launch_pw_vec<__nv_bfloat16, __nv_bfloat16, 8, OP>(a, st);
      else if (x.dtype == ESN_F32 && y.dtype == ESN_F32) launch_pw_vec<float, float, 4, OP>(a, st);
      else if (x.dtype == ESN_F32) launch_pw_vec<float, __nv_bfloat16, 4, OP>(a, st);
      else if (y.dtype == ESN_F32) launch_pw_vec<__nv_bfloat16, float, 4, OP>(a, st);
      else launch_pw_vec<__nv_bfloat16, __nv_bfloat16, 4, OP>(a, st);
      ESN_CHECK_LAUNCH();
      return ESN_OK;
    }
  }
  static const bool pool_noshift = getenv("ESN_POOL_NOSHIFT") != nullptr;
  if (OP == OP_MAXPOOL2 && !padded && !nchw && x.dtype == ESN_BF16 && y.dtype == ESN_BF16 && !p->ep.residual.ptr && x.c_stride % 8 == 0 &&
      y.c_stride % 8 == 0 && (uintptr_t)x.ptr % 16 == 0 && (uintptr_t)y.ptr % 2 == 0 && !pool_noshift) {
    // unaligned output slice: aligned-vector kernel shifted by OFF channels (elements outside the slice are never touched)
    const int off = (int)(((uintptr_t)y.ptr / 2) % 8);
    const int nv = (y.c + off + 7) / 8;
    const long long total = (long long)y.n * y.h * y.w * nv;
    const int grid = esn_cdiv(total, 256);
    switch (off) {
      case 1: maxpool2x2_shift_kernel<1><<<grid, 256, 0, st>>>(a, nv); break;
      case 2: maxpool2x2_shift_kernel<2><<<grid, 256, 0, st>>>(a, nv); break;
      case 3: maxpool2x2_shift_kernel<3><<<grid, 256, 0, st>>>(a, nv); break;
      case 4: maxpool2x2_shift_kernel<4><<<grid, 256, 0, st>>>(a, nv); break;
      case 5: maxpool2x2_shift_kernel<5><<<grid, 256, 0, st>>>(a, nv); break;
      case 6: maxpool2x2_shift_kernel<6><<<grid, 256, 0, st>>>(a, nv); break;
      case 7: maxpool2x2_shift_kernel<7><<<grid, 256, 0, st>>>(a, nv); break;
      default: maxpool2x2_shift_kernel<0><<<grid, 256, 0, st>>>(a, nv); break;
    }
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  const bool v4 = (y.c % 4 == 0) && (y.c_stride % 4 == 0) && ((uintptr_t)y.ptr % (4 * ysz) == 0) &&
                  (nchw || ((x.c_stride % 4 == 0) && ((uintptr_t)x.ptr % (4 * xsz) == 0)));
  const int V = v4 ? 4 : 1;
  const long long total = (long long)y.n * y.h * y.w * ((y.c + V - 1) / V);
  const int block = 256, grid = esn_cdiv(total, block);
#define ESN_PW_LAUNCH(TI, TO)                                    \
  do {                                                           \
    if (v4)                                                      \
      pw_kernel<TI, TO, 4, OP><<<grid, block, 0, st>>>(a);       \
    else                                                         \
      pw_kernel<TI, TO, 1, OP><<<grid, block, 0, st>>>(a);       \
  } while (0)
  if (x.dtype == ESN_F32 && y.dtype == ESN_F32)
    ESN_PW_LAUNCH(float, float);
  else if (x.dtype == ESN_F32 && y.dtype == ESN_BF16)
    ESN_PW_LAUNCH(float, __nv_bfloat16);
  else if (x.dtype == ESN_BF16 && y.dtype == ESN_BF16)
    ESN_PW_LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (x.dtype == ESN_BF16 && y.dtype == ESN_F32)
    ESN_PW_LAUNCH(__nv_bfloat16, float);
  else
    return ESN_ERR_BAD_ARG;
#undef ESN_PW_LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

// ---- NCHW f32/bf16 <-> NHWC f32/bf16 through a 32x32 smem transpose over (C, W) per (n,h)
template <typename TI, typename TO>
__global__ void nchw_to_nhwc_kernel(const TI* __restrict__ x, TO* __restrict__ y, int N, int C, int H, int W, int y_cs) {
  __shared__ float tile[32][33];
  const int nh = blockIdx.z;
  const int n = nh / H, h = nh % H;
  const int c0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, w = w0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && w < W) ? ld1<TI>(x + ((size_t)((size_t)n * C + c) * H + h) * W + w) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int w = w0 + i, c = c0 + threadIdx.x;
    if (w < W && c < C) st1<TO>(y + ((size_t)((size_t)n * H + h) * W + w) * y_cs + c, tile[threadIdx.x][i]);
  }
}

template <typename TI, typename TO>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ x, TO* __restrict__ y, int N, int C, int H, int W, int x_cs) {
  __shared__ float tile[32][33];
  const int nh = blockIdx.z;
  const int n = nh / H, h = nh % H;
  const int c0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int w = w0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && w < W) ? ld1<TI>(x + ((size_t)((size_t)n * H + h) * W + w) * x_cs + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, w = w0 + threadIdx.x;
    if (w < W && c < C) st1<TO>(y + ((size_t)((size_t)n * C + c) * H + h) * W + w, tile[threadIdx.x][i]);
  }
}

}  // namespace

// ---- concat tail: y[:, 0:c) = act(x * scale + shift), y[:, c:tail_c) = 0, eight channels (16 bytes of bf16) per thread.
// Every pixel's tail is written as whole vectors (whole 32-byte sectors), so the concat buffer needs no zero fill and the
// three injected channels cost no partial-sector read-modify-write.
namespace {
template <typename TO>
__global__ void __launch_bounds__(256) concat_tail_kernel(const float* __restrict__ x, TO* __restrict__ y, long long total,
                                                          int c, int y_cs, int vshift, EpiArgs ep) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long pix = idx >> vshift;
  const int v = (int)(idx & ((1 << vshift) - 1));
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (v == 0) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x) + pix);
    const float in[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < c) {
        const float r = in[j] * (ep.scale ? __ldg(ep.scale + j) : 1.f) + (ep.shift ? __ldg(ep.shift + j) : 0.f);
        f[j] = apply_act(r, ep.act, ep.act == ESN_ACT_PRELU ? __ldg(ep.alpha + j) : 0.f);
      }
  }
  TO* yp = y + pix * y_cs + 8 * v;
  st4<TO>(yp, make_float4(f[0], f[1], f[2], f[3]));
  st4<TO>(yp + 4, make_float4(f[4], f[5], f[6], f[7]));
}
template <>
__global__ void __launch_bounds__(256) concat_tail_kernel<__nv_bfloat16>(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                                         long long total, int c, int y_cs, int vshift, EpiArgs ep) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long pix = idx >> vshift;
  const int v = (int)(idx & ((1 << vshift) - 1));
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (v == 0) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x) + pix);
    const float in[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < c) {
        const float r = in[j] * (ep.scale ? __ldg(ep.scale + j) : 1.f) + (ep.shift ? __ldg(ep.shift + j) : 0.f);
        f[j] = apply_act(r, ep.act, ep.act == ESN_ACT_PRELU ? __ldg(ep.alpha + j) : 0.f);
      }
  }
  *reinterpret_cast<uint4*>(y + pix * y_cs + 8 * v) = float_to_bf16x8(f);
}
}  // namespace

extern "C" int esn_concat_tail(const EsnPool* p, int32_t tail_c, void* stream) {
  if (!p || !esn_valid_nhwc(p->x) || !esn_valid_nhwc(p->y)) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (x.n != y.n || x.h != y.h || x.w != y.w || x.c != y.c) return ESN_ERR_BAD_SHAPE;
  // whole pixels of 8-channel vectors, a power of two of them (DABNet: 32 or 64 channels)
  if (x.dtype != ESN_F32 || x.c > 4 || x.c_stride != 4 || tail_c < 8 || tail_c > y.c_stride || (tail_c & (tail_c - 1)))
    return ESN_ERR_UNSUPPORTED;
  if (p->ep.residual.ptr) return ESN_ERR_UNSUPPORTED;
  int rc = esn_check_epilogue(p->ep, y);
  if (rc) return rc;
  const size_t ysz = y.dtype == ESN_F32 ? 4 : 2;
  if (((uintptr_t)x.ptr % 16) || ((uintptr_t)y.ptr % (8 * ysz)) || y.c_stride % 8) return ESN_ERR_ALIGN;
  int vshift = 0;
  while ((8 << vshift) < tail_c) ++vshift;
  const long long total = ((long long)x.n * x.h * x.w) << vshift;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const EpiArgs ep = make_epi(p->ep);
  if (y.dtype == ESN_F32)
    concat_tail_kernel<float><<<grid, 256, 0, st>>>((const float*)x.ptr, (float*)y.ptr, total, x.c, y.c_stride, vshift, ep);
  else
    concat_tail_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x.ptr, (__nv_bfloat16*)y.ptr, total, x.c, y.c_stride, vshift, ep);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_maxpool2x2_affine_act(const EsnPool* p, void* stream) { return run_pw<OP_MAXPOOL2>(p, stream); }
extern "C" int esn_avgpool3x3s2_affine_act(const EsnPool* p, void* stream) { return run_pw<OP_AVGPOOL3S2>(p, stream); }
extern "C" int esn_affine_act(const EsnPool* p, void* stream) { return run_pw<OP_AFFINE>(p, stream); }

extern "C" int esn_convert_layout(const EsnTensor* x, const EsnTensor* y, void* stream) {
  if (!x || !y || !x->ptr || !y->ptr) return ESN_ERR_BAD_ARG;
  if (x->n != y->n || x->h != y->h || x->w != y->w || x->c != y->c) return ESN_ERR_BAD_SHAPE;
  if ((x->dtype != ESN_F32 && x->dtype != ESN_BF16) || (y->dtype != ESN_F32 && y->dtype != ESN_BF16))
    return ESN_ERR_BAD_ARG;
  const int N = x->n, C = x->c, H = x->h, W = x->w;
  if ((long long)N * H > 65535LL * 32768) return ESN_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 block(32, 8), grid(esn_cdiv(W, 32), esn_cdiv(C, 32), N * H);
  if (grid.z > 65535) {  // fold rows: the kernels only need (n,h) as one index
    return ESN_ERR_UNSUPPORTED;
  }
#define ESN_CVT(K, TI, TO, CS) K<TI, TO><<<grid, block, 0, st>>>((const TI*)x->ptr, (TO*)y->ptr, N, C, H, W, CS)
  if (x->layout == ESN_NCHW && y->layout == ESN_NHWC) {
    if (x->dtype == ESN_F32 && y->dtype == ESN_F32) ESN_CVT(nchw_to_nhwc_kernel, float, float, y->c_stride);
    else if (x->dtype == ESN_F32) ESN_CVT(nchw_to_nhwc_kernel, float, __nv_bfloat16, y->c_stride);
    else if (y->dtype == ESN_F32) ESN_CVT(nchw_to_nhwc_kernel, __nv_bfloat16, float, y->c_stride);
    else ESN_CVT(nchw_to_nhwc_kernel, __nv_bfloat16, __nv_bfloat16, y->c_stride);
  } else if (x->layout == ESN_NHWC && y->layout == ESN_NCHW) {
    if (x->dtype == ESN_F32 && y->dtype == ESN_F32) ESN_CVT(nhwc_to_nchw_kernel, float, float, x->c_stride);
    else if (x->dtype == ESN_F32) ESN_CVT(nhwc_to_nchw_kernel, float, __nv_bfloat16, x->c_stride);
    else if (y->dtype == ESN_F32) ESN_CVT(nhwc_to_nchw_kernel, __nv_bfloat16, float, x->c_stride);
    else ESN_CVT(nhwc_to_nchw_kernel, __nv_bfloat16, __nv_bfloat16, x->c_stride);
  } else {
    return ESN_ERR_UNSUPPORTED;
  }
#undef ESN_CVT
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
