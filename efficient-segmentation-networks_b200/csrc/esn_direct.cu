// CUDA-core direct convolution (fp32 accumulate), NHWC, with the fused epilogue of esn.h.
// This is the exact-arithmetic path (fp32 parity against the reference) and the path for the
// shapes the tcgen05 kernel does not take: Cin = 3 stems read straight from the caller's NCHW
// fp32 image, depthwise convs, odd channel counts.
#include "esn_common.cuh"

namespace {

struct DirectArgs {
  const void* x;
  void* y;
  const float* w;
  int N, Hi, Wi, Cin, x_cs;
  int Ho, Wo, Cout, y_cs;
  int kh, kw, stride, pad_h, pad_w, dil_h, dil_w, transposed, dw, x_nchw;
  EpiArgs ep;
};

template <typename TI, typename TO, int COV, int CIV>
__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectArgs a) {
  const int ncog = (a.Cout + COV - 1) / COV;
  const long long total = (long long)a.N * a.Ho * a.Wo * ncog;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cog = (int)(idx % ncog);
  const long long pix = idx / ncog;
  const int wo = (int)(pix % a.Wo);
  const int ho = (int)((pix / a.Wo) % a.Ho);
  const int n = (int)(pix / ((long long)a.Wo * a.Ho));
  const int co = cog * COV;
  const TI* __restrict__ x = reinterpret_cast<const TI*>(a.x);

  float acc[COV];
#pragma unroll
  for (int j = 0; j < COV; ++j) acc[j] = 0.f;

  for (int r = 0; r < a.kh; ++r) {
    int hi;
    if (!a.transposed) {
      hi = ho * a.stride - a.pad_h + r * a.dil_h;
    } else {
      const int t = ho + a.pad_h - r * a.dil_h;
      if (t < 0 || (t % a.stride) != 0) continue;
      hi = t / a.stride;
    }
    if (hi < 0 || hi >= a.Hi) continue;
    for (int s = 0; s < a.kw; ++s) {
      int wi;
      if (!a.transposed) {
        wi = wo * a.stride - a.pad_w + s * a.dil_w;
      } else {
        const int t = wo + a.pad_w - s * a.dil_w;
        if (t < 0 || (t % a.stride) != 0) continue;
        wi = t / a.stride;
      }
      if (wi < 0 || wi >= a.Wi) continue;
      const int tap = r * a.kw + s;
      if (a.dw) {
        const float* wt = a.w + (size_t)tap * a.Cout + co;
        const TI* xp = x + ((size_t)((size_t)n * a.Hi + hi) * a.Wi + wi) * a.x_cs + co;
        if (COV == 4) {
          const float4 xv = ld4<TI>(xp);
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wt));
          acc[0] += xv.x * wv.x;
          acc[1 % COV] += xv.y * wv.y;
          acc[2 % COV] += xv.z * wv.z;
          acc[3 % COV] += xv.w * wv.w;
        } else {
          acc[0] += ld1<TI>(xp) * __ldg(wt);
        }
      } else {
        const float* wt = a.w + (size_t)tap * a.Cin * a.Cout + co;
        if (a.x_nchw) {
          for (int ci = 0; ci < a.Cin; ++ci) {
            const float xv = ld1<TI>(x + ((size_t)((size_t)n * a.Cin + ci) * a.Hi + hi) * a.Wi + wi);
            const float* wr = wt + (size_t)ci * a.Cout;
#pragma unroll
            for (int j = 0; j < COV; ++j)
              if (co + j < a.Cout) acc[j] += xv * __ldg(wr + j);
          }
        } else {
          const TI* xp = x + ((size_t)((size_t)n * a.Hi + hi) * a.Wi + wi) * a.x_cs;
          for (int ci = 0; ci < a.Cin; ci += CIV) {
            float xv[CIV];
            if (CIV == 4) {
              const float4 t = ld4<TI>(xp + ci);
              xv[0] = t.x;
              xv[1 % CIV] = t.y;
              xv[2 % CIV] = t.z;
              xv[3 % CIV] = t.w;
            } else {
              xv[0] = ld1<TI>(xp + ci);
            }
#pragma unroll
            for (int i = 0; i < CIV; ++i) {
              const float* wr = wt + (size_t)(ci + i) * a.Cout;
              if (COV == 4) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(wr));
                acc[0] += xv[i] * wv.x;
                acc[1 % COV] += xv[i] * wv.y;
                acc[2 % COV] += xv[i] * wv.z;
                acc[3 % COV] += xv[i] * wv.w;
              } else {
                acc[0] += xv[i] * __ldg(wr);
              }
            }
          }
        }
      }
    }
  }

  // ---- fused epilogue
  const size_t opix = ((size_t)((size_t)n * a.Ho + ho) * a.Wo + wo);
  float v[COV];
#pragma unroll
  for (int j = 0; j < COV; ++j) {
    const int c = co + j;
    if (c < a.Cout) {
      const float sc = a.ep.scale ? __ldg(a.ep.scale + c) : 1.f;
      const float sh = a.ep.shift ? __ldg(a.ep.shift + c) : 0.f;
      float t = acc[j] * sc + sh;
      if (a.ep.res) {
        if (a.ep.pre_act) t = apply_act(t, a.ep.act, (a.ep.act == ESN_ACT_PRELU) ? __ldg(a.ep.alpha + c) : 0.f);
        const size_t ri = opix * a.ep.res_cstride + c;
        t += (a.ep.res_dtype == ESN_BF16) ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.ep.res)[ri])
                                           : reinterpret_cast<const float*>(a.ep.res)[ri];
      }
      const float al = (a.ep.act == ESN_ACT_PRELU) ? __ldg(a.ep.alpha + c) : 0.f;
      v[j] = apply_act(t, a.ep.act, al);
    } else {
      v[j] = 0.f;
    }
  }
  TO* yp = reinterpret_cast<TO*>(a.y) + opix * a.y_cs + co;
  if (COV == 4) {
    st4<TO>(yp, make_float4(v[0], v[1 % COV], v[2 % COV], v[3 % COV]));
  } else {
    st1<TO>(yp, v[0]);
  }
}

// ConvTranspose2d(k=2, stride 2, pad 0), Cout <= 32: every input pixel owns its 2x2 output block, so one
// thread reads its Cin values once per tap (L1) against shared-memory weights and writes four NHWC pixels
// through the fused epilogue.  (ESPNet's up_l3 / up_l2, ESPNet.py:346-350.)
template <typename TI, typename TO>
__global__ void __launch_bounds__(128) convt2x2_nhwc_kernel(const DirectArgs a, const int cp) {
  extern __shared__ float sw[];  // [4][Cin][cp]
  for (int i = threadIdx.x; i < 4 * a.Cin * cp; i += blockDim.x) {
    const int co = i % cp, tc = i / cp;
    sw[i] = co < a.Cout ? a.w[(size_t)tc * a.Cout + co] : 0.f;
  }
  __syncthreads();
  const long long total = (long long)a.N * a.Hi * a.Wi;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % a.Wi);
  const int i = (int)((idx / a.Wi) % a.Hi);
  const int n = (int)(idx / ((long long)a.Wi * a.Hi));
  const TI* xp = reinterpret_cast<const TI*>(a.x) + (size_t)idx * a.x_cs;
  const int nq = cp / 4;
#pragma unroll 1
  for (int tap = 0; tap < 4; ++tap) {
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    const float* wt = sw + (size_t)tap * a.Cin * cp;
    for (int c = 0; c < a.Cin; ++c) {
      const float v = ld1<TI>(xp + c);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nq) {
          const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)c * cp + 4 * q);
          acc[4 * q] = fmaf(v, w4.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
        }
    }
    const int ho = 2 * i + (tap >> 1), wo = 2 * j + (tap & 1);
    const size_t opix = ((size_t)((size_t)n * a.Ho + ho) * a.Wo + wo);
    TO* yp = reinterpret_cast<TO*>(a.y) + opix * a.y_cs;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < a.Cout) {
        const float sc = a.ep.scale ? __ldg(a.ep.scale + k) : 1.f;
        const float sh = a.ep.shift ? __ldg(a.ep.shift + k) : 0.f;
        const float al = (a.ep.act == ESN_ACT_PRELU) ? __ldg(a.ep.alpha + k) : 0.f;
        float t = acc[k] * sc + sh;
        if (a.ep.res) {
          if (a.ep.pre_act) t = apply_act(t, a.ep.act, al);
          const size_t ri = opix * a.ep.res_cstride + k;
          t += (a.ep.res_dtype == ESN_BF16) ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.ep.res)[ri])
                                             : reinterpret_cast<const float*>(a.ep.res)[ri];
        }
        st1<TO>(yp + k, apply_act(t, a.ep.act, al));
      }
  }
}

template <typename TI, typename TO>
int launch_direct(const DirectArgs& a, bool cov4, bool civ4, cudaStream_t st) {
  const int cov = cov4 ? 4 : 1;
  const long long total = (long long)a.N * a.Ho * a.Wo * ((a.Cout + cov - 1) / cov);
  const int block = 256;
  const int grid = esn_cdiv(total, block);
  if (cov4 && civ4)
    conv_direct_kernel<TI, TO, 4, 4><<<grid, block, 0, st>>>(a);
  else if (cov4)
    conv_direct_kernel<TI, TO, 4, 1><<<grid, block, 0, st>>>(a);
  else if (civ4)
    conv_direct_kernel<TI, TO, 1, 4><<<grid, block, 0, st>>>(a);
  else
    conv_direct_kernel<TI, TO, 1, 1><<<grid, block, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

}  // namespace

int esn_check_epilogue(const EsnEpilogue& e, const EsnTensor& y, bool allow_residual_first) {
  if (e.act < ESN_ACT_NONE || e.act > ESN_ACT_PRELU) return ESN_ERR_BAD_ARG;
  if ((e.flags & ESN_EP_RESIDUAL_FIRST) && !allow_residual_first) return ESN_ERR_UNSUPPORTED;
  if (e.act == ESN_ACT_PRELU && !e.alpha) return ESN_ERR_BAD_ARG;
  if (e.residual.ptr) {
    const EsnTensor& r = e.residual;
    if (!esn_valid_nhwc(r)) return ESN_ERR_BAD_ARG;
    if (r.n != y.n || r.h != y.h || r.w != y.w || r.c != y.c) return ESN_ERR_BAD_SHAPE;
  }
  return ESN_OK;
}

bool esn_dwconv_try(const EsnConv* p, void* stream, int* rc);   // esn_stencil.cu: vectorised depthwise path (gather)
bool esn_dw_strip_try(const EsnConv* p, void* stream, int* rc); // esn_dw_strip.cu: register-strip depthwise path (stride 1, 'same')

namespace {

// Transposed depthwise 3x3 / stride 2 (the input gradient of Fast-SCNN's / ESPNetv2's strided depthwise convs,
// FastSCNN.py:28-45,62-82, ESPNet_v2/Model.py:15-99): y[ho, wo, c] = sum over the taps (r, s) with ho + pad - r and wo + pad - s even
// of x[(ho + pad - r) / 2, (wo + pad - s) / 2, c] * w[r][s][c] (+ residual).  An output pixel has 1, 2 or 4 such taps, decided by
// its parities: a thread owns one output pixel x 8 channels and visits exactly those (the generic kernel above walks all nine
// with a modulo per tap, 4 channels per thread: 1.1 ms per launch on Fast-SCNN's 535 MB gradients, 5.6 ms of its 30.6 ms step).
__global__ void __launch_bounds__(256) dwT3x3s2_v8_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          const float* __restrict__ w, const __nv_bfloat16* __restrict__ res,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          long long total, int Hi, int Wi, int Ho, int Wo, int C, int x_cs, int y_cs,
                                                          int res_cs, int pad_h, int pad_w) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ng = C / 8;
  const int c = (int)(idx % ng) * 8;
  const long long pix = idx / ng;
  const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho);
  const long long n = pix / ((long long)Wo * Ho);
  const int th = ho + pad_h, tw = wo + pad_w;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  // rows: th odd -> r = 1; th even -> r = 0 and r = 2
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int r = (th & 1) ? 1 : 2 * a;
    if ((th & 1) && a == 1) break;
    const int t = th - r;
    if (t < 0) continue;
    const int hi = t >> 1;
    if (hi >= Hi) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int s = (tw & 1) ? 1 : 2 * b;
      if ((tw & 1) && b == 1) break;
      const int u = tw - s;
      if (u < 0) continue;
      const int wi = u >> 1;
      if (wi >= Wi) continue;
      float xv[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + ((size_t)(n * Hi + hi) * Wi + wi) * x_cs + c)), xv);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(r * 3 + s) * C + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(r * 3 + s) * C + c + 4));
      acc[0] = fmaf(xv[0], w0.x, acc[0]); acc[1] = fmaf(xv[1], w0.y, acc[1]);
      acc[2] = fmaf(xv[2], w0.z, acc[2]); acc[3] = fmaf(xv[3], w0.w, acc[3]);
      acc[4] = fmaf(xv[4], w1.x, acc[4]); acc[5] = fmaf(xv[5], w1.y, acc[5]);
      acc[6] = fmaf(xv[6], w1.z, acc[6]); acc[7] = fmaf(xv[7], w1.w, acc[7]);
    }
  }
  if (scale || shift) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(acc[j], scale ? __ldg(scale + c + j) : 1.f, shift ? __ldg(shift + c + j) : 0.f);
  }
  if (res) {
    float rv[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(res + (size_t)pix * res_cs + c), rv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += rv[j];
  }
  *reinterpret_cast<uint4*>(y + (size_t)pix * y_cs + c) = float_to_bf16x8(acc);
}

inline bool al16(const void* p, int cs) { return p && cs % 8 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int esn_conv2d_direct(const EsnConv* p, void* stream) {
  if (!p || !p->w) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (!esn_valid_nhwc(y)) return ESN_ERR_BAD_ARG;
  const bool nchw = x.layout == ESN_NCHW;
  if (nchw) {
    if (!x.ptr || x.dtype != ESN_F32) return ESN_ERR_BAD_ARG;
  } else if (!esn_valid_nhwc(x)) {
    return ESN_ERR_BAD_ARG;
  }
  if (p->kh < 1 || p->kw < 1 || p->stride < 1 || p->dil_h < 1 || p->dil_w < 1) return ESN_ERR_BAD_ARG;
  const bool dw = p->groups != 1;
  if (dw && (p->groups != x.c || x.c != y.c || nchw)) return ESN_ERR_UNSUPPORTED;
  if (x.n != y.n) return ESN_ERR_BAD_SHAPE;
  int eh, ew;
  if (!p->transposed) {
    eh = (x.h + 2 * p->pad_h - p->dil_h * (p->kh - 1) - 1) / p->stride + 1;
    ew = (x.w + 2 * p->pad_w - p->dil_w * (p->kw - 1) - 1) / p->stride + 1;
    if (eh != y.h || ew != y.w) return ESN_ERR_BAD_SHAPE;
  } else {
    eh = (x.h - 1) * p->stride - 2 * p->pad_h + p->dil_h * (p->kh - 1) + 1;
    ew = (x.w - 1) * p->stride - 2 * p->pad_w + p->dil_w * (p->kw - 1) + 1;
    if (y.h < eh || y.h >= eh + p->stride || y.w < ew || y.w >= ew + p->stride) return ESN_ERR_BAD_SHAPE;
  }
  int rc = esn_check_epilogue(p->ep, y);
  if (rc) return rc;
  if (dw && p->transposed && p->kh == 3 && p->kw == 3 && p->stride == 2 && p->dil_h == 1 && p->dil_w == 1 && x.dtype == ESN_BF16 &&
      y.dtype == ESN_BF16 && x.c % 8 == 0 && al16(x.ptr, x.c_stride) && al16(y.ptr, y.c_stride) && ((uintptr_t)p->w % 16) == 0 &&
      p->ep.act == ESN_ACT_NONE && p->ep.flags == 0 &&
      (!p->ep.residual.ptr || (p->ep.residual.dtype == ESN_BF16 && al16(p->ep.residual.ptr, p->ep.residual.c_stride)))) {
    const long long total = (long long)y.n * y.h * y.w * (y.c / 8);
    dwT3x3s2_v8_kernel<<<esn_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const __nv_bfloat16*)x.ptr, (__nv_bfloat16*)y.ptr, reinterpret_cast<const float*>(p->w),
        (const __nv_bfloat16*)p->ep.residual.ptr, p->ep.scale, p->ep.shift, total, x.h, x.w, y.h, y.w, y.c, x.c_stride, y.c_stride, p->ep.residual.c_stride,
        p->pad_h, p->pad_w);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  if (dw && !getenv("ESN_DISABLE_DW_STRIP") && esn_dw_strip_try(p, stream, &rc)) return rc;
  if (dw && esn_dwconv_try(p, stream, &rc)) return rc;

  DirectArgs a;
  a.x = x.ptr;
  a.y = y.ptr;
  a.w = reinterpret_cast<const float*>(p->w);
  a.N = x.n;
  a.Hi = x.h;
  a.Wi = x.w;
  a.Cin = x.c;
  a.x_cs = nchw ? 0 : x.c_stride;
  a.Ho = y.h;
  a.Wo = y.w;
  a.Cout = y.c;
  a.y_cs = y.c_stride;
  a.kh = p->kh;
  a.kw = p->kw;
  a.stride = p->stride;
  a.pad_h = p->pad_h;
  a.pad_w = p->pad_w;
  a.dil_h = p->dil_h;
  a.dil_w = p->dil_w;
  a.transposed = p->transposed;
  a.dw = dw;
  a.x_nchw = nchw;
  a.ep = make_epi(p->ep);

  const size_t ysz = y.dtype == ESN_F32 ? 4 : 2, xsz = x.dtype == ESN_F32 ? 4 : 2;
  const bool scale_al = (!p->ep.scale || ((uintptr_t)p->ep.scale % 16) == 0);
  bool cov4 = (y.c % 4 == 0) && (y.c_stride % 4 == 0) && ((uintptr_t)y.ptr % (4 * ysz) == 0) &&
              ((uintptr_t)p->w % 16 == 0) && scale_al;
  bool civ4 = !nchw && !dw && (x.c % 4 == 0) && (x.c_stride % 4 == 0) && ((uintptr_t)x.ptr % (4 * xsz) == 0);
  if (dw) {  // depthwise reads x with the same vector width as it writes y
    cov4 = cov4 && (x.c_stride % 4 == 0) && ((uintptr_t)x.ptr % (4 * xsz) == 0);
    civ4 = false;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->transposed && !dw && !nchw && p->kh == 2 && p->kw == 2 && p->stride == 2 && p->pad_h == 0 && p->pad_w == 0 &&
      p->dil_h == 1 && p->dil_w == 1 && y.c <= 32 && y.h == 2 * x.h && y.w == 2 * x.w) {
    const int cp = (y.c + 3) / 4 * 4;
    const size_t smem = (size_t)4 * x.c * cp * sizeof(float);
    if (smem <= 48 * 1024) {
      const int grid = esn_cdiv((long long)x.n * x.h * x.w, 128);
      if (x.dtype == ESN_F32 && y.dtype == ESN_F32) convt2x2_nhwc_kernel<float, float><<<grid, 128, smem, st>>>(a, cp);
      else if (x.dtype == ESN_F32) convt2x2_nhwc_kernel<float, __nv_bfloat16><<<grid, 128, smem, st>>>(a, cp);
      else if (y.dtype == ESN_F32) convt2x2_nhwc_kernel<__nv_bfloat16, float><<<grid, 128, smem, st>>>(a, cp);
      else convt2x2_nhwc_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 128, smem, st>>>(a, cp);
      ESN_CHECK_LAUNCH();
      return ESN_OK;
    }
  }
  if (x.dtype == ESN_F32 && y.dtype == ESN_F32) return launch_direct<float, float>(a, cov4, civ4, st);
  if (x.dtype == ESN_F32 && y.dtype == ESN_BF16) return launch_direct<float, __nv_bfloat16>(a, cov4, civ4, st);
  if (x.dtype == ESN_BF16 && y.dtype == ESN_BF16)
    return launch_direct<__nv_bfloat16, __nv_bfloat16>(a, cov4, civ4, st);
  if (x.dtype == ESN_BF16 && y.dtype == ESN_F32) return launch_direct<__nv_bfloat16, float>(a, cov4, civ4, st);
  return ESN_ERR_BAD_ARG;
}
