// DABNet depthwise asymmetric pair: both branches, their four BN+PReLU stages, the branch add and
// bn_relu_2 in ONE pass over the tensor (DABNet.py:73-80).  HBM-bound: 1 read + 1 write of a
// C x h x w tensor; the 17 neighbour reads per output come from L1/L2.
//
// Parameter block prm, fp32 [27][C]:
//   rows  0-2  dconv3x1 taps      3-5  dconv1x3 taps      (branch 1, dilation 1)
//   rows  6-8  ddconv3x1 taps     9-11 ddconv1x3 taps     (branch 2, dilation d)
//   rows 12-14 scale,shift,alpha after dconv3x1   15-17 after dconv1x3
//   rows 18-20 after ddconv3x1    21-23 after ddconv1x3   24-26 bn_relu_2 (after the add)
#include "esn_common.cuh"

namespace {

struct DabArgs {
  const void* x;
  void* y;
  const float* prm;
  int N, H, W, C, x_cs, y_cs, d;
};

__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
  const float2 lo = ffma2(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  const float2 hi = ffma2(make_float2(a.z, a.w), make_float2(b.z, b.w), make_float2(c.z, c.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float prelu1(float v, float al) { return v >= 0.f ? v : v * al; }
__device__ __forceinline__ float4 affine_prelu(float4 v, float4 sc, float4 sh, float4 al) {
  const float4 t = f4_fma(v, sc, sh);
  return make_float4(prelu1(t.x, al.x), prelu1(t.y, al.y), prelu1(t.z, al.z), prelu1(t.w, al.w));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) dab_dw_pair_kernel(const DabArgs a) {
  const int ncg = a.C / 4;
  const long long total = (long long)a.N * a.H * a.W * ncg;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % ncg) * 4;
  const long long pix = idx / ncg;
  const int w = (int)(pix % a.W);
  const int h = (int)((pix / a.W) % a.H);
  const int n = (int)(pix / ((long long)a.W * a.H));
  const TI* __restrict__ x = reinterpret_cast<const TI*>(a.x) + (size_t)n * a.H * a.W * a.x_cs + c;
  auto P = [&](int row) { return __ldg(reinterpret_cast<const float4*>(a.prm + (size_t)row * a.C + c)); };
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  // one branch: taps at rows wbase..wbase+2 (3x1) and wbase+3..wbase+5 (1x3), affines at abase..abase+5
  auto branch = [&](int wbase, int abase, int d) -> float4 {
    const float4 wa0 = P(wbase), wa1 = P(wbase + 1), wa2 = P(wbase + 2);
    const float4 sa = P(abase), ba = P(abase + 1), aa = P(abase + 2);
    float4 acc = zero;
#pragma unroll
    for (int s = -1; s <= 1; ++s) {
      const int ww = w + s * d;
      if (ww < 0 || ww >= a.W) continue;  // zero padding applies to the intermediate tensor
      float4 t = zero;
      if (h - d >= 0) t = f4_fma(ld4<TI>(x + ((size_t)(h - d) * a.W + ww) * a.x_cs), wa0, t);
      t = f4_fma(ld4<TI>(x + ((size_t)h * a.W + ww) * a.x_cs), wa1, t);
      if (h + d < a.H) t = f4_fma(ld4<TI>(x + ((size_t)(h + d) * a.W + ww) * a.x_cs), wa2, t);
      t = affine_prelu(t, sa, ba, aa);
      acc = f4_fma(t, P(wbase + 4 + s), acc);
    }
    return affine_prelu(acc, P(abase + 3), P(abase + 4), P(abase + 5));
  };
  const float4 b1 = branch(0, 12, 1);
  const float4 b2 = branch(6, 18, a.d);
  const float4 sum = make_float4(b1.x + b2.x, b1.y + b2.y, b1.z + b2.z, b1.w + b2.w);
  const float4 out = affine_prelu(sum, P(24), P(25), P(26));
  st4<TO>(reinterpret_cast<TO*>(a.y) + (size_t)pix * a.y_cs + c, out);
}

// Strip variant: one thread = 4 channels x P consecutive pixels of one row.  The 27 parameter rows are read
// once per strip instead of once per pixel (they were 70 % of the L1 traffic), branch 1 (dilation 1) slides
// its three stage-1 columns along the strip (3 instead of 9 loads per pixel), and the two branches run one
// after the other so only one branch's parameters are live in registers.
template <typename TI, typename TO, int P>
__global__ void __launch_bounds__(128) dab_dw_pair_strip_kernel(const DabArgs a) {
  const int ncg = a.C / 4;
  const int S = (a.W + P - 1) / P;
  const long long total = (long long)a.N * a.H * S * ncg;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % ncg) * 4;
  const long long t1 = idx / ncg;
  const int w0 = (int)(t1 % S) * P;
  const int h = (int)((t1 / S) % a.H);
  const int n = (int)(t1 / ((long long)S * a.H));
  const TI* __restrict__ x = reinterpret_cast<const TI*>(a.x) + (size_t)n * a.H * a.W * a.x_cs + c;
  auto PR = [&](int row) { return __ldg(reinterpret_cast<const float4*>(a.prm + (size_t)row * a.C + c)); };
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 r[P];

  auto run_branch = [&](const int wbase, const int abase, const int d, const bool first) {
    const float4 wa0 = PR(wbase), wa1 = PR(wbase + 1), wa2 = PR(wbase + 2);
    const float4 sa = PR(abase), ba = PR(abase + 1), aa = PR(abase + 2);
    const float4 wb0 = PR(wbase + 3), wb1 = PR(wbase + 4), wb2 = PR(wbase + 5);
    const float4 sb = PR(abase + 3), bb = PR(abase + 4), ab = PR(abase + 5);
    const bool up = h - d >= 0, dn = h + d < a.H;
    const TI* xu = x + (size_t)(h - d) * a.W * a.x_cs;
    const TI* xc = x + (size_t)h * a.W * a.x_cs;
    const TI* xd = x + (size_t)(h + d) * a.W * a.x_cs;
    // stage 1 at one column: vertical 3-tap conv + BN + PReLU; zero outside the row (padding of the intermediate)
    auto T = [&](const int col) -> float4 {
      if (col < 0 || col >= a.W) return zero;
      const size_t o = (size_t)col * a.x_cs;
      float4 t = zero;
      if (up) t = f4_fma(ld4<TI>(xu + o), wa0, t);
      t = f4_fma(ld4<TI>(xc + o), wa1, t);
      if (dn) t = f4_fma(ld4<TI>(xd + o), wa2, t);
      return affine_prelu(t, sa, ba, aa);
    };
    auto S2 = [&](const float4 tl, const float4 tc, const float4 tr) -> float4 {
      float4 acc = f4_fma(tl, wb0, zero);
      acc = f4_fma(tc, wb1, acc);
      acc = f4_fma(tr, wb2, acc);
      return affine_prelu(acc, sb, bb, ab);
    };
    if (d == 1) {
      float4 tl = T(w0 - 1), tc = T(w0);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const float4 tr = T(w0 + p + 1);
        const float4 v = S2(tl, tc, tr);
        r[p] = first ? v : make_float4(r[p].x + v.x, r[p].y + v.y, r[p].z + v.z, r[p].w + v.w);
        tl = tc;
        tc = tr;
      }
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const int w = w0 + p;
        const float4 v = (w < a.W) ? S2(T(w - d), T(w), T(w + d)) : zero;
        r[p] = first ? v : make_float4(r[p].x + v.x, r[p].y + v.y, r[p].z + v.z, r[p].w + v.w);
      }
    }
  };
  run_branch(0, 12, 1, true);
  run_branch(6, 18, a.d, false);
  const float4 s2 = PR(24), b2 = PR(25), a2 = PR(26);
  TO* y = reinterpret_cast<TO*>(a.y) + ((size_t)((size_t)n * a.H + h) * a.W + w0) * a.y_cs + c;
#pragma unroll
  for (int p = 0; p < P; ++p)
    if (w0 + p < a.W) st4<TO>(y + (size_t)p * a.y_cs, affine_prelu(r[p], s2, b2, a2));
}

// Row variant (default): one CTA = one image row x a chunk of CC channels.  Stage 1 of both branches (vertical 3-tap
// conv + BN + PReLU) is evaluated ONCE per pixel into shared memory (fp32, zero columns either side = the padding of
// the intermediate tensor), stage 2 reads its three (dilated) columns from there: 5 coalesced global loads and 2
// stage-1 evaluations per pixel instead of 12 and 4, no bounds tests in the inner loops.  The rows above / below are
// the centre rows of neighbouring CTAs, so the vertical halo is served by L2 and DRAM traffic stays 1 read + 1 write.
// BN scales are folded into the taps once per thread (a thread's 4 channels are loop-invariant).
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_prelu(float4 t, float4 al) {
  return make_float4(prelu1(t.x, al.x), prelu1(t.y, al.y), prelu1(t.z, al.z), prelu1(t.w, al.w));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) dab_dw_pair_row_kernel(const DabArgs a, const int CC, const int nchunk) {
  extern __shared__ __align__(16) float4 dab_sm[];
  const int G = CC >> 2, gshift = 31 - __clz(G);          // 4-channel groups per pixel in this chunk (power of two)
  const int chunk = blockIdx.x % nchunk;
  const int row = blockIdx.x / nchunk;                     // n * H + h
  const int h = row % a.H;
  const int d = a.d, W = a.W;
  float4* __restrict__ T1 = dab_sm;                        // [(W + 2)][G]   column w lives at (w + 1)
  float4* __restrict__ T2 = dab_sm + (size_t)(W + 2) * G;  // [(W + 2d)][G]  column w lives at (w + d)
  const int items = W * G;
  const int g = threadIdx.x & (G - 1);                     // loop-invariant: blockDim.x % G == 0
  const int c = chunk * CC + g * 4;
  auto PR = [&](int r) { return __ldg(reinterpret_cast<const float4*>(a.prm + (size_t)r * a.C + c)); };
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < G; i += blockDim.x) { T1[i] = zero; T1[(W + 1) * G + i] = zero; }
  for (int i = threadIdx.x; i < d * G; i += blockDim.x) { T2[i] = zero; T2[(W + d) * G + i] = zero; }
  const TI* __restrict__ xrow = reinterpret_cast<const TI*>(a.x) + (size_t)row * W * a.x_cs + c;   // pixel (h, 0)

  auto stage1 = [&](const int wbase, const int abase, const int dd, float4* __restrict__ T, const int toff) {
    const float4 sc = PR(abase), sh = PR(abase + 1), al = PR(abase + 2);
    const float4 w0 = f4_mul(PR(wbase), sc), w1 = f4_mul(PR(wbase + 1), sc), w2 = f4_mul(PR(wbase + 2), sc);
    const bool up = h - dd >= 0, dn = h + dd < a.H;
    const ptrdiff_t ro = (ptrdiff_t)dd * W * a.x_cs;
#pragma unroll 4
    for (int i = threadIdx.x; i < items; i += 256) {
      const TI* p = xrow + (size_t)(i >> gshift) * a.x_cs;
      float4 t = f4_fma(ld4<TI>(p), w1, sh);
      if (up) t = f4_fma(ld4<TI>(p - ro), w0, t);
      if (dn) t = f4_fma(ld4<TI>(p + ro), w2, t);
      T[i + toff] = f4_prelu(t, al);
    }
  };
  stage1(0, 12, 1, T1, G);
  stage1(6, 18, d, T2, d * G);
  __syncthreads();

  const float4 sb1 = PR(15), hb1 = PR(16), ab1 = PR(17);
  const float4 u0 = f4_mul(PR(3), sb1), u1 = f4_mul(PR(4), sb1), u2 = f4_mul(PR(5), sb1);
  const float4 sb2 = PR(21), hb2 = PR(22), ab2 = PR(23);
  const float4 v0 = f4_mul(PR(9), sb2), v1 = f4_mul(PR(10), sb2), v2 = f4_mul(PR(11), sb2);
  const float4 s3 = PR(24), h3 = PR(25), a3 = PR(26);
  TO* __restrict__ yrow = reinterpret_cast<TO*>(a.y) + (size_t)row * W * a.y_cs + c;
  const int dG = d * G;
#pragma unroll 4
  for (int i = threadIdx.x; i < items; i += 256) {
    float4 b1 = f4_fma(T1[i], u0, hb1);
    b1 = f4_fma(T1[i + G], u1, b1);
    b1 = f4_prelu(f4_fma(T1[i + 2 * G], u2, b1), ab1);
    float4 b2 = f4_fma(T2[i], v0, hb2);
    b2 = f4_fma(T2[i + dG], v1, b2);
    b2 = f4_prelu(f4_fma(T2[i + 2 * dG], v2, b2), ab2);
    const float4 sum = make_float4(b1.x + b2.x, b1.y + b2.y, b1.z + b2.z, b1.w + b2.w);
    st4<TO>(yrow + (size_t)(i >> gshift) * a.y_cs, f4_prelu(f4_fma(sum, s3, h3), a3));
  }
}

template <typename TI, typename TO>
static int launch_dab_row(const DabArgs& a, int CC, size_t smem, cudaStream_t st) {
  auto kfn = dab_dw_pair_row_kernel<TI, TO>;
  static std::atomic<bool> attr_done[kEsnMaxDevices];   // per instantiation and device; idempotent, a race only repeats the call
  const int dev = esn_current_device();
  if (!attr_done[dev].load(std::memory_order_acquire)) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) return ESN_ERR_CUDA;
    attr_done[dev].store(true, std::memory_order_release);
  }
  const int nchunk = a.C / CC;
  kfn<<<(unsigned)((long long)a.N * a.H * nchunk), 256, smem, st>>>(a, CC, nchunk);
  return ESN_OK;
}

}  // namespace

extern "C" int esn_dab_dw_pair(const EsnDabPair* p, void* stream) {
  if (!p || !p->prm || !esn_valid_nhwc(p->x) || !esn_valid_nhwc(p->y)) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (x.n != y.n || x.h != y.h || x.w != y.w || x.c != y.c) return ESN_ERR_BAD_SHAPE;
  if (p->dilation < 1) return ESN_ERR_BAD_ARG;
  if (x.c % 4 || x.c_stride % 4 || y.c_stride % 4) return ESN_ERR_UNSUPPORTED;
  const size_t xsz = x.dtype == ESN_F32 ? 4 : 2, ysz = y.dtype == ESN_F32 ? 4 : 2;
  if (((uintptr_t)x.ptr % (4 * xsz)) || ((uintptr_t)y.ptr % (4 * ysz)) || ((uintptr_t)p->prm % 16)) return ESN_ERR_ALIGN;
  DabArgs a;
  a.x = x.ptr;
  a.y = y.ptr;
  a.prm = p->prm;
  a.N = x.n;
  a.H = x.h;
  a.W = x.w;
  a.C = x.c;
  a.x_cs = x.c_stride;
  a.y_cs = y.c_stride;
  a.d = p->dilation;
  const long long total = (long long)x.n * x.h * x.w * (x.c / 4);
  const int block = 256, grid = esn_cdiv(total, block);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool no_row = getenv("ESN_DAB_NOROW") != nullptr;
  if (!no_row && (long long)x.n * x.h * (x.c / 4) < (1ll << 31)) {
    // channel chunk: the largest power of two dividing C whose two stage-1 rows fit 72 KB (3 CTAs per SM)
    int CC = 4;
    while (CC * 2 <= 64 && x.c % (CC * 2) == 0) CC *= 2;
    auto bytes = [&](int cc) { return (size_t)((x.w + 2) + (x.w + 2 * p->dilation)) * (cc / 4) * sizeof(float4); };
    while (CC > 4 && bytes(CC) > 72 * 1024) CC /= 2;
    if (bytes(CC) <= 96 * 1024) {
      int rc;
      if (x.dtype == ESN_F32 && y.dtype == ESN_F32) rc = launch_dab_row<float, float>(a, CC, bytes(CC), st);
      else if (x.dtype == ESN_BF16 && y.dtype == ESN_BF16) rc = launch_dab_row<__nv_bfloat16, __nv_bfloat16>(a, CC, bytes(CC), st);
      else if (x.dtype == ESN_F32) rc = launch_dab_row<float, __nv_bfloat16>(a, CC, bytes(CC), st);
      else rc = launch_dab_row<__nv_bfloat16, float>(a, CC, bytes(CC), st);
      if (rc) return rc;
      ESN_CHECK_LAUNCH();
      return ESN_OK;
    }
  }
  static const bool no_strip = getenv("ESN_DAB_NOSTRIP") != nullptr;
  if (!no_strip && x.w >= 8) {
    constexpr int P = 8;
    const long long tot = (long long)x.n * x.h * ((x.w + P - 1) / P) * (x.c / 4);
    const int g2 = esn_cdiv(tot, 128);
    if (x.dtype == ESN_F32 && y.dtype == ESN_F32) dab_dw_pair_strip_kernel<float, float, P><<<g2, 128, 0, st>>>(a);
    else if (x.dtype == ESN_BF16 && y.dtype == ESN_BF16) dab_dw_pair_strip_kernel<__nv_bfloat16, __nv_bfloat16, P><<<g2, 128, 0, st>>>(a);
    else if (x.dtype == ESN_F32) dab_dw_pair_strip_kernel<float, __nv_bfloat16, P><<<g2, 128, 0, st>>>(a);
    else dab_dw_pair_strip_kernel<__nv_bfloat16, float, P><<<g2, 128, 0, st>>>(a);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  if (x.dtype == ESN_F32 && y.dtype == ESN_F32)
    dab_dw_pair_kernel<float, float><<<grid, block, 0, st>>>(a);
  else if (x.dtype == ESN_BF16 && y.dtype == ESN_BF16)
    dab_dw_pair_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>(a);
  else if (x.dtype == ESN_F32)
    dab_dw_pair_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>(a);
  else
    dab_dw_pair_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

// ------------------------------------------------------------------------------------------------
// Depthwise k_h x k_w convolution (k <= 3, any dilation, stride 1 or 2), NHWC, 128-bit accesses:
// one thread = one output pixel x 8 channels (bf16) or 4 channels (fp32); per-channel taps and the
// fused epilogue parameters come from shared memory.  HBM-bound: 1 read + 1 write of the tensor, the
// neighbour re-reads hit L1/L2.  Used by CGNet (F_loc / F_sur, CGNet.py:106-171), DABNet's training path,
// Fast-SCNN / ESPNetv2 depthwise layers.
namespace {

struct DwArgs {
  const void* x;
  void* y;
  const float* w;   // [taps][C]
  int N, Hi, Wi, C, x_cs, Ho, Wo, y_cs;
  int kh, kw, stride, pad_h, pad_w, dil_h, dil_w;
  EpiArgs ep;
};

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p)), f);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    *reinterpret_cast<uint4*>(p) = float_to_bf16x8(f);
  }
};

template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(const DwArgs a) {
  constexpr int V = Vec<T>::N;
  extern __shared__ __align__(16) float sm[];   // w[taps][C] | scale[C] | shift[C] | alpha[C]
  const int taps = a.kh * a.kw;
  float* sw = sm;
  float* sp = sm + taps * a.C;
  for (int i = threadIdx.x; i < taps * a.C; i += blockDim.x) sw[i] = a.w[i];
  for (int i = threadIdx.x; i < a.C; i += blockDim.x) {
    sp[i] = a.ep.scale ? a.ep.scale[i] : 1.f;
    sp[a.C + i] = a.ep.shift ? a.ep.shift[i] : 0.f;
    sp[2 * a.C + i] = (a.ep.act == ESN_ACT_PRELU) ? a.ep.alpha[i] : 0.f;
  }
  __syncthreads();
  const int ncg = a.C / V;
  const long long total = (long long)a.N * a.Ho * a.Wo * ncg;
  // grid-stride loop: the per-CTA parameter staging above is amortised over many pixels
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
  const int c = (int)(idx % ncg) * V;
  const long long pix = idx / ncg;
  const int wo = (int)(pix % a.Wo);
  const int ho = (int)((pix / a.Wo) % a.Ho);
  const int n = (int)(pix / ((long long)a.Wo * a.Ho));
  const T* x = reinterpret_cast<const T*>(a.x) + (size_t)n * a.Hi * a.Wi * a.x_cs + c;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  for (int r = 0; r < a.kh; ++r) {
    const int hi = ho * a.stride - a.pad_h + r * a.dil_h;
    if (hi < 0 || hi >= a.Hi) continue;
    for (int s = 0; s < a.kw; ++s) {
      const int wi = wo * a.stride - a.pad_w + s * a.dil_w;
      if (wi < 0 || wi >= a.Wi) continue;
      float xv[V];
      Vec<T>::load(x + ((size_t)hi * a.Wi + wi) * a.x_cs, xv);
      const float* wt = sw + (r * a.kw + s) * a.C + c;
      float wv[V];
#pragma unroll
      for (int j = 0; j < V; j += 4) {     // 128-bit shared-memory reads of the per-channel taps
        const float4 t = *reinterpret_cast<const float4*>(wt + j);
        wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = fmaf(xv[j], wv[j], acc[j]);
    }
  }
  float res[V];
  if (a.ep.res) Vec<T>::load(reinterpret_cast<const T*>(a.ep.res) + (size_t)pix * a.ep.res_cstride + c, res);
  float psc[V], psh[V], pal[V];
#pragma unroll
  for (int j = 0; j < V; j += 4) {
    const float4 t0 = *reinterpret_cast<const float4*>(sp + c + j);
    const float4 t1 = *reinterpret_cast<const float4*>(sp + a.C + c + j);
    const float4 t2 = *reinterpret_cast<const float4*>(sp + 2 * a.C + c + j);
    psc[j] = t0.x; psc[j + 1] = t0.y; psc[j + 2] = t0.z; psc[j + 3] = t0.w;
    psh[j] = t1.x; psh[j + 1] = t1.y; psh[j + 2] = t1.z; psh[j + 3] = t1.w;
    pal[j] = t2.x; pal[j + 1] = t2.y; pal[j + 2] = t2.z; pal[j + 3] = t2.w;
  }
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float t = fmaf(acc[j], psc[j], psh[j]);
    if (a.ep.res) {
      if (a.ep.pre_act) t = apply_act(t, a.ep.act, pal[j]);
      t += res[j];
    }
    acc[j] = apply_act(t, a.ep.act, pal[j]);
  }
  Vec<T>::store(reinterpret_cast<T*>(a.y) + (size_t)pix * a.y_cs + c, acc);
  }
}

}  // namespace

// Called by esn_conv2d_direct for depthwise convs whose views are 16-byte vectorisable; returns false if not.
bool esn_dwconv_try(const EsnConv* p, void* stream, int* rc) {
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (x.layout != ESN_NHWC || x.dtype != y.dtype || p->transposed || p->groups != x.c || x.c != y.c) return false;
  const int V = x.dtype == ESN_BF16 ? 8 : 4;
  if (x.c % V || x.c_stride % V || y.c_stride % V || ((uintptr_t)x.ptr % 16) || ((uintptr_t)y.ptr % 16)) return false;
  if (p->ep.residual.ptr && (p->ep.residual.dtype != x.dtype || p->ep.residual.c_stride % V || ((uintptr_t)p->ep.residual.ptr % 16)))
    return false;
  const int taps = p->kh * p->kw;
  const size_t smem = (size_t)(taps + 3) * x.c * sizeof(float);
  if (smem > 48 * 1024) return false;
  DwArgs a;
  a.x = x.ptr; a.y = y.ptr; a.w = reinterpret_cast<const float*>(p->w);
  a.N = x.n; a.Hi = x.h; a.Wi = x.w; a.C = x.c; a.x_cs = x.c_stride; a.Ho = y.h; a.Wo = y.w; a.y_cs = y.c_stride;
  a.kh = p->kh; a.kw = p->kw; a.stride = p->stride; a.pad_h = p->pad_h; a.pad_w = p->pad_w; a.dil_h = p->dil_h; a.dil_w = p->dil_w;
  a.ep = make_epi(p->ep);
  const long long total = (long long)y.n * y.h * y.w * (x.c / V);
  int grid = esn_cdiv(total, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x.dtype == ESN_BF16) dwconv_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a);
  else dwconv_kernel<float><<<grid, 256, smem, st>>>(a);
  g_esn_launches.fetch_add(1, std::memory_order_relaxed);
  *rc = (cudaPeekAtLastError() == cudaSuccess) ? ESN_OK : ESN_ERR_CUDA;
  if (*rc != ESN_OK) cudaGetLastError();
  return true;
}
