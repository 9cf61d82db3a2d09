// Register-strip depthwise convolution: host side + the device helpers of csrc/esn_dw_strip_kernel.cuh.
// Reached through esn_conv2d_direct (include/esn.h) for depthwise stride-1 "same" convs with 1x3 / 3x1 / 3x3 taps.
#include "esn_common.cuh"

namespace {
template <typename T> struct DwsRaw;
template <> struct DwsRaw<float> {
  typedef float4 type;
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct DwsRaw<__nv_bfloat16> {
  typedef uint4 type;
  static __device__ __forceinline__ uint4 zero() { return make_uint4(0u, 0u, 0u, 0u); }
};
__device__ __forceinline__ float4 ldw4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldraw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint4 ldraw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack(const float4& r, float2 (&f)[2]) {
  f[0] = make_float2(r.x, r.y);
  f[1] = make_float2(r.z, r.w);
}
__device__ __forceinline__ void unpack(const uint4& r, float2 (&f)[4]) {
  // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked (two ALU ops per pair, no conversion unit)
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = make_float2(__uint_as_float(u[i] << 16), __uint_as_float(u[i] & 0xffff0000u));
}
__device__ __forceinline__ void stv(float* p, float2 (&v)[2]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
}
__device__ __forceinline__ void stv(__nv_bfloat16* p, float2 (&v)[4]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[i].x, v[i].y);
  *reinterpret_cast<uint4*>(p) = r;
}
template <int ACT> __device__ __forceinline__ float2 dws_act2(float2 v, float2 al) {
  if (ACT == ESN_ACT_RELU) return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
  if (ACT == ESN_ACT_PRELU) return make_float2(v.x >= 0.f ? v.x : v.x * al.x, v.y >= 0.f ? v.y : v.y * al.y);
  return v;
}
}  // namespace

#include "esn_dw_strip_kernel.cuh"

namespace {
template <typename T, int V, int KH, int KW, int TW, int ACT, int RES>
int launch_strip2(const DwsArgs& a, cudaStream_t st) {
  const long long grid = (a.total + 127) / 128;
  if (grid > 0x7fffffffLL) return ESN_ERR_UNSUPPORTED;
  // the 3x3 kernel on 8-channel vectors holds 72 taps + 48 accumulators + 3 row buffers: ~210 registers, 2 CTAs per SM
  // (measured: capping it at 168 registers for a third CTA spills and is 5-10 % slower)
  constexpr int MINB = (KH == 3 && KW == 3 && V == 8) ? 2 : 3;
  dw_strip_kernel<T, V, KH, KW, TW, ACT, RES, MINB><<<(unsigned)grid, 128, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
template <typename T, int V, int KH, int KW, int TW>
int launch_strip(DwsArgs a, cudaStream_t st) {
  // 3 CTAs x 128 threads resident per SM; keep >= ~4 waves of them before chains are allowed to be long (each chain pays two
  // halo rows, so longer is cheaper per output row, but a tail of idle SMs costs more)
  dws_plan(a, V, KH, KW, TW, 148LL * 256 * 4);
  const bool res = a.res != nullptr;
  switch (a.act) {
    case ESN_ACT_RELU:
      return res ? launch_strip2<T, V, KH, KW, TW, ESN_ACT_RELU, 1>(a, st) : launch_strip2<T, V, KH, KW, TW, ESN_ACT_RELU, 0>(a, st);
    case ESN_ACT_PRELU:
      return res ? launch_strip2<T, V, KH, KW, TW, ESN_ACT_PRELU, 1>(a, st) : launch_strip2<T, V, KH, KW, TW, ESN_ACT_PRELU, 0>(a, st);
    default:
      return res ? launch_strip2<T, V, KH, KW, TW, ESN_ACT_NONE, 1>(a, st) : launch_strip2<T, V, KH, KW, TW, ESN_ACT_NONE, 0>(a, st);
  }
}
}  // namespace

// Called by esn_conv2d_direct before the gather kernel; returns false when the shape is not a strip shape.
bool esn_dw_strip_try(const EsnConv* p, void* stream, int* rc) {
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (x.layout != ESN_NHWC || y.layout != ESN_NHWC || x.dtype != y.dtype || p->transposed || p->groups != x.c || x.c != y.c)
    return false;
  if (p->stride != 1 || x.h != y.h || x.w != y.w || x.n != y.n) return false;
  const int kh = p->kh, kw = p->kw;
  if (!((kh == 3 || kh == 1) && (kw == 3 || kw == 1)) || (kh == 1 && kw == 1)) return false;
  if (p->pad_h != (kh / 2) * p->dil_h && !(kh == 1 && p->pad_h == 0)) return false;
  if (p->pad_w != (kw / 2) * p->dil_w && !(kw == 1 && p->pad_w == 0)) return false;
  if (kh == 1 && p->pad_h != 0) return false;
  if (kw == 1 && p->pad_w != 0) return false;
  const int V = x.dtype == ESN_BF16 ? 8 : 4;
  if (x.c % V || x.c_stride % V || y.c_stride % V || ((uintptr_t)x.ptr % 16) || ((uintptr_t)y.ptr % 16)) return false;
  const EsnTensor& r = p->ep.residual;
  if (r.ptr && (r.dtype != x.dtype || r.layout != ESN_NHWC || r.c_stride % V || ((uintptr_t)r.ptr % 16) || r.n != y.n || r.h != y.h ||
                r.w != y.w || r.c != y.c))
    return false;
  if (p->ep.flags & ESN_EP_RESIDUAL_FIRST) return false;     // (acc + res) * scale: the scale cannot be folded into the taps
  if (p->ep.act == ESN_ACT_PRELU && !p->ep.alpha) return false;
  if (((uintptr_t)p->w | (uintptr_t)p->ep.scale | (uintptr_t)p->ep.shift | (uintptr_t)p->ep.alpha) % 16) return false;
  if (p->dil_h < 1 || p->dil_w < 1 || p->dil_h > 64 || p->dil_w > 64) return false;
  DwsArgs a;
  a.x = x.ptr; a.y = y.ptr; a.w = reinterpret_cast<const float*>(p->w);
  a.scale = p->ep.scale; a.shift = p->ep.shift; a.alpha = p->ep.alpha;
  a.res = r.ptr; a.res_cs = r.c_stride;
  a.act = p->ep.act; a.pre_act = (p->ep.flags & ESN_EP_ACT_BEFORE_RESIDUAL) ? 1 : 0;
  a.N = x.n; a.H = x.h; a.W = x.w; a.C = x.c; a.x_cs = x.c_stride; a.y_cs = y.c_stride;
  a.dil_h = p->dil_h; a.dil_w = p->dil_w;
  a.seg = a.QW = a.VS = 0; a.total = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x.dtype == ESN_BF16) {
    if (kh == 3 && kw == 3) *rc = launch_strip<__nv_bfloat16, 8, 3, 3, 2>(a, st);
    else if (kh == 3) *rc = launch_strip<__nv_bfloat16, 8, 3, 1, 2>(a, st);
    else *rc = launch_strip<__nv_bfloat16, 8, 1, 3, 4>(a, st);
  } else {
    if (kh == 3 && kw == 3) *rc = launch_strip<float, 4, 3, 3, 4>(a, st);
    else if (kh == 3) *rc = launch_strip<float, 4, 3, 1, 4>(a, st);
    else *rc = launch_strip<float, 4, 1, 3, 4>(a, st);
  }
  return true;
}
