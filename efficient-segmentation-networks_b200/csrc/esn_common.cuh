// Shared helpers for libesn_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>

#include <cstdio>
#include <cstdlib>

#include "esn.h"

extern std::atomic<long long> g_esn_launches;

#define ESN_CHECK_LAUNCH()                                   \
  do {                                                       \
    g_esn_launches.fetch_add(1, std::memory_order_relaxed);  \
    if (cudaPeekAtLastError() != cudaSuccess) {              \
      const cudaError_t e_ = cudaGetLastError();             \
      if (getenv("ESN_DEBUG"))                               \
        fprintf(stderr, "esn: %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return ESN_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

static inline int esn_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// One-time initialisation is per DEVICE (cudaFuncSetAttribute applies to the current device only, and under the
// reference's nn.DataParallel one process drives several GPUs from one thread each, train.py:166-168).
constexpr int kEsnMaxDevices = 64;
static inline int esn_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kEsnMaxDevices) dev = 0;
  return dev;
}

// ---- element access: T in {float, __nv_bfloat16} ---------------------------------
template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// 4 consecutive channels (pointer must be 16 B (f32) / 8 B (bf16) aligned)
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// 8 bf16 channels <-> 8 floats (16-byte access)
__device__ __forceinline__ void bf16x8_to_float(const uint4& r, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float* f) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act == ESN_ACT_RELU) return fmaxf(v, 0.f);
  if (act == ESN_ACT_PRELU) return v >= 0.f ? v : v * alpha;
  return v;
}

// Blackwell packed fp32 FMA (FFMA2): two lanes per instruction -- for the fp32-issue-bound kernels
// (DABNet depthwise pair, ERFNet 2x2 transposed-conv head) halving the FMA instruction count is the lever; each lane
// is an ordinary fma.rn.f32, so results are bit-identical to the scalar form.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
// Device-side copy of the epilogue description.
struct EpiArgs {
  const float* scale;
  const float* shift;
  const float* alpha;
  const void* res;
  int res_cstride;
  int res_dtype;
  int act;
  int pre_act;   // ESN_EP_ACT_BEFORE_RESIDUAL
  int res_first; // ESN_EP_RESIDUAL_FIRST
};

static inline EpiArgs make_epi(const EsnEpilogue& e) {
  EpiArgs a;
  a.scale = e.scale;
  a.shift = e.shift;
  a.alpha = e.alpha;
  a.res = e.residual.ptr;
  a.res_cstride = e.residual.c_stride;
  a.res_dtype = e.residual.dtype;
  a.act = e.act;
  a.pre_act = (e.flags & ESN_EP_ACT_BEFORE_RESIDUAL) ? 1 : 0;
  a.res_first = (e.flags & ESN_EP_RESIDUAL_FIRST) ? 1 : 0;
  return a;
}

static inline bool esn_valid_nhwc(const EsnTensor& t) {
  return t.ptr && t.layout == ESN_NHWC && (t.dtype == ESN_F32 || t.dtype == ESN_BF16) && t.n > 0 && t.h > 0 &&
         t.w > 0 && t.c > 0 && t.c_stride >= t.c;
}

int esn_check_epilogue(const EsnEpilogue& e, const EsnTensor& y, bool allow_residual_first = false);
