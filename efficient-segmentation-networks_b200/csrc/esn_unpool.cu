// ENet's index-carrying pooling pair (model/ENet.py:126-130, 225, 262):
//   esn_maxpool3x3s2_idx : MaxPool2d(3, stride 2, padding 1, return_indices=True) on NHWC; the index is the
//                          flat position h*W + w inside the un-padded input plane, FIRST maximum in window
//                          raster order wins (torch max_pool2d_with_indices); int32 suffices.
//   esn_max_unpool2x2    : MaxUnpool2d(2) written as a GATHER so it is deterministic: every output pixel
//                          inspects the <= 4 pooled cells whose 3x3 window covers it and takes the LAST one in
//                          raster order whose index points at it (the CPU reference scatters sequentially in
//                          raster order, so the last writer wins; torch's CUDA scatter is a data race --
//                          SURVEY.md H5).  Fused with the block's "+ ext, activation" (ENet.py:268-272).
#include "esn_common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_idx_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                               int32_t* __restrict__ idx, int N, int Hi, int Wi, int C,
                                                               int x_cs, int Ho, int Wo, int y_cs) {
  const long long total = (long long)N * Ho * Wo * C;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
  float best = -INFINITY;
  int bi = -1;
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * ho - 1 + r;
    if (h < 0 || h >= Hi) continue;
    for (int s = 0; s < 3; ++s) {
      const int w = 2 * wo - 1 + s;
      if (w < 0 || w >= Wi) continue;
      const float v = ld1<T>(x + ((size_t)((size_t)n * Hi + h) * Wi + w) * x_cs + c);
      if (v > best || bi < 0) {   // strict > keeps the first maximum; bi<0 admits the first valid element
        best = v;
        bi = h * Wi + w;
      }
    }
  }
  st1<T>(y + (size_t)p * y_cs + c, best);
  idx[(size_t)p * C + c] = bi;
}

// 16-byte variant (bf16, C % 8 == 0): one thread = one pooled pixel x 8 channels.
__global__ void __launch_bounds__(256) maxpool3x3s2_idx_v8_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                                  int32_t* __restrict__ idx, int N, int Hi, int Wi, int C, int x_cs,
                                                                  int Ho, int Wo, int y_cs) {
  const int ncg = C >> 3;
  const long long total = (long long)N * Ho * Wo * ncg;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % ncg) << 3;
  const long long p = i / ncg;
  const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
  float best[8];
  int bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = -1; }
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * ho - 1 + r;
    if (h < 0 || h >= Hi) continue;
    for (int s = 0; s < 3; ++s) {
      const int w = 2 * wo - 1 + s;
      if (w < 0 || w >= Wi) continue;
      float v[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + ((size_t)((size_t)n * Hi + h) * Wi + w) * x_cs + c)), v);
      const int pos = h * Wi + w;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > best[j] || bi[j] < 0) { best[j] = v[j]; bi[j] = pos; }
    }
  }
  *reinterpret_cast<uint4*>(y + (size_t)p * y_cs + c) = float_to_bf16x8(best);
  int4* ip = reinterpret_cast<int4*>(idx + (size_t)p * C + c);
  ip[0] = make_int4(bi[0], bi[1], bi[2], bi[3]);
  ip[1] = make_int4(bi[4], bi[5], bi[6], bi[7]);
}

struct UnpoolArgs {
  const void* v;      // pooled-resolution values [N,Hp,Wp,C]
  const int32_t* idx; // [N,Hp,Wp,C]
  const void* ext;    // optional [N,Ho,Wo,C], added before the activation
  void* y;            // [N,Ho,Wo,C]
  int N, Hp, Wp, C, v_cs, Ho, Wo, ext_cs, y_cs, act;
  const float* alpha;
};

template <typename T>
__global__ void __launch_bounds__(256) max_unpool_kernel(const UnpoolArgs a) {
  const long long total = (long long)a.N * a.Ho * a.Wo * a.C;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % a.C);
  const long long p = i / a.C;
  const int x = (int)(p % a.Wo), yv = (int)((p / a.Wo) % a.Ho), n = (int)(p / ((long long)a.Wo * a.Ho));
  const int target = yv * a.Wo + x;
  // pooled cells (ci, cj) whose window rows 2ci-1..2ci+1 / cols 2cj-1..2cj+1 contain (yv, x), raster order
  const int i0 = yv >> 1, i1 = (yv + 1) >> 1, j0 = x >> 1, j1 = (x + 1) >> 1;
  float val = 0.f;
  const T* v = reinterpret_cast<const T*>(a.v);
  for (int ci = i0; ci <= i1; ++ci) {
    if (ci >= a.Hp) continue;
    for (int cj = j0; cj <= j1; ++cj) {
      if (cj >= a.Wp) continue;
      const size_t q = ((size_t)((size_t)n * a.Hp + ci) * a.Wp + cj);
      if (a.idx[q * a.C + c] == target) val = ld1<T>(v + q * a.v_cs + c);   // later cells overwrite earlier ones
    }
  }
  if (a.ext) val += ld1<T>(reinterpret_cast<const T*>(a.ext) + (size_t)p * a.ext_cs + c);
  val = apply_act(val, a.act, a.act == ESN_ACT_PRELU ? __ldg(a.alpha + c) : 0.f);
  st1<T>(reinterpret_cast<T*>(a.y) + (size_t)p * a.y_cs + c, val);
}

// 16-byte variant (bf16, C % 8 == 0): one thread = one output pixel x 8 channels.
__global__ void __launch_bounds__(256) max_unpool_v8_kernel(const UnpoolArgs a) {
  const int ncg = a.C >> 3;
  const long long total = (long long)a.N * a.Ho * a.Wo * ncg;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % ncg) << 3;
  const long long p = i / ncg;
  const int x = (int)(p % a.Wo), yv = (int)((p / a.Wo) % a.Ho), n = (int)(p / ((long long)a.Wo * a.Ho));
  const int target = yv * a.Wo + x;
  const int i0 = yv >> 1, i1 = (yv + 1) >> 1, j0 = x >> 1, j1 = (x + 1) >> 1;
  float val[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) val[j] = 0.f;
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
  for (int ci = i0; ci <= i1; ++ci) {
    if (ci >= a.Hp) continue;
    for (int cj = j0; cj <= j1; ++cj) {
      if (cj >= a.Wp) continue;
      const size_t q = ((size_t)((size_t)n * a.Hp + ci) * a.Wp + cj);
      const int4 ia = __ldg(reinterpret_cast<const int4*>(a.idx + q * a.C + c));
      const int4 ib = __ldg(reinterpret_cast<const int4*>(a.idx + q * a.C + c + 4));
      const int id[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
      bool any = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) any |= id[j] == target;
      if (any) {
        float f[8];
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(v + q * a.v_cs + c)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (id[j] == target) val[j] = f[j];        // later cells overwrite earlier ones
      }
    }
  }
  if (a.ext) {
    float e[8];
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.ext) + (size_t)p * a.ext_cs + c)), e);
#pragma unroll
    for (int j = 0; j < 8; ++j) val[j] += e[j];
  }
  if (a.act == ESN_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) val[j] = fmaxf(val[j], 0.f);
  } else if (a.act == ESN_ACT_PRELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) val[j] = val[j] >= 0.f ? val[j] : val[j] * __ldg(a.alpha + c + j);
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)p * a.y_cs + c) = float_to_bf16x8(val);
}

}  // namespace

extern "C" int esn_maxpool3x3s2_idx(const EsnTensor* x, const EsnTensor* y, int32_t* idx, void* stream) {
  if (!x || !y || !idx || !esn_valid_nhwc(*x) || !esn_valid_nhwc(*y) || x->dtype != y->dtype) return ESN_ERR_BAD_ARG;
  if (y->n != x->n || y->c != x->c || y->h != (x->h - 1) / 2 + 1 || y->w != (x->w - 1) / 2 + 1) return ESN_ERR_BAD_SHAPE;
  if ((long long)x->h * x->w > 2147483647LL) return ESN_ERR_UNSUPPORTED;
  const long long total = (long long)y->n * y->h * y->w * y->c;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = esn_cdiv(total, 256);
  if (x->dtype == ESN_BF16 && x->c % 8 == 0 && x->c_stride % 8 == 0 && y->c_stride % 8 == 0 && (uintptr_t)x->ptr % 16 == 0 &&
      (uintptr_t)y->ptr % 16 == 0 && (uintptr_t)idx % 16 == 0)
    maxpool3x3s2_idx_v8_kernel<<<esn_cdiv(total / 8, 256), 256, 0, st>>>((const __nv_bfloat16*)x->ptr, (__nv_bfloat16*)y->ptr, idx, x->n,
                                                                         x->h, x->w, x->c, x->c_stride, y->h, y->w, y->c_stride);
  else if (x->dtype == ESN_F32)
    maxpool3x3s2_idx_kernel<float><<<grid, 256, 0, st>>>((const float*)x->ptr, (float*)y->ptr, idx, x->n, x->h, x->w, x->c,
                                                         x->c_stride, y->h, y->w, y->c_stride);
  else
    maxpool3x3s2_idx_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x->ptr, (__nv_bfloat16*)y->ptr, idx,
                                                                 x->n, x->h, x->w, x->c, x->c_stride, y->h, y->w, y->c_stride);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_max_unpool2x2(const EsnUnpool* p, void* stream) {
  if (!p || !p->idx || !esn_valid_nhwc(p->v) || !esn_valid_nhwc(p->y) || p->v.dtype != p->y.dtype) return ESN_ERR_BAD_ARG;
  if (p->y.n != p->v.n || p->y.c != p->v.c || p->y.h != 2 * p->v.h || p->y.w != 2 * p->v.w) return ESN_ERR_BAD_SHAPE;
  if (p->ext.ptr && (!esn_valid_nhwc(p->ext) || p->ext.dtype != p->y.dtype || p->ext.h != p->y.h || p->ext.w != p->y.w ||
                     p->ext.c != p->y.c))
    return ESN_ERR_BAD_ARG;
  if (p->act == ESN_ACT_PRELU && !p->alpha) return ESN_ERR_BAD_ARG;
  UnpoolArgs a;
  a.v = p->v.ptr; a.idx = p->idx; a.ext = p->ext.ptr; a.y = p->y.ptr;
  a.N = p->v.n; a.Hp = p->v.h; a.Wp = p->v.w; a.C = p->v.c; a.v_cs = p->v.c_stride;
  a.Ho = p->y.h; a.Wo = p->y.w; a.ext_cs = p->ext.c_stride; a.y_cs = p->y.c_stride;
  a.act = p->act; a.alpha = p->alpha;
  const long long total = (long long)a.N * a.Ho * a.Wo * a.C;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = esn_cdiv(total, 256);
  const bool v8 = p->v.dtype == ESN_BF16 && a.C % 8 == 0 && a.v_cs % 8 == 0 && a.y_cs % 8 == 0 && (!a.ext || a.ext_cs % 8 == 0) &&
                  (uintptr_t)a.v % 16 == 0 && (uintptr_t)a.y % 16 == 0 && (uintptr_t)a.ext % 16 == 0 && (uintptr_t)a.idx % 16 == 0;
  if (v8) max_unpool_v8_kernel<<<esn_cdiv(total / 8, 256), 256, 0, st>>>(a);
  else if (p->v.dtype == ESN_F32) max_unpool_kernel<float><<<grid, 256, 0, st>>>(a);
  else max_unpool_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
