// Backward of ENet's pooling pair (training path; reference: autograd of nn.MaxPool2d(3, 2, 1, return_indices=True) and
// nn.MaxUnpool2d(2), model/ENet.py:126-130, 225, 262 reached from loss.backward() at train.py:353).
//   esn_maxpool3x3s2_idx_bwd : dx[n,h,w,c] (+)= sum of dy over the (at most four) windows whose recorded arg-max is (h,w) --
//                              written as a GATHER over the input pixels, so it is deterministic and needs no atomics
//                              although the 3x3 / stride-2 windows overlap
//   esn_max_unpool2x2_bwd    : dv[n,i,j,c] = dy[n, idx / W2, idx % W2, c] -- every pooled cell reads the gradient at the
//                              position it was scattered to (cells that lost a collision in the forward still receive it:
//                              that is what torch's max_unpool2d backward does)
// HBM-bound elementwise passes, 8 channels (16 bytes of bf16) per thread where the views allow.
#include "esn_common.cuh"

namespace {

template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool3x3s2_idx_bwd_kernel(const T* __restrict__ dy, const int32_t* __restrict__ idx,
                                                                   T* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo,
                                                                   int dy_cs, int dx_cs, int accumulate) {
  const int cg = C / V;
  const long long total = (long long)N * H * W * cg;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % cg) * V;
  const long long pix = t / cg;
  const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
  const int target = h * W + w;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = accumulate ? ld1<T>(dx + pix * dx_cs + c + k) : 0.f;
  // windows (i, j) cover input rows 2i-1 .. 2i+1: i in {h/2, (h+1)/2} (equal for even h), same for columns
  const int i0 = h >> 1, i1 = (h + 1) >> 1, j0 = w >> 1, j1 = (w + 1) >> 1;
  for (int ii = 0; ii < 2; ++ii) {
    const int i = ii ? i1 : i0;
    if ((ii && i1 == i0) || i >= Ho) continue;
    for (int jj = 0; jj < 2; ++jj) {
      const int j = jj ? j1 : j0;
      if ((jj && j1 == j0) || j >= Wo) continue;
      const long long q = ((long long)n * Ho + i) * Wo + j;
#pragma unroll
      for (int k = 0; k < V; ++k)
        if (__ldg(idx + q * C + c + k) == target) acc[k] += ld1<T>(dy + q * dy_cs + c + k);
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) st1<T>(dx + pix * dx_cs + c + k, acc[k]);
}

template <typename T, int V>
__global__ void __launch_bounds__(256) max_unpool2x2_bwd_kernel(const T* __restrict__ dy, const int32_t* __restrict__ idx,
                                                                T* __restrict__ dv, long long npix_lo, int C, int Hp, int Wp,
                                                                int dy_cs, int dv_cs) {
  const int cg = C / V;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= npix_lo * cg) return;
  const int c = (int)(t % cg) * V;
  const long long q = t / cg;
  const long long n = q / ((long long)Hp * Wp);
  const long long plane = n * (4LL * Hp * Wp);          // output plane of this image: (2Hp) x (2Wp) pixels
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int p = __ldg(idx + q * C + c + k);
    st1<T>(dv + q * dv_cs + c + k, ld1<T>(dy + (plane + p) * dy_cs + c + k));
  }
}

}  // namespace

extern "C" int esn_maxpool3x3s2_idx_bwd(const EsnTensor* dy, const int32_t* idx, const EsnTensor* dx, int32_t accumulate,
                                        void* stream) {
  if (!dy || !dx || !idx || !esn_valid_nhwc(*dy) || !esn_valid_nhwc(*dx) || dy->dtype != dx->dtype) return ESN_ERR_BAD_ARG;
  if (dy->n != dx->n || dy->c != dx->c || dy->h != (dx->h - 1) / 2 + 1 || dy->w != (dx->w - 1) / 2 + 1) return ESN_ERR_BAD_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C = dx->c;
  const int V = (C % 8 == 0) ? 8 : (C % 4 == 0 ? 4 : 1);
  const long long total = (long long)dx->n * dx->h * dx->w * (C / V);
  const int grid = esn_cdiv(total, 256);
#define LAUNCH(T, VV)                                                                                                      \
  maxpool3x3s2_idx_bwd_kernel<T, VV><<<grid, 256, 0, st>>>((const T*)dy->ptr, idx, (T*)dx->ptr, dx->n, dx->h, dx->w, C, dy->h, \
                                                           dy->w, dy->c_stride, dx->c_stride, accumulate)
  if (dx->dtype == ESN_BF16) {
    if (V == 8) LAUNCH(__nv_bfloat16, 8); else if (V == 4) LAUNCH(__nv_bfloat16, 4); else LAUNCH(__nv_bfloat16, 1);
  } else {
    if (V == 8) LAUNCH(float, 8); else if (V == 4) LAUNCH(float, 4); else LAUNCH(float, 1);
  }
#undef LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_max_unpool2x2_bwd(const EsnTensor* dy, const int32_t* idx, const EsnTensor* dv, void* stream) {
  if (!dy || !dv || !idx || !esn_valid_nhwc(*dy) || !esn_valid_nhwc(*dv) || dy->dtype != dv->dtype) return ESN_ERR_BAD_ARG;
  if (dy->n != dv->n || dy->c != dv->c || dy->h != 2 * dv->h || dy->w != 2 * dv->w) return ESN_ERR_BAD_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C = dv->c;
  const int V = (C % 8 == 0) ? 8 : (C % 4 == 0 ? 4 : 1);
  const long long npix = (long long)dv->n * dv->h * dv->w;
  const int grid = esn_cdiv(npix * (C / V), 256);
#define LAUNCH(T, VV)                                                                                                       \
  max_unpool2x2_bwd_kernel<T, VV><<<grid, 256, 0, st>>>((const T*)dy->ptr, idx, (T*)dv->ptr, npix, C, dv->h, dv->w, dy->c_stride, \
                                                        dv->c_stride)
  if (dv->dtype == ESN_BF16) {
    if (V == 8) LAUNCH(__nv_bfloat16, 8); else if (V == 4) LAUNCH(__nv_bfloat16, 4); else LAUNCH(__nv_bfloat16, 1);
  } else {
    if (V == 8) LAUNCH(float, 8); else if (V == 4) LAUNCH(float, 4); else LAUNCH(float, 1);
  }
#undef LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of CGNet's global-context gate FGlo (CGNet.py:173-191, y = x * g[n][c] with g = sigmoid(MLP(mean_hw x))):
//   esn_dot_nc        : out[n][c] += sum_hw a[n,h,w,c] * b[n,h,w,c]   (dg = sum dy * x; fp32 atomics, out zeroed by the caller)
//   esn_scale_add_nc  : y = a * s[n][c] + t[n][c] (+ extra)            (dx = dy * g + dpooled / HW, plus an earlier gradient)
// Both HBM-bound passes over the activation; the two-layer MLP on the (N, C) vectors is host-side torch (a few kFLOP).
namespace {

template <typename T, int V>
__global__ void __launch_bounds__(256) dot_nc_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out,
                                                     int HW, int C, int a_cs, int b_cs, int rows_per_cta) {
  __shared__ float red[256 * V];
  const int cg = C / V;                       // channel groups; 256 % cg == 0 is checked on the host
  const int lane_c = threadIdx.x % cg, lane_p = threadIdx.x / cg, np = 256 / cg;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * rows_per_cta;
  const int p1 = min(HW, p0 + rows_per_cta);
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (int p = p0 + lane_p; p < p1; p += np) {
    const size_t pix = (size_t)n * HW + p;
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += ld1<T>(a + pix * a_cs + lane_c * V + k) * ld1<T>(b + pix * b_cs + lane_c * V + k);
  }
#pragma unroll
  for (int k = 0; k < V; ++k) red[threadIdx.x * V + k] = acc[k];
  __syncthreads();
  if (lane_p == 0) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float s = 0.f;
      for (int q = 0; q < np; ++q) s += red[(q * cg + lane_c) * V + k];
      atomicAdd(out + (size_t)n * C + lane_c * V + k, s);
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) scale_add_nc_kernel(const T* __restrict__ a, const float* __restrict__ s, const float* __restrict__ t,
                                                           const T* __restrict__ extra, T* __restrict__ y, long long npix, int HW,
                                                           int C, int a_cs, int e_cs, int y_cs) {
  const int cg = C / V;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= npix * cg) return;
  const int c = (int)(i % cg) * V;
  const long long pix = i / cg;
  const long long n = pix / HW;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    float v = ld1<T>(a + pix * a_cs + c + k) * __ldg(s + n * C + c + k) + (t ? __ldg(t + n * C + c + k) : 0.f);
    if (extra) v += ld1<T>(extra + pix * e_cs + c + k);
    st1<T>(y + pix * y_cs + c + k, v);
  }
}

}  // namespace

extern "C" int esn_dot_nc(const EsnTensor* a, const EsnTensor* b, float* out, void* stream) {
  if (!a || !b || !out || !esn_valid_nhwc(*a) || !esn_valid_nhwc(*b) || a->dtype != b->dtype) return ESN_ERR_BAD_ARG;
  if (a->n != b->n || a->h != b->h || a->w != b->w || a->c != b->c) return ESN_ERR_BAD_SHAPE;
  const int C = a->c;
  const int V = (C % 8 == 0 && 256 % (C / 8) == 0) ? 8 : ((C % 4 == 0 && 256 % (C / 4) == 0) ? 4 : 0);
  if (!V) return ESN_ERR_UNSUPPORTED;
  const int HW = a->h * a->w;
  const int rows = 2048;
  dim3 grid((unsigned)esn_cdiv(HW, rows), (unsigned)a->n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == ESN_BF16) {
    if (V == 8) dot_nc_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a->ptr, (const __nv_bfloat16*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
    else dot_nc_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a->ptr, (const __nv_bfloat16*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
  } else {
    if (V == 8) dot_nc_kernel<float, 8><<<grid, 256, 0, st>>>((const float*)a->ptr, (const float*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
    else dot_nc_kernel<float, 4><<<grid, 256, 0, st>>>((const float*)a->ptr, (const float*)b->ptr, out, HW, C, a->c_stride, b->c_stride, rows);
  }
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_scale_add_nc(const EsnTensor* a, const float* s, const float* t, const EsnTensor* extra, const EsnTensor* y,
                                void* stream) {
  if (!a || !s || !y || !esn_valid_nhwc(*a) || !esn_valid_nhwc(*y) || a->dtype != y->dtype) return ESN_ERR_BAD_ARG;
  if (a->n != y->n || a->h != y->h || a->w != y->w || a->c != y->c) return ESN_ERR_BAD_SHAPE;
  const bool has_e = extra && extra->ptr;
  if (has_e && (!esn_valid_nhwc(*extra) || extra->dtype != a->dtype || extra->n != a->n || extra->h != a->h || extra->w != a->w ||
                extra->c != a->c))
    return ESN_ERR_BAD_ARG;
  const int C = a->c;
  const int V = (C % 8 == 0) ? 8 : (C % 4 == 0 ? 4 : 1);
  const long long npix = (long long)a->n * a->h * a->w;
  const int grid = esn_cdiv(npix * (C / V), 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define LAUNCH(T, VV)                                                                                                          \
  scale_add_nc_kernel<T, VV><<<grid, 256, 0, st>>>((const T*)a->ptr, s, t, has_e ? (const T*)extra->ptr : nullptr, (T*)y->ptr, npix, \
                                                   a->h * a->w, C, a->c_stride, has_e ? extra->c_stride : 0, y->c_stride)
  if (a->dtype == ESN_BF16) {
    if (V == 8) LAUNCH(__nv_bfloat16, 8); else if (V == 4) LAUNCH(__nv_bfloat16, 4); else LAUNCH(__nv_bfloat16, 1);
  } else {
    if (V == 8) LAUNCH(float, 8); else if (V == 4) LAUNCH(float, 4); else LAUNCH(float, 1);
  }
#undef LAUNCH
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
