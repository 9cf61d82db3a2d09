// CGNet global-context gate FGlo (model/CGNet.py:173-191): y = x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2)
//   esn_global_avgpool : per-image, per-channel mean over H*W  -> fp32 [N][C]          (HBM-bound: one read of x)
//   esn_fglo_gate      : the two tiny Linear layers + ReLU + Sigmoid -> fp32 gate [N][C] (one CTA per image)
//   esn_scale_nc       : y = x * gate[n][c] (+ residual)  -- also the ContextGuidedBlock's "input + output"
#include "esn_common.cuh"

namespace {

constexpr int kGapThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kGapThreads) global_sum_kernel(const T* __restrict__ x, int HW, int C, int cs,
                                                                  float* __restrict__ sums, int px_per_cta) {
  __shared__ float red[kGapThreads][4];
  const int n = blockIdx.z;
  const int CG = min((C + 3) / 4, 64);
  const int lanes = kGapThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * 64 + cg) * 4;
  const bool vec = (cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(x) % (4 * sizeof(T))) == 0);
  float s[4] = {0, 0, 0, 0};
  const int p0 = blockIdx.x * px_per_cta, p1 = min(HW, p0 + px_per_cta);
  const T* xb = x + (size_t)n * HW * cs;
  if (pl < lanes && c < C) {
    for (int p = p0 + pl; p < p1; p += lanes) {
      if (vec) {
        const float4 t = ld4<T>(xb + (size_t)p * cs + c);
        s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < C) s[j] += ld1<T>(xb + (size_t)p * cs + c + j);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  if (pl == 0 && c < C) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] += red[l * CG + cg][j];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < C) sums[((size_t)blockIdx.x * gridDim.z + n) * C + c + j] = s[j];   // partial of this chunk (deterministic)
  }
}

// one CTA per image: hidden = relu(W1 (sum/HW) + b1), gate = sigmoid(W2 hidden + b2)
__global__ void fglo_gate_kernel(const float* __restrict__ sums, int chunks, int N, float inv_hw, const float* __restrict__ w1,
                                 const float* __restrict__ b1, const float* __restrict__ w2,
                                 const float* __restrict__ b2, float* __restrict__ gate, int C, int R) {
  extern __shared__ float sm[];   // mean[C] | hidden[R]
  float* mean = sm;
  float* hid = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < chunks; ++k) t += sums[((size_t)k * N + n) * C + c];   // fixed order: run-to-run identical
    mean[c] = t * inv_hw;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float acc = b1[r];
    for (int c = 0; c < C; ++c) acc = fmaf(w1[(size_t)r * C + c], mean[c], acc);
    hid[r] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = b2[c];
    for (int r = 0; r < R; ++r) acc = fmaf(w2[(size_t)c * R + r], hid[r], acc);
    gate[(size_t)n * C + c] = 1.f / (1.f + expf(-acc));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_nc_kernel(const T* __restrict__ x, const float* __restrict__ gate,
                                                        const T* __restrict__ res, T* __restrict__ y, int N, int HW, int C,
                                                        int x_cs, int res_cs, int y_cs) {
  const int ncg = C / 4;
  const long long total = (long long)N * HW * ncg;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % ncg) * 4;
  const long long p = i / ncg;
  const int n = (int)(p / HW);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)n * C + c));
  float4 v = ld4<T>(x + (size_t)p * x_cs + c);
  v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
  if (res) {
    const float4 r = ld4<T>(res + (size_t)p * res_cs + c);
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  }
  st4<T>(y + (size_t)p * y_cs + c, v);
}

}  // namespace

static void gap_plan(const EsnTensor* x, int* per, int* chunks) {
  const int HW = x->h * x->w;
  const int cblocks = esn_cdiv(esn_cdiv(x->c, 4), 64);
  int want = esn_cdiv(4 * 148, x->n * cblocks);
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  *per = esn_cdiv(HW, want);
  if (*per < 256) *per = 256;
  *chunks = esn_cdiv(HW, *per);
}

extern "C" int esn_global_avgpool_chunks(const EsnTensor* x) {
  if (!x || x->h < 1 || x->w < 1 || x->c < 1 || x->n < 1) return ESN_ERR_BAD_ARG;
  int per, chunks;
  gap_plan(x, &per, &chunks);
  return chunks;
}

extern "C" int esn_global_avgpool(const EsnTensor* x, float* sums, void* stream) {
  if (!x || !sums || !esn_valid_nhwc(*x)) return ESN_ERR_BAD_ARG;
  const int HW = x->h * x->w;
  const int cblocks = esn_cdiv(esn_cdiv(x->c, 4), 64);
  int per, chunks;
  gap_plan(x, &per, &chunks);
  dim3 grid(chunks, cblocks, x->n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x->dtype == ESN_F32)
    global_sum_kernel<float><<<grid, kGapThreads, 0, st>>>((const float*)x->ptr, HW, x->c, x->c_stride, sums, per);
  else
    global_sum_kernel<__nv_bfloat16><<<grid, kGapThreads, 0, st>>>((const __nv_bfloat16*)x->ptr, HW, x->c, x->c_stride, sums, per);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_fglo_gate(const EsnFGlo* p, void* stream) {
  if (!p || !p->sums || !p->w1 || !p->b1 || !p->w2 || !p->b2 || !p->gate || p->n < 1 || p->channels < 1 || p->hidden < 1 ||
      p->hw < 1 || p->chunks < 1)
    return ESN_ERR_BAD_ARG;
  if (p->channels > 4096 || p->hidden > 1024) return ESN_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  fglo_gate_kernel<<<p->n, 128, (p->channels + p->hidden) * sizeof(float), st>>>(p->sums, p->chunks, p->n, 1.f / (float)p->hw, p->w1, p->b1, p->w2,
                                                                                 p->b2, p->gate, p->channels, p->hidden);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_scale_nc(const EsnTensor* x, const float* gate, const EsnTensor* residual, const EsnTensor* y, void* stream) {
  if (!x || !y || !gate || !esn_valid_nhwc(*x) || !esn_valid_nhwc(*y) || x->dtype != y->dtype) return ESN_ERR_BAD_ARG;
  if (x->n != y->n || x->h != y->h || x->w != y->w || x->c != y->c) return ESN_ERR_BAD_SHAPE;
  const bool has_res = residual && residual->ptr;
  if (has_res && (!esn_valid_nhwc(*residual) || residual->dtype != x->dtype || residual->c != x->c || residual->h != x->h ||
                  residual->w != x->w))
    return ESN_ERR_BAD_ARG;
  const size_t sz = x->dtype == ESN_F32 ? 4 : 2;
  if (x->c % 4 || x->c_stride % 4 || y->c_stride % 4 || ((uintptr_t)x->ptr % (4 * sz)) || ((uintptr_t)y->ptr % (4 * sz)) ||
      ((uintptr_t)gate % 16) || (has_res && (residual->c_stride % 4 || ((uintptr_t)residual->ptr % (4 * sz)))))
    return ESN_ERR_ALIGN;
  const int HW = x->h * x->w;
  const long long total = (long long)x->n * HW * (x->c / 4);
  const int grid = esn_cdiv(total, 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x->dtype == ESN_F32)
    scale_nc_kernel<float><<<grid, 256, 0, st>>>((const float*)x->ptr, gate, has_res ? (const float*)residual->ptr : nullptr,
                                                 (float*)y->ptr, x->n, HW, x->c, x->c_stride, has_res ? residual->c_stride : 0,
                                                 y->c_stride);
  else
    scale_nc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x->ptr, gate,
                                                         has_res ? (const __nv_bfloat16*)residual->ptr : nullptr,
                                                         (__nv_bfloat16*)y->ptr, x->n, HW, x->c, x->c_stride,
                                                         has_res ? residual->c_stride : 0, y->c_stride);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
