// Device source of esn_soft_ce (csrc/esn_softce.cu), free of CUDA headers so the CPU test suite can compile it with g++
// behind tests/cuda_cpu_shim.h.
#pragma once
#include <math.h>
#include <stdint.h>

namespace {

constexpr int kSoftCeThreads = 256;

// One thread per pixel of NCHW fp32 logits (consecutive threads = consecutive pixels: every class plane is read coalesced).
// Per pixel, with soft targets t_c = (1 - eps) * [c == label] + eps / C and a_c = w_c * t_c:
//   loss = -sum_c a_c * log softmax(x)_c            d loss / d x_k = (sum_c a_c) * softmax(x)_k - a_k
// *sum += sum over pixels of loss (block tree reduction, one atomic per CTA); dx (optional) = gradient * scale * (*gout).
__global__ void __launch_bounds__(kSoftCeThreads) soft_ce_kernel(const float* __restrict__ x, const long long* __restrict__ y,
                                                                  const float* __restrict__ w, float* __restrict__ sum,
                                                                  float* __restrict__ dx, const float* __restrict__ gout,
                                                                  const int C, const long long hw, const long long npix,
                                                                  const float eps, const int ignore, const float scale) {
  __shared__ float red[kSoftCeThreads];
  float acc = 0.f;
  const float uni = eps / (float)C;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += step) {
    const long long n = i / hw, s = i - n * hw;
    const float* xp = x + n * C * hw + s;
    const long long lab = y[i];
    const bool valid = lab != ignore && lab >= 0 && lab < C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, xp[c * hw]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(xp[c * hw] - m);
    const float lse = m + logf(se);
    float sa = 0.f, l = 0.f;
    if (valid) {
      for (int c = 0; c < C; ++c) {
        const float a = (w ? w[c] : 1.f) * ((c == lab ? 1.f - eps : 0.f) + uni);
        sa += a;
        l -= a * (xp[c * hw] - lse);
      }
      acc += l;
    }
    if (dx) {
      const float g = scale * (gout ? *gout : 1.f);
      float* dp = dx + n * C * hw + s;
      for (int c = 0; c < C; ++c) {
        float v = 0.f;
        if (valid) {
          const float a = (w ? w[c] : 1.f) * ((c == lab ? 1.f - eps : 0.f) + uni);
          v = (sa * expf(xp[c * hw] - lse) - a) * g;
        }
        dp[c * hw] = v;
      }
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kSoftCeThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(sum, red[0]);
}

}  // namespace
