// Device source of esn_gate_bcast (csrc/esn_gate.cu), free of CUDA headers so the CPU test suite can compile it with g++
// behind tests/cuda_cpu_shim.h.  ld1<T> / st1<T> come from esn_common.cuh on the device and from the shim on the CPU.
#pragma once
#include <stdint.h>

namespace {

// y[n,h,w,c] = g[n,h,w] * x[n,h,w,c] + b[n,c]      (all NHWC with their own pixel strides; b may be null)
// TG: the gate's own element type -- LEDNet keeps its single-channel pyramid in fp32 while the class scores are bf16
template <typename T, typename TG = T>
__global__ void __launch_bounds__(256) gate_bcast_kernel(const TG* __restrict__ g, const int g_cs, const T* __restrict__ x,
                                                         const int x_cs, const T* __restrict__ b, const int b_cs,
                                                         T* __restrict__ y, const int y_cs, const long long npix,
                                                         const int hw, const int C) {
  const long long total = npix * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    const long long p = i / C;
    float v = ld1<TG>(g + p * g_cs) * ld1<T>(x + p * x_cs + c);
    if (b) v += ld1<T>(b + (p / hw) * b_cs + c);
    st1<T>(y + p * y_cs + c, v);
  }
}

}  // namespace
