// The device code of esn_input.cu, kept free of CUDA headers so that tests/test_pipeline_cpu.py can compile THIS SOURCE with g++
// behind a thread-per-CUDA-thread shim (tests/cuda_cpu_shim.h: pthread barrier for __syncthreads, static storage for
// __shared__) and check its indexing, tails and alignment branches against the oracle without a GPU.
#pragma once
#include <stdint.h>

namespace {

constexpr int kInThreads = 256;
constexpr int kInTilePx = 4 * kInThreads;          // 1024 pixels, 3072 bytes per tile
constexpr int kInTileWords = kInTilePx * 3 / 4;    // 768
constexpr int kInTileVec = kInTilePx * 3 / 16;     // 192

template <bool kReverse>
__global__ void __launch_bounds__(kInThreads) image_u8hwc_to_f32nchw_kernel(const uint8_t* __restrict__ img, float* __restrict__ out,
                                                                            const long long plane, const long long tiles_per_img,
                                                                            const long long total_tiles, const float m0,
                                                                            const float m1, const float m2) {
  __shared__ __align__(16) uint32_t tile[kInTileWords];
  const int tid = threadIdx.x;
  const float mean[3] = {m0, m1, m2};
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const long long n = t / tiles_per_img;
    const long long p0 = (t - n * tiles_per_img) * kInTilePx;
    const long long left = plane - p0;
    const int npx = left < kInTilePx ? (int)left : kInTilePx;
    const uint8_t* src = img + (n * plane + p0) * 3;
    if (npx == kInTilePx && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      if (tid < kInTileVec) reinterpret_cast<uint4*>(tile)[tid] = __ldg(reinterpret_cast<const uint4*>(src) + tid);
    } else {      // ragged last tile of an image, or an image whose byte offset is not 16-byte aligned
      uint8_t* tb = reinterpret_cast<uint8_t*>(tile);
      for (int i = tid; i < npx * 3; i += kInThreads) tb[i] = src[i];
    }
    __syncthreads();
    const int px = tid * 4;
    if (px < npx) {
      const uint32_t w[3] = {tile[3 * tid], tile[3 * tid + 1], tile[3 * tid + 2]};
      float v[3][4];      // [input channel][pixel]
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int k = 3 * j + c;      // byte index inside the 12-byte group
          v[c][j] = (float)((w[k >> 2] >> (8 * (k & 3))) & 0xffu) - mean[c];
        }
      float* o = out + n * 3 * plane + p0 + px;
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        const int ci = kReverse ? 2 - co : co;      // compile-time: v[][] stays in registers
        float* dst = o + co * plane;
        if (px + 3 < npx && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[ci][0], v[ci][1], v[ci][2], v[ci][3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (px + j < npx) dst[j] = v[ci][j];
        }
      }
    }
    __syncthreads();      // the tile is overwritten by the next iteration
  }
}

}  // namespace
