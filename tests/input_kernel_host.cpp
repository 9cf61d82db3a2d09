// TEST INFRASTRUCTURE ONLY: runs the device source of esn_image_u8hwc_to_f32nchw (csrc/esn_input_kernel.cuh) on the CPU
// through tests/cuda_cpu_shim.h, with the launch arithmetic of the C-ABI entry (csrc/esn_input.cu) restated below.
//   usage: input_kernel_host n h w reverse in_offset out_offset grid_cap < raw uint8 image bytes > raw float32 output
// in_offset / out_offset shift the buffers off their 16-byte alignment (bytes / floats) to reach the unaligned branches;
// grid_cap bounds the grid so that a CTA walks several tiles.
#include <stdio.h>
#include <stdlib.h>

#include "cuda_cpu_shim.h"
#include "esn_input_kernel.cuh"

int main(int argc, char** argv) {
  if (argc != 8) return 2;
  const int n = atoi(argv[1]), h = atoi(argv[2]), w = atoi(argv[3]), reverse = atoi(argv[4]);
  const int in_off = atoi(argv[5]), out_off = atoi(argv[6]), grid_cap = atoi(argv[7]);
  const long long plane = (long long)h * w;
  const size_t in_bytes = (size_t)n * plane * 3, out_elems = (size_t)n * plane * 3;
  uint8_t* in_base = static_cast<uint8_t*>(aligned_alloc(64, in_bytes + 128));
  float* out_base = static_cast<float*>(aligned_alloc(64, (out_elems + 32) * sizeof(float)));
  uint8_t* img = in_base + in_off;
  float* out = out_base + out_off;
  if (fread(img, 1, in_bytes, stdin) != in_bytes) return 3;
  for (size_t i = 0; i < out_elems; ++i) out[i] = -12345.0f;      // every element must be overwritten
  const float mean[3] = {72.3924f, 82.90902f, 73.158325f};
  const long long tiles_per_img = (plane + kInTilePx - 1) / kInTilePx;
  const long long total = tiles_per_img * n;
  long long grid = total < grid_cap ? total : grid_cap;
  if (total > 0) {
    if (reverse)
      shim_launch((unsigned)grid, kInThreads, [&] { image_u8hwc_to_f32nchw_kernel<true>(img, out, plane, tiles_per_img, total, mean[0], mean[1], mean[2]); });
    else
      shim_launch((unsigned)grid, kInThreads, [&] { image_u8hwc_to_f32nchw_kernel<false>(img, out, plane, tiles_per_img, total, mean[0], mean[1], mean[2]); });
  }
  fwrite(out, sizeof(float), out_elems, stdout);
  return 0;
}
