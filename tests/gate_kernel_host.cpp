// TEST INFRASTRUCTURE ONLY: runs the device source of esn_gate_bcast (csrc/esn_gate_kernel.cuh, float instantiation) on the
// CPU through tests/cuda_cpu_shim.h.
//   usage: gate_kernel_host n h w C g_cs x_cs b_cs y_cs has_b grid < floats: g | x | b > floats: y (n*h*w*y_cs, untouched = -12345)
#include <stdio.h>
#include <stdlib.h>

#include "cuda_cpu_shim.h"
#include "esn_gate_kernel.cuh"

int main(int argc, char** argv) {
  if (argc != 11) return 2;
  const int n = atoi(argv[1]), h = atoi(argv[2]), w = atoi(argv[3]), C = atoi(argv[4]);
  const int g_cs = atoi(argv[5]), x_cs = atoi(argv[6]), b_cs = atoi(argv[7]), y_cs = atoi(argv[8]), has_b = atoi(argv[9]);
  const unsigned grid = (unsigned)atoi(argv[10]);
  const long long npix = (long long)n * h * w;
  std::vector<float> g(npix * g_cs), x(npix * x_cs), b((size_t)n * b_cs + 1), y(npix * y_cs, -12345.0f);
  if (fread(g.data(), 4, g.size(), stdin) != g.size()) return 3;
  if (fread(x.data(), 4, x.size(), stdin) != x.size()) return 3;
  if (has_b && fread(b.data(), 4, (size_t)n * b_cs, stdin) != (size_t)n * b_cs) return 3;
  const float* bp = has_b ? b.data() : nullptr;
  shim_launch(grid, 256, [&] { gate_bcast_kernel<float>(g.data(), g_cs, x.data(), x_cs, bp, b_cs, y.data(), y_cs, npix, h * w, C); });
  fwrite(y.data(), 4, y.size(), stdout);
  return 0;
}
