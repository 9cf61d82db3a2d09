// TEST INFRASTRUCTURE ONLY: runs the device source of esn_augment_u8 (csrc/esn_augment_kernel.cuh) on the CPU, every CUDA
// thread as a plain loop iteration (no shared memory, no barriers).
//   usage: augment_kernel_host h w crop_h crop_w rh rw scale h_off w_off flip do_scale ignore m0 m1 m2
//          < bytes: img (h*w*3) | label (h*w)      > float32 (3*crop_h*crop_w) | int64 (crop_h*crop_w)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cuda_cpu_shim.h"
namespace {
inline double aug_dmul(double a, double b) { volatile double r = a * b; return r; }     // volatile: no contraction into an FMA
inline double aug_dsub(double a, double b) { volatile double r = a - b; return r; }
inline int aug_rint(float v) { return (int)lrintf(v); }                                 // default rounding mode: half to even
inline int aug_floorf(float v) { return (int)floorf(v); }
inline double aug_floord(double v) { return floor(v); }
}  // namespace
#include "esn_augment_kernel.cuh"

int main(int argc, char** argv) {
  if (argc != 16) return 2;
  AugArgs a;
  AugItem& it = a.it[0];
  it.h = atoi(argv[1]); it.w = atoi(argv[2]);
  a.crop_h = atoi(argv[3]); a.crop_w = atoi(argv[4]);
  it.rh = atoi(argv[5]); it.rw = atoi(argv[6]);
  it.scale = strtod(argv[7], nullptr);
  it.h_off = atoi(argv[8]); it.w_off = atoi(argv[9]); it.flip = atoi(argv[10]); it.do_scale = atoi(argv[11]);
  a.ignore_label = atoi(argv[12]);
  for (int c = 0; c < 3; ++c) a.mean[c] = strtof(argv[13 + c], nullptr);
  a.n = 1;
  std::vector<uint8_t> img((size_t)it.h * it.w * 3), lab((size_t)it.h * it.w);
  if (fread(img.data(), 1, img.size(), stdin) != img.size() || fread(lab.data(), 1, lab.size(), stdin) != lab.size()) return 3;
  it.img = img.data(); it.label = lab.data();
  std::vector<float> out((size_t)3 * a.crop_h * a.crop_w, -12345.f);
  std::vector<long long> ol((size_t)a.crop_h * a.crop_w, -1);
  a.out_img = out.data(); a.out_label = ol.data();
  for (int y = 0; y < a.crop_h; ++y)
    for (int x = 0; x < a.crop_w; ++x) aug_pixel(a, 0, y, x);
  fwrite(out.data(), 4, out.size(), stdout);
  fwrite(ol.data(), 8, ol.size(), stdout);
  return 0;
}
