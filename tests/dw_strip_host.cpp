// TEST INFRASTRUCTURE ONLY: runs the device source of the register-strip depthwise conv (csrc/esn_dw_strip_kernel.cuh, float
// instantiation: 4 channels per vector) on the CPU, every CUDA thread as a plain loop iteration (the kernel has no shared
// memory and no barriers).
//   usage: dw_strip_host n h w C x_cs y_cs kh kw dil_h dil_w act has_res pre_act tw seg_max min_threads
//          < floats: x (n*h*w*x_cs) | w (kh*kw*C) | scale C | shift C | alpha C | res (n*h*w*y_cs, if has_res)
//          > floats: y (n*h*w*y_cs, untouched = -12345)
#include <stdio.h>
#include <stdlib.h>

#include "cuda_cpu_shim.h"
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float2 ffma2(float2 a, float2 b, float2 c) { return float2{__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y)}; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
namespace {
template <typename T> struct DwsRaw;
template <> struct DwsRaw<float> {
  typedef float4 type;
  static float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
inline float4 ldw4(const float* p) { return float4{p[0], p[1], p[2], p[3]}; }
inline float4 ldraw(const float* p) { return float4{p[0], p[1], p[2], p[3]}; }
inline void unpack(const float4& r, float2 (&f)[2]) { f[0] = make_float2(r.x, r.y); f[1] = make_float2(r.z, r.w); }
inline void stv(float* p, float2 (&v)[2]) { p[0] = v[0].x; p[1] = v[0].y; p[2] = v[1].x; p[3] = v[1].y; }
template <int ACT> inline float2 dws_act2(float2 v, float2 al) {
  if (ACT == 1) return make_float2(fmaxf_(v.x, 0.f), fmaxf_(v.y, 0.f));
  if (ACT == 2) return make_float2(v.x >= 0.f ? v.x : v.x * al.x, v.y >= 0.f ? v.y : v.y * al.y);
  return v;
}
}  // namespace
#include "esn_dw_strip_kernel.cuh"

template <int KH, int KW, int TW, int ACT, int RES>
static void run2(const DwsArgs& a) {
  const long long slack = 37;      // threads beyond `total` must do nothing
  for (long long t = 0; t < a.total + slack; ++t)
    if (t < a.total) dw_strip_thread<float, 4, KH, KW, TW, ACT, RES>(a, t);
}
template <int KH, int KW, int TW>
static void run(DwsArgs a, int seg_max, long long min_threads) {
  dws_plan(a, 4, KH, KW, TW, min_threads, seg_max);
  const bool res = a.res != nullptr;
  if (a.act == 1) res ? run2<KH, KW, TW, 1, 1>(a) : run2<KH, KW, TW, 1, 0>(a);
  else if (a.act == 2) res ? run2<KH, KW, TW, 2, 1>(a) : run2<KH, KW, TW, 2, 0>(a);
  else res ? run2<KH, KW, TW, 0, 1>(a) : run2<KH, KW, TW, 0, 0>(a);
}

int main(int argc, char** argv) {
  if (argc != 17) return 2;
  int v[16];
  for (int i = 0; i < 16; ++i) v[i] = atoi(argv[i + 1]);
  const int n = v[0], h = v[1], w = v[2], C = v[3], x_cs = v[4], y_cs = v[5], kh = v[6], kw = v[7];
  const int has_res = v[11], tw = v[13], seg_max = v[14];
  const long long min_threads = atoll(argv[16]);
  const size_t npix = (size_t)n * h * w;
  std::vector<float> x(npix * x_cs), wt((size_t)kh * kw * C), sc(C), sh(C), al(C), res(npix * y_cs + 1), y(npix * y_cs, -12345.0f);
  auto rd = [](std::vector<float>& b, size_t cnt) { return fread(b.data(), 4, cnt, stdin) == cnt; };
  if (!rd(x, x.size()) || !rd(wt, wt.size()) || !rd(sc, C) || !rd(sh, C) || !rd(al, C)) return 3;
  if (has_res && !rd(res, npix * y_cs)) return 3;
  DwsArgs a;
  a.x = x.data(); a.y = y.data(); a.w = wt.data(); a.scale = sc.data(); a.shift = sh.data(); a.alpha = al.data();
  a.res = has_res ? res.data() : nullptr; a.res_cs = y_cs; a.act = v[10]; a.pre_act = v[12];
  a.N = n; a.H = h; a.W = w; a.C = C; a.x_cs = x_cs; a.y_cs = y_cs; a.dil_h = v[8]; a.dil_w = v[9];
  if (kh == 3 && kw == 3 && tw == 2) run<3, 3, 2>(a, seg_max, min_threads);
  else if (kh == 3 && kw == 3 && tw == 4) run<3, 3, 4>(a, seg_max, min_threads);
  else if (kh == 3 && kw == 1 && tw == 2) run<3, 1, 2>(a, seg_max, min_threads);
  else if (kh == 3 && kw == 1 && tw == 4) run<3, 1, 4>(a, seg_max, min_threads);
  else if (kh == 1 && kw == 3 && tw == 4) run<1, 3, 4>(a, seg_max, min_threads);
  else return 4;
  fwrite(y.data(), 4, y.size(), stdout);
  return 0;
}
