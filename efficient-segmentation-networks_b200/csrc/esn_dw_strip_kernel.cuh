// Device source of the register-strip depthwise convolution (csrc/esn_dw_strip.cu), free of CUDA headers so the CPU test
// suite can compile the float instantiation with g++ behind tests/cuda_cpu_shim.h.
//
// Depthwise KH x KW conv (KH, KW in {1, 3}), stride 1, "same" padding (pad = dilation * (K-1)/2), any dilation, NHWC,
// fused per-channel scale / shift / residual / ReLU|PReLU epilogue -- CGNet F_loc / F_sur (CGNet.py:106-171), Fast-SCNN
// _DWConv / _DSConv (FastSCNN.py:28-59), ESPNetv2 EESP branches (ESPNet_v2/Model.py:15-99), DABNet 3x1 / 1x3 (DABNet.py:51-83)
// in the training path, ContextNet.
//
// HBM-bound on paper: 1 read + 1 write of the tensor.  The round-1 kernel gathered 9 input vectors per output vector from
// L1/L2 and re-read the taps from shared memory (18 LDS.128 per output): 0.15 of HBM peak.  Here a thread owns one
// 16-byte channel vector, TW output columns spaced one dilation step apart, and walks DOWN an image column in steps of the
// vertical dilation: the taps live in registers (pre-multiplied by the epilogue scale), every input row passes through
// registers once and feeds the three output rows it belongs to (accumulators rotate), the TW + 2 column vectors of a row
// serve all TW outputs.  Loads per output vector: (TW + 2) / TW instead of 9; no shared memory, no barriers.  The fp32
// multiply-adds are issued as packed FFMA2 (two channels per instruction).  The vectors of the next input row are
// requested before the current row's arithmetic, so two rows of loads are in flight per thread.
//
// DwsRaw<T>, ldraw / unpack / stv (one 16-byte channel vector <-> float2[V/2]), ffma2, dws_act2 come from the including file.
#pragma once
#include <stdint.h>

namespace {

struct DwsArgs {
  const void* x;
  void* y;
  const float* w;       // [KH*KW][C] fp32 taps
  const float* scale;   // [C] or null (= 1)
  const float* shift;   // [C] or null (= 0)
  const float* alpha;   // [C] PReLU slopes (act == PRELU)
  const void* res;      // residual (same dtype as y) or null
  int res_cs;
  int act;              // 0 none, 1 relu, 2 prelu  (ESN_ACT_*)
  int pre_act;          // activation before AND after the residual add (ESN_EP_ACT_BEFORE_RESIDUAL)
  int N, H, W, C, x_cs, y_cs;
  int dil_h, dil_w;
  int seg;              // output rows per thread (along its dilated column chain)
  int QW;               // column groups per image row
  int VS;               // vertical (chain, segment) pairs per image
  long long total;      // threads with work
};

// Work decomposition (host side; also used by the CPU test harness): fills seg / QW / VS / total.  `min_threads` = how many
// threads the grid should at least have before chains are kept long (long chains amortise the two halo rows).
static inline void dws_plan(DwsArgs& a, int V, int KH, int KW, int TW, long long min_threads, int seg_max = 32) {
  const int dw = (KW == 1) ? 1 : a.dil_w, dh = (KH == 1) ? 1 : a.dil_h;
  a.QW = (a.W + TW * dw - 1) / (TW * dw) * dw;
  const long long per_row = (long long)a.N * a.QW * (a.C / V);
  const int chain_len = (a.H + dh - 1) / dh;
  int seg = seg_max;
  while (seg > 4 && per_row * dh * ((chain_len + seg - 1) / seg) < min_threads) seg >>= 1;
  a.seg = seg;
  a.VS = dh * ((chain_len + seg - 1) / seg);
  a.total = per_row * a.VS;
}

// One thread.  V = channels per 16-byte vector of T.
template <typename T, int V, int KH, int KW, int TW>
__device__ __forceinline__ void dw_strip_thread(const DwsArgs& a, long long t) {
  constexpr int NCOL = TW + KW - 1;
  constexpr int H2 = V / 2;
  const int CG = a.C / V;
  const int cg = (int)(t % CG);
  t /= CG;
  const int q = (int)(t % a.QW);
  t /= a.QW;
  const int vs = (int)(t % a.VS);
  const int n = (int)(t / a.VS);
  const int dw = (KW == 1) ? 1 : a.dil_w;
  const int dh = (KH == 1) ? 1 : a.dil_h;
  const int w0 = (q / dw) * TW * dw + (q % dw);
  const int h_first = (vs % dh) + (vs / dh) * a.seg * dh;
  if (w0 >= a.W || h_first >= a.H) return;
  const int c = cg * V;
  int nrows = (a.H - h_first + dh - 1) / dh;      // rows left in this chain
  if (nrows > a.seg) nrows = a.seg;

  // taps (x epilogue scale), shift and slopes in registers
  float2 wr[KH * KW][H2];
  float2 sh[H2], al[H2];
#pragma unroll
  for (int i = 0; i < H2; i += 2) {       // 16-byte parameter loads: channels c + 2i .. c + 2i + 3
    const float4 sc = a.scale ? ldw4(a.scale + c + 2 * i) : make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int tap = 0; tap < KH * KW; ++tap) {
      const float4 wv = ldw4(a.w + (size_t)tap * a.C + c + 2 * i);
      wr[tap][i] = make_float2(wv.x * sc.x, wv.y * sc.y);
      wr[tap][i + 1] = make_float2(wv.z * sc.z, wv.w * sc.w);
    }
    const float4 sv = a.shift ? ldw4(a.shift + c + 2 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 av = (a.act == 2) ? ldw4(a.alpha + c + 2 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    sh[i] = make_float2(sv.x, sv.y);
    sh[i + 1] = make_float2(sv.z, sv.w);
    al[i] = make_float2(av.x, av.y);
    al[i + 1] = make_float2(av.z, av.w);
  }
  // column validity (zero padding left / right, ragged right edge)
  unsigned cmask = 0, omask = 0;
#pragma unroll
  for (int j = 0; j < NCOL; ++j) {
    const int col = w0 + (j - (KW - 1) / 2) * dw;
    if (col >= 0 && col < a.W) cmask |= 1u << j;
  }
#pragma unroll
  for (int j = 0; j < TW; ++j)
    if (w0 + j * dw < a.W) omask |= 1u << j;
  const T* xb = reinterpret_cast<const T*>(a.x) + ((long long)n * a.H * a.W + (w0 - ((KW - 1) / 2) * dw)) * a.x_cs + c;
  T* yb = reinterpret_cast<T*>(a.y) + ((size_t)n * a.H * a.W + w0) * a.y_cs + c;
  const T* rb = a.res ? reinterpret_cast<const T*>(a.res) + ((size_t)n * a.H * a.W + w0) * a.res_cs + c : nullptr;
  const size_t xcol = (size_t)dw * a.x_cs, ycol = (size_t)dw * a.y_cs, rcol = (size_t)dw * a.res_cs;

  typedef typename DwsRaw<T>::type raw_t;
  auto load_row = [&](int row, raw_t (&r)[NCOL]) {
    const bool ok = row >= 0 && row < a.H;
    const T* p = xb + (long long)row * a.W * a.x_cs;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) r[j] = (ok && (cmask >> j & 1)) ? ldraw(p + j * xcol) : DwsRaw<T>::zero();
  };
  auto store_row = [&](int row, float2 (&acc)[TW][H2]) {
#pragma unroll
    for (int j = 0; j < TW; ++j) {
      if (!(omask >> j & 1)) continue;
      float2 v[H2];
#pragma unroll
      for (int i = 0; i < H2; ++i) v[i] = make_float2(acc[j][i].x + sh[i].x, acc[j][i].y + sh[i].y);
      if (rb) {
        float2 r[H2];
        unpack(ldraw(rb + (size_t)row * a.W * a.res_cs + j * rcol), r);
#pragma unroll
        for (int i = 0; i < H2; ++i) {
          if (a.pre_act) v[i] = dws_act2(v[i], a.act, al[i]);
          v[i].x += r[i].x;
          v[i].y += r[i].y;
        }
      }
#pragma unroll
      for (int i = 0; i < H2; ++i) v[i] = dws_act2(v[i], a.act, al[i]);
      stv(yb + (size_t)row * a.W * a.y_cs + j * ycol, v);
    }
  };

  raw_t cur[NCOL], nxt[NCOL];
  if (KH == 1) {
    load_row(h_first, cur);
    for (int k = 0; k < nrows; ++k) {
      const int row = h_first + k;
      if (k + 1 < nrows) load_row(row + 1, nxt);
      float2 xf[NCOL][H2];
#pragma unroll
      for (int j = 0; j < NCOL; ++j) unpack(cur[j], xf[j]);
      float2 acc[TW][H2];
#pragma unroll
      for (int j = 0; j < TW; ++j)
#pragma unroll
        for (int i = 0; i < H2; ++i) {
          float2 s = make_float2(xf[j][i].x * wr[0][i].x, xf[j][i].y * wr[0][i].y);
#pragma unroll
          for (int tp = 1; tp < KW; ++tp) s = ffma2(xf[j + tp][i], wr[tp][i], s);
          acc[j][i] = s;
        }
      store_row(row, acc);
#pragma unroll
      for (int j = 0; j < NCOL; ++j) cur[j] = nxt[j];
    }
    return;
  }
  // KH == 3: input row I_k = h_first + k*dh feeds output rows k-1 (bottom tap), k (centre), k+1 (top tap)
  float2 A[TW][H2], B[TW][H2], Cc[TW][H2];
#pragma unroll
  for (int j = 0; j < TW; ++j)
#pragma unroll
    for (int i = 0; i < H2; ++i) A[j][i] = B[j][i] = Cc[j][i] = make_float2(0.f, 0.f);
  auto step = [&](int k, float2 (&prev)[TW][H2], float2 (&mid)[TW][H2], float2 (&next)[TW][H2]) {
    const int row = h_first + k * dh;
    if (k < nrows) load_row(row + dh, nxt);
    float2 xf[NCOL][H2];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) unpack(cur[j], xf[j]);
#pragma unroll
    for (int j = 0; j < TW; ++j)
#pragma unroll
      for (int i = 0; i < H2; ++i) {
        float2 nx = make_float2(xf[j][i].x * wr[0][i].x, xf[j][i].y * wr[0][i].y);
#pragma unroll
        for (int tp = 1; tp < KW; ++tp) nx = ffma2(xf[j + tp][i], wr[tp][i], nx);
        next[j][i] = nx;
#pragma unroll
        for (int tp = 0; tp < KW; ++tp) {
          mid[j][i] = ffma2(xf[j + tp][i], wr[KW + tp][i], mid[j][i]);
          prev[j][i] = ffma2(xf[j + tp][i], wr[2 * KW + tp][i], prev[j][i]);
        }
      }
    if (k >= 1) store_row(row - dh, prev);
#pragma unroll
    for (int j = 0; j < NCOL; ++j) cur[j] = nxt[j];
  };
  load_row(h_first - dh, cur);
  for (int k = -1; k <= nrows; k += 3) {
    step(k, A, B, Cc);
    if (k + 1 > nrows) break;
    step(k + 1, B, Cc, A);
    if (k + 2 > nrows) break;
    step(k + 2, Cc, A, B);
  }
}

template <typename T, int V, int KH, int KW, int TW>
__global__ void __launch_bounds__(128) dw_strip_kernel(const DwsArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < a.total) dw_strip_thread<T, V, KH, KW, TW>(a, t);
}

}  // namespace
