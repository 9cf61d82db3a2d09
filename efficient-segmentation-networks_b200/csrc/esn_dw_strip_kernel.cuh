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
// DwsRaw<T>, ldraw / unpack / stv (one 16-byte channel vector <-> float2[V/2]), ldw4, ffma2, dws_act2<ACT> come from the
// including file.
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

// One thread.  V = channels per 16-byte vector of T.  ACT / RES: compile-time epilogue (ACT: ESN_ACT_* value; RES: 0 no residual,
// 1 residual, with a.pre_act deciding whether the activation also runs before the add).
template <typename T, int V, int KH, int KW, int TW, int ACT, int RES>
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

  // taps (x epilogue scale) in registers; the shift is the accumulators' initial value
  float2 wr[KH * KW][H2];
  float2 sh[H2];
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
    sh[i] = make_float2(sv.x, sv.y);
    sh[i + 1] = make_float2(sv.z, sv.w);
  }
  // column validity (zero padding left / right, ragged right edge); interior threads take the unchecked path
  unsigned cmask = 0, omask = 0;
#pragma unroll
  for (int j = 0; j < NCOL; ++j) {
    const int col = w0 + (j - (KW - 1) / 2) * dw;
    if (col >= 0 && col < a.W) cmask |= 1u << j;
  }
#pragma unroll
  for (int j = 0; j < TW; ++j)
    if (w0 + j * dw < a.W) omask |= 1u << j;
  const bool interior = cmask == (1u << NCOL) - 1u && omask == (1u << TW) - 1u;
  const long long xrow = (long long)a.W * a.x_cs, yrow = (long long)a.W * a.y_cs, rrow = (long long)a.W * a.res_cs;
  const long long xcol = (long long)dw * a.x_cs, ycol = (long long)dw * a.y_cs, rcol = (long long)dw * a.res_cs;
  // running row pointers: input row about to be loaded, output row about to be stored
  const int first_in = (KH == 1) ? h_first : h_first - dh;
  const T* xp = reinterpret_cast<const T*>(a.x) + ((long long)n * a.H + first_in) * xrow + (long long)(w0 - ((KW - 1) / 2) * dw) * a.x_cs + c;
  T* yp = reinterpret_cast<T*>(a.y) + ((long long)n * a.H + h_first) * yrow + (long long)w0 * a.y_cs + c;
  const T* rp = RES ? reinterpret_cast<const T*>(a.res) + ((long long)n * a.H + h_first) * rrow + (long long)w0 * a.res_cs + c : nullptr;
  int in_row = first_in;

  typedef typename DwsRaw<T>::type raw_t;
  auto load_row = [&](raw_t (&r)[NCOL]) {          // loads input row `in_row`, then advances to the next row of the chain
    const bool ok = in_row >= 0 && in_row < a.H;
    if (ok && interior) {
#pragma unroll
      for (int j = 0; j < NCOL; ++j) r[j] = ldraw(xp + j * xcol);
    } else {
#pragma unroll
      for (int j = 0; j < NCOL; ++j) r[j] = (ok && (cmask >> j & 1)) ? ldraw(xp + j * xcol) : DwsRaw<T>::zero();
    }
    xp += dh * xrow;
    in_row += dh;
  };
  auto store_row = [&](float2 (&acc)[TW][H2]) {    // epilogue of one finished output row, then advances the output row
    float2 al[H2];
#pragma unroll
    for (int i = 0; i < H2; ++i) al[i] = make_float2(0.f, 0.f);
    if (ACT == 2) {       // PReLU slopes: L1-resident 16-byte loads, not worth 8 registers across the whole walk
#pragma unroll
      for (int i = 0; i < H2; i += 2) {
        const float4 av = ldw4(a.alpha + c + 2 * i);
        al[i] = make_float2(av.x, av.y);
        al[i + 1] = make_float2(av.z, av.w);
      }
    }
#pragma unroll
    for (int j = 0; j < TW; ++j) {
      if (!interior && !(omask >> j & 1)) continue;
      float2 v[H2];
#pragma unroll
      for (int i = 0; i < H2; ++i) v[i] = acc[j][i];
      if (RES) {
        float2 r[H2];
        unpack(ldraw(rp + j * rcol), r);
#pragma unroll
        for (int i = 0; i < H2; ++i) {
          if (a.pre_act) v[i] = dws_act2<ACT>(v[i], al[i]);
          v[i].x += r[i].x;
          v[i].y += r[i].y;
        }
      }
#pragma unroll
      for (int i = 0; i < H2; ++i) v[i] = dws_act2<ACT>(v[i], al[i]);
      stv(yp + j * ycol, v);
    }
    yp += dh * yrow;
    if (RES) rp += dh * rrow;
  };

  raw_t R0[NCOL], R1[NCOL];       // row buffers swap roles every step (no register copies)
  if constexpr (KH == 1) {
    auto row1 = [&](int k, raw_t (&cur)[NCOL], raw_t (&nxt)[NCOL]) {
      if (k + 1 < nrows) load_row(nxt);
      float2 xf[NCOL][H2];
#pragma unroll
      for (int j = 0; j < NCOL; ++j) unpack(cur[j], xf[j]);
      float2 acc[TW][H2];
#pragma unroll
      for (int j = 0; j < TW; ++j)
#pragma unroll
        for (int i = 0; i < H2; ++i) {
          float2 s = sh[i];
#pragma unroll
          for (int tp = 0; tp < KW; ++tp) s = ffma2(xf[j + tp][i], wr[tp][i], s);
          acc[j][i] = s;
        }
      store_row(acc);
    };
    load_row(R0);
    for (int k = 0; k < nrows; k += 2) {
      row1(k, R0, R1);
      if (k + 1 >= nrows) break;
      row1(k + 1, R1, R0);
    }
  } else {
  // KH == 3: input row I_k = h_first + k*dh feeds output rows k-1 (bottom tap), k (centre), k+1 (top tap)
  float2 A[TW][H2], B[TW][H2], Cc[TW][H2];
#pragma unroll
  for (int j = 0; j < TW; ++j)
#pragma unroll
    for (int i = 0; i < H2; ++i) A[j][i] = B[j][i] = Cc[j][i] = make_float2(0.f, 0.f);
  // Three row buffers: while row k is consumed, rows k+1 and k+2 are in flight -- with two resident CTAs per SM (the 3x3
  // kernel needs ~200 registers) one step of arithmetic is shorter than a DRAM round trip, two steps are not.
  raw_t R2[NCOL];
  auto step = [&](int k, raw_t (&cur)[NCOL], raw_t (&fill)[NCOL], float2 (&prev)[TW][H2], float2 (&mid)[TW][H2],
                  float2 (&next)[TW][H2]) {
    if (k + 2 <= nrows) load_row(fill);
    float2 xf[NCOL][H2];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) unpack(cur[j], xf[j]);
#pragma unroll
    for (int j = 0; j < TW; ++j)
#pragma unroll
      for (int i = 0; i < H2; ++i) {
        float2 nx = sh[i];
#pragma unroll
        for (int tp = 0; tp < KW; ++tp) nx = ffma2(xf[j + tp][i], wr[tp][i], nx);
        next[j][i] = nx;
#pragma unroll
        for (int tp = 0; tp < KW; ++tp) {
          mid[j][i] = ffma2(xf[j + tp][i], wr[KW + tp][i], mid[j][i]);
          prev[j][i] = ffma2(xf[j + tp][i], wr[2 * KW + tp][i], prev[j][i]);
        }
      }
    if (k >= 1) store_row(prev);
  };
  load_row(R0);       // input rows I_-1 and I_0
  load_row(R1);
  for (int k = -1; k <= nrows; k += 3) {      // accumulator roles and row-buffer roles both rotate with period 3
    step(k, R0, R2, A, B, Cc);
    if (k + 1 > nrows) break;
    step(k + 1, R1, R0, B, Cc, A);
    if (k + 2 > nrows) break;
    step(k + 2, R2, R1, Cc, A, B);
  }
  }
}

template <typename T, int V, int KH, int KW, int TW, int ACT, int RES, int MINB>
__global__ void __launch_bounds__(128, MINB) dw_strip_kernel(const DwsArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < a.total) dw_strip_thread<T, V, KH, KW, TW, ACT, RES>(a, t);
}

}  // namespace
