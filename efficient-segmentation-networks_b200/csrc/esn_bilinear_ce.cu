// Close of a training iteration for the nets whose head is a bilinear up-sampling of low-resolution class scores
// (DABNet.py:181, FastSCNN.py:233, ContextNet, CGNet, EDANet: F.interpolate(scores, input.size()[2:], mode='bilinear',
// align_corners=False)) followed by CrossEntropyLoss2d (utils/losses/loss.py:15-32, called at train.py:351-353):
//
//     logits = interpolate(scores);  loss = sum_p w[t_p] (lse(logits_p) - logits_p[t_p]) / sum_p w[t_p]
//     d scores = interpolate^T ( w[t_p] (softmax(logits_p) - onehot(t_p)) )                    (times 1 / sum w, applied later)
//
// in ONE pass that never writes a full-resolution tensor.  The separate kernels write the fp32 logits (319 MB at
// 8 x 19 x 512 x 1024), read them twice for the loss and its gradient, write the gradient and read it back for the transposed
// interpolation: 1.6 GB of traffic and 0.55 ms of a 7.4 ms DABNet step for 2.5 MB of scores and 33 MB of labels.
//
// Geometry: an output row y reads source rows k = floor(src(y)) and min(k + 1, h - 1), with src(y) as ATen computes it in fp32
// (align_corners = False: max(0, (y + 0.5) h / H - 0.5); True: y (h - 1) / (H - 1)); the rows with the same k form "cell" k,
// likewise for columns -- any scale, both modes.  A CTA OWNS a kCR x kCC block of source pixels, walks the (kCR + 1) x
// (kCC + 1) cells that touch them (the border cells are recomputed by the neighbouring CTA: arithmetic is cheap here) and keeps
// only the gradient contributions to its own source pixels, in shared memory -- so the score gradient is written with plain
// stores: no global atomics, no zero fill.  A pixel's loss is counted by the CTA that owns the cell's upper-left source pixel.
//
// Inside a warp: lane = (cell q of four adjacent cells, row r of the cell); a thread blends its row's two source vectors
// once (A, B), walks the pixels of its cell row (logit_c = A_c + lx (B_c - A_c), soft-max in the log2 domain: one FFMA and
// one ex2.approx per class), sums the soft-max part of the gradient against (1 - lx) and lx in registers -- the onehot part
// goes straight into the CTA's tile, once per run of equal labels -- and the eight rows of a cell are combined by a
// reduce-scatter over the four source pixels (four shuffles per class), after which four lanes per cell add into the tile.
#include "esn_common.cuh"

namespace {

constexpr int kCR = 4, kCC = 16;          // source pixels owned by a CTA
constexpr int kThreads = 256;

struct BceArgs {
  const void* scores;
  const long long* target;
  const float* weight;
  float* sums;
  float* dscores;
  int N, h, w, C, cs, dcs, H, W, align, ignore;
  float sy, sx;       // source step per output row / column, as ATen's area_pixel_compute_scale gives it
};

// source coordinate of output index o (ATen's area_pixel_compute_source_index, fp32)
__device__ __forceinline__ float src_coord(int o, float scale, int align) {
  return align ? scale * (float)o : fmaxf(scale * ((float)o + 0.5f) - 0.5f, 0.f);
}
// first output index whose source cell floor(src) is >= k (n_out when there is none): an estimate from the inverse map,
// corrected against src_coord itself so that cells and per-pixel weights can never disagree
__device__ __forceinline__ int first_out(int k, float scale, int align, int n_out) {
  if (k <= 0) return 0;
  if (!(scale > 0.f)) return n_out;
  int o = (int)ceilf(align ? (float)k / scale : ((float)k + 0.5f) / scale - 0.5f);
  o = min(max(o, 0), n_out);
  while (o > 0 && (int)src_coord(o - 1, scale, align) >= k) --o;
  while (o < n_out && (int)src_coord(o, scale, align) < k) ++o;
  return o;
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename TS, int CP>
__global__ void __launch_bounds__(kThreads, CP <= 20 ? 2 : 1) bilinear_ce_kernel(const BceArgs a) {
  __shared__ float S[(kCR + 2) * (kCC + 2) * CP];      // source scores: tile row tr = source row clamp(k0 - 1 + tr)
  __shared__ float dS[kCR * kCC * CP];                 // gradient of the owned source pixels
  __shared__ float red[2][kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = lane & 7, q = lane >> 3;
  const int bw = (a.w + kCC - 1) / kCC, bh = (a.h + kCR - 1) / kCR;
  int b = blockIdx.x;
  const int bl = b % bw;
  b /= bw;
  const int bk = b % bh, n = b / bh;
  const int k0 = bk * kCR, l0 = bl * kCC;
  const int kr = min(kCR, a.h - k0), lc = min(kCC, a.w - l0);       // owned extent
  const TS* sc = reinterpret_cast<const TS*>(a.scores);
  for (int i = tid; i < (kCR + 2) * (kCC + 2) * CP; i += kThreads) {
    const int c = i % CP, p = i / CP;
    const int tc = p % (kCC + 2), tr = p / (kCC + 2);
    const int sr = min(max(k0 - 1 + tr, 0), a.h - 1), scol = min(max(l0 - 1 + tc, 0), a.w - 1);
    S[i] = c < a.C ? ld1<TS>(sc + ((size_t)(n * a.h + sr) * a.w + scol) * a.cs + c) : -1e30f;   // padded classes: exp() == 0
  }
  for (int i = tid; i < kCR * kCC * CP; i += kThreads) dS[i] = 0.f;
  __syncthreads();

  const int ncell = (kr + 1) * (lc + 1);                // cells (k0 - 1 + kc, l0 - 1 + lci), kc = 0 .. kr, lci = 0 .. lc
  float loss_acc = 0.f, w_acc = 0.f;
  for (int it = warp; it * 4 < ncell; it += kThreads / 32) {
    const int ci = it * 4 + q;                          // four consecutive cells of the flattened walk per warp
    const bool cell_ok = ci < ncell;
    const int kc = min(ci, ncell - 1) / (lc + 1), lci = min(ci, ncell - 1) - kc * (lc + 1);
    const int k = k0 - 1 + kc, l = l0 - 1 + lci;
    const bool live = cell_ok && k >= 0 && l >= 0;                 // cells start at source row / column 0
    const bool own_cell = live && kc >= 1 && lci >= 1;
    // output rows / columns of this cell
    const int ya = live ? first_out(k, a.sy, a.align, a.H) : 0, yb = live ? (k + 1 < a.h ? first_out(k + 1, a.sy, a.align, a.H) : a.H) : 0;
    const int xa = live ? first_out(l, a.sx, a.align, a.W) : 0, xb = live ? (l + 1 < a.w ? first_out(l + 1, a.sx, a.align, a.W) : a.W) : 0;
    const int max_rows = __reduce_max_sync(0xffffffffu, yb - ya);
    // tile rows kc, kc + 1 / columns lci, lci + 1 hold the (clamped) source pixels of this cell; which of them are ours
    const int rsA = min(max(k, 0), a.h - 1), rsB = min(max(k + 1, 0), a.h - 1);
    const int csA = min(max(l, 0), a.w - 1), csB = min(max(l + 1, 0), a.w - 1);
    const bool orA = rsA >= k0 && rsA < k0 + kr, orB = rsB >= k0 && rsB < k0 + kr;
    const bool ocA = csA >= l0 && csA < l0 + lc, ocB = csB >= l0 && csB < l0 + lc;
    float* dAA = dS + ((rsA - k0) * kCC + (csA - l0)) * CP;
    float* dAB = dS + ((rsA - k0) * kCC + (csB - l0)) * CP;
    float* dBA = dS + ((rsB - k0) * kCC + (csA - l0)) * CP;
    float* dBB = dS + ((rsB - k0) * kCC + (csB - l0)) * CP;
    const float* sAA = S + (kc * (kCC + 2) + lci) * CP;
    const float* sAB = sAA + CP;
    const float* sBA = sAA + (kCC + 2) * CP;
    const float* sBB = sBA + CP;
    const bool row_b = (r & 4) != 0, col_b = (r & 2) != 0;        // the source pixel this lane collects after the reduction
    float* dmine = row_b ? (col_b ? dBB : dBA) : (col_b ? dAB : dAA);
    const bool own_mine = live && (r & 1) == 0 && (row_b ? orB : orA) && (col_b ? ocB : ocA);
    for (int yy0 = 0; yy0 < max_rows; yy0 += 8) {          // eight rows of every cell per pass (warp-uniform trip count)
      const int y = ya + yy0 + r;
      const bool row_ok = y < yb;
      float G0[CP], G1[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) G0[c] = G1[c] = 0.f;
      float ly = 0.f;
      if (row_ok) {
        ly = src_coord(y, a.sy, a.align) - (float)k;
        // the row's two blended source vectors, in units of log2 (one FFMA + one EX2 per class and pixel below)
        constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
        float A[CP], D[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          const float va = sAA[c] + ly * (sBA[c] - sAA[c]);
          const float vb = sAB[c] + ly * (sBB[c] - sAB[c]);
          A[c] = va * kLog2e;
          D[c] = (vb - va) * kLog2e;
        }
        const long long* trow = a.target + ((size_t)n * a.H + y) * a.W;
        // onehot part of the gradient: summed per run of equal labels and added straight into the owned tile when the label
        // changes (four shared-memory atomics per run; per pixel when every pixel has another label)
        const float wyA = 1.f - ly;
        auto flush = [&](int cls, float oa, float ob) {
          if (orA && ocA) atomicAdd(dAA + cls, -(wyA * oa));
          if (orA && ocB) atomicAdd(dAB + cls, -(wyA * ob));
          if (orB && ocA) atomicAdd(dBA + cls, -(ly * oa));
          if (orB && ocB) atomicAdd(dBB + cls, -(ly * ob));
        };
        int cur = -1;
        float oa = 0.f, ob = 0.f;
        for (int x = xa; x < xb; ++x) {
          const long long t = __ldg(trow + x);
          if (t == a.ignore || t < 0 || t >= a.C) continue;
          const int ti = (int)t;
          const float wy = a.weight ? __ldg(a.weight + ti) : 1.f;
          const float lx = src_coord(x, a.sx, a.align) - (float)l;
          float v[CP];
          float m = -INFINITY;
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            v[c] = fmaf(lx, D[c], A[c]);
            m = fmaxf(m, v[c]);
          }
          float z = 0.f;
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            v[c] = ex2_ftz(v[c] - m);
            z += v[c];
          }
          if (own_cell) {
            // the labelled class's logit from the tile (a dynamic index into v[] would spill it)
            const float ta = sAA[ti] + ly * (sBA[ti] - sAA[ti]), tb = sAB[ti] + ly * (sBB[ti] - sAB[ti]);
            loss_acc += wy * ((m + __log2f(z)) * kLn2 - fmaf(lx, tb - ta, ta));
            w_acc += wy;
          }
          const float inv = __fdividef(wy, z);
          const float wa = 1.f - lx;
          const float i0 = (col_b ? lx : wa) * inv, i1 = (col_b ? wa : lx) * inv;      // G0 = this lane's column, G1 = the other
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            G0[c] = fmaf(i0, v[c], G0[c]);
            G1[c] = fmaf(i1, v[c], G1[c]);
          }
          if (ti != cur) {
            if (cur >= 0) flush(cur, oa, ob);
            cur = ti;
            oa = ob = 0.f;
          }
          oa = fmaf(wa, wy, oa);
          ob = fmaf(lx, wy, ob);
        }
        if (cur >= 0) flush(cur, oa, ob);
      }
      // combine the eight rows of a cell by reduce-scatter: a lane ends up with ONE of the four source pixels -- row B when
      // r & 4, column B when r & 2 -- so it multiplies by the row weight it keeps (wm) and the one it hands over (wt), swaps
      // the row halves with lane ^ 4, the column halves with lane ^ 2 and sums the last pair with lane ^ 1: four shuffles
      // per class instead of twelve.  Idle lanes carry zeros.
      {
        const float wyA = 1.f - ly;
        const float wm = row_b ? ly : wyA, wt = row_b ? wyA : ly;
#pragma unroll
        for (int c = 0; c < CP; ++c) {
          float p0 = wm * G0[c], p1 = wm * G1[c];
          p0 += __shfl_xor_sync(0xffffffffu, wt * G0[c], 4);
          p1 += __shfl_xor_sync(0xffffffffu, wt * G1[c], 4);
          p0 += __shfl_xor_sync(0xffffffffu, p1, 2);
          p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
          if (own_mine) atomicAdd(dmine + c, p0);
        }
      }
    }
  }
  // loss sums: warp shuffle -> shared -> one atomic pair per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    w_acc += __shfl_xor_sync(0xffffffffu, w_acc, o);
  }
  if (lane == 0) {
    red[0][warp] = loss_acc;
    red[1][warp] = w_acc;
  }
  __syncthreads();
  if (warp == 0) {
    loss_acc = lane < kThreads / 32 ? red[0][lane] : 0.f;
    w_acc = lane < kThreads / 32 ? red[1][lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
      w_acc += __shfl_xor_sync(0xffffffffu, w_acc, o);
    }
    if (lane == 0) {
      atomicAdd(a.sums, loss_acc);
      atomicAdd(a.sums + 1, w_acc);
    }
  }
  // the owned gradients, every lane of the padded pixel (zeros behind the classes)
  for (int i = tid; i < kr * lc * a.dcs; i += kThreads) {
    const int c = i % a.dcs, p = i / a.dcs;
    const int tc = p % lc, tr = p / lc;
    a.dscores[((size_t)(n * a.h + k0 + tr) * a.w + l0 + tc) * a.dcs + c] = c < CP ? dS[(tr * kCC + tc) * CP + c] : 0.f;
  }
}

template <typename TS>
void launch_bce(const BceArgs& a, int grid, cudaStream_t st) {
  if (a.C <= 12)
    bilinear_ce_kernel<TS, 12><<<grid, kThreads, 0, st>>>(a);
  else if (a.C <= 20)
    bilinear_ce_kernel<TS, 20><<<grid, kThreads, 0, st>>>(a);
  else
    bilinear_ce_kernel<TS, 32><<<grid, kThreads, 0, st>>>(a);
}

}  // namespace

extern "C" int esn_bilinear_ce(const EsnBilinearCE* p, void* stream) {
  if (!p || !p->target || !p->sums || !esn_valid_nhwc(p->scores) || !esn_valid_nhwc(p->dscores)) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->scores;
  const EsnTensor& g = p->dscores;
  if (x.dtype != ESN_F32 && x.dtype != ESN_BF16) return ESN_ERR_BAD_ARG;
  if (g.dtype != ESN_F32) return ESN_ERR_BAD_ARG;
  if (g.n != x.n || g.h != x.h || g.w != x.w || g.c != x.c) return ESN_ERR_BAD_SHAPE;
  if (x.c < 1 || x.c > 32) return ESN_ERR_UNSUPPORTED;
  if (p->out_h < 1 || p->out_w < 1) return ESN_ERR_BAD_SHAPE;
  BceArgs a;
  a.scores = x.ptr;
  a.target = reinterpret_cast<const long long*>(p->target);
  a.weight = p->weight;
  a.sums = p->sums;
  a.dscores = reinterpret_cast<float*>(g.ptr);
  a.N = x.n; a.h = x.h; a.w = x.w; a.C = x.c; a.cs = x.c_stride; a.dcs = g.c_stride;
  a.H = p->out_h; a.W = p->out_w; a.ignore = p->ignore_label;
  a.align = p->align_corners ? 1 : 0;
  // ATen's area_pixel_compute_scale<float>: align_corners ? (in - 1) / (out - 1) (0 for out == 1) : in / out
  a.sy = a.align ? (a.H > 1 ? (float)(x.h - 1) / (float)(a.H - 1) : 0.f) : (float)x.h / (float)a.H;
  a.sx = a.align ? (a.W > 1 ? (float)(x.w - 1) / (float)(a.W - 1) : 0.f) : (float)x.w / (float)a.W;
  const long long grid = (long long)x.n * ((x.h + kCR - 1) / kCR) * ((x.w + kCC - 1) / kCC);
  if (grid > 0x7fffffffLL) return ESN_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x.dtype == ESN_F32)
    launch_bce<float>(a, (int)grid, st);
  else
    launch_bce<__nv_bfloat16>(a, (int)grid, st);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
