// Network stem: 3x3 stride-2 pad-1 convolution read straight from the caller's NCHW fp32 image
// (Cin = 3), optionally concatenated with MaxPool2d(2,2) of the same image, then the per-channel
// affine (bias + folded BatchNorm) and activation, written as NHWC (bf16 or fp32).
//   ERFNet.py:16-27  DownsamplerBlock(3,16): 13 conv + 3 pool channels, BN, ReLU
//   DABNet.py:132    Conv(3,32,3,2) + BNPReLU
// HBM-bound: reads 12 B/input pixel once (neighbour re-reads hit L1), writes Ctot*2 B per output pixel.
#include "esn_common.cuh"

namespace {

struct StemArgs {
  const float* x;
  void* y;
  const float* w;  // [9][3][cconv]
  int N, H, W, Ho, Wo, cconv, ctot, y_cs, with_pool, pad;
  EpiArgs ep;
};

template <typename TO, int CPAD>
__global__ void __launch_bounds__(128) stem_kernel(const StemArgs a) {
  __shared__ float sw[27 * CPAD];
  __shared__ float sp[3 * CPAD];
  for (int i = threadIdx.x; i < 27 * CPAD; i += blockDim.x) {
    const int t = i / CPAD, c = i % CPAD;
    sw[i] = c < a.cconv ? a.w[t * a.cconv + c] : 0.f;
  }
  for (int i = threadIdx.x; i < CPAD; i += blockDim.x) {
    const bool in = i < a.ctot;
    sp[i] = (in && a.ep.scale) ? a.ep.scale[i] : 1.f;
    sp[CPAD + i] = (in && a.ep.shift) ? a.ep.shift[i] : 0.f;
    sp[2 * CPAD + i] = (in && a.ep.act == ESN_ACT_PRELU) ? a.ep.alpha[i] : 0.f;
  }
  __syncthreads();
  const long long total = (long long)a.N * a.Ho * a.Wo;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % a.Wo);
  const int ho = (int)((idx / a.Wo) % a.Ho);
  const int n = (int)(idx / ((long long)a.Wo * a.Ho));
  const size_t plane = (size_t)a.H * a.W;
  const float* xb = a.x + (size_t)n * 3 * plane;

  float v[27];
  unsigned okmask = 0;   // bit (r*3+s): tap inside the image (needed by the 3x3 pool: padding is -inf there)
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = 2 * ho - a.pad + r;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int wi = 2 * wo - a.pad + s;
      const bool ok = hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
      okmask |= (ok ? 1u : 0u) << (r * 3 + s);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = ok ? __ldg(xb + c * plane + (size_t)hi * a.W + wi) : 0.f;
    }
  }
  float acc[CPAD];
#pragma unroll
  for (int c = 0; c < CPAD; ++c) acc[c] = 0.f;
#pragma unroll
  for (int t = 0; t < 27; ++t) {
#pragma unroll
    for (int c4 = 0; c4 < CPAD; c4 += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(sw + t * CPAD + c4);
      acc[c4] = fmaf(v[t], wv.x, acc[c4]);
      acc[c4 + 1] = fmaf(v[t], wv.y, acc[c4 + 1]);
      acc[c4 + 2] = fmaf(v[t], wv.z, acc[c4 + 2]);
      acc[c4 + 3] = fmaf(v[t], wv.w, acc[c4 + 3]);
    }
  }
  if (a.with_pool) {  // 1: taps (1,1),(1,2),(2,1),(2,2) = the 2x2 window; 2: all valid taps = MaxPool2d(3,2,1)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float m;
      if (a.with_pool == 2) {
        m = -INFINITY;
#pragma unroll
        for (int t = 0; t < 9; ++t)
          if (okmask & (1u << t)) m = fmaxf(m, v[t * 3 + c]);
      } else {
        m = fmaxf(fmaxf(v[(1 * 3 + 1) * 3 + c], v[(1 * 3 + 2) * 3 + c]),
                  fmaxf(v[(2 * 3 + 1) * 3 + c], v[(2 * 3 + 2) * 3 + c]));
      }
#pragma unroll
      for (int k = 0; k < CPAD; ++k)
        if (k == a.cconv + c) acc[k] = m;
    }
  }
#pragma unroll
  for (int c = 0; c < CPAD; ++c) {
    const float t = fmaf(acc[c], sp[c], sp[CPAD + c]);
    acc[c] = apply_act(t, a.ep.act, sp[2 * CPAD + c]);
  }
  TO* yp = reinterpret_cast<TO*>(a.y) + (size_t)idx * a.y_cs;
#pragma unroll
  for (int c4 = 0; c4 < CPAD; c4 += 4)
    if (c4 < a.ctot) st4<TO>(yp + c4, make_float4(acc[c4], acc[c4 + 1], acc[c4 + 2], acc[c4 + 3]));
}

// ------------------------------------------------------------------------------------------------------------------
// bf16-output stem on the tensor cores.  The CUDA-core kernel above spends 27 * Cout fp32 FMAs per output pixel (864 at
// Cout = 32): it is FMA-issue-bound at 0.18-0.26 of HBM peak.  Here the 3x3x3 patch of an output pixel is one 27-element
// (padded to 32) row of an im2col A operand and the conv is D[16 px][Cout] = A[16 px][32] x B[32][Cout] per warp on
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32).  K = 27 is far too small for tcgen05 to matter: with the arithmetic on the tensor
// cores the kernel is bound by the image read + output write, which is the point.  The fp32 image is split into bf16 hi + lo
// parts (two MMAs per tile) so the input keeps 16 mantissa bits; the weights are bf16 as in every other bf16 layer.
//   * a CTA walks (image, output row, 64-pixel segment) work items; the three input rows x three channel planes of a segment
//     (3 x 3 x 129 floats) are staged in shared memory by coalesced loads -- each image element is read from global once per
//     output row it feeds -- zero-filled outside the image (the conv's zero padding);
//   * each warp builds its A fragments from shared memory (16 conflict-light LDS.32 per lane), issues 4 * NT MMAs, applies the
//     folded BN / activation, and the tile leaves through a shared-memory transpose as 16-byte NHWC stores;
//   * the max-pool channels of ERFNet / ENet (cat[conv, pool]) are computed from the same staged rows.
struct StemFrag {
  uint32_t b[2][4][2];     // [k-step][n-tile][reg]: bf16x2 B fragments (weights), built once per thread
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

constexpr int kStemSeg = 64;                 // output pixels per work item (4 warps x 16)
constexpr int kStemCols = 2 * kStemSeg + 1;  // input columns a segment touches
constexpr int kStemPitch = kStemCols + 3;    // shared-memory row pitch in floats (132: keeps the 9 rows' banks apart)

// SPLIT: 0 = image rounded to bf16; 1 = image as bf16 hi + lo (16 mantissa bits, 2 MMAs per tile)
template <int NT, int SPLIT>
__global__ void __launch_bounds__(128) stem_mma_kernel(const StemArgs a, const int segs_per_row, const int items) {
  constexpr int CP = NT * 8;
  __shared__ float sx[9 * kStemPitch];                 // [row r][channel c] -> sx[(r * 3 + c) * pitch + col]
  __shared__ __align__(16) __nv_bfloat16 so[4][16 * CP];   // per-warp output tile [16 px][CP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  // epilogue parameters of this lane's columns (8t + 2q, 8t + 2q + 1) in registers
  float psc[NT][2], psh[NT][2], pal[NT][2];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 8 * t + 2 * q + e;
      const bool in = c < a.ctot;
      psc[t][e] = (in && a.ep.scale) ? __ldg(a.ep.scale + c) : 1.f;
      psh[t][e] = (in && a.ep.shift) ? __ldg(a.ep.shift + c) : 0.f;
      pal[t][e] = (in && a.ep.act == ESN_ACT_PRELU) ? __ldg(a.ep.alpha + c) : 0.f;
    }
  // B fragments: B[k][n] = w[k][n] (k = tap * 3 + cin < 27, n < cconv), zero elsewhere; lane holds k = ks*16 + 2q (+1, +8, +9), n = 8t + g
  StemFrag fr;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int n = 8 * t + g;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k0 = ks * 16 + 2 * q + 8 * h;
        const float w0 = (k0 < 27 && n < a.cconv) ? __ldg(a.w + k0 * a.cconv + n) : 0.f;
        const float w1 = (k0 + 1 < 27 && n < a.cconv) ? __ldg(a.w + (k0 + 1) * a.cconv + n) : 0.f;
        fr.b[ks][t][h] = pack_bf16(w0, w1);
      }
    }
  // shared-memory offsets of this lane's A elements: k -> (r, s, c) = (k / 9, (k / 3) % 3, k % 3); pixel p reads column 2p + s.
  // K padding (k >= 27) reads element 0 of the window: finite, and multiplied by a zero weight row.
  int aoff[2][2][2];       // [k-step][k-half (k, k + 8)][element (k, k + 1)]
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = ks * 16 + 2 * q + 8 * h + e;
        aoff[ks][h][e] = k < 27 ? ((k / 9) * 3 + (k % 3)) * kStemPitch + (k / 3) % 3 + 2 * (warp * 16 + g) : 0;
      }
  const int plane = a.H * a.W;
  // staging slots of this thread: element i = tid + 128 j of the [9][129] window -> (row r, channel c, column); fixed for the
  // whole kernel.  The NEXT item's elements are fetched into registers before the current item's arithmetic, so the global
  // round trip overlaps the MMAs / epilogue instead of sitting between two barriers.
  constexpr int kSlots = (9 * kStemCols + 127) / 128;      // 10
  int soff[kSlots], goff[kSlots];
  short srow[kSlots], scol[kSlots];
#pragma unroll
  for (int j = 0; j < kSlots; ++j) {
    const int i = threadIdx.x + 128 * j;
    const int rc = i / kStemCols, col = i - rc * kStemCols;
    const bool used = i < 9 * kStemCols;
    soff[j] = used ? rc * kStemPitch + col : -1;
    srow[j] = (short)(rc / 3);
    scol[j] = (short)col;
    goff[j] = used ? (rc - 3 * (rc / 3)) * plane + (rc / 3) * a.W + col : 0;     // channel plane + row + column
  }
  float pre[kSlots];
  auto fetch = [&](int item) {
    const int seg = item % segs_per_row;
    const int t2 = item / segs_per_row;
    const int ho = t2 % a.Ho, n = t2 / a.Ho;
    const int col0 = 2 * seg * kStemSeg - a.pad, row0 = 2 * ho - a.pad;
    const float* xb = a.x + (size_t)n * 3 * plane + (long long)row0 * a.W + col0;
    if (row0 >= 0 && row0 + 2 < a.H && col0 >= 0 && col0 + kStemCols <= a.W) {      // interior: no per-element checks
#pragma unroll
      for (int j = 0; j < kSlots; ++j) pre[j] = __ldg(xb + goff[j]);
    } else {
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        const int hi = row0 + srow[j], wi = col0 + scol[j];
        pre[j] = (soff[j] >= 0 && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) ? __ldg(xb + goff[j]) : 0.f;
      }
    }
  };
  if ((int)blockIdx.x < items) fetch(blockIdx.x);
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int seg = item % segs_per_row;
    const int t2 = item / segs_per_row;
    const int ho = t2 % a.Ho, n = t2 / a.Ho;
    const int wo0 = seg * kStemSeg;
    const int col0 = 2 * wo0 - a.pad, row0 = 2 * ho - a.pad;
    __syncthreads();       // previous item's readers are done with sx
#pragma unroll
    for (int j = 0; j < kSlots; ++j)
      if (soff[j] >= 0) sx[soff[j]] = pre[j];
    __syncthreads();
    if (item + (int)gridDim.x < items) fetch(item + gridDim.x);
    const int p0 = warp * 16;                       // this warp's first pixel inside the segment
    if (wo0 + p0 < a.Wo) {
      float d[NT][4];
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) d[t][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        // A fragment regs: {row g, k 2q..}, {row g + 8, k 2q..}, {row g, k 2q + 8..}, {row g + 8, k 2q + 8..}
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const float v0 = sx[aoff[ks][h][0] + 16 * rr], v1 = sx[aoff[ks][h][1] + 16 * rr];      // row g + 8: 16 columns on
            const uint32_t hp = pack_bf16(v0, v1);
            ahi[2 * h + rr] = hp;
            if (SPLIT) alo[2 * h + rr] = pack_bf16(v0 - __uint_as_float(hp << 16), v1 - __uint_as_float(hp & 0xffff0000u));
          }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          mma16816(d[t], ahi, fr.b[ks][t]);
          if (SPLIT) mma16816(d[t], alo, fr.b[ks][t]);
        }
      }
      // epilogue: affine + activation, bf16, into the warp's [16][CP] tile (lane holds rows g, g + 8, columns 8t + 2q, + 1)
      __nv_bfloat16* tile = so[warp];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int c = 8 * t + 2 * q;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const float v0 = apply_act(fmaf(d[t][2 * rr], psc[t][0], psh[t][0]), a.ep.act, pal[t][0]);
          const float v1 = apply_act(fmaf(d[t][2 * rr + 1], psc[t][1], psh[t][1]), a.ep.act, pal[t][1]);
          *reinterpret_cast<uint32_t*>(tile + (g + 8 * rr) * CP + c) = pack_bf16(v0, v1);
        }
      }
      __syncwarp();
      if (a.with_pool) {      // pooled image channels land behind the conv channels: 16 px x 3 channels = 48 values per warp
        for (int i = lane; i < 48; i += 32) {
          const int px = i / 3, c = i - 3 * px;
          const int cb = 2 * (p0 + px);
          float m = -INFINITY;
          if (a.with_pool == 2) {       // MaxPool2d(3, 2, 1): all taps inside the image (padding is -inf there)
            for (int r = 0; r < 3; ++r)
              for (int sidx = 0; sidx < 3; ++sidx) {
                const int hi = row0 + r, wi = col0 + cb + sidx;
                if (hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) m = fmaxf(m, sx[(r * 3 + c) * kStemPitch + cb + sidx]);
              }
          } else {                      // MaxPool2d(2, 2): taps (1,1), (1,2), (2,1), (2,2) of the padded 3x3 window
            m = fmaxf(fmaxf(sx[(3 + c) * kStemPitch + cb + 1], sx[(3 + c) * kStemPitch + cb + 2]),
                      fmaxf(sx[(6 + c) * kStemPitch + cb + 1], sx[(6 + c) * kStemPitch + cb + 2]));
          }
          const int k = a.cconv + c;
          const float sc = a.ep.scale ? __ldg(a.ep.scale + k) : 1.f, sh = a.ep.shift ? __ldg(a.ep.shift + k) : 0.f;
          const float al = a.ep.act == ESN_ACT_PRELU ? __ldg(a.ep.alpha + k) : 0.f;
          tile[px * CP + k] = __float2bfloat16_rn(apply_act(fmaf(m, sc, sh), a.ep.act, al));
        }
        __syncwarp();
      }
      // 16-byte stores: 16 px x ctot channels; consecutive pixels are y_cs elements apart
      __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(a.y) + (((size_t)n * a.Ho + ho) * a.Wo + wo0 + p0) * a.y_cs;
      const int vpp = a.ctot >> 3;            // 16-byte vectors per pixel (ctot % 8 == 0 is checked on the host)
      const int npx = min(16, a.Wo - wo0 - p0);
      for (int i = lane; i < npx * vpp; i += 32) {
        const int px = i / vpp, v = i - px * vpp;
        *reinterpret_cast<uint4*>(yb + (size_t)px * a.y_cs + 8 * v) = *reinterpret_cast<const uint4*>(tile + px * CP + 8 * v);
      }
    }
  }
}

}  // namespace

extern "C" int esn_stem_conv3x3s2(const EsnStem* p, void* stream) {
  if (!p || !p->w || !p->x.ptr || !esn_valid_nhwc(p->y)) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (x.layout != ESN_NCHW || x.dtype != ESN_F32 || x.c != 3) return ESN_ERR_UNSUPPORTED;
  const int pad = (p->with_pool & ESN_STEM_PAD0) ? 0 : 1;
  const int pool = p->with_pool & 3;
  if (x.n != y.n || y.h != (x.h + 2 * pad - 3) / 2 + 1 || y.w != (x.w + 2 * pad - 3) / 2 + 1) return ESN_ERR_BAD_SHAPE;
  if (pool == 1 && ((x.h | x.w) & 1)) return ESN_ERR_UNSUPPORTED;
  if (pool && !pad) return ESN_ERR_UNSUPPORTED;
  const int ctot = p->cconv + (pool ? 3 : 0);
  if (y.c != ctot || ctot > 32 || ctot % 4 || y.c_stride % 4) return ESN_ERR_UNSUPPORTED;
  const size_t ysz = y.dtype == ESN_F32 ? 4 : 2;
  if ((uintptr_t)y.ptr % (4 * ysz)) return ESN_ERR_ALIGN;
  int rc = esn_check_epilogue(p->ep, y);
  if (rc) return rc;
  if (p->ep.residual.ptr) return ESN_ERR_UNSUPPORTED;
  StemArgs a;
  a.x = reinterpret_cast<const float*>(x.ptr);
  a.y = y.ptr;
  a.w = p->w;
  a.N = x.n;
  a.H = x.h;
  a.W = x.w;
  a.Ho = y.h;
  a.Wo = y.w;
  a.cconv = p->cconv;
  a.ctot = ctot;
  a.y_cs = y.c_stride;
  a.with_pool = pool;
  a.pad = pad;
  a.ep = make_epi(p->ep);
  const long long total = (long long)y.n * y.h * y.w;
  const int block = 128, grid = esn_cdiv(total, block);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool c16 = ctot <= 16;
  static const bool no_mma = [] { const char* e = getenv("ESN_STEM_FP32"); return e && e[0] == '1'; }();
  // measured on B200 (profiles/r02_bench_stem.json): 32 output channels 0.79 -> 0.42 ms on 16 x 1024 x 2048; with 16 channels
  // (13 conv + 3 pool) the CUDA-core kernel's 27 * 13 FMAs per pixel are as fast as the MMA kernel's fixed per-tile work, so
  // those stems keep the exact-fp32-weight kernel
  if (y.dtype == ESN_BF16 && !no_mma && ctot > 16 && ctot % 8 == 0 && y.c_stride % 8 == 0 && (uintptr_t)y.ptr % 16 == 0) {
    // tensor-core stem (bf16 output): persistent CTAs over (image, output row, 64-pixel segment) items
    const int segs = esn_cdiv(y.w, kStemSeg);
    const long long items64 = (long long)y.n * y.h * segs;
    if (items64 > 0x7fffffffLL || (long long)x.h * x.w * 3 > 0x7fffffffLL) return ESN_ERR_UNSUPPORTED;
    const int items = (int)items64;
    const int g2 = items < 148 * 12 ? items : 148 * 12;
    static const int split = [] { const char* e = getenv("ESN_STEM_SPLIT"); return e ? atoi(e) : 1; }();
    if (c16) {
      if (split) stem_mma_kernel<2, 1><<<g2, 128, 0, st>>>(a, segs, items);
      else stem_mma_kernel<2, 0><<<g2, 128, 0, st>>>(a, segs, items);
    } else {
      if (split) stem_mma_kernel<4, 1><<<g2, 128, 0, st>>>(a, segs, items);
      else stem_mma_kernel<4, 0><<<g2, 128, 0, st>>>(a, segs, items);
    }
  } else if (y.dtype == ESN_BF16) {
    if (c16) stem_kernel<__nv_bfloat16, 16><<<grid, block, 0, st>>>(a);
    else stem_kernel<__nv_bfloat16, 32><<<grid, block, 0, st>>>(a);
  } else {
    if (c16) stem_kernel<float, 16><<<grid, block, 0, st>>>(a);
    else stem_kernel<float, 32><<<grid, block, 0, st>>>(a);
  }
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
