// ConvTranspose2d(16, classes, 3, stride 2, padding 1, output_padding 1) fused with the argmax over classes on the warp-level
// tensor cores: ENet's `transposed_conv` + the CPU argmax of test.py:79-82 (ENet.py:229-236, 271).  The two-kernel path wrote
// the 24-channel bf16 scores of the FULL-resolution image (3.2 GB at 32 x 1024 x 2048) and read them back for the argmax:
// 2.8 + 2.3 ms of a 23.4 ms step.  Here the scores never leave registers.
//
//   y[2i+a, 2j+b, co] = sum over (dy, dx) in {0,1}^2 of  x[i+dy, j+dx, :] . W[:, co, a+1-2dy, b+1-2dx]      (kernel index in 0..2)
//
// i.e. output position (a, b) of input pixel (i, j) sees 1 / 2 / 2 / 4 of the four neighbours: nine (position, neighbour)
// pairs = the nine filter taps.  A warp takes 16 consecutive input pixels: for every pair one mma.m16n8k16 per 8 classes
// (M = 16 pixels, K = 16 input channels, N = 24 = classes padded): 27 MMAs, fp32 accumulators, bias as their initial value,
// -inf in the padded classes.  A fragments are 4-byte loads straight from the NHWC activations; B fragments (bf16 weights,
// packed on the host in fragment order) stay in registers for the kernel's lifetime.  Argmax: six candidates per lane in
// ascending class order, then two shuffle rounds inside the quad (ties -> lower class: first maximum wins); lane t of a quad
// stores the two mask bytes of output row a = t >> 1 for pixel g + 8 * (t & 1).
#include "esn_common.cuh"

namespace {

struct HeadT3Args {
  const __nv_bfloat16* x;
  const uint32_t* wfrag;    // [9 pairs][3 n-tiles][32 lanes][2]
  const float* bias;        // [classes] or null
  uint8_t* mask;            // (N, 2h, 2w)
  int N, H, W, x_cs, classes;
  long long tiles;
};

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128, 3) head_convt3x3s2_mask_kernel(const HeadT3Args a) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  uint32_t bf[9][3][2];
#pragma unroll
  for (int p = 0; p < 9; ++p)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(a.wfrag) + (p * 3 + nt) * 32 + lane);
      bf[p][nt][0] = v.x;
      bf[p][nt][1] = v.y;
    }
  float binit[3][2];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cls = nt * 8 + 2 * t + e;
      binit[nt][e] = cls < a.classes ? (a.bias ? __ldg(a.bias + cls) : 0.f) : -INFINITY;
    }
  const int wt = a.W >> 4;
  for (long long tile = warp; tile < a.tiles; tile += nwarps) {
    const int j0 = (int)(tile % wt) * 16;
    const int i = (int)((tile / wt) % a.H);
    const long long n = tile / ((long long)wt * a.H);
    // A fragments of the four neighbours (dy, dx): rows g / g + 8 = pixels j0 + g (+ 8) + dx of image row i + dy
    uint32_t af[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int dy = nb >> 1, dx = nb & 1;
      const bool rok = i + dy < a.H;
#pragma unroll
      for (int hs = 0; hs < 2; ++hs) {
        const int col = j0 + g + 8 * hs + dx;
        const bool ok = rok && col < a.W;
        const uint32_t* px = reinterpret_cast<const uint32_t*>(a.x + ((size_t)(n * a.H + (rok ? i + dy : i)) * a.W + (ok ? col : j0)) * a.x_cs);
        af[nb][hs] = ok ? __ldg(px + t) : 0u;            // channels 2t, 2t+1
        af[nb][hs + 2] = ok ? __ldg(px + t + 4) : 0u;    // channels 2t+8, 2t+9
      }
    }
    float acc[4][3][4];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        acc[pos][nt][0] = acc[pos][nt][2] = binit[nt][0];
        acc[pos][nt][1] = acc[pos][nt][3] = binit[nt][1];
      }
    // pair p -> (output position a*2+b, neighbour dy*2+dx)
    constexpr int kPos[9] = {0, 1, 1, 2, 2, 3, 3, 3, 3};
    constexpr int kNb[9] = {0, 0, 1, 0, 2, 0, 1, 2, 3};
#pragma unroll
    for (int p = 0; p < 9; ++p)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) mma_bf16_16816(acc[kPos[p]][nt], af[kNb[p]], bf[p][nt][0], bf[p][nt][1]);
    // argmax per (position, pixel half): c0,c1 = row g (classes nt*8 + 2t, +1), c2,c3 = row g + 8
    int res[4][2];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos)
#pragma unroll
      for (int hs = 0; hs < 2; ++hs) {
        float bv = acc[pos][0][hs * 2];
        int bi = 2 * t;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (nt == 0 && e == 0) continue;
            const float v = acc[pos][nt][hs * 2 + e];
            if (v > bv) { bv = v; bi = nt * 8 + 2 * t + e; }
          }
#pragma unroll
        for (int off = 1; off < 4; off <<= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        res[pos][hs] = bi;
      }
    const int hs = t & 1, ra = t >> 1;
    const int s0 = hs ? res[0][1] : res[0][0], s1 = hs ? res[1][1] : res[1][0];
    const int s2 = hs ? res[2][1] : res[2][0], s3 = hs ? res[3][1] : res[3][0];
    const uchar2 out = make_uchar2((uint8_t)(ra ? s2 : s0), (uint8_t)(ra ? s3 : s1));
    uint8_t* o = a.mask + ((size_t)(n * 2 * a.H + 2 * i + ra) * (2 * a.W) + 2 * (j0 + g + 8 * hs));
    *reinterpret_cast<uchar2*>(o) = out;
  }
}

// ---- ConvTranspose2d(16, classes, 2, stride 2) + argmax: the close of ERFNet / ESNet (ERFNet.py:112,128; ESNet.py:182) --------
//
//   y[2i+a, 2j+b, co] = x[i, j, :] . W[:, co, a, b] + bias[co]
//
// Every input pixel owns its four output pixels, so a warp takes 16 consecutive input pixels and runs, per output position
// (a, b), one mma.m16n8k16 per 8 classes (M = 16 pixels, K = 16 channels, N = 24 = classes padded).  The CUDA-core kernel
// (esn_head_convt2x2) spends 1216 FMAs per input pixel and is FP32-issue-bound at 0.09 of the HBM roofline; here the
// arithmetic is 24 MMAs per 16 pixels and the kernel streams.  The weights stay fp32-accurate: W = hi + lo in bf16, two
// MMAs per tile (the activations are bf16 either way).  The K slots of the MMA are a permutation of the channels chosen so
// that lane (g, t) needs channels 4t .. 4t+3 of pixel g (one 8-byte load: four lanes read a whole 32-byte pixel); the host
// packs the B fragments with the same permutation.  Two tiles are prefetched ahead of the one being multiplied.
struct HeadT2Args {
  const __nv_bfloat16* x;
  const uint32_t* wfrag;    // [2 hi/lo][4 positions][3 n-tiles][32 lanes][2]
  const float* bias;
  uint8_t* mask;            // (N, 2h, 2w)
  int W, x_cs, classes;
  long long tiles;
};

__global__ void __launch_bounds__(128, 4) head_convt2x2_mask_kernel(const HeadT2Args a) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  uint32_t bf[2][4][3][2];
#pragma unroll
  for (int hl = 0; hl < 2; ++hl)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(a.wfrag) + ((hl * 4 + p) * 3 + nt) * 32 + lane);
        bf[hl][p][nt][0] = v.x;
        bf[hl][p][nt][1] = v.y;
      }
  float binit[3][2];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cls = nt * 8 + 2 * t + e;
      binit[nt][e] = cls < a.classes ? (a.bias ? __ldg(a.bias + cls) : 0.f) : -INFINITY;
    }
  // rows g and g + 8 of a tile: channels 4t .. 4t+3 of pixels tile * 16 + g (+ 8)
  auto load = [&](long long tile, uint2& r0, uint2& r1) {
    const __nv_bfloat16* px = a.x + (size_t)(tile * 16 + g) * a.x_cs + 4 * t;
    r0 = __ldg(reinterpret_cast<const uint2*>(px));
    r1 = __ldg(reinterpret_cast<const uint2*>(px + (size_t)8 * a.x_cs));
  };
  const uint2 z2 = make_uint2(0u, 0u);
  long long tile = warp, nxt = warp + nwarps, nn = warp + 2 * nwarps;
  uint2 c0 = z2, c1 = z2, n0 = z2, n1 = z2;
  if (tile < a.tiles) load(tile, c0, c1);
  if (nxt < a.tiles) load(nxt, n0, n1);
  for (; tile < a.tiles; tile = nxt, nxt = nn, nn += nwarps) {
    uint2 f0 = z2, f1 = z2;
    if (nn < a.tiles) load(nn, f0, f1);
    const uint32_t af[4] = {c0.x, c1.x, c0.y, c1.y};     // (row g, slots 2t..), (row g+8, slots 2t..), (row g, 2t+8..), (row g+8, 2t+8..)
    float acc[4][3][4];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        acc[pos][nt][0] = acc[pos][nt][2] = binit[nt][0];
        acc[pos][nt][1] = acc[pos][nt][3] = binit[nt][1];
      }
#pragma unroll
    for (int pos = 0; pos < 4; ++pos)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        mma_bf16_16816(acc[pos][nt], af, bf[1][pos][nt][0], bf[1][pos][nt][1]);      // low parts first
        mma_bf16_16816(acc[pos][nt], af, bf[0][pos][nt][0], bf[0][pos][nt][1]);
      }
    int res[4][2];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos)
#pragma unroll
      for (int hs = 0; hs < 2; ++hs) {
        float bv = acc[pos][0][hs * 2];
        int bi = 2 * t;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (nt == 0 && e == 0) continue;
            const float v = acc[pos][nt][hs * 2 + e];
            if (v > bv) { bv = v; bi = nt * 8 + 2 * t + e; }
          }
#pragma unroll
        for (int off = 1; off < 4; off <<= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        res[pos][hs] = bi;
      }
    // lane t of a quad stores the two mask bytes of output row a = t >> 1 for pixel g + 8 * (t & 1)
    const int hs = t & 1, ra = t >> 1;
    const int s0 = hs ? res[0][1] : res[0][0], s1 = hs ? res[1][1] : res[1][0];
    const int s2 = hs ? res[2][1] : res[2][0], s3 = hs ? res[3][1] : res[3][0];
    const uchar2 out = make_uchar2((uint8_t)(ra ? s2 : s0), (uint8_t)(ra ? s3 : s1));
    const long long pix = tile * 16 + g + 8 * hs;
    const long long row = pix / a.W;                     // n * H + i
    const int j = (int)(pix - row * a.W);
    uint8_t* o = a.mask + ((size_t)(2 * row + ra) * (2 * a.W) + 2 * j);
    *reinterpret_cast<uchar2*>(o) = out;
    c0 = n0; c1 = n1; n0 = f0; n1 = f1;
  }
}

}  // namespace

extern "C" int esn_head_convt3x3s2_mask(const EsnHeadT3* p, void* stream) {
  if (!p || !p->wfrag || !p->mask || !esn_valid_nhwc(p->x)) return ESN_ERR_BAD_ARG;
  if (p->classes < 1 || p->classes > 24) return ESN_ERR_UNSUPPORTED;
  const EsnTensor& x = p->x;
  if (x.dtype != ESN_BF16 || x.c != 16 || x.w % 16 || x.c_stride % 2) return ESN_ERR_UNSUPPORTED;
  if (((uintptr_t)x.ptr % 4) || ((uintptr_t)p->wfrag % 8) || ((uintptr_t)p->mask % 2)) return ESN_ERR_ALIGN;
  HeadT3Args a;
  a.x = (const __nv_bfloat16*)x.ptr;
  a.wfrag = p->wfrag;
  a.bias = p->bias;
  a.mask = p->mask;
  a.N = x.n; a.H = x.h; a.W = x.w; a.x_cs = x.c_stride; a.classes = p->classes;
  a.tiles = (long long)x.n * x.h * (x.w / 16);
  long long ctas = (a.tiles + 3) / 4;                   // four warps per CTA
  if (ctas > 148 * 3 * 4) ctas = 148 * 3 * 4;           // grid-stride beyond ~4 waves of resident CTAs
  head_convt3x3s2_mask_kernel<<<(unsigned)ctas, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_head_convt2x2_mask(const EsnHeadT3* p, void* stream) {
  if (!p || !p->wfrag || !p->mask || !esn_valid_nhwc(p->x)) return ESN_ERR_BAD_ARG;
  if (p->classes < 1 || p->classes > 24) return ESN_ERR_UNSUPPORTED;
  const EsnTensor& x = p->x;
  if (x.dtype != ESN_BF16 || x.c != 16 || x.w % 16 || x.c_stride % 4) return ESN_ERR_UNSUPPORTED;
  if (((uintptr_t)x.ptr % 8) || ((uintptr_t)p->wfrag % 8) || ((uintptr_t)p->mask % 2)) return ESN_ERR_ALIGN;
  HeadT2Args a;
  a.x = (const __nv_bfloat16*)x.ptr;
  a.wfrag = p->wfrag;
  a.bias = p->bias;
  a.mask = p->mask;
  a.W = x.w; a.x_cs = x.c_stride; a.classes = p->classes;
  a.tiles = (long long)x.n * x.h * (x.w / 16);
  long long ctas = (a.tiles + 3) / 4;                   // four warps per CTA
  if (ctas > 148 * 4 * 2) ctas = 148 * 4 * 2;           // grid-stride beyond two waves of resident CTAs
  head_convt2x2_mask_kernel<<<(unsigned)ctas, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
