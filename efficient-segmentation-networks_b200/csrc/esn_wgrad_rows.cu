// Weight gradient of dense 3x3 convolutions (stride 1 or 2, any dilation) on the warp-level tensor-core path, all nine taps
// per pass (round 2):
//     dW[tap][ci][co] += sum_p  X[p + delta_tap][ci] * dY[p][co]
// The pixel index is the GEMM's K dimension, so in NHWC memory BOTH operands are "transposed" (channels contiguous).  The
// tcgen05 kernel (esn_wgrad_umma.cu) feeds them to the tensor core as MN-major descriptors, which B200 fetches at ~16 B/clk
// (measured): 222 us for DABNet's 32 -> 32 convs at 8 x 256 x 512, 10x the HBM time of the two tensors; the wmma
// kernel (esn_wgrad_mma.cu) that served stride 2 runs one tap per CTA and reads both tensors nine times.  Here
//   * a work unit is one output row segment of TW pixels; a CTA walks DOWN a column strip, so with vertical dilation 1 an
//     input row segment is staged ONCE (16-byte cp.async, zero fill = the conv's padding) into a ring of 3 + stride slots and
//     serves up to three output rows; the dY segment is double buffered;
//   * `ldmatrix.trans` delivers both operands in mma.sync's fragment layout straight from the [pixel][channel] tiles, so
//     nothing is transposed by threads; the 80-byte pixel pitch keeps the eight row addresses of a ldmatrix in different
//     banks, and for stride 2 even / odd input pixels are stored in two planes so a tap's 16 pixels are contiguous again;
//   * warp w owns filter tap w: a 32 (ci) x 32 (co) fp32 accumulator tile (2 x 4 mma.m16n8k16 per 16 pixels) that stays in
//     registers over the CTA's whole unit range and is added to dW once, with 16-byte atomics;
//   * larger channel counts are tiled 32 x 32 over blockIdx.y (x slices are disjoint, dY is re-read from L2 per ci tile).
// bf16 NHWC operands, channel counts multiples of 8 (the callers' zero-padded widths); anything else stays on the other
// kernels.  Replaces the weight branch of aten::convolution_backward for DABNet's 3x3 convs and down-samplers
// (DABNet.py:69-71,101-110,132-136; train.py:353).
#include "esn_common.cuh"

namespace {

constexpr int kWarps = 9;                 // one per filter tap
constexpr int kThreads = kWarps * 32;
constexpr int kCT = 32;                   // channel tile (ci and co)
constexpr int kPitch = 80;                // bytes per staged pixel: 64 of data + 16 of padding (bank spread for ldmatrix)

struct RowsArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  float* dw;
  int N, Hi, Wi, Cin, x_cs;
  int Ho, Wo, Cout, dy_cs;
  int stride, pad_h, pad_w, dil_h, dil_w;
  int tw;             // output pixels per unit (multiple of 16)
  int ntw;            // units per output row
  int npx;            // staged input pixels per row segment
  int plane;          // stride 2: pixels per parity plane (npx = 2 * plane); stride 1: unused
  int rows_per_cta;   // output rows of one column strip (n, wt) per CTA
  int chunks;         // CTAs per strip
  int nci;            // ci tiles
  int ring;           // x-row slots in shared memory: 3 + stride (vertical dilation 1: a row is staged ONCE and serves up to
                      // three output rows) or 6 (two independent sets of three)
  int reuse;          // 1: ring mode
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// staged position of input pixel q (relative to the segment start) in a row segment
template <int STRIDE>
__device__ __forceinline__ int xpos(int q, int plane) {
  return STRIDE == 1 ? q : (q & 1) * plane + (q >> 1);
}

template <int STRIDE>
__global__ void __launch_bounds__(kThreads, 2) wgrad_rows_kernel(const RowsArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = (blockIdx.y % a.nci) * kCT, co0 = (blockIdx.y / a.nci) * kCT;
  // CTA -> (column strip (n, wt), chunk of output rows [h0, h1)): consecutive units of a CTA are consecutive output rows
  const int strip = blockIdx.x / a.chunks, chunk = blockIdx.x % a.chunks;
  const int n = strip / a.ntw, wt = strip % a.ntw;
  const int h0 = chunk * a.rows_per_cta, h1 = min(a.Ho, h0 + a.rows_per_cta);
  const int wo0 = wt * a.tw;
  const int wi0 = wo0 * a.stride - a.pad_w;                   // input column of staged pixel 0
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const int xrow_bytes = a.npx * kPitch;
  const uint32_t dybase = sbase + a.ring * xrow_bytes;
  const int dy_bytes = a.tw * kPitch;
  // slot of tap row r of output row ho: ring mode -- input row hi = ho*stride - pad + r lives in slot (ho*stride + r) % ring, so
  // the rows shared with the previous output row are already there; otherwise two sets of three slots
  auto slot = [&](int ho, int r) { return a.reuse ? (ho * a.stride + r) % a.ring : ((ho - h0) & 1) * 3 + r; };

  // ---- loader: every thread issues 16-byte copies; thread -> (16-byte chunk of 8 channels, pixel lane): no divisions
  auto load_unit = [&](int ho, bool first) {
    const int ch = threadIdx.x & 3, pl = threadIdx.x >> 2;
    constexpr int kLanes = kThreads / 4;
    const int cx = ci0 + ch * 8;
    const bool cx_ok = cx < a.Cin;
    const int r_lo = (a.reuse && !first) ? 3 - a.stride : 0;  // rows not staged by the previous unit
    for (int r = r_lo; r < 3; ++r) {
      const int hi = ho * a.stride - a.pad_h + r * a.dil_h;
      const bool row_ok = cx_ok && hi >= 0 && hi < a.Hi;
      const __nv_bfloat16* rowp = a.x + (size_t)((size_t)n * a.Hi + (row_ok ? hi : 0)) * a.Wi * a.x_cs + cx;
      const uint32_t dst = sbase + slot(ho, r) * xrow_bytes + ch * 16;
      for (int q = pl; q < a.npx; q += kLanes) {
        const int wi = wi0 + q;
        const bool ok = row_ok && wi >= 0 && wi < a.Wi;
        cp_async16(dst + xpos<STRIDE>(q, a.plane) * kPitch, ok ? rowp + (size_t)wi * a.x_cs : a.x, ok);
      }
    }
    const uint32_t sd = dybase + ((ho - h0) & 1) * dy_bytes + ch * 16;
    const int cd = co0 + ch * 8;
    const bool cd_ok = cd < a.Cout;
    const __nv_bfloat16* dyp = a.dy + (size_t)((size_t)n * a.Ho + ho) * a.Wo * a.dy_cs + cd;
    for (int j = pl; j < a.tw; j += kLanes) {
      const int wo = wo0 + j;
      const bool ok = cd_ok && wo < a.Wo;
      cp_async16(sd + j * kPitch, ok ? dyp + (size_t)wo * a.dy_cs : a.dy, ok);
    }
  };

  float acc[2][4][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;

  // tap of this warp and its lane-constant ldmatrix offsets
  const int tr = warp / 3, ts = warp % 3;
  const int mi = lane >> 3, rr = lane & 7;
  // A (x^T): matrix mi -> k half (mi >> 1), m half (mi & 1); lane row rr = pixel k within the half
  const int a_k = (mi >> 1) * 8 + rr, a_c = (mi & 1) * 8;
  // B (dy): matrix mi -> k half (mi & 1), n tile of the pair (mi >> 1)
  const int b_k = (mi & 1) * 8 + rr, b_c = (mi >> 1) * 8;

  if (h0 < h1) load_unit(h0, true);
  cp_async_commit();
  for (int ho = h0; ho < h1; ++ho) {
    if (ho + 1 < h1) load_unit(ho + 1, false);      // into slots the current output row does not read
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sx = sbase + slot(ho, tr) * xrow_bytes;
    const uint32_t sd = dybase + ((ho - h0) & 1) * dy_bytes;
    const int ksteps = a.tw >> 4;
    // fragments of k-step ks + 1 are loaded before the MMAs of k-step ks are issued (two register sets, loop unrolled by two)
    auto ldfr = [&](int ks, uint32_t (&af)[2][4], uint32_t (&bf)[4][2]) {
      const int j = ks * 16;                                  // first output pixel of the k-step
      const int q = (j + a_k) * STRIDE + ts * a.dil_w;        // input pixel of output pixel j + a_k for this tap
      const uint32_t xa = sx + xpos<STRIDE>(q, a.plane) * kPitch + a_c * 2;
      ldsm_x4_trans(xa, af[0][0], af[0][1], af[0][2], af[0][3]);            // ci 0..15 of the tile
      ldsm_x4_trans(xa + 32, af[1][0], af[1][1], af[1][2], af[1][3]);       // ci 16..31
      const uint32_t da = sd + (j + b_k) * kPitch + b_c * 2;
      ldsm_x4_trans(da, bf[0][0], bf[0][1], bf[1][0], bf[1][1]);            // co 0..15
      ldsm_x4_trans(da + 32, bf[2][0], bf[2][1], bf[3][0], bf[3][1]);       // co 16..31
    };
    auto domma = [&](const uint32_t (&af)[2][4], const uint32_t (&bf)[4][2]) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[m][nt], af[m][0], af[m][1], af[m][2], af[m][3], bf[nt][0], bf[nt][1]);
    };
    uint32_t afA[2][4], bfA[4][2], afB[2][4], bfB[4][2];
    ldfr(0, afA, bfA);
    for (int ks = 0; ks < ksteps; ks += 2) {
      const bool two = ks + 1 < ksteps;
      if (two) ldfr(ks + 1, afB, bfB);
      domma(afA, bfA);
      if (two) {
        if (ks + 2 < ksteps) ldfr(ks + 2, afA, bfA);
        domma(afB, bfB);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // ---- dW[tap][ci][co] += acc.  Fragment: c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g + 8, same cols).  Lanes t and t^1
  // swap halves so that the even lane holds four consecutive co of row g and the odd lane four of row g + 8: one 16-byte
  // atomic per lane and fragment instead of two 8-byte ones
  const int g = lane >> 2, t = lane & 3;
  float* dwt = a.dw + (size_t)warp * a.Cin * a.Cout;
  const bool quad_ok = (a.Cout & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.dw) & 15) == 0);
  const bool odd = t & 1;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float* c = acc[m][nt];
      const float s0 = odd ? c[0] : c[2], s1 = odd ? c[1] : c[3];
      const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
      const float4 v = odd ? make_float4(r0, r1, c[2], c[3]) : make_float4(c[0], c[1], r0, r1);
      const int ci = ci0 + m * 16 + g + (odd ? 8 : 0);
      const int co = co0 + nt * 8 + 2 * (t & 2);
      if (ci >= a.Cin) continue;
      float* dst = dwt + (size_t)ci * a.Cout + co;
      if (quad_ok && co + 3 < a.Cout) {
        atomicAdd(reinterpret_cast<float4*>(dst), v);
      } else {
        if (co < a.Cout) atomicAdd(dst, v.x);
        if (co + 1 < a.Cout) atomicAdd(dst + 1, v.y);
        if (co + 2 < a.Cout) atomicAdd(dst + 2, v.z);
        if (co + 3 < a.Cout) atomicAdd(dst + 3, v.w);
      }
    }
}

int g_smem_set[2][kEsnMaxDevices];

}  // namespace

// called by esn_conv2d_wgrad; returns false when the problem is not this kernel's
bool esn_wgrad_rows_try(const EsnConv* p, void* stream, int* rc) {
  static const int mode = getenv("ESN_WGRAD_ROWS") ? atoi(getenv("ESN_WGRAD_ROWS")) : 1;     // 0: off, 2: every 3x3 conv it can take
  if (mode == 0) return false;
  const EsnTensor& x = p->x;
  const EsnTensor& dy = p->y;
  if (x.layout != ESN_NHWC || x.dtype != ESN_BF16 || dy.dtype != ESN_BF16 || p->groups != 1 || p->transposed) return false;
  if (p->kh != 3 || p->kw != 3 || (p->stride != 1 && p->stride != 2)) return false;
  if (x.c % 8 || dy.c % 8 || x.c_stride % 8 || dy.c_stride % 8 || (reinterpret_cast<uintptr_t>(x.ptr) & 15) ||
      (reinterpret_cast<uintptr_t>(dy.ptr) & 15))
    return false;
  if (x.c > 256 || dy.c > 256 || dy.w < 16) return false;
  // measured on B200 (DABNet shapes): stride 1 with >= 128 input channels is faster on the tcgen05 kernel (43 vs 55 us for
  // 128 -> 64 at 8 x 64 x 128: eight 32 x 32 tiles re-read dY four times); everything else is 2-4x faster here
  if (mode == 1 && p->stride == 1 && x.c > 64) return false;
  RowsArgs a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy.ptr);
  a.dw = reinterpret_cast<float*>(const_cast<void*>(p->w));
  a.N = x.n; a.Hi = x.h; a.Wi = x.w; a.Cin = x.c; a.x_cs = x.c_stride;
  a.Ho = dy.h; a.Wo = dy.w; a.Cout = dy.c; a.dy_cs = dy.c_stride;
  a.stride = p->stride; a.pad_h = p->pad_h; a.pad_w = p->pad_w; a.dil_h = p->dil_h; a.dil_w = p->dil_w;
  static const int tw_env = getenv("ESN_WGRAD_ROWS_TW") ? atoi(getenv("ESN_WGRAD_ROWS_TW")) : 128;
  int tw_want = tw_env;
  if (p->stride == 2 && tw_want > 64) tw_want = 64;                    // measured: stride 1 128 > 64 > 32 (fixed cost per unit); stride 2 64 > 128 (two CTAs per SM)
  a.tw = dy.w >= tw_want ? tw_want : (dy.w + 15) / 16 * 16;
  a.ntw = esn_cdiv(dy.w, a.tw);
  int npx = (a.tw - 1) * p->stride + 2 * p->dil_w + 1;
  if (p->stride == 2) {
    a.plane = (npx + 1) / 2;
    npx = 2 * a.plane;
  } else {
    a.plane = 0;
  }
  a.npx = npx;
  a.reuse = p->dil_h == 1 ? 1 : 0;
  a.ring = a.reuse ? 3 + p->stride : 6;
  const int smem = (a.ring * npx + 2 * a.tw) * kPitch;
  if (smem > 200 * 1024) return false;
  a.nci = esn_cdiv(x.c, kCT);
  const int tiles = a.nci * esn_cdiv(dy.c, kCT);
  const int strips = dy.n * a.ntw;
  const int dev = esn_current_device();
  const int si = p->stride - 1;
  if (g_smem_set[si][dev] < smem) {
    const cudaError_t e = si == 0 ? cudaFuncSetAttribute(wgrad_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                                  : cudaFuncSetAttribute(wgrad_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    g_smem_set[si][dev] = 200 * 1024;
  }
  // CTAs: ~ (CTAs that fit per SM, at most 2 by registers) x 148 in total; a CTA owns >= 8 output rows of one column strip
  // (the first row of a CTA stages three input rows, every further one `stride`)
  int per_sm = (220 * 1024) / (smem + 1024);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int want = (148 * per_sm + tiles - 1) / tiles;              // CTAs along x
  int chunks = want / strips;                                 // rounded down: one full wave rather than one and a bit
  if (chunks > dy.h / 8) chunks = dy.h / 8;
  if (chunks < 1) chunks = 1;
  a.rows_per_cta = esn_cdiv(dy.h, chunks);
  a.chunks = esn_cdiv(dy.h, a.rows_per_cta);
  const int gx = strips * a.chunks;
  dim3 grid(gx, tiles);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (si == 0) wgrad_rows_kernel<1><<<grid, kThreads, smem, st>>>(a);
  else wgrad_rows_kernel<2><<<grid, kThreads, smem, st>>>(a);
  g_esn_launches.fetch_add(1, std::memory_order_relaxed);
  *rc = (cudaPeekAtLastError() == cudaSuccess) ? ESN_OK : ESN_ERR_CUDA;
  if (*rc != ESN_OK) cudaGetLastError();
  return true;
}
