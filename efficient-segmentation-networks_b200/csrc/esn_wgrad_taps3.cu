// Weight gradient of dense three-tap convolutions (3x1 / 1x3, stride 1, any dilation: ERFNet's factorized convs,
// ERFNet.py:34-42) on the warp-level tensor-core path, all three taps per pass -- the three-tap sibling of esn_wgrad_rows.cu:
//     dW[tap][ci][co] += sum_p  X[p + delta_tap][ci] * dY[p][co]
// ERFNet's 68 such weight gradients per training step ran on the tcgen05 kernel with MN-major operands (esn_wgrad_umma.cu,
// operand-fetch-bound: 220 us for a 16 -> 16 conv at 8 x 256 x 512, 0.3 TB/s; 3.7 ms of a 12.8 ms step).  Same scheme as the
// 3x3 kernel -- row segments staged with 16-byte cp.async (zero fill = padding), `ldmatrix.trans` fragments straight from the
// [pixel][channel] tiles, fp32 accumulators in registers over the CTA's whole row range, 16-byte atomics at the end -- with
//   * warp -> (tap, 32 x 32 channel tile): a CTA covers NCI x NCO tiles (64 -> 64: 2 x 2, twelve warps; 128 -> 128: 1 x 4 with
//     the four ci tiles over blockIdx.y; 16 -> 16: one tile, three warps);
//   * 3x1: the taps are three staged input rows (a ring of four slots when the vertical dilation is 1, else two sets of
//     three); 1x3: ONE staged row, the taps are pixel offsets of 0 / d / 2d into it.
#include "esn_common.cuh"

namespace {

struct Taps3Args {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  float* dw;
  int N, Hi, Wi, Cin, x_cs;
  int Ho, Wo, Cout, dy_cs;
  int pad_h, pad_w, dil_h, dil_w;
  int tw, ntw, npx;
  int rows_per_cta, chunks;
  int ncig;           // ci tile groups (of NCI tiles) over blockIdx.y
  int ring, reuse;
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


template <bool VERT, int NCI, int NCO>
__global__ void __launch_bounds__(3 * NCI * NCO * 32) wgrad_taps3_kernel(const Taps3Args a) {
  constexpr int kThreads = 3 * NCI * NCO * 32;
  constexpr int PX = NCI * 64 + 16, PD = NCO * 64 + 16;      // bytes per staged pixel (16 of padding: bank spread for ldmatrix)
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = warp % 3, tile = warp / 3;
  const int tci = tile % NCI, tco = tile / NCI;
  const int ci0 = (blockIdx.y % a.ncig) * NCI * 32, co0 = (blockIdx.y / a.ncig) * NCO * 32;
  const int strip = blockIdx.x / a.chunks, chunk = blockIdx.x % a.chunks;
  const int n = strip / a.ntw, wt = strip % a.ntw;
  const int h0 = chunk * a.rows_per_cta, h1 = min(a.Ho, h0 + a.rows_per_cta);
  const int wo0 = wt * a.tw;
  const int wi0 = wo0 - a.pad_w;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const int xrow_bytes = a.npx * PX;
  const uint32_t dybase = sbase + a.ring * xrow_bytes;
  const int dy_bytes = a.tw * PD;
  constexpr int NR = VERT ? 3 : 1;                            // staged input rows per output row
  auto slot = [&](int ho, int r) {
    if (!VERT) return (ho - h0) & 1;
    return a.reuse ? (ho + r) % a.ring : ((ho - h0) & 1) * 3 + r;
  };

  auto load_unit = [&](int ho, bool first) {
    {
      constexpr int CPP = NCI * 4;                            // 16-byte chunks per staged x pixel
      constexpr int kLanes = kThreads / CPP;
      const int ch = threadIdx.x % CPP, pl = threadIdx.x / CPP;
      const int cx = ci0 + ch * 8;
      const bool cx_ok = cx < a.Cin;
      const int r_lo = (VERT && a.reuse && !first) ? 2 : 0;
      for (int r = r_lo; r < NR; ++r) {
        const int hi = ho - a.pad_h + r * a.dil_h;
        const bool row_ok = cx_ok && hi >= 0 && hi < a.Hi;
        const __nv_bfloat16* rowp = a.x + (size_t)((size_t)n * a.Hi + (row_ok ? hi : 0)) * a.Wi * a.x_cs + cx;
        const uint32_t dst = sbase + slot(ho, r) * xrow_bytes + ch * 16;
        if (pl < kLanes)
          for (int q = pl; q < a.npx; q += kLanes) {
            const int wi = wi0 + q;
            const bool ok = row_ok && wi >= 0 && wi < a.Wi;
            cp_async16(dst + q * PX, ok ? rowp + (size_t)wi * a.x_cs : a.x, ok);
          }
      }
    }
    {
      constexpr int CPP = NCO * 4;
      constexpr int kLanes = kThreads / CPP;
      const int ch = threadIdx.x % CPP, pl = threadIdx.x / CPP;
      const uint32_t sd = dybase + ((ho - h0) & 1) * dy_bytes + ch * 16;
      const int cd = co0 + ch * 8;
      const bool cd_ok = cd < a.Cout;
      const __nv_bfloat16* dyp = a.dy + (size_t)((size_t)n * a.Ho + ho) * a.Wo * a.dy_cs + cd;
      if (pl < kLanes)
        for (int j = pl; j < a.tw; j += kLanes) {
          const int wo = wo0 + j;
          const bool ok = cd_ok && wo < a.Wo;
          cp_async16(sd + j * PD, ok ? dyp + (size_t)wo * a.dy_cs : a.dy, ok);
        }
    }
  };

  float acc[2][4][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;

  const int mi = lane >> 3, rr = lane & 7;
  const int a_k = (mi >> 1) * 8 + rr, a_c = (mi & 1) * 8;     // A (x^T): matrix mi -> k half (mi >> 1), m half (mi & 1)
  const int b_k = (mi & 1) * 8 + rr, b_c = (mi >> 1) * 8;     // B (dy): matrix mi -> k half (mi & 1), n tile of the pair (mi >> 1)
  const int coloff = VERT ? 0 : tap * a.dil_w;                // 1x3: the tap is a pixel offset into the one staged row

  if (h0 < h1) load_unit(h0, true);
  cp_async_commit();
  for (int ho = h0; ho < h1; ++ho) {
    if (ho + 1 < h1) load_unit(ho + 1, false);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sx = sbase + slot(ho, VERT ? tap : 0) * xrow_bytes + (tci * 32 + a_c) * 2;
    const uint32_t sd = dybase + ((ho - h0) & 1) * dy_bytes + (tco * 32 + b_c) * 2;
    const int ksteps = a.tw >> 4;
    for (int ks = 0; ks < ksteps; ++ks) {
      const int j = ks * 16;
      const uint32_t xa = sx + (j + a_k + coloff) * PX;
      uint32_t af[2][4], bf[4][2];
      ldsm_x4_trans(xa, af[0][0], af[0][1], af[0][2], af[0][3]);
      ldsm_x4_trans(xa + 32, af[1][0], af[1][1], af[1][2], af[1][3]);
      const uint32_t da = sd + (j + b_k) * PD;
      ldsm_x4_trans(da, bf[0][0], bf[0][1], bf[1][0], bf[1][1]);
      ldsm_x4_trans(da + 32, bf[2][0], bf[2][1], bf[3][0], bf[3][1]);
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[m][nt], af[m][0], af[m][1], af[m][2], af[m][3], bf[nt][0], bf[nt][1]);
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // dW[tap][ci][co] += acc; lanes t / t^1 swap halves so every lane issues one 16-byte atomic per fragment (esn_wgrad_rows.cu)
  const int g = lane >> 2, t = lane & 3;
  float* dwt = a.dw + (size_t)tap * a.Cin * a.Cout;
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
      const int ci = ci0 + tci * 32 + m * 16 + g + (odd ? 8 : 0);
      const int co = co0 + tco * 32 + nt * 8 + 2 * (t & 2);
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

template <bool VERT, int NCI, int NCO>
bool launch_taps3(Taps3Args a, const EsnConv* p, cudaStream_t st, int* rc) {
  constexpr int kThreads = 3 * NCI * NCO * 32;
  constexpr int PX = NCI * 64 + 16, PD = NCO * 64 + 16;
  const int Wo = a.Wo;
  a.reuse = (VERT && p->dil_h == 1) ? 1 : 0;
  a.ring = VERT ? (a.reuse ? 4 : 6) : 2;
  int smem = 0;
  for (int tw = 128; tw >= 32; tw >>= 1) {                    // the widest row segment that leaves two CTAs per SM
    a.tw = Wo >= tw ? tw : (Wo + 15) / 16 * 16;
    a.npx = a.tw + (VERT ? 0 : 2 * p->dil_w);
    smem = a.ring * a.npx * PX + 2 * a.tw * PD;
    if (smem <= 100 * 1024 || tw == 32) break;
  }
  if (smem > 200 * 1024) return false;
  a.ntw = esn_cdiv(Wo, a.tw);
  a.ncig = esn_cdiv(a.Cin, NCI * 32);
  const int tiles = a.ncig * esn_cdiv(a.Cout, NCO * 32);
  const int strips = a.N * a.ntw;
  static int set[kEsnMaxDevices];
  const int dev = esn_current_device();
  if (set[dev] < smem) {
    if (cudaFuncSetAttribute(wgrad_taps3_kernel<VERT, NCI, NCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    set[dev] = 200 * 1024;
  }
  int per_sm = (220 * 1024) / (smem + 1024);
  const int by_threads = 2048 / kThreads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  const int want = (148 * per_sm + tiles - 1) / tiles;
  int chunks = want / strips;
  if (chunks > a.Ho / 8) chunks = a.Ho / 8;
  if (chunks < 1) chunks = 1;
  a.rows_per_cta = esn_cdiv(a.Ho, chunks);
  a.chunks = esn_cdiv(a.Ho, a.rows_per_cta);
  dim3 grid(strips * a.chunks, tiles);
  wgrad_taps3_kernel<VERT, NCI, NCO><<<grid, kThreads, smem, st>>>(a);
  g_esn_launches.fetch_add(1, std::memory_order_relaxed);
  *rc = (cudaPeekAtLastError() == cudaSuccess) ? ESN_OK : ESN_ERR_CUDA;
  if (*rc != ESN_OK) cudaGetLastError();
  return true;
}

}  // namespace

// called by esn_conv2d_wgrad; returns false when the problem is not this kernel's
bool esn_wgrad_taps3_try(const EsnConv* p, void* stream, int* rc) {
  static const int mode = getenv("ESN_WGRAD_TAPS3") ? atoi(getenv("ESN_WGRAD_TAPS3")) : 1;     // 0: off, 2: also >= 128 channels
  if (mode == 0) return false;
  const EsnTensor& x = p->x;
  const EsnTensor& dy = p->y;
  if (x.layout != ESN_NHWC || x.dtype != ESN_BF16 || dy.dtype != ESN_BF16 || p->groups != 1 || p->transposed || p->stride != 1)
    return false;
  const bool vert = p->kh == 3 && p->kw == 1, horz = p->kh == 1 && p->kw == 3;
  if (!vert && !horz) return false;
  if (x.c % 8 || dy.c % 8 || x.c_stride % 8 || dy.c_stride % 8 || (reinterpret_cast<uintptr_t>(x.ptr) & 15) ||
      (reinterpret_cast<uintptr_t>(dy.ptr) & 15))
    return false;
  if (x.c > 256 || dy.c > 256 || dy.w < 16 || p->dil_w > 32 || p->dil_h > 64) return false;
  if (mode == 1 && (x.c > 64 || dy.c > 64)) return false;
  Taps3Args a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy.ptr);
  a.dw = reinterpret_cast<float*>(const_cast<void*>(p->w));
  a.N = x.n; a.Hi = x.h; a.Wi = x.w; a.Cin = x.c; a.x_cs = x.c_stride;
  a.Ho = dy.h; a.Wo = dy.w; a.Cout = dy.c; a.dy_cs = dy.c_stride;
  a.pad_h = p->pad_h; a.pad_w = p->pad_w; a.dil_h = p->dil_h; a.dil_w = p->dil_w;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool small = x.c <= 32 && dy.c <= 32, mid = x.c <= 64 && dy.c <= 64;
  if (vert) {
    if (small) return launch_taps3<true, 1, 1>(a, p, st, rc);
    if (mid) return launch_taps3<true, 2, 2>(a, p, st, rc);
    return launch_taps3<true, 1, 4>(a, p, st, rc);
  }
  if (small) return launch_taps3<false, 1, 1>(a, p, st, rc);
  if (mid) return launch_taps3<false, 2, 2>(a, p, st, rc);
  return launch_taps3<false, 1, 4>(a, p, st, rc);
}
