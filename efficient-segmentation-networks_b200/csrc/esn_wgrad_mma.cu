// Tensor-core weight gradient (bf16 operands, fp32 accumulate) for dense convolutions:
//     dW[tap][ci][co] += sum_p  X[p + delta_tap, ci] * dY[p, co]
// i.e. D(co x ci) = dY^T (co x P) . X (P x ci) with the pixel index P as the GEMM K dimension.
// Both operands are "MN-major" in NHWC memory (channels contiguous, pixels strided), which the
// warp-level mma path reads directly with layout tags (matrix_a col_major, matrix_b row_major) from a
// shared-memory staging tile, so no transposition is needed.  HBM-bound (reads |x| + |dy| once per
// tap group from L2/HBM); round 2 replaces this with a tcgen05 MN-major descriptor kernel.
#include <mma.h>

#include "esn_common.cuh"

namespace {

using namespace nvcuda;

struct WgMmaArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  float* dw;
  int N, Hi, Wi, Cin, x_cs;
  int Ho, Wo, Cout, dy_cs;
  int kh, kw, stride, pad_h, pad_w, dil_h, dil_w;
  long long px_per_cta;
};

constexpr int kKT = 64;       // pixels per smem stage
constexpr int kWgThreads = 256;

// CO_T x CI_T output tile per CTA (multiples of 16), one filter tap per blockIdx.y
template <int CO_T, int CI_T>
__global__ void __launch_bounds__(kWgThreads) wgrad_mma_kernel(const WgMmaArgs a) {
  constexpr int FR = CO_T / 16, FC = CI_T / 16, NF = FR * FC;
  constexpr int PER_WARP = (NF + 7) / 8;
  __shared__ __align__(32) __nv_bfloat16 sA[kKT][CO_T + 8];   // dY tile  [pixel][co]
  __shared__ __align__(32) __nv_bfloat16 sB[kKT][CI_T + 8];   // X tile   [pixel][ci]
  const int tap = blockIdx.y;
  const int r = tap / a.kw, s = tap % a.kw;
  const int nci = (a.Cin + CI_T - 1) / CI_T;
  const int ci0 = (blockIdx.z % nci) * CI_T, co0 = (blockIdx.z / nci) * CO_T;
  const int warp = threadIdx.x >> 5;
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[PER_WARP];
#pragma unroll
  for (int f = 0; f < PER_WARP; ++f) wmma::fill_fragment(acc[f], 0.f);

  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(M, p0 + a.px_per_cta);
  const bool va = (a.dy_cs % 8 == 0) && (co0 % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.dy) & 15) == 0);
  const bool vb = (a.x_cs % 8 == 0) && (ci0 % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
  // Software pipeline: the global loads of pixel block k+1 are issued into registers before the MMAs of block k,
  // so their latency hides behind the tensor work (one __syncthreads pair per block as before).
  constexpr int EA = (kKT * (CO_T / 8) + kWgThreads - 1) / kWgThreads;
  constexpr int EB = (kKT * (CI_T / 8) + kWgThreads - 1) / kWgThreads;
  uint4 ra[EA], rb[EB];
  auto fetch = [&](const long long pb) {
#pragma unroll
    for (int i = 0; i < EA; ++i) {
      const int e = threadIdx.x + i * kWgThreads;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (e < kKT * (CO_T / 8)) {
        const int pp = e / (CO_T / 8), c8 = (e % (CO_T / 8)) * 8;
        const long long p = pb + pp;
        if (p < p1) {
          const __nv_bfloat16* src = a.dy + p * a.dy_cs + co0 + c8;
          if (va && co0 + c8 + 8 <= a.Cout) {
            v = __ldg(reinterpret_cast<const uint4*>(src));
          } else {
            __nv_bfloat16 t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = (co0 + c8 + j < a.Cout) ? src[j] : __float2bfloat16(0.f);
            v = *reinterpret_cast<uint4*>(t);
          }
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < EB; ++i) {
      const int e = threadIdx.x + i * kWgThreads;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (e < kKT * (CI_T / 8)) {
        const int pp = e / (CI_T / 8), c8 = (e % (CI_T / 8)) * 8;
        const long long p = pb + pp;
        if (p < p1) {
          const int wo = (int)(p % a.Wo);
          const int ho = (int)((p / a.Wo) % a.Ho);
          const int n = (int)(p / ((long long)a.Wo * a.Ho));
          const int hi = ho * a.stride - a.pad_h + r * a.dil_h, wi = wo * a.stride - a.pad_w + s * a.dil_w;
          if (hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) {
            const __nv_bfloat16* src = a.x + ((size_t)((size_t)n * a.Hi + hi) * a.Wi + wi) * a.x_cs + ci0 + c8;
            if (vb && ci0 + c8 + 8 <= a.Cin) {
              v = __ldg(reinterpret_cast<const uint4*>(src));
            } else {
              __nv_bfloat16 t[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) t[j] = (ci0 + c8 + j < a.Cin) ? src[j] : __float2bfloat16(0.f);
              v = *reinterpret_cast<uint4*>(t);
            }
          }
        }
      }
      rb[i] = v;
    }
  };
  if (p0 < p1) fetch(p0);
  for (long long pb = p0; pb < p1; pb += kKT) {
#pragma unroll
    for (int i = 0; i < EA; ++i) {
      const int e = threadIdx.x + i * kWgThreads;
      if (e < kKT * (CO_T / 8)) *reinterpret_cast<uint4*>(&sA[e / (CO_T / 8)][(e % (CO_T / 8)) * 8]) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < EB; ++i) {
      const int e = threadIdx.x + i * kWgThreads;
      if (e < kKT * (CI_T / 8)) *reinterpret_cast<uint4*>(&sB[e / (CI_T / 8)][(e % (CI_T / 8)) * 8]) = rb[i];
    }
    __syncthreads();
    if (pb + kKT < p1) fetch(pb + kKT);
#pragma unroll
    for (int kk = 0; kk < kKT; kk += 16) {
#pragma unroll
      for (int f = 0; f < PER_WARP; ++f) {
        const int fi = warp * PER_WARP + f;
        if (fi < NF) {
          const int fr = fi / FC, fc = fi % FC;
          wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> af;   // A(co, p) = sA[p][co]
          wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> bf;   // B(p, ci) = sB[p][ci]
          wmma::load_matrix_sync(af, &sA[kk][fr * 16], CO_T + 8);
          wmma::load_matrix_sync(bf, &sB[kk][fc * 16], CI_T + 8);
          wmma::mma_sync(acc[f], af, bf, acc[f]);
        }
      }
    }
    __syncthreads();
  }
  // ---- accumulate the tile into dW[tap][ci][co] (fp32 atomics), through a per-warp smem patch
  __shared__ __align__(32) float patch[8][16][20];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int f = 0; f < PER_WARP; ++f) {
    const int fi = warp * PER_WARP + f;
    if (fi < NF) {
      const int fr = fi / FC, fc = fi % FC;
      wmma::store_matrix_sync(&patch[warp][0][0], acc[f], 20, wmma::mem_row_major);   // patch[co][ci]
      __syncwarp();
      for (int e = lane; e < 256; e += 32) {
        const int ci_l = e / 16, co_l = e % 16;          // consecutive lanes -> consecutive co (coalesced)
        const int co = co0 + fr * 16 + co_l, ci = ci0 + fc * 16 + ci_l;
        const float v = patch[warp][co_l][ci_l];
        if (co < a.Cout && ci < a.Cin && v != 0.f) atomicAdd(a.dw + ((size_t)tap * a.Cin + ci) * a.Cout + co, v);
      }
      __syncwarp();
    }
  }
}

template <int CO_T, int CI_T>
void launch(const WgMmaArgs& a, int taps, long long M, cudaStream_t st, long long (*pick)(long long, int)) {
  WgMmaArgs b = a;
  const int tiles = ((a.Cin + CI_T - 1) / CI_T) * ((a.Cout + CO_T - 1) / CO_T);
  long long chunk = pick(M, taps * tiles);
  chunk = (chunk + kKT - 1) / kKT * kKT;
  b.px_per_cta = chunk;
  dim3 grid((unsigned)((M + chunk - 1) / chunk), taps, tiles);
  wgrad_mma_kernel<CO_T, CI_T><<<grid, kWgThreads, 0, st>>>(b);
}

long long pick_chunk_mma(long long M, int other_ctas) {
  long long want = (2LL * 148 + other_ctas - 1) / other_ctas;
  if (want < 1) want = 1;
  long long chunk = (M + want - 1) / want;
  if (chunk < 512) chunk = 512;
  return chunk;
}

}  // namespace

// called by esn_conv2d_wgrad for bf16 x / bf16 dy dense convs; returns false when not applicable
bool esn_wgrad_mma_try(const EsnConv* p, void* stream, int* rc) {
  const EsnTensor& x = p->x;
  const EsnTensor& dy = p->y;
  if (x.layout != ESN_NHWC || x.dtype != ESN_BF16 || dy.dtype != ESN_BF16 || p->groups != 1) return false;
  if (x.c < 8 || dy.c < 8) return false;
  WgMmaArgs a;
  a.x = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy.ptr);
  a.dw = reinterpret_cast<float*>(const_cast<void*>(p->w));
  a.N = x.n; a.Hi = x.h; a.Wi = x.w; a.Cin = x.c; a.x_cs = x.c_stride;
  a.Ho = dy.h; a.Wo = dy.w; a.Cout = dy.c; a.dy_cs = dy.c_stride;
  a.kh = p->kh; a.kw = p->kw; a.stride = p->stride; a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.dil_h = p->dil_h; a.dil_w = p->dil_w;
  a.px_per_cta = 0;
  const long long M = (long long)dy.n * dy.h * dy.w;
  const int taps = p->kh * p->kw;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool co_big = dy.c > 64, ci_big = x.c > 64;
  const bool co_small = dy.c <= 32, ci_small = x.c <= 32;
  if (co_big && ci_big) launch<128, 128>(a, taps, M, st, pick_chunk_mma);
  else if (co_big) launch<128, 64>(a, taps, M, st, pick_chunk_mma);
  else if (ci_big) launch<64, 128>(a, taps, M, st, pick_chunk_mma);
  else if (co_small && ci_small) launch<32, 32>(a, taps, M, st, pick_chunk_mma);
  else launch<64, 64>(a, taps, M, st, pick_chunk_mma);
  g_esn_launches.fetch_add(1, std::memory_order_relaxed);
  *rc = (cudaPeekAtLastError() == cudaSuccess) ? ESN_OK : ESN_ERR_CUDA;
  if (*rc != ESN_OK) cudaGetLastError();
  return true;
}
