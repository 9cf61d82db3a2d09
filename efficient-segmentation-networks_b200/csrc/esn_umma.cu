// tcgen05 implicit-GEMM convolution for sm_100a.
//
// A convolution is computed as a sum over filter taps of shifted GEMMs
//     Y[p, :] = sum_t  X[p + delta_t, :] . W_t^T          p = output position, X/Y NHWC bf16
// * M tile   = 128 output positions (a bh x bw patch of one image), one TMEM lane each
// * K blocks = min(Cin,64) channels of one tap; the shifted X patch is fetched by ONE 5-D TMA box
//              load per (tap, K block); out-of-image coordinates are zero-filled by TMA, which is
//              exactly the conv's zero padding (no halo code, no predicates)
// * N        = Cout (padded to 16) <= 256; all taps' weights stay resident in shared memory for the
//              lifetime of the persistent CTA
// * fp32 accumulators live in TMEM, double buffered, so the epilogue of tile i overlaps the MMAs
//   of tile i+1.  Epilogue: scale/shift (bias + folded BatchNorm), residual, ReLU/PReLU, bf16.
// Stride-2 convs read X through a (pixel-pair, W/2, row-parity, H/2, N) view so every tap is still a
// dense box; stride-2 transposed convs run as 4 output-parity phases of stride-1 taps.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include <cuda.h>

#include <mutex>

#include "esn_common.cuh"

namespace {

constexpr int kMaxTaps = 9;
constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr long long kSpinLimitCycles = 4000000000LL;  // ~2 s

struct alignas(64) UmmaArgs {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int ntaps, nkb, kb_elems, N, cout;
  int bw, bh, tiles_w, tiles_h, ntiles, gh, gw;
  int tap_dx[kMaxTaps], tap_dy[kMaxTaps], tap_par[kMaxTaps], tap_coff[kMaxTaps], tap_wrow[kMaxTaps];
  __nv_bfloat16* y;
  int Hy, Wy, y_cs, sy, oy, sx, ox;
  EpiArgs ep;
  int stages;
  uint32_t stage_bytes, wblock_bytes, w_region_bytes;
  uint32_t idesc, desc_hi;  // instruction descriptor; upper 32 bits of the smem descriptors
  uint32_t tmem_cols;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > kSpinLimitCycles) __trap();  // a protocol bug must fail loudly, never hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (lower 32 bits): start address and LBO (unused for
// swizzled K-major layouts; 1 by convention).  Upper 32 bits (SBO, version, swizzle) are in desc_hi.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t desc_hi) {
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
  return ((uint64_t)desc_hi << 32) | lo;
}

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ UmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t w_base = base;
  const uint32_t a_base = base + a.w_region_bytes;
  const uint32_t bar_base = a_base + (uint32_t)a.stages * a.stage_bytes;
  const int S = a.stages;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * S);
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * S + 1 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * S + 3 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmA);
    tma_prefetch_desc(&a.tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(wfull_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kiters = a.ntaps * a.nkb;  // smem stages consumed per tile

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: weights once, then the A ring
      mbar_expect_tx(wfull_bar, (uint32_t)kiters * a.wblock_bytes);
      for (int t = 0; t < a.ntaps; ++t)
        for (int kb = 0; kb < a.nkb; ++kb)
          tma_load_2d(w_base + (uint32_t)(t * a.nkb + kb) * a.wblock_bytes, &a.tmB, wfull_bar, kb * a.kb_elems,
                      a.tap_wrow[t]);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int tw = tile % a.tiles_w;
        const int th = (tile / a.tiles_w) % a.tiles_h;
        const int n = tile / (a.tiles_w * a.tiles_h);
        for (int t = 0; t < a.ntaps; ++t) {
          const int cw = tw * a.bw + a.tap_dx[t];
          const int ch = th * a.bh + a.tap_dy[t];
          for (int kb = 0; kb < a.nkb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            mbar_expect_tx(full_bar(s), a.stage_bytes);
            tma_load_5d(a_base + (uint32_t)s * a.stage_bytes, &a.tmA, full_bar(s), a.tap_coff[t] + kb * a.kb_elems, cw,
                        a.tap_par[t], ch, n);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      const int ksteps = a.kb_elems / 16;
      uint32_t it = 0, tc = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tc) {
        const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)a.N;
        for (int ki = 0; ki < kiters; ++ki, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = a_base + (uint32_t)s * a.stage_bytes;
          const uint32_t b_addr = w_base + (uint32_t)ki * a.wblock_bytes;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = make_desc(a_addr + 32u * k, a.desc_hi);
            const uint64_t bd = make_desc(b_addr + 32u * k, a.desc_hi);
            umma_bf16(d_tmem, ad, bd, a.idesc, (ki | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
        }
        umma_commit(tfull_bar(acc));  // accumulator ready for the epilogue
      }
    }
  } else {
    // ---------------- epilogue warps: TMEM -> registers -> bf16 NHWC
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ri = row / a.bw, rj = row % a.bw;
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tc) {
      const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
      const int tw = tile % a.tiles_w;
      const int th = (tile / a.tiles_w) % a.tiles_h;
      const int n = tile / (a.tiles_w * a.tiles_h);
      const int gi = th * a.bh + ri, gj = tw * a.bw + rj;
      const bool valid = gi < a.gh && gj < a.gw;
      const size_t opix = ((size_t)n * a.Hy + (size_t)(gi * a.sy + a.oy)) * a.Wy + (size_t)(gj * a.sx + a.ox);
      __nv_bfloat16* yp = a.y + opix * a.y_cs;
      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)a.N;
      for (int c0 = 0; c0 < a.N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cb = c0 + 8 * h;
            if (cb < a.cout) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = min(cb + j, a.cout - 1);
                const float sc = a.ep.scale ? __ldg(a.ep.scale + c) : 1.f;
                const float sh = a.ep.shift ? __ldg(a.ep.shift + c) : 0.f;
                f[j] = __uint_as_float(r[8 * h + j]) * sc + sh;
              }
              const bool full8 = cb + 8 <= a.cout;
              if (a.ep.res) {
                const __nv_bfloat16* rp =
                    reinterpret_cast<const __nv_bfloat16*>(a.ep.res) + opix * a.ep.res_cstride + cb;
                if (full8) {
                  float g[8];
                  bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(rp)), g);
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] += g[j];
                } else {
                  for (int j = 0; j < 8; ++j)
                    if (cb + j < a.cout) f[j] += __bfloat162float(rp[j]);
                }
              }
              if (a.ep.act == ESN_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
              } else if (a.ep.act == ESN_ACT_PRELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float al = __ldg(a.ep.alpha + min(cb + j, a.cout - 1));
                  f[j] = f[j] >= 0.f ? f[j] : f[j] * al;
                }
              }
              if (full8) {
                *reinterpret_cast<uint4*>(yp + cb) = float_to_bf16x8(f);
              } else {
                for (int j = 0; j < 8; ++j)
                  if (cb + j < a.cout) yp[cb + j] = __float2bfloat16_rn(f[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int floordiv2(int t) { return (t >= 0) ? t / 2 : -((-t + 1) / 2); }

struct DeviceLimits {
  int sms = 0;
  int max_smem = 0;
};
const DeviceLimits& limits() {
  static DeviceLimits l;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&l.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&l.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, conv_umma_kernel) == cudaSuccess) l.max_smem -= (int)fa.sharedSizeBytes;
    cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, l.max_smem);
  });
  return l;
}

}  // namespace

extern "C" int esn_conv2d_umma(const EsnConv* p, void* stream) {
  if (!p || !p->w) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  if (!esn_valid_nhwc(x) || !esn_valid_nhwc(y)) return ESN_ERR_BAD_ARG;
  if (x.dtype != ESN_BF16 || y.dtype != ESN_BF16 || p->groups != 1) return ESN_ERR_UNSUPPORTED;
  if (p->ep.residual.ptr && p->ep.residual.dtype != ESN_BF16) return ESN_ERR_UNSUPPORTED;
  int rc = esn_check_epilogue(p->ep, y);
  if (rc) return rc;
  const int Cin = x.c, Cout = y.c, N = p->cout_pad;
  if (N % 16 || N < 16 || N > 256 || N < Cout) return ESN_ERR_UNSUPPORTED;
  int KB;
  if (Cin == 16 || Cin == 32 || Cin == 64)
    KB = Cin;
  else if (Cin % 64 == 0)
    KB = 64;
  else
    return ESN_ERR_UNSUPPORTED;
  const int nkb = Cin / KB;
  if (x.c_stride % 8 || y.c_stride % 8 || ((uintptr_t)x.ptr % 16) || ((uintptr_t)y.ptr % 16) ||
      ((uintptr_t)p->w % 16))
    return ESN_ERR_ALIGN;
  if (p->ep.residual.ptr && (p->ep.residual.c_stride % 8 || ((uintptr_t)p->ep.residual.ptr % 16))) return ESN_ERR_ALIGN;
  if (x.n != y.n) return ESN_ERR_BAD_SHAPE;
  const int ntaps_all = p->kh * p->kw;
  if (ntaps_all > kMaxTaps) return ESN_ERR_UNSUPPORTED;
  if (p->stride != 1 && p->stride != 2) return ESN_ERR_UNSUPPORTED;
  if (p->transposed && (p->stride != 2 || p->dil_h != 1 || p->dil_w != 1)) return ESN_ERR_UNSUPPORTED;
  if (!p->transposed) {
    const int eh = (x.h + 2 * p->pad_h - p->dil_h * (p->kh - 1) - 1) / p->stride + 1;
    const int ew = (x.w + 2 * p->pad_w - p->dil_w * (p->kw - 1) - 1) / p->stride + 1;
    if (eh != y.h || ew != y.w) return ESN_ERR_BAD_SHAPE;
    if (p->stride == 2 && ((x.h | x.w) & 1)) return ESN_ERR_UNSUPPORTED;
  } else {
    if (y.h != 2 * x.h || y.w != 2 * x.w) return ESN_ERR_UNSUPPORTED;
  }
  EncodeTiledFn encode = get_encode();
  if (!encode) return ESN_ERR_CUDA;
  const DeviceLimits& lim = limits();
  if (lim.sms <= 0) return ESN_ERR_CUDA;

  UmmaArgs a;
  memset(&a, 0, sizeof(a));
  const CUtensorMapSwizzle swz = KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (KB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const uint32_t row_bytes = KB * 2;
  const uint32_t layout_type = KB == 64 ? 2u : (KB == 32 ? 4u : 6u);
  const uint32_t sbo = (8u * row_bytes) >> 4;
  a.desc_hi = sbo | (1u << 14) | (layout_type << 29);  // SBO [32,46), version=1 [46,48), layout [61,64)
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
  a.kb_elems = KB;
  a.nkb = nkb;
  a.N = N;
  a.cout = Cout;
  a.stage_bytes = kTileM * row_bytes;
  a.wblock_bytes = N * row_bytes;

  // iteration grid: output positions (conv) or input positions (one transposed-conv phase)
  const int gh = p->transposed ? x.h : y.h, gw = p->transposed ? x.w : y.w;
  int bw = 128;
  if (gw <= 64) {
    bw = 8;
    while (bw < gw) bw <<= 1;
  }
  a.bw = bw;
  a.bh = kTileM / bw;
  a.gh = gh;
  a.gw = gw;
  a.tiles_w = esn_cdiv(gw, a.bw);
  a.tiles_h = esn_cdiv(gh, a.bh);
  a.ntiles = x.n * a.tiles_w * a.tiles_h;
  a.y = reinterpret_cast<__nv_bfloat16*>(y.ptr);
  a.Hy = y.h;
  a.Wy = y.w;
  a.y_cs = y.c_stride;
  a.ep = make_epi(p->ep);

  // ---- activation tensor map (5-D)
  {
    const cuuint64_t cs = (cuuint64_t)x.c_stride;
    cuuint64_t dims[5], strides[4];
    if (!p->transposed && p->stride == 2) {
      dims[0] = 2 * cs; dims[1] = x.w / 2; dims[2] = 2; dims[3] = x.h / 2; dims[4] = x.n;
      strides[0] = 2 * cs * 2; strides[1] = (cuuint64_t)x.w * cs * 2; strides[2] = 2 * (cuuint64_t)x.w * cs * 2;
      strides[3] = (cuuint64_t)x.h * x.w * cs * 2;
    } else {
      dims[0] = (cuuint64_t)x.c; dims[1] = x.w; dims[2] = 1; dims[3] = x.h; dims[4] = x.n;
      strides[0] = cs * 2; strides[1] = (cuuint64_t)x.w * cs * 2; strides[2] = (cuuint64_t)x.w * cs * 2;
      strides[3] = (cuuint64_t)x.h * x.w * cs * 2;
    }
    const cuuint32_t box[5] = {(cuuint32_t)KB, (cuuint32_t)a.bw, 1, (cuuint32_t)a.bh, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (encode(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ESN_ERR_CUDA;
  }
  // ---- weight tensor map (2-D): rows = tap*N + cout index, cols = Cin
  {
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)ntaps_all * N};
    const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)N};
    const cuuint32_t es[2] = {1, 1};
    if (encode(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->w), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ESN_ERR_CUDA;
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nphase = p->transposed ? 4 : 1;
  for (int ph = 0; ph < nphase; ++ph) {
    const int pa = ph >> 1, pb = ph & 1;
    int nt = 0;
    for (int r = 0; r < p->kh; ++r)
      for (int s = 0; s < p->kw; ++s) {
        int dy, dx, par = 0, coff = 0;
        if (p->transposed) {
          const int th = pa + p->pad_h - r, tw = pb + p->pad_w - s;
          if ((th & 1) || (tw & 1)) continue;
          dy = floordiv2(th);
          dx = floordiv2(tw);
        } else if (p->stride == 2) {
          const int th = r * p->dil_h - p->pad_h, tw = s * p->dil_w - p->pad_w;
          par = th & 1;
          dy = floordiv2(th - par);
          const int wpar = tw & 1;
          dx = floordiv2(tw - wpar);
          coff = wpar * x.c_stride;
        } else {
          dy = r * p->dil_h - p->pad_h;
          dx = s * p->dil_w - p->pad_w;
        }
        a.tap_dy[nt] = dy;
        a.tap_dx[nt] = dx;
        a.tap_par[nt] = par;
        a.tap_coff[nt] = coff;
        a.tap_wrow[nt] = (r * p->kw + s) * N;
        ++nt;
      }
    if (nt == 0) return ESN_ERR_UNSUPPORTED;  // a phase with no taps would need a bias-only fill
    a.ntaps = nt;
    a.sy = p->transposed ? 2 : 1;
    a.sx = a.sy;
    a.oy = p->transposed ? pa : 0;
    a.ox = p->transposed ? pb : 0;
    const uint32_t wbytes = (uint32_t)nt * nkb * a.wblock_bytes;
    a.w_region_bytes = (wbytes + 1023u) & ~1023u;
    // shared memory plan: resident weights + A ring + barriers (+1 KB alignment slack)
    const uint32_t fixed = a.w_region_bytes + 1024u + 256u;
    const bool two_ctas = (fixed + 4u * a.stage_bytes <= 100u * 1024u) && (2 * N <= 256);
    const uint32_t budget = two_ctas ? 110u * 1024u : (uint32_t)lim.max_smem;
    if (fixed + 2u * a.stage_bytes > budget) return ESN_ERR_UNSUPPORTED;
    int stages = (int)((budget - fixed) / a.stage_bytes);
    if (stages > 8) stages = 8;
    a.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * N)) cols <<= 1;
    a.tmem_cols = cols;
    const size_t smem = fixed + (size_t)stages * a.stage_bytes;
    int grid = lim.sms * (two_ctas ? 2 : 1);
    if (grid > a.ntiles) grid = a.ntiles;
    conv_umma_kernel<<<grid, kThreads, smem, st>>>(a);
    ESN_CHECK_LAUNCH();
  }
  return ESN_OK;
}
