// Fused factorized pair for sm_100a:   y = act2( conv_1xk( act1( conv_kx1(x)*s1 + b1 ) )*s2 + b2 (+ residual) )
//
// The two halves of ERFNet's non_bottleneck_1d (ERFNet.py:30-65: conv3x1 -> ReLU -> conv1x3 -> BN -> ReLU and
// the dilated second pair with the block's residual) are each ONE kernel: the intermediate tensor never leaves
// the SM.  One CTA walks whole image rows:
//   * conv1 (k x 1, vertical taps) of a 128*MT-pixel tile: k TMA boxes (rows h + (t-1)d) -> tcgen05.mma -> TMEM
//   * epilogue 1: TMEM -> bias/ReLU -> bf16 -> shared memory, written in the swizzled K-major layout the
//     UMMA A operand wants; the buffer holds the whole intermediate row with zero pixels either side
//   * conv2 (1 x k, horizontal taps) reads that row through descriptors shifted by whole pixels (the
//     swizzle is a function of absolute address bits, so any whole-row shift is a legal operand start);
//     it is issued two tiles behind conv1, so a tile's right-hand neighbour already exists
//   * epilogue 2: TMEM -> scale/shift (+ residual prefetched by TMA) -> act -> staging -> TMA store
// HBM traffic per pair: x read once (+ halo rows through L2), residual read once, y written once --
// half of the two-kernel path.  Channel counts: C = Cin = Cout in {16, 64}; row width a multiple of the tile.
#include "esn_umma_ptx.cuh"

namespace {

constexpr int kPadPx = 8;      // zero pixels either side of the intermediate row (>= dilation)
constexpr int kMaxTR = 8;      // tiles per image row

struct alignas(64) PairArgs {
  CUtensorMap tmA, tmB1, tmB2, tmY, tmR;
  int nrows, H, TR, MT, BW, d, ntaps;
  int a_boxw, a_nbox;
  int N, has_res, act1, act2, stages;
  uint32_t stage_bytes, wblock_bytes, inter_bytes, out_buf_bytes, swz_mask;
  uint32_t idesc, desc_hi, tmem_cols;
  const float *scale1, *shift1, *scale2, *shift2, *alpha2;
};

template <int KB>
__global__ void __launch_bounds__(kThreads, 1) conv_pair_kernel(const __grid_constant__ PairArgs a) {
  constexpr int KSTEPS = KB / 16;
  constexpr uint32_t RB = KB * 2u;
  constexpr uint32_t SUB16 = (kTileM * RB) >> 4;
  constexpr int NS = 2, NA = 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int S = a.stages, ntaps = a.ntaps, MT = a.MT, TR = a.TR;
  const uint32_t N = (uint32_t)a.N;
  const uint32_t w1_base = base;
  const uint32_t w2_base = base + (uint32_t)ntaps * a.wblock_bytes;
  const uint32_t w_region = (2u * ntaps * a.wblock_bytes + 1023u) & ~1023u;
  const uint32_t a_base = base + w_region;
  const uint32_t i_base = a_base + (uint32_t)S * a.stage_bytes;
  const uint32_t o_base = i_base + a.inter_bytes;
  const uint32_t prm_base = o_base + (uint32_t)NS * a.out_buf_bytes;
  const uint32_t bar_base = prm_base + 5u * 64u * 4u;
  // barriers
  const uint32_t full0 = bar_base, empty0 = full0 + 8u * 8;
  const uint32_t wfull_bar = empty0 + 8u * 8;
  const uint32_t t1full0 = wfull_bar + 8u, t1empty0 = t1full0 + 8u * NA;
  const uint32_t t2full0 = t1empty0 + 8u * NA, t2empty0 = t2full0 + 8u * NA;
  const uint32_t ifull0 = t2empty0 + 8u * NA, ifree0 = ifull0 + 8u * kMaxTR;
  const uint32_t sfull0 = ifree0 + 8u * kMaxTR, sfree0 = sfull0 + 8u * NS;
  const uint32_t tmem_slot = sfree0 + 8u * NS;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* prm = reinterpret_cast<float*>(smem_raw + (prm_base - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmA);
    tma_prefetch_desc(&a.tmB1);
    tma_prefetch_desc(&a.tmB2);
    tma_prefetch_desc(&a.tmY);
    if (a.has_res) tma_prefetch_desc(&a.tmR);
    for (int s = 0; s < 8; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    mbar_init(wfull_bar, 1);
    for (int b = 0; b < NA; ++b) {
      mbar_init(t1full0 + 8u * b, 1);
      mbar_init(t1empty0 + 8u * b, kEpiThreads / 32);
      mbar_init(t2full0 + 8u * b, 1);
      mbar_init(t2empty0 + 8u * b, kEpiThreads / 32);
    }
    for (int j = 0; j < kMaxTR; ++j) {
      mbar_init(ifull0 + 8u * j, kEpiThreads / 32);
      mbar_init(ifree0 + 8u * j, 1);
    }
    for (int b = 0; b < NS; ++b) {
      mbar_init(sfull0 + 8u * b, 1);
      mbar_init(sfree0 + 8u * b, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < a.N; i += kThreads) {
    prm[i] = a.scale1 ? a.scale1[i] : 1.f;
    prm[64 + i] = a.shift1 ? a.shift1[i] : 0.f;
    prm[128 + i] = a.scale2 ? a.scale2[i] : 1.f;
    prm[192 + i] = a.shift2 ? a.shift2[i] : 0.f;
    prm[256 + i] = (a.act2 == ESN_ACT_PRELU && a.alpha2) ? a.alpha2[i] : 0.f;
  }
  // zero pixels either side of the intermediate row (the 1 x k conv's zero padding)
  {
    const uint32_t pad16 = (uint32_t)kPadPx * RB / 16u;
    const uint32_t tail = i_base + (uint32_t)(kPadPx + TR * a.BW) * RB;   // first pixel past the row
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = threadIdx.x; i < 2u * pad16; i += kThreads)
      sts128(i < pad16 ? i_base + 16u * i : tail + 16u * (i - pad16), z);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // rows blockIdx.x, blockIdx.x + grid, ...: rows in flight on different SMs are neighbours, so the vertical
  // taps' halo rows are shared through L2
  const int nrow_mine = ((int)blockIdx.x < a.nrows) ? (a.nrows - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = nrow_mine * TR;
  const uint32_t acc_cols = (uint32_t)MT * N;

  if (warp == 0) {
    if (G > 0) {
      const bool leader = elect_one();
      if (leader) {
        mbar_expect_tx(wfull_bar, 2u * ntaps * a.wblock_bytes);
        for (int t = 0; t < ntaps; ++t) {
          tma_load_2d(w1_base + (uint32_t)t * a.wblock_bytes, &a.tmB1, wfull_bar, 0, t * (int)N);
          tma_load_2d(w2_base + (uint32_t)t * a.wblock_bytes, &a.tmB2, wfull_bar, 0, t * (int)N);
        }
      }
      int s = 0;
      uint32_t ph = 0;
      int row = blockIdx.x;
      for (int r = 0; r < nrow_mine; ++r, row += gridDim.x) {
        const int n = row / a.H, h = row - n * a.H;
        for (int j = 0; j < TR; ++j) {
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(empty0 + 8u * s, ph ^ 1u);
            if (leader) {
              mbar_expect_tx(full0 + 8u * s, a.stage_bytes);
              const uint32_t dst = a_base + (uint32_t)s * a.stage_bytes;
              for (int q = 0; q < a.a_nbox; ++q)
                tma_load_5d(dst + (uint32_t)(q * a.a_boxw) * RB, &a.tmA, full0 + 8u * s, 0, j * a.BW + q * a.a_boxw, 0,
                            h + (t - (ntaps >> 1)) * a.d, n);
            }
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (G > 0) {
      const bool leader = elect_one();
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      const uint32_t dhi = a.desc_hi, idesc = a.idesc;
      const uint32_t w1_lo = desc_lo(w1_base), w2_lo = desc_lo(w2_base), wblk16 = a.wblock_bytes >> 4;
      const uint32_t a_lo0 = desc_lo(a_base), stage16 = a.stage_bytes >> 4;
      const uint32_t i_lo0 = desc_lo(i_base);
      int s = 0;
      uint32_t ph = 0;
      for (int g = 0; g < G + 2; ++g) {
        if (g < G) {
          // ---- conv1 of tile g
          const uint32_t slot = (uint32_t)g & (NA - 1), use = (uint32_t)g / NA;
          mbar_wait(t1empty0 + 8u * slot, (use & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + slot * acc_cols;
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(full0 + 8u * s, ph);
            tc_fence_after();
            const uint32_t al = a_lo0 + (uint32_t)s * stage16, bl = w1_lo + (uint32_t)t * wblk16;
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                if (leader)
                  umma_bf16_lo(d_tmem + (uint32_t)m * N, al + (uint32_t)m * SUB16 + 2u * k, bl + 2u * k, dhi, idesc,
                               (t | k) != 0 ? 1u : 0u);
            }
            if (leader) umma_commit(empty0 + 8u * s);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
          if (leader) umma_commit(t1full0 + 8u * slot);
          __syncwarp();
        }
        if (g >= 2) {
          // ---- conv2 of tile g-2: its right-hand neighbour (same row) has been written by epilogue 1
          const int gg = g - 2;
          const int r = gg / TR, j = gg - r * TR;
          const int jn = j + 1 < TR ? j + 1 : j;
          mbar_wait(ifull0 + 8u * jn, (uint32_t)r & 1u);
          const uint32_t slot = (uint32_t)gg & (NA - 1), use = (uint32_t)gg / NA;
          mbar_wait(t2empty0 + 8u * slot, (use & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (NA + slot) * acc_cols;
          for (int t = 0; t < ntaps; ++t) {
            const int px = kPadPx + j * a.BW + (t - (ntaps >> 1)) * a.d;
            const uint32_t al = i_lo0 + (((uint32_t)px * RB) >> 4), bl = w2_lo + (uint32_t)t * wblk16;
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                if (leader)
                  umma_bf16_lo(d_tmem + (uint32_t)m * N, al + (uint32_t)m * SUB16 + 2u * k, bl + 2u * k, dhi, idesc,
                               (t | k) != 0 ? 1u : 0u);
            }
          }
          if (leader) {
            umma_commit(t2full0 + 8u * slot);
            umma_commit(ifree0 + 8u * j);   // tile j of the intermediate row: its last reader is conv2(j+1) -- see epilogue 1
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0 && a.has_res) {
      int row = blockIdx.x;
      uint32_t tc = 0;
      for (int r = 0; r < nrow_mine; ++r, row += gridDim.x) {
        const int n = row / a.H, h = row - n * a.H;
        for (int j = 0; j < TR; ++j, ++tc) {
          const uint32_t b = tc & (NS - 1), use = tc / NS;
          mbar_wait(sfree0 + 8u * b, (use & 1u) ^ 1u);
          mbar_expect_tx(sfull0 + 8u * b, a.out_buf_bytes);
          const uint32_t dst = o_base + b * a.out_buf_bytes;
          for (int q = 0; q < a.a_nbox; ++q)
            tma_load_4d(dst + (uint32_t)(q * a.a_boxw) * RB, &a.tmR, sfull0 + 8u * b, 0, j * a.BW + q * a.a_boxw, h, n);
        }
      }
    }
  } else {
    // ---------------- epilogue warps
    const int q = warp & 3;
    const int grp = (warp - kEpiWarp0) >> 2;
    const int nchunk = a.N >> 4;
    const uint32_t swz = a.swz_mask;
    const int has_res = a.has_res, act1 = a.act1, act2 = a.act2;
    int row2 = blockIdx.x, j2 = 0;     // (row, tile) of the next epilogue-2 tile
    for (int g = 0; g < G + 2; ++g) {
      if (g >= 2) {
        // ---- epilogue 2 of tile g-2
        const uint32_t tc = (uint32_t)(g - 2);
        const uint32_t b = tc & (NS - 1), use = tc / NS;
        const uint32_t obuf = o_base + b * a.out_buf_bytes;
        if (has_res)
          mbar_wait(sfull0 + 8u * b, use & 1u);
        else
          mbar_wait(sfree0 + 8u * b, (use & 1u) ^ 1u);
        const uint32_t slot = tc & (NA - 1), ause = tc / NA;
        mbar_wait(t2full0 + 8u * slot, ause & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (NA + slot) * acc_cols;
        int m = 0, chn = grp;
        while (chn >= nchunk) { chn -= nchunk; ++m; }
        while (m < MT) {
          const int c0 = chn << 4;
          uint32_t rr[16];
          tmem_ld16(taddr + (uint32_t)(m * a.N + c0), rr);
          tmem_ld_wait();
          const int R = m * kTileM + q * 32 + lane;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int cb8 = c0 + 8 * hh;
            float f[8];
            const uint32_t pa = prm_base + 512u + 4u * (uint32_t)cb8;
            const float4 s0 = lds_f4(pa), s1 = lds_f4(pa + 16u);
            const float4 h0 = lds_f4(pa + 256u), h1 = lds_f4(pa + 272u);
            f[0] = fmaf(__uint_as_float(rr[8 * hh + 0]), s0.x, h0.x);
            f[1] = fmaf(__uint_as_float(rr[8 * hh + 1]), s0.y, h0.y);
            f[2] = fmaf(__uint_as_float(rr[8 * hh + 2]), s0.z, h0.z);
            f[3] = fmaf(__uint_as_float(rr[8 * hh + 3]), s0.w, h0.w);
            f[4] = fmaf(__uint_as_float(rr[8 * hh + 4]), s1.x, h1.x);
            f[5] = fmaf(__uint_as_float(rr[8 * hh + 5]), s1.y, h1.y);
            f[6] = fmaf(__uint_as_float(rr[8 * hh + 6]), s1.z, h1.z);
            f[7] = fmaf(__uint_as_float(rr[8 * hh + 7]), s1.w, h1.w);
            uint32_t off = (uint32_t)R * RB + (uint32_t)cb8 * 2u;
            off ^= ((off >> 7) & swz) << 4;
            const uint32_t saddr = obuf + off;
            if (has_res) {
              float gq[8];
              bf16x8_to_float(lds128(saddr), gq);
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) f[jj] += gq[jj];
            }
            if (act2 == ESN_ACT_RELU) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) f[jj] = fmaxf(f[jj], 0.f);
            } else if (act2 == ESN_ACT_PRELU) {
              const float4 a0 = lds_f4(pa + 512u), a1 = lds_f4(pa + 528u);
              const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) f[jj] = f[jj] >= 0.f ? f[jj] : f[jj] * al[jj];
            }
            sts128(saddr, float_to_bf16x8(f));
          }
          chn += kEpiThreads / 128;
          while (chn >= nchunk) { chn -= nchunk; ++m; }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t2empty0 + 8u * slot);
        fence_proxy_async();
        epi_bar_sync();
        if (threadIdx.x == kEpiWarp0 * 32) {
          const int n = row2 / a.H, h = row2 - n * a.H;
          for (int qb = 0; qb < a.a_nbox; ++qb)
            tma_store_4d(&a.tmY, obuf + (uint32_t)(qb * a.a_boxw) * RB, 0, j2 * a.BW + qb * a.a_boxw, h, n);
          tma_store_commit();
          tma_store_wait_read<NS - 1>();
          if (tc + 1 >= (uint32_t)NS) mbar_arrive(sfree0 + 8u * ((tc + 1) & (NS - 1)));
        }
        if (++j2 == TR) { j2 = 0; row2 += gridDim.x; }
      }
      if (g < G) {
        // ---- epilogue 1 of tile g: TMEM -> act1(acc*s1 + b1) -> bf16 -> intermediate row (UMMA operand layout)
        const int r = g / TR, j = g - r * TR;
        if (r > 0) {
          // the previous row's conv2 of tiles j-1, j, j+1 read this part of the buffer
          const int jl = j + 1 < TR ? j + 1 : j;
          mbar_wait(ifree0 + 8u * jl, (uint32_t)(r - 1) & 1u);
        }
        const uint32_t slot = (uint32_t)g & (NA - 1), use = (uint32_t)g / NA;
        mbar_wait(t1full0 + 8u * slot, use & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * acc_cols;
        int m = 0, chn = grp;
        while (chn >= nchunk) { chn -= nchunk; ++m; }
        while (m < MT) {
          const int c0 = chn << 4;
          uint32_t rr[16];
          tmem_ld16(taddr + (uint32_t)(m * a.N + c0), rr);
          tmem_ld_wait();
          const int P = kPadPx + j * a.BW + m * kTileM + q * 32 + lane;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int cb8 = c0 + 8 * hh;
            float f[8];
            const uint32_t pa = prm_base + 4u * (uint32_t)cb8;
            const float4 s0 = lds_f4(pa), s1 = lds_f4(pa + 16u);
            const float4 h0 = lds_f4(pa + 256u), h1 = lds_f4(pa + 272u);
            f[0] = fmaf(__uint_as_float(rr[8 * hh + 0]), s0.x, h0.x);
            f[1] = fmaf(__uint_as_float(rr[8 * hh + 1]), s0.y, h0.y);
            f[2] = fmaf(__uint_as_float(rr[8 * hh + 2]), s0.z, h0.z);
            f[3] = fmaf(__uint_as_float(rr[8 * hh + 3]), s0.w, h0.w);
            f[4] = fmaf(__uint_as_float(rr[8 * hh + 4]), s1.x, h1.x);
            f[5] = fmaf(__uint_as_float(rr[8 * hh + 5]), s1.y, h1.y);
            f[6] = fmaf(__uint_as_float(rr[8 * hh + 6]), s1.z, h1.z);
            f[7] = fmaf(__uint_as_float(rr[8 * hh + 7]), s1.w, h1.w);
            if (act1 == ESN_ACT_RELU) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) f[jj] = fmaxf(f[jj], 0.f);
            }
            uint32_t off = (uint32_t)P * RB + (uint32_t)cb8 * 2u;
            off ^= ((off >> 7) & swz) << 4;
            sts128(i_base + off, float_to_bf16x8(f));
          }
          chn += kEpiThreads / 128;
          while (chn >= nchunk) { chn -= nchunk; ++m; }
        }
        tc_fence_before();
        fence_proxy_async();      // the intermediate is read by tcgen05.mma through the async proxy
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t1empty0 + 8u * slot);
          mbar_arrive(ifull0 + 8u * j);
        }
      }
    }
    if (threadIdx.x == kEpiWarp0 * 32) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

struct PairLimits {
  int sms = 0, max_smem = 0;
};
const PairLimits& pair_limits() {
  static PairLimits l;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&l.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&l.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, conv_pair_kernel<64>) == cudaSuccess) l.max_smem -= (int)fa.sharedSizeBytes;
    cudaFuncSetAttribute(conv_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.max_smem);
    cudaFuncSetAttribute(conv_pair_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.max_smem);
  });
  return l;
}

}  // namespace

extern "C" int esn_conv_pair_umma(const EsnConvPair* p, void* stream) {
  if (!p || !p->w1 || !p->w2) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& y = p->y;
  const EsnTensor& res = p->ep2.residual;
  if (!esn_valid_nhwc(x) || !esn_valid_nhwc(y)) return ESN_ERR_BAD_ARG;
  if (x.dtype != ESN_BF16 || y.dtype != ESN_BF16 || (res.ptr && res.dtype != ESN_BF16)) return ESN_ERR_UNSUPPORTED;
  int rc = esn_check_epilogue(p->ep2, y);
  if (rc) return rc;
  if (p->ep1.residual.ptr || p->ep1.flags || p->ep2.flags) return ESN_ERR_UNSUPPORTED;
  if (p->ep1.act != ESN_ACT_NONE && p->ep1.act != ESN_ACT_RELU) return ESN_ERR_UNSUPPORTED;
  const int C = x.c;
  if ((C != 16 && C != 64) || y.c != C) return ESN_ERR_UNSUPPORTED;
  if (x.n != y.n || x.h != y.h || x.w != y.w) return ESN_ERR_BAD_SHAPE;
  if (p->taps != 3 || p->dilation < 1 || p->dilation > kPadPx) return ESN_ERR_UNSUPPORTED;
  if (x.c_stride % 8 || y.c_stride % 8 || ((uintptr_t)x.ptr % 16) || ((uintptr_t)y.ptr % 16) || ((uintptr_t)p->w1 % 16) ||
      ((uintptr_t)p->w2 % 16))
    return ESN_ERR_ALIGN;
  if (res.ptr && (res.c_stride % 8 || ((uintptr_t)res.ptr % 16))) return ESN_ERR_ALIGN;
  const int KB = C, N = C;
  const int MT = 64 / KB;
  const int BW = MT * kTileM;
  if (x.w % BW) return ESN_ERR_UNSUPPORTED;
  const int TR = x.w / BW;
  if (TR > kMaxTR) return ESN_ERR_UNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return ESN_ERR_CUDA;
  const PairLimits& lim = pair_limits();
  if (lim.sms <= 0) return ESN_ERR_CUDA;

  PairArgs a;
  memset(&a, 0, sizeof(a));
  const CUtensorMapSwizzle swz = KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  const uint32_t row_bytes = KB * 2;
  const uint32_t layout_type = KB == 64 ? 2u : 6u;
  a.desc_hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout_type << 29);
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
  a.swz_mask = KB == 64 ? 7u : 1u;
  a.nrows = x.n * x.h;
  a.H = x.h;
  a.TR = TR;
  a.MT = MT;
  a.BW = BW;
  a.d = p->dilation;
  a.ntaps = p->taps;
  a.a_boxw = BW > 256 ? 256 : BW;
  a.a_nbox = BW / a.a_boxw;
  a.N = N;
  a.has_res = res.ptr != nullptr;
  a.act1 = p->ep1.act;
  a.act2 = p->ep2.act;
  a.scale1 = p->ep1.scale;
  a.shift1 = p->ep1.shift;
  a.scale2 = p->ep2.scale;
  a.shift2 = p->ep2.shift;
  a.alpha2 = p->ep2.alpha;
  if (a.act2 == ESN_ACT_PRELU && !a.alpha2) return ESN_ERR_BAD_ARG;
  a.stage_bytes = (uint32_t)BW * row_bytes;
  a.wblock_bytes = (uint32_t)N * row_bytes;
  a.inter_bytes = (uint32_t)(x.w + 2 * kPadPx) * row_bytes;
  a.inter_bytes = (a.inter_bytes + 1023u) & ~1023u;
  a.out_buf_bytes = (uint32_t)BW * row_bytes;
  const uint32_t w_region = (2u * a.ntaps * a.wblock_bytes + 1023u) & ~1023u;
  const uint32_t fixed = 1024u + w_region + a.inter_bytes + 2u * a.out_buf_bytes + 5u * 64u * 4u + 1024u;
  if (fixed >= (uint32_t)lim.max_smem) return ESN_ERR_UNSUPPORTED;
  int stages = (int)(((uint32_t)lim.max_smem - fixed) / a.stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < a.ntaps + 1) return ESN_ERR_UNSUPPORTED;
  a.stages = stages;
  a.tmem_cols = 512;
  if (8 * MT * N > 512) return ESN_ERR_UNSUPPORTED;

  {  // activations: (C, W, 1, H, N); one box = a_boxw pixels of one row
    const cuuint64_t cs = (cuuint64_t)x.c_stride;
    const cuuint64_t dims[5] = {(cuuint64_t)x.c, (cuuint64_t)x.w, 1, (cuuint64_t)x.h, (cuuint64_t)x.n};
    const cuuint64_t strides[4] = {cs * 2, (cuuint64_t)x.w * cs * 2, (cuuint64_t)x.w * cs * 2, (cuuint64_t)x.h * x.w * cs * 2};
    const cuuint32_t box[5] = {(cuuint32_t)KB, (cuuint32_t)a.a_boxw, 1, 1, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (encode(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ESN_ERR_CUDA;
  }
  for (int which = 0; which < 2; ++which) {  // weights: rows = tap*N + cout, cols = Cin
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)a.ntaps * N};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)N};
    const cuuint32_t es[2] = {1, 1};
    if (encode(which ? &a.tmB2 : &a.tmB1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(which ? p->w2 : p->w1), dims,
               strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ESN_ERR_CUDA;
  }
  for (int which = 0; which < (a.has_res ? 2 : 1); ++which) {  // output / residual: (C, W, H, N)
    const EsnTensor& t = which ? res : y;
    const cuuint64_t cs = (cuuint64_t)t.c_stride;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)y.w, (cuuint64_t)y.h, (cuuint64_t)y.n};
    const cuuint64_t strides[3] = {cs * 2, (cuuint64_t)y.w * cs * 2, (cuuint64_t)y.h * y.w * cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)a.a_boxw, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(which ? &a.tmR : &a.tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ESN_ERR_CUDA;
  }
  const size_t smem = (size_t)fixed + (size_t)a.stages * a.stage_bytes;
  int grid = lim.sms;
  if (grid > a.nrows) grid = a.nrows;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (KB == 64) conv_pair_kernel<64><<<grid, kThreads, smem, st>>>(a);
  else conv_pair_kernel<16><<<grid, kThreads, smem, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
