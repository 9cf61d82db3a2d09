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
#include <cstdio>

#include "esn_umma_ptx.cuh"

namespace {

__device__ __forceinline__ void tma_prefetch_l2_5d(const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

#ifdef ESN_PAIR_TIMING
#define TWAIT(slot, stmt) do { const long long t0_ = clock64(); stmt; tw[slot] += clock64() - t0_; } while (0)
#else
#define TWAIT(slot, stmt) stmt
#endif

constexpr int kPadPx = 8;      // zero pixels either side of the intermediate row (>= dilation)
constexpr int kMaxTR = 8;      // tiles per image row

struct alignas(64) PairArgs {
  CUtensorMap tmA, tmB1, tmB2, tmY, tmR;   // tmR: residual, used for L2 prefetch only
  int nrows, H, TR, MT, BW, d, ntaps;
  int a_boxw, a_nbox;
  int N, has_res, act1, act2, stages, NI, NS;   // NI: intermediate row buffers (1 or 2); NS: output staging buffers
  const __nv_bfloat16* res;
  int res_cs, W, pf;   // pf: L2 prefetch distance in tiles
  int issuers;         // 1: one MMA-issuing warp; 2: conv1 (+ residual MMA) on warp 1, conv2 on warp 2
  uint32_t stage_bytes, wblock_bytes, inter_bytes, out_buf_bytes, swz_mask;
  uint32_t idesc, desc_hi, tmem_cols;
  const float *scale1, *shift1, *scale2, *shift2, *alpha2;
};

template <int KB>
__global__ void __launch_bounds__(kThreads, 1) conv_pair_kernel(const __grid_constant__ PairArgs a) {
  constexpr int KSTEPS = KB / 16;
  constexpr uint32_t RB = KB * 2u;
  constexpr uint32_t SUB16 = (kTileM * RB) >> 4;
  constexpr int NA = 4;
  const int NS = a.NS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int S = a.stages, ntaps = a.ntaps, MT = a.MT, TR = a.TR;
  const uint32_t N = (uint32_t)a.N;
  const uint32_t w1_base = base;
  const uint32_t w2_base = base + (uint32_t)ntaps * a.wblock_bytes;
  const uint32_t id_base = base + 2u * (uint32_t)ntaps * a.wblock_bytes;   // identity matrix (residual MMA)
  const uint32_t w_region = ((2u * ntaps + 1u) * a.wblock_bytes + 1023u) & ~1023u;
  const uint32_t a_base = base + w_region;
  const uint32_t i_base = a_base + (uint32_t)S * a.stage_bytes;
  const uint32_t o_base = i_base + (uint32_t)a.NI * a.inter_bytes;
  const int nishift = a.NI == 2 ? 1 : 0;
  const uint32_t prm_base = o_base + (uint32_t)NS * a.out_buf_bytes;
  const uint32_t bar_base = prm_base + 5u * 64u * 4u;
  // barriers
  const uint32_t full0 = bar_base, empty0 = full0 + 8u * 8;
  const uint32_t wfull_bar = empty0 + 8u * 8;
  const uint32_t t1full0 = wfull_bar + 8u, t1empty0 = t1full0 + 8u * NA;
  const uint32_t t2full0 = t1empty0 + 8u * NA, t2empty0 = t2full0 + 8u * NA;
  const uint32_t ifull0 = t2empty0 + 8u * NA, ifree0 = ifull0 + 8u * 2 * kMaxTR;   // [buffer][tile]
  const uint32_t sfree0 = ifree0 + 8u * 2 * kMaxTR;
  const uint32_t rres0 = sfree0 + 8u * 2;          // residual MMA issued into accumulator slot (two-issuer mode)
  const uint32_t tmem_slot = rres0 + 8u * NA;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* prm = reinterpret_cast<float*>(smem_raw + (prm_base - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef ESN_PAIR_TIMING
  long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = clock64();
#endif
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmA);
    tma_prefetch_desc(&a.tmB1);
    tma_prefetch_desc(&a.tmB2);
    tma_prefetch_desc(&a.tmY);
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
    for (int j = 0; j < 2 * kMaxTR; ++j) {
      mbar_init(ifull0 + 8u * j, kEpiThreads / 32);
      mbar_init(ifree0 + 8u * j, 1);
    }
    for (int b = 0; b < 2; ++b) mbar_init(sfree0 + 8u * b, 1);
    for (int b = 0; b < NA; ++b) mbar_init(rres0 + 8u * b, 1);
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
    const uint4 z = make_uint4(0, 0, 0, 0);
    // identity B operand (N x KB, K-major, swizzled like the weights): "+ residual" is one more MMA whose A
    // operand is the residual tile TMA drops into the ring -- no thread ever touches the residual
    for (uint32_t i = threadIdx.x; i < a.wblock_bytes / 16u; i += kThreads) sts128(id_base + 16u * i, z);
    __syncthreads();
    if ((int)threadIdx.x < a.N) {
      uint32_t o = threadIdx.x * RB + threadIdx.x * 2u;
      o ^= ((o >> 7) & a.swz_mask) << 4;
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(id_base + o), "h"((unsigned short)0x3F80) : "memory");
    }
    for (int bi = 0; bi < a.NI; ++bi) {
      const uint32_t ib = i_base + (uint32_t)bi * a.inter_bytes;
      const uint32_t tail = ib + (uint32_t)(kPadPx + TR * a.BW) * RB;   // first pixel past the row
      for (uint32_t i = threadIdx.x; i < 2u * pad16; i += kThreads)
        sts128(i < pad16 ? ib + 16u * i : tail + 16u * (i - pad16), z);
    }
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
      // L2 prefetch runs a.pf tiles ahead of the shared-memory ring: the ring is too shallow (shared memory
      // holds the weights and the intermediate row) to cover DRAM latency on its own.  Only the centre row is
      // prefetched: the rows above / below are the centre rows of the CTAs working next to this one.
      int pf_row = blockIdx.x, pf_j = 0, pf_g = 0;
      auto prefetch_tile = [&]() {
        if (leader) {
          const int n = pf_row / a.H, h = pf_row - n * a.H;
          for (int q = 0; q < a.a_nbox; ++q) {
            tma_prefetch_l2_5d(&a.tmA, 0, pf_j * a.BW + q * a.a_boxw, 0, h, n);
            if (a.has_res) tma_prefetch_l2_4d(&a.tmR, 0, pf_j * a.BW + q * a.a_boxw, h, n);
          }
        }
        ++pf_g;
        if (++pf_j == TR) { pf_j = 0; pf_row += gridDim.x; }
      };
      while (pf_g < a.pf && pf_g < G) prefetch_tile();
      // ring order of step g (mirrors the MMA warp): the k taps of tile g, then the residual of tile g-3
      int n = (int)blockIdx.x / a.H, h = (int)blockIdx.x - n * a.H, j = 0;       // tile g
      int n3 = n, h3 = h, j3 = 0;                                                // tile g-3
      for (int g = 0; g < G + 3; ++g) {
        if (g < G) {
          if (pf_g < G && a.pf > 0) prefetch_tile();
          for (int t = 0; t < ntaps; ++t) {
            TWAIT(0, mbar_wait(empty0 + 8u * s, ph ^ 1u));
            if (leader) {
              mbar_expect_tx(full0 + 8u * s, a.stage_bytes);
              const uint32_t dst = a_base + (uint32_t)s * a.stage_bytes;
              for (int q = 0; q < a.a_nbox; ++q)
                tma_load_5d(dst + (uint32_t)(q * a.a_boxw) * RB, &a.tmA, full0 + 8u * s, 0, j * a.BW + q * a.a_boxw, 0,
                            h + (t - (ntaps >> 1)) * a.d, n);
            }
            if (++s == S) { s = 0; ph ^= 1u; }
          }
          if (++j == TR) {
            j = 0;
            h += gridDim.x;
            while (h >= a.H) { h -= a.H; ++n; }
          }
        }
        if (g >= 3 && a.has_res) {
          TWAIT(1, mbar_wait(empty0 + 8u * s, ph ^ 1u));
          if (leader) {
            mbar_expect_tx(full0 + 8u * s, a.stage_bytes);
            const uint32_t dst = a_base + (uint32_t)s * a.stage_bytes;
            for (int q = 0; q < a.a_nbox; ++q)
              tma_load_4d(dst + (uint32_t)(q * a.a_boxw) * RB, &a.tmR, full0 + 8u * s, 0, j3 * a.BW + q * a.a_boxw, h3, n3);
          }
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (g >= 3 && ++j3 == TR) {
          j3 = 0;
          h3 += gridDim.x;
          while (h3 >= a.H) { h3 -= a.H; ++n3; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer.  Step g issues conv1 (k x 1, A ring) of tile g and conv2 (1 x k, intermediate
    // row) of tile g-3 interleaved tap by tap -- two independent accumulator chains in flight -- then
    // "+ residual" as one more MMA against the identity.  conv2 trails by three tiles: its inputs (epilogue 1
    // of tiles <= g-2) never depend on MMAs that were only just issued.
    // This single-issuer loop is the fallback (ESN_PAIR_ISSUERS=1).  The default splits the work over two issuing
    // warps (below): ~20 % faster per launch because the issue stream itself -- waits, tcgen05.mma, ~100-cycle
    // tcgen05.commits -- was the critical path.  The split that works keeps ONE consumer on the TMA ring (warp 1 also
    // issues the residual MMA); letting warp 2 consume ring slots as well produced a rare "unspecified launch failure".
    if (G > 0 && a.issuers == 2 && elect_one()) {
      // ---- two-issuer mode, warp 1: conv1 of tile g from the ring, then (residual pairs) the residual tile of
      // g-3 as the FIRST MMA of conv2's accumulator (accumulate = 0); warp 2 adds the 1 x k taps on top once
      // rres[slot] says that MMA has been issued.  The ring has a single consumer.
      constexpr bool leader = true;   // the whole loop runs on the one elected thread (see esn_umma.cu)
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      const uint32_t dhi = a.desc_hi, idesc = a.idesc;
      const uint32_t w1_lo = desc_lo(w1_base), wblk16 = a.wblock_bytes >> 4;
      const uint32_t a_lo0 = desc_lo(a_base), stage16 = a.stage_bytes >> 4, id_lo = desc_lo(id_base);
      int s = 0;
      uint32_t ph = 0;
      for (int g = 0; g < G + 3; ++g) {
        if (g < G) {
          const uint32_t slot1 = (uint32_t)g & (NA - 1);
          mbar_wait(t1empty0 + 8u * slot1, ((((uint32_t)g / NA) & 1u) ^ 1u));
          tc_fence_after();
          const uint32_t d1 = tmem_base + slot1 * acc_cols;
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(full0 + 8u * s, ph);
            const uint32_t al1 = a_lo0 + (uint32_t)s * stage16, bl1 = w1_lo + (uint32_t)t * wblk16;
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                if (leader)
                  umma_bf16_lo(d1 + (uint32_t)m * N, al1 + (uint32_t)m * SUB16 + 2u * k, bl1 + 2u * k, dhi, idesc,
                               (t | k) != 0 ? 1u : 0u);
            }
            if (leader) umma_commit(empty0 + 8u * s);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
          if (leader) umma_commit(t1full0 + 8u * slot1);
        }
        if (g >= 3 && a.has_res) {
          const int gg = g - 3;
          const uint32_t slot2 = (uint32_t)gg & (NA - 1);
          mbar_wait(t2empty0 + 8u * slot2, ((((uint32_t)gg / NA) & 1u) ^ 1u));
          tc_fence_after();
          mbar_wait(full0 + 8u * s, ph);
          const uint32_t d2 = tmem_base + (NA + slot2) * acc_cols, alr = a_lo0 + (uint32_t)s * stage16;
          for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              if (leader) umma_bf16_lo(d2 + (uint32_t)m * N, alr + (uint32_t)m * SUB16 + 2u * k, id_lo + 2u * k, dhi, idesc, k != 0 ? 1u : 0u);
          }
          if (leader) {
            umma_commit(empty0 + 8u * s);
            mbar_arrive(rres0 + 8u * slot2);
          }
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    } else if (G > 0 && a.issuers != 2 && elect_one()) {
      constexpr bool leader = true;
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      const uint32_t dhi = a.desc_hi, idesc = a.idesc;
      const uint32_t w1_lo = desc_lo(w1_base), w2_lo = desc_lo(w2_base), wblk16 = a.wblock_bytes >> 4;
      const uint32_t a_lo0 = desc_lo(a_base), stage16 = a.stage_bytes >> 4;
      const uint32_t i_lo0 = desc_lo(i_base), id_lo = desc_lo(id_base);
      const uint32_t tap16 = ((uint32_t)a.d * RB) >> 4;
      int s = 0, r2m = 0, j2m = 0;
      uint32_t ph = 0;
      for (int g = 0; g < G + 3; ++g) {
        const bool has1 = g < G, has2 = g >= 3;
        uint32_t d1 = 0, d2 = 0, slot1 = 0, slot2 = 0, ib = 0, i_lo = 0;
        int j = 0;
        if (has2) {
          const int gg = g - 3;
          const int r = r2m;
          j = j2m;
          if (++j2m == TR) { j2m = 0; ++r2m; }
          const int jn = j + 1 < TR ? j + 1 : j;
          ib = (uint32_t)r & (uint32_t)(a.NI - 1);          // intermediate buffer of this row
          TWAIT(0, mbar_wait(ifull0 + 8u * (ib * kMaxTR + jn), (uint32_t)(r >> nishift) & 1u));
          slot2 = (uint32_t)gg & (NA - 1);
          TWAIT(1, mbar_wait(t2empty0 + 8u * slot2, ((((uint32_t)gg / NA) & 1u) ^ 1u)));
          d2 = tmem_base + (NA + slot2) * acc_cols;
          i_lo = i_lo0 + ((ib * a.inter_bytes + (uint32_t)(kPadPx + j * a.BW - (ntaps >> 1) * a.d) * RB) >> 4);
        }
        if (has1) {
          slot1 = (uint32_t)g & (NA - 1);
          TWAIT(2, mbar_wait(t1empty0 + 8u * slot1, ((((uint32_t)g / NA) & 1u) ^ 1u)));
          d1 = tmem_base + slot1 * acc_cols;
        }
        tc_fence_after();
        for (int t = 0; t < ntaps; ++t) {
          uint32_t al1 = 0;
          if (has1) {
            TWAIT(3, mbar_wait(full0 + 8u * s, ph));
            al1 = a_lo0 + (uint32_t)s * stage16;
          }
          const uint32_t bl1 = w1_lo + (uint32_t)t * wblk16, bl2 = w2_lo + (uint32_t)t * wblk16;
          const uint32_t al2 = i_lo + (uint32_t)t * tap16;
          for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              if (leader && has1)
                umma_bf16_lo(d1 + (uint32_t)m * N, al1 + (uint32_t)m * SUB16 + 2u * k, bl1 + 2u * k, dhi, idesc,
                             (t | k) != 0 ? 1u : 0u);
              if (leader && has2)
                umma_bf16_lo(d2 + (uint32_t)m * N, al2 + (uint32_t)m * SUB16 + 2u * k, bl2 + 2u * k, dhi, idesc,
                             (t | k) != 0 ? 1u : 0u);
            }
          }
          if (has1) {
            if (leader) umma_commit(empty0 + 8u * s);
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        }
        if (has2 && a.has_res) {
          // + residual: D += R . I, R = the residual tile in the next ring slot
          TWAIT(3, mbar_wait(full0 + 8u * s, ph));
          const uint32_t alr = a_lo0 + (uint32_t)s * stage16;
          for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              if (leader) umma_bf16_lo(d2 + (uint32_t)m * N, alr + (uint32_t)m * SUB16 + 2u * k, id_lo + 2u * k, dhi, idesc, 1u);
          }
          if (leader) umma_commit(empty0 + 8u * s);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (leader) {
          if (has1) umma_commit(t1full0 + 8u * slot1);
          if (has2) {
            umma_commit(t2full0 + 8u * slot2);
            umma_commit(ifree0 + 8u * (ib * kMaxTR + j));   // epilogue 1 waits on tile j+1's commit before rewriting tile j
          }
        }
      }
    }
  } else if (warp == 2) {
    if (G > 0 && a.issuers == 2 && elect_one()) {
      // ---- two-issuer mode, warp 2: conv2 (1 x k) of tile g-3 from the intermediate row
      constexpr bool leader = true;   // the whole loop runs on the one elected thread (see esn_umma.cu)
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      const uint32_t dhi = a.desc_hi, idesc = a.idesc;
      const uint32_t w2_lo = desc_lo(w2_base), wblk16 = a.wblock_bytes >> 4;
      const uint32_t i_lo0 = desc_lo(i_base);
      const uint32_t tap16 = ((uint32_t)a.d * RB) >> 4;
      int r2m = 0, j2m = 0;
      for (int gg = 0; gg < G; ++gg) {
        const int r = r2m, j = j2m;
        if (++j2m == TR) { j2m = 0; ++r2m; }
        const int jn = j + 1 < TR ? j + 1 : j;
        const uint32_t ib = (uint32_t)r & (uint32_t)(a.NI - 1);
        mbar_wait(ifull0 + 8u * (ib * kMaxTR + jn), (uint32_t)(r >> nishift) & 1u);
        const uint32_t slot2 = (uint32_t)gg & (NA - 1), use = (uint32_t)gg / NA;
        if (a.has_res)
          mbar_wait(rres0 + 8u * slot2, use & 1u);               // accumulator holds the residual (issued by warp 1)
        else
          mbar_wait(t2empty0 + 8u * slot2, (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d2 = tmem_base + (NA + slot2) * acc_cols;
        const uint32_t i_lo = i_lo0 + ((ib * a.inter_bytes + (uint32_t)(kPadPx + j * a.BW - (ntaps >> 1) * a.d) * RB) >> 4);
        const uint32_t acc0 = a.has_res ? 1u : 0u;
        for (int t = 0; t < ntaps; ++t) {
          const uint32_t bl2 = w2_lo + (uint32_t)t * wblk16, al2 = i_lo + (uint32_t)t * tap16;
          for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              if (leader)
                umma_bf16_lo(d2 + (uint32_t)m * N, al2 + (uint32_t)m * SUB16 + 2u * k, bl2 + 2u * k, dhi, idesc,
                             (t | k) != 0 ? 1u : acc0);
          }
        }
        if (leader) {
          umma_commit(t2full0 + 8u * slot2);
          umma_commit(ifree0 + 8u * (ib * kMaxTR + j));
        }
      }
    }
  } else {
    // ---------------- epilogue warps.  Every thread owns ONE (sub-tile, 16-channel) item of each tile -- row
    // R of the tile, channels c0..c0+15 -- so all of its addresses are loop constants: the swizzled offsets
    // into the intermediate row and the staging tile move by whole tiles (multiples of 1 KB), the TMEM
    // address by whole accumulator slots.
    const int q = warp & 3;
    const int grp = (warp - kEpiWarp0) >> 2;
    const int my_m = KB == 64 ? 0 : grp;
    const int my_c0 = KB == 64 ? grp * 16 : 0;
    const int R = my_m * kTileM + q * 32 + lane;
    const uint32_t swz = a.swz_mask;
    const int act1 = a.act1, act2 = a.act2;
    const uint32_t tile_bytes = (uint32_t)a.BW * RB;
    uint32_t off_i[2], off_s[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t o = (uint32_t)(kPadPx + R) * RB + (uint32_t)(my_c0 + 8 * hh) * 2u;
      off_i[hh] = o ^ (((o >> 7) & swz) << 4);
      o = (uint32_t)R * RB + (uint32_t)(my_c0 + 8 * hh) * 2u;
      off_s[hh] = o ^ (((o >> 7) & swz) << 4);
    }
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(my_m * a.N + my_c0);
    const uint32_t pa1 = prm_base + 4u * (uint32_t)my_c0;            // scale1 | +256 shift1
    const uint32_t pa2 = pa1 + 512u;                                 // scale2 | +256 shift2 | +512 alpha2

    // running state of the three tile streams (epilogue 1: tile g, residual prefetch: g-3, epilogue 2: g-4)
    int r1 = 0, j1 = 0;
    int n2 = (int)blockIdx.x / a.H, h2 = (int)blockIdx.x - n2 * a.H, j2 = 0;

    for (int g = 0; g < G + 4; ++g) {
      if (g < G) {
        // ---- epilogue 1 of tile g: TMEM -> act1(acc*s1 + b1) -> bf16 -> intermediate row (UMMA operand layout)
        const uint32_t ib = (uint32_t)r1 & (uint32_t)(a.NI - 1);
        if (r1 >= a.NI) {
          // the previous user of this buffer (row r - NI): its conv2 of tiles j-1, j, j+1 read this part
          const int jl = j1 + 1 < TR ? j1 + 1 : j1;
          TWAIT(0, mbar_wait(ifree0 + 8u * (ib * kMaxTR + jl), (uint32_t)((r1 >> nishift) - 1) & 1u));
        }
        const uint32_t slot = (uint32_t)g & (NA - 1), use = (uint32_t)g / NA;
        TWAIT(1, mbar_wait(t1full0 + 8u * slot, use & 1u));
        tc_fence_after();
        uint32_t rr[16];
        TWAIT(5, tmem_ld16(taddr0 + slot * acc_cols, rr); tmem_ld_wait());
        const uint32_t ibase = i_base + ib * a.inter_bytes + (uint32_t)j1 * tile_bytes;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float f[8];
          const float4 s0 = lds_f4(pa1 + 32u * hh), s1 = lds_f4(pa1 + 32u * hh + 16u);
          const float4 h0 = lds_f4(pa1 + 32u * hh + 256u), h1 = lds_f4(pa1 + 32u * hh + 272u);
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
          sts128(ibase + off_i[hh], float_to_bf16x8(f));
        }
        tc_fence_before();
        fence_proxy_async();      // the intermediate is read by tcgen05.mma through the async proxy
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t1empty0 + 8u * slot);
          mbar_arrive(ifull0 + 8u * (ib * kMaxTR + j1));
        }
        if (++j1 == TR) { j1 = 0; ++r1; }
      }
      if (g >= 4) {
        // ---- epilogue 2 of tile g-4: conv2(g-4) was issued a full step earlier, so this never waits on an
        // MMA that was only just issued
        const uint32_t tc = (uint32_t)(g - 4);
        const uint32_t b = NS == 2 ? (tc & 1u) : 0u, use = NS == 2 ? (tc >> 1) : tc;
        const uint32_t obuf = o_base + b * a.out_buf_bytes;
        TWAIT(2, mbar_wait(sfree0 + 8u * b, (use & 1u) ^ 1u));   // the previous store out of this buffer has been read
        const uint32_t slot = tc & (NA - 1), ause = tc / NA;
        TWAIT(3, mbar_wait(t2full0 + 8u * slot, ause & 1u));
        tc_fence_after();
        uint32_t rr[16];
        TWAIT(6, tmem_ld16(taddr0 + (NA + slot) * acc_cols, rr); tmem_ld_wait());
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float f[8];
          const float4 s0 = lds_f4(pa2 + 32u * hh), s1 = lds_f4(pa2 + 32u * hh + 16u);
          const float4 h0 = lds_f4(pa2 + 32u * hh + 256u), h1 = lds_f4(pa2 + 32u * hh + 272u);
          f[0] = fmaf(__uint_as_float(rr[8 * hh + 0]), s0.x, h0.x);
          f[1] = fmaf(__uint_as_float(rr[8 * hh + 1]), s0.y, h0.y);
          f[2] = fmaf(__uint_as_float(rr[8 * hh + 2]), s0.z, h0.z);
          f[3] = fmaf(__uint_as_float(rr[8 * hh + 3]), s0.w, h0.w);
          f[4] = fmaf(__uint_as_float(rr[8 * hh + 4]), s1.x, h1.x);
          f[5] = fmaf(__uint_as_float(rr[8 * hh + 5]), s1.y, h1.y);
          f[6] = fmaf(__uint_as_float(rr[8 * hh + 6]), s1.z, h1.z);
          f[7] = fmaf(__uint_as_float(rr[8 * hh + 7]), s1.w, h1.w);
          if (act2 == ESN_ACT_RELU) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) f[jj] = fmaxf(f[jj], 0.f);
          } else if (act2 == ESN_ACT_PRELU) {
            const float4 a0 = lds_f4(pa2 + 32u * hh + 512u), a1 = lds_f4(pa2 + 32u * hh + 528u);
            const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) f[jj] = f[jj] >= 0.f ? f[jj] : f[jj] * al[jj];
          }
          sts128(obuf + off_s[hh], float_to_bf16x8(f));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t2empty0 + 8u * slot);
        fence_proxy_async();
        TWAIT(4, epi_bar_sync());
        if (threadIdx.x == kEpiWarp0 * 32) {
          for (int qb = 0; qb < a.a_nbox; ++qb)
            tma_store_4d(&a.tmY, obuf + (uint32_t)(qb * a.a_boxw) * RB, 0, j2 * a.BW + qb * a.a_boxw, h2, n2);
          tma_store_commit();
          TWAIT(7, if (NS == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>());
          if (tc + 1 >= (uint32_t)NS) mbar_arrive(sfree0 + 8u * (NS == 2 ? ((tc + 1) & 1u) : 0u));
        }
        if (++j2 == TR) {
          j2 = 0;
          h2 += gridDim.x;
          while (h2 >= a.H) { h2 -= a.H; ++n2; }
        }
      }
    }
    if (threadIdx.x == kEpiWarp0 * 32) tma_store_wait_all();
  }

#ifdef ESN_PAIR_TIMING
  if (blockIdx.x == 1 && lane == 0 && (warp == 0 || warp == 1 || warp == 3 || warp == 10))
    printf("pair timing warp %d tiles %d total %lld | w0 %lld w1 %lld w2 %lld w3 %lld w4 %lld w5 %lld w6 %lld w7 %lld\n", warp, G,
           clock64() - t_begin, tw[0], tw[1], tw[2], tw[3], tw[4], tw[5], tw[6], tw[7]);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();   // the issuer role ran on one lane: reconverge before the .sync.aligned instruction
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

struct PairLimits {
  int sms = 0, max_smem = 0;
};
const PairLimits& pair_limits() {
  static PairLimits ls[kEsnMaxDevices];
  static std::once_flag once[kEsnMaxDevices];
  const int dev = esn_current_device();
  std::call_once(once[dev], [dev] {
    PairLimits& l = ls[dev];
    cudaDeviceGetAttribute(&l.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&l.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, conv_pair_kernel<64>) == cudaSuccess) l.max_smem -= (int)fa.sharedSizeBytes;
    cudaFuncSetAttribute(conv_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.max_smem);
    cudaFuncSetAttribute(conv_pair_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.max_smem);
  });
  return ls[dev];
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
  if (a.has_res && a.scale2) return ESN_ERR_UNSUPPORTED;   // the residual rides the accumulator: fold scale2 into w2
  a.stage_bytes = (uint32_t)BW * row_bytes;
  a.wblock_bytes = (uint32_t)N * row_bytes;
  a.inter_bytes = (uint32_t)(x.w + 2 * kPadPx) * row_bytes;
  a.inter_bytes = (a.inter_bytes + 1023u) & ~1023u;
  a.out_buf_bytes = (uint32_t)BW * row_bytes;
  const uint32_t w_region = ((2u * a.ntaps + 1u) * a.wblock_bytes + 1023u) & ~1023u;
  // two intermediate row buffers (no wait on the previous row's conv2) when they leave room for >= 6 A stages
  uint32_t fixed = 0;
  int stages = 0;
  static const int force_ns = getenv("ESN_PAIR_NS") ? atoi(getenv("ESN_PAIR_NS")) : 0;
  for (a.NI = 2; a.NI >= 1; --a.NI) {
    // staging buffers and A stages are both BW-pixel tiles: share what is left (one staging buffer when tight)
    const uint32_t base_fixed = 1024u + w_region + (uint32_t)a.NI * a.inter_bytes + 5u * 64u * 4u + 1024u;
    const int units = base_fixed < (uint32_t)lim.max_smem ? (int)(((uint32_t)lim.max_smem - base_fixed) / a.stage_bytes) : 0;
    a.NS = units >= 2 * a.ntaps + 4 ? 2 : 1;
    if (force_ns == 1 || force_ns == 2) a.NS = force_ns;
    stages = units - a.NS;
    if (stages > 8) stages = 8;
    fixed = base_fixed + (uint32_t)a.NS * a.out_buf_bytes;
    if (stages >= 2 * a.ntaps || a.NI == 1) break;
  }
  if (stages < a.ntaps + 1) return ESN_ERR_UNSUPPORTED;
  if ((N >> 4) * MT != 4) return ESN_ERR_UNSUPPORTED;   // one (sub-tile, 16-channel) item per epilogue thread
  a.res = reinterpret_cast<const __nv_bfloat16*>(res.ptr);
  a.res_cs = res.c_stride;
  a.W = x.w;
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
  static const int pf_env = getenv("ESN_PAIR_PF") ? atoi(getenv("ESN_PAIR_PF")) : 0;
  a.pf = pf_env;
  static const int issuers_env = getenv("ESN_PAIR_ISSUERS") ? atoi(getenv("ESN_PAIR_ISSUERS")) : 2;
  a.issuers = issuers_env == 1 ? 1 : 2;
  for (int which = 0; which < (a.has_res ? 2 : 1); ++which) {  // output, residual (prefetch only): (C, W, H, N)
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
  if (getenv("ESN_DEBUG") && getenv("ESN_DEBUG")[0] == '2')
    fprintf(stderr, "esn pair: C=%d n=%d h=%d w=%d d=%d res=%d act2=%d NI=%d NS=%d stages=%d TR=%d smem=%zu grid=%d x=%p y=%p r=%p xcs=%d ycs=%d rcs=%d -> %s\n",
            C, x.n, x.h, x.w, a.d, a.has_res, a.act2, a.NI, a.NS, a.stages, a.TR, smem, grid, x.ptr, y.ptr, res.ptr, x.c_stride,
            y.c_stride, res.c_stride, cudaGetErrorString(cudaPeekAtLastError()));
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
