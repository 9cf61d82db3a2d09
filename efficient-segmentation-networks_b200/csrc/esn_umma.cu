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
#include "esn_umma_ptx.cuh"

namespace {

constexpr int kMaxTaps = 9;

struct alignas(64) UmmaArgs {
  CUtensorMap tmA;   // activations, 5-D, main box
  CUtensorMap tmAh;  // activations, 5-D, tail box of the horizontal-reuse window
  CUtensorMap tmB;   // weights, 2-D
  CUtensorMap tmY;   // output, 4-D   (staged epilogue only)
  CUtensorMap tmY2;  // phase-fused transposed conv: output rows 2i+1 (tmY: rows 2i)

  CUtensorMap tmR;   // residual, 4-D (staged epilogue with residual only)
  CUtensorMap tmZ;   // second output, 4-D (dual == 2)
  // second epilogue stage: v2 = act2(bf16(v) * scale2 + shift2).  dual == 1: only v2 is stored (through tmY, built over
  // y2); dual == 2: v leaves through tmY and v2 through tmZ out of a second set of staging buffers
  int dual, act2;
  int warp_issue;    // MMA issuer: 1 = warp-uniform loop with an elected lane, 0 = one thread runs the whole loop
  const float* scale2;
  const float* shift2;
  const float* alpha2;
  int mode, nunits;
  int g_dw, g_dh, g_dn;              // MODE_GENERIC: (tw, th, n) increment of one grid stride (interleaved tiles)
  int ntaps, nkb, kb_elems, N, cout, MT;
  int bw, bh, tiles_w, tiles_h, gh, gw;
  int a_boxw, a_nbox, o_boxw, o_nbox;
  int hs_d, hs_pad;                  // MODE_HREUSE: tap spacing / left padding in pixels
  int hr_kh, hr_dh, hr_ph;           // MODE_HREUSE: tap rows (1 for 1 x k convs), their spacing and the top padding in rows
  int vr_d, vr_pad, vr_L, vr_nseg;   // MODE_VREUSE: row stride, top padding, outputs per unit, segments
  int vr_cnt, vr_rem;                // rows of the longest residue class; residues >= vr_rem have one less
  int vr_kh, vr_kw;                  // MODE_VREUSE: tap rows (ring slots per output) x taps per row (shifted windows of one slot)
  int tap_dx[kMaxTaps], tap_dy[kMaxTaps], tap_par[kMaxTaps], tap_coff[kMaxTaps], tap_wrow[kMaxTaps];
  __nv_bfloat16* y;
  int Hy, Wy, y_cs, sy, oy, sx, ox;
  EpiArgs ep;
  int staged, has_res, stages, NS, NA;   // NS staging buffers, NA TMEM accumulator buffers
  int shuf_bpa;                      // phase-fused transposed conv: column blocks per output-row parity (0 = off)
  int cbo, ncb;                      // staged epilogue: channels per 128B-wide column block, #blocks
  uint32_t load_bytes, stage_bytes, wblock_bytes, w_region_bytes, out_block_bytes, out_buf_bytes, out_swz_mask;
  uint32_t idesc, desc_hi;           // instruction descriptor; upper 32 bits of the smem descriptors
  uint32_t tmem_cols;
};

// ------------------------------------------------------------------ kernel
// Work decomposition.  A "unit" is what one CTA walks before moving on:
//   MODE_GENERIC : one output tile (bh x bw positions), one TMA box per (tap, K block)
//   MODE_HREUSE  : one output row tile (1 x bw); ONE box of bw+(k-1)d pixels per K block, the 1xk
//                  taps read it through row-shifted UMMA descriptors (the swizzle is a function of
//                  absolute smem address bits, so any whole-row shift of the start address is legal
//                  with base_offset = 0 -- measured on B200)
//   MODE_VREUSE  : L output row tiles h0, h0+d, h0+2d, ... of one column strip; input rows go
//                  through the smem ring once each and serve as tap 0/1/2 of three outputs
// Every CTA owns a contiguous range of units and walks it with an incremental iterator (no
// divisions on the hot path); stage / phase counters are running registers.
// Shared memory map (base aligned to 1024 B):
//   [ weights: ntaps*nkb blocks of N x KB ]  [ A ring: stages x stage_bytes ]
//   [ staging: NS x (ncb blocks of rows x cbo) -- residual lands here, output leaves from here ]
//   [ epilogue params: scale|shift|alpha ]  [ mbarriers ]
enum { MODE_GENERIC = 0, MODE_HREUSE = 1, MODE_VREUSE = 2 };

template <int MODE>
struct UnitIter {
  int n, tw, th, res, seg;   // th: generic/hreuse tile row index; res/seg: vreuse residue and segment
  int w0, h0, hstep, len;
  __device__ __forceinline__ void set(const UmmaArgs& a) {
    w0 = tw * a.bw;
    if (MODE == MODE_VREUSE) {
      const int cnt = a.vr_cnt - (res >= a.vr_rem ? 1 : 0);
      const int first = seg * a.vr_L;
      len = min(cnt - first, a.vr_L);
      h0 = res + first * a.vr_d;
      hstep = a.vr_d;
    } else {
      h0 = th * a.bh;
      hstep = 0;
      len = 1;
    }
  }
  __device__ __forceinline__ void init(const UmmaArgs& a, int u) {
    th = res = seg = 0;
    if (MODE == MODE_VREUSE) {
      seg = u % a.vr_nseg;
      int t = u / a.vr_nseg;
      res = t % a.vr_d;
      t /= a.vr_d;
      tw = t % a.tiles_w;
      n = t / a.tiles_w;
    } else {
      tw = u % a.tiles_w;
      const int t = u / a.tiles_w;
      th = t % a.tiles_h;
      n = t / a.tiles_h;
    }
    set(a);
  }
  __device__ __forceinline__ void next(const UmmaArgs& a) {
    if (MODE == MODE_VREUSE) {
      if (++seg == a.vr_nseg) {
        seg = 0;
        if (++res == a.vr_d) {
          res = 0;
          if (++tw == a.tiles_w) { tw = 0; ++n; }
        }
      }
    } else if (MODE == MODE_HREUSE) {
      if (++tw == a.tiles_w) {
        tw = 0;
        if (++th == a.tiles_h) { th = 0; ++n; }
      }
    } else {
      // interleaved assignment (unit += gridDim.x): tiles in flight on different SMs are neighbours in
      // memory, so the halo rows of multi-row taps are shared through L2 instead of re-read from HBM
      tw += a.g_dw;
      if (tw >= a.tiles_w) { tw -= a.tiles_w; ++th; }
      th += a.g_dh;
      if (th >= a.tiles_h) { th -= a.tiles_h; ++n; }
      n += a.g_dn;
    }
    set(a);
  }
};

template <int KB, int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ UmmaArgs a) {
  constexpr int KSTEPS = KB / 16;
  constexpr uint32_t RB = KB * 2u;                 // bytes per A/B smem row
  constexpr uint32_t SUB16 = (kTileM * RB) >> 4;   // one 128-row M sub-tile, in 16 B descriptor units
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int S = a.stages, NS = a.NS;
  const uint32_t w_base = base;
  const uint32_t a_base = base + a.w_region_bytes;
  const uint32_t o_base = a_base + (uint32_t)S * a.stage_bytes;
  const uint32_t prm_base = o_base + (uint32_t)NS * a.out_buf_bytes;
  const uint32_t bar_base = prm_base + 3u * 256u * 4u;
  const uint32_t full0 = bar_base, empty0 = bar_base + 8u * S;
  const uint32_t wfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tfull0 = bar_base + 8u * (2 * S + 1), tempty0 = bar_base + 8u * (2 * S + 5);
  const uint32_t sfull0 = bar_base + 8u * (2 * S + 9), sfree0 = bar_base + 8u * (2 * S + 13);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 17);
  const uint32_t prm2_base = (bar_base + 512u + 1023u) & ~1023u;   // dual only: scale2|shift2|alpha2, then the y2 staging set
  const uint32_t o2_delta = prm2_base + 3072u - o_base;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* prm = reinterpret_cast<float*>(smem_raw + (prm_base - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmA);
    tma_prefetch_desc(&a.tmB);
    if (MODE == MODE_HREUSE || (MODE == MODE_VREUSE && a.vr_kw > 1)) tma_prefetch_desc(&a.tmAh);
    if (a.staged) tma_prefetch_desc(&a.tmY);
    if (a.staged && a.shuf_bpa) tma_prefetch_desc(&a.tmY2);
    if (a.has_res && a.staged) tma_prefetch_desc(&a.tmR);
    if (a.dual == 2) tma_prefetch_desc(&a.tmZ);
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    mbar_init(wfull_bar, 1);
    for (int b = 0; b < 4; ++b) {
      mbar_init(tfull0 + 8u * b, 1);
      mbar_init(tempty0 + 8u * b, kEpiThreads / 32);
    }
    for (int b = 0; b < 4; ++b) {
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
  // per-channel epilogue parameters -> smem (read back as broadcast LDS.128)
  for (int i = threadIdx.x; i < a.N; i += kThreads) {
    const bool in = i < a.cout;
    prm[i] = (in && a.ep.scale) ? a.ep.scale[i] : 1.f;
    prm[256 + i] = (in && a.ep.shift) ? a.ep.shift[i] : 0.f;
    prm[512 + i] = (in && a.ep.act == ESN_ACT_PRELU) ? a.ep.alpha[i] : 0.f;
  }
  if (a.dual) {
    float* prm2 = reinterpret_cast<float*>(smem_raw + (prm2_base - raw));
    for (int i = threadIdx.x; i < a.N; i += kThreads) {
      const bool in = i < a.cout;
      prm2[i] = (in && a.scale2) ? a.scale2[i] : 1.f;
      prm2[256 + i] = (in && a.shift2) ? a.shift2[i] : 0.f;
      prm2[512 + i] = (in && a.act2 == ESN_ACT_PRELU) ? a.alpha2[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // contiguous unit range of this CTA
  const int u_step = MODE == MODE_GENERIC ? (int)gridDim.x : 1;
  const int u_begin = MODE == MODE_GENERIC ? (int)blockIdx.x : (int)(((long long)a.nunits * blockIdx.x) / gridDim.x);
  const int u_end = MODE == MODE_GENERIC ? a.nunits : (int)(((long long)a.nunits * (blockIdx.x + 1)) / gridDim.x);
  const uint32_t acc_cols = (uint32_t)(a.MT * a.N);
  const int ntaps = a.ntaps, nkb = a.nkb, MT = a.MT;
  const uint32_t NA = (uint32_t)a.NA;

  if (warp == 0) {
    if (u_begin < u_end) {
      // ---------------- TMA producer (warp-uniform control flow; one elected lane issues):
      // weights once, then the A ring
      const bool leader = elect_one();
      if (leader) {
        mbar_expect_tx(wfull_bar, (uint32_t)(ntaps * nkb) * a.wblock_bytes);
        for (int t = 0; t < ntaps; ++t)
          for (int kb = 0; kb < nkb; ++kb)
            tma_load_2d(w_base + (uint32_t)(t * nkb + kb) * a.wblock_bytes, &a.tmB, wfull_bar, kb * KB, a.tap_wrow[t]);
      }
      int s = 0;
      uint32_t ph = 0;   // parity of the phase the consumer completes next on stage s
      UnitIter<MODE> un;
      un.init(a, u_begin);
      for (int u = u_begin; u < u_end; u += u_step, un.next(a)) {
        if (un.len <= 0) continue;
        if (MODE == MODE_VREUSE) {
          // rows h0-pad + j*d, j = 0..len+ntaps-2: one ring slot each, each row loaded exactly once
          int row = un.h0 - a.vr_pad;
          const int wl = un.w0 - a.hs_pad;     // k_h x k_w convs: each slot holds the row window [w0-pad, w0+bw+(k_w-1)d-pad)
          for (int j = 0; j < un.len + a.vr_kh - 1; ++j, row += un.hstep) {
            mbar_wait(empty0 + 8u * s, ph ^ 1u);
            if (leader) {
              mbar_expect_tx(full0 + 8u * s, a.load_bytes);
              const uint32_t dst = a_base + (uint32_t)s * a.stage_bytes;
              for (int q = 0; q < a.a_nbox; ++q)
                tma_load_5d(dst + (uint32_t)(q * a.a_boxw) * RB, &a.tmA, full0 + 8u * s, 0, wl + q * a.a_boxw, 0,
                            row, un.n);
              if (a.vr_kw > 1) tma_load_5d(dst + (uint32_t)a.bw * RB, &a.tmAh, full0 + 8u * s, 0, wl + a.bw, 0, row, un.n);
            }
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        } else if (MODE == MODE_HREUSE) {
          const int wl = un.w0 - a.hs_pad;   // window [w0 - pad, w0 + bw + (k-1)d - pad)
          // one window per (tap row, K block); rows outside the image are zero-filled by TMA (the conv's padding)
          int row = un.h0 - a.hr_ph;
          for (int r = 0; r < a.hr_kh; ++r, row += a.hr_dh) {
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait(empty0 + 8u * s, ph ^ 1u);
              if (leader) {
                mbar_expect_tx(full0 + 8u * s, a.load_bytes);
                const uint32_t dst = a_base + (uint32_t)s * a.stage_bytes;
                for (int q = 0; q < a.a_nbox; ++q)
                  tma_load_5d(dst + (uint32_t)(q * a.a_boxw) * RB, &a.tmA, full0 + 8u * s, kb * KB, wl + q * a.a_boxw, 0,
                              row, un.n);
                tma_load_5d(dst + (uint32_t)a.bw * RB, &a.tmAh, full0 + 8u * s, kb * KB, wl + a.bw, 0, row, un.n);
              }
              if (++s == S) { s = 0; ph ^= 1u; }
            }
          }
        } else {
          for (int t = 0; t < ntaps; ++t) {
            const int cw = un.w0 + a.tap_dx[t];
            const int ch = un.h0 + a.tap_dy[t];
            for (int kb = 0; kb < nkb; ++kb) {
              mbar_wait(empty0 + 8u * s, ph ^ 1u);
              if (leader) {
                mbar_expect_tx(full0 + 8u * s, a.load_bytes);
                tma_load_5d(a_base + (uint32_t)s * a.stage_bytes, &a.tmA, full0 + 8u * s, a.tap_coff[t] + kb * KB, cw,
                            a.tap_par[t], ch, un.n);
              }
              if (++s == S) { s = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && a.warp_issue) {
    if (u_begin < u_end) {
      // ---------------- MMA issuer, warp-uniform form (ESN_UMMA_ISSUE=warp; kept for A/B measurements): warp-uniform loops (operands stay in uniform registers), one
      // elected lane issues tcgen05.mma / tcgen05.commit; descriptors advance with 32-bit adds
      const bool leader = elect_one();
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      const uint32_t dhi = a.desc_hi, idesc = a.idesc;
      const uint32_t w_lo = desc_lo(w_base), wblk16 = a.wblock_bytes >> 4;
      const uint32_t a_lo0 = desc_lo(a_base), stage16 = a.stage_bytes >> 4;
      const uint32_t N = (uint32_t)a.N;
      int s = 0;
      uint32_t ph = 0, acc = 0, aph = 0;
      UnitIter<MODE> un;
      un.init(a, u_begin);
      for (int u = u_begin; u < u_end; u += u_step, un.next(a)) {
        if (un.len <= 0) continue;
        for (int i = 0; i < un.len; ++i) {
          mbar_wait(tempty0 + 8u * acc, aph ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * acc_cols;
          if (MODE == MODE_VREUSE) {
            // tap t of this output lives in ring slot s+t (s = slot of the oldest live row)
            int st = s;
            uint32_t pt = ph;
            const int kh = a.vr_kh, kw = a.vr_kw;
            const uint32_t shift16 = ((uint32_t)a.hs_d * RB) >> 4;   // horizontal tap spacing in descriptor units
            uint32_t bl = w_lo;
            for (int t = 0; t < kh; ++t) {
              if (i == 0 || t == kh - 1) {
                mbar_wait(full0 + 8u * st, pt);
              }
              uint32_t al = a_lo0 + (uint32_t)st * stage16;
              for (int q = 0; q < kw; ++q, al += shift16, bl += wblk16) {
                for (int m = 0; m < MT; ++m) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    if (leader)
                      umma_bf16_lo(d_tmem + (uint32_t)m * N, al + (uint32_t)m * SUB16 + 2u * k, bl + 2u * k, dhi, idesc,
                                   (t | q | k) != 0 ? 1u : 0u);
                }
              }
              if (++st == S) { st = 0; pt ^= 1u; }
            }
            if (leader) umma_commit(empty0 + 8u * s);   // the oldest row is dead once these MMAs retire
            if (++s == S) { s = 0; ph ^= 1u; }
            if (i == un.len - 1) {          // unit done: release the kh-1 rows still held
              for (int t = 1; t < kh; ++t) {
                if (leader) umma_commit(empty0 + 8u * s);
                if (++s == S) { s = 0; ph ^= 1u; }
              }
            }
          } else if (MODE == MODE_HREUSE) {
            const uint32_t shift16 = ((uint32_t)a.hs_d * RB) >> 4;   // tap spacing in descriptor units
            const int kwt = ntaps / a.hr_kh;                          // taps per tap row
            for (int r = 0; r < a.hr_kh; ++r) {
              for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full0 + 8u * s, ph);
                uint32_t al = a_lo0 + (uint32_t)s * stage16;
                uint32_t bl = w_lo + (uint32_t)(r * kwt * nkb + kb) * wblk16;
                for (int t = 0; t < kwt; ++t, al += shift16, bl += (uint32_t)nkb * wblk16) {
                  for (int m = 0; m < MT; ++m) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                      if (leader)
                        umma_bf16_lo(d_tmem + (uint32_t)m * N, al + (uint32_t)m * SUB16 + 2u * k, bl + 2u * k, dhi,
                                     idesc, (r | kb | t | k) != 0 ? 1u : 0u);
                  }
                }
                if (leader) umma_commit(empty0 + 8u * s);
                if (++s == S) { s = 0; ph ^= 1u; }
              }
            }
          } else {
            const int kiters = ntaps * nkb;
            uint32_t bl = w_lo;
            for (int ki = 0; ki < kiters; ++ki, bl += wblk16) {
              mbar_wait(full0 + 8u * s, ph);
              const uint32_t al = a_lo0 + (uint32_t)s * stage16;
              for (int m = 0; m < MT; ++m) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  if (leader)
                    umma_bf16_lo(d_tmem + (uint32_t)m * N, al + (uint32_t)m * SUB16 + 2u * k, bl + 2u * k, dhi, idesc,
                                 (ki | k) != 0 ? 1u : 0u);
              }
              if (leader) umma_commit(empty0 + 8u * s);  // frees the smem stage when these MMAs retire
              if (++s == S) { s = 0; ph ^= 1u; }
            }
          }
          if (leader) umma_commit(tfull0 + 8u * acc);  // accumulators ready for the epilogue
          __syncwarp();
          if (++acc == NA) { acc = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (u_begin < u_end && elect_one()) {
      // ---------------- MMA issuer: ONE thread runs the whole issue loop.  (A warp-uniform loop with the elected lane
      // branching around every tcgen05.mma cost ~130 cycles per instruction -- reconvergence points, re-read kernel
      // parameters -- and that issue rate, not the tensor pipe or HBM, bounded every conv: r02_prof_conv3x3_*.txt.)
      // Loop bounds and strides live in registers, descriptors advance by 32-bit adds, the K steps are unrolled.
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
      const uint32_t dhi = a.desc_hi, idesc = a.idesc;
      const uint32_t w_lo = desc_lo(w_base), wblk16 = a.wblock_bytes >> 4;
      const uint32_t a_lo0 = desc_lo(a_base), stage16 = a.stage_bytes >> 4;
      const uint32_t N = (uint32_t)a.N;
      const uint32_t shift16 = ((uint32_t)a.hs_d * RB) >> 4;   // horizontal tap spacing in descriptor units
      const int kh = a.vr_kh, kw = a.vr_kw;
      const uint32_t tapstep16 = (uint32_t)nkb * wblk16;
      // all MT sub-tiles x KSTEPS of one (A window, weight block) pair; first == 0 starts a new accumulation
      auto issue = [&](uint32_t d, uint32_t al, uint32_t bl, uint32_t first) {
        for (int m = 0; m < MT; ++m, d += N, al += SUB16) {
          umma_bf16_lo(d, al, bl, dhi, idesc, first);
#pragma unroll
          for (int k = 1; k < KSTEPS; ++k) umma_bf16_lo(d, al + 2u * k, bl + 2u * k, dhi, idesc, 1u);
        }
      };
      int s = 0;
      uint32_t ph = 0, acc = 0, aph = 0;
      UnitIter<MODE> un;
      un.init(a, u_begin);
      for (int u = u_begin; u < u_end; u += u_step, un.next(a)) {
        if (un.len <= 0) continue;
        for (int i = 0; i < un.len; ++i) {
          mbar_wait(tempty0 + 8u * acc, aph ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * acc_cols;
          if (MODE == MODE_VREUSE) {
            // tap row t of this output lives in ring slot s+t (s = slot of the oldest live row)
            int st = s;
            uint32_t pt = ph;
            uint32_t bl = w_lo;
            for (int t = 0; t < kh; ++t) {
              if (i == 0 || t == kh - 1) mbar_wait(full0 + 8u * st, pt);
              uint32_t al = a_lo0 + (uint32_t)st * stage16;
              for (int q = 0; q < kw; ++q, al += shift16, bl += wblk16) issue(d_tmem, al, bl, (t | q) != 0 ? 1u : 0u);
              if (++st == S) { st = 0; pt ^= 1u; }
            }
            umma_commit(empty0 + 8u * s);   // the oldest row is dead once these MMAs retire
            if (++s == S) { s = 0; ph ^= 1u; }
            if (i == un.len - 1) {          // unit done: release the kh-1 rows still held
              for (int t = 1; t < kh; ++t) {
                umma_commit(empty0 + 8u * s);
                if (++s == S) { s = 0; ph ^= 1u; }
              }
            }
          } else if (MODE == MODE_HREUSE) {
            const int hkh = a.hr_kh, kwt = ntaps / hkh;
            for (int r = 0; r < hkh; ++r) {
              for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full0 + 8u * s, ph);
                uint32_t al = a_lo0 + (uint32_t)s * stage16;
                uint32_t bl = w_lo + (uint32_t)(r * kwt * nkb + kb) * wblk16;
                for (int t = 0; t < kwt; ++t, al += shift16, bl += tapstep16) issue(d_tmem, al, bl, (r | kb | t) != 0 ? 1u : 0u);
                umma_commit(empty0 + 8u * s);
                if (++s == S) { s = 0; ph ^= 1u; }
              }
            }
          } else {
            const int kiters = ntaps * nkb;
            uint32_t bl = w_lo;
            for (int ki = 0; ki < kiters; ++ki, bl += wblk16) {
              mbar_wait(full0 + 8u * s, ph);
              issue(d_tmem, a_lo0 + (uint32_t)s * stage16, bl, ki != 0 ? 1u : 0u);
              umma_commit(empty0 + 8u * s);  // frees the smem stage when these MMAs retire
              if (++s == S) { s = 0; ph ^= 1u; }
            }
          }
          umma_commit(tfull0 + 8u * acc);  // accumulators ready for the epilogue
          if (++acc == NA) { acc = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0 && a.staged && a.has_res) {
      // ---------------- residual producer: the residual tile lands in the staging buffer the
      // epilogue will overwrite in place with the output tile
      const uint32_t nsmask = (uint32_t)NS - 1u, nsshift = NS == 4 ? 2u : (NS == 2 ? 1u : 0u);
      uint32_t tc = 0;
      UnitIter<MODE> un;
      if (u_begin < u_end) un.init(a, u_begin);
      for (int u = u_begin; u < u_end; u += u_step, un.next(a)) {
        int h = un.h0;
        for (int i = 0; i < un.len; ++i, ++tc, h += un.hstep) {
          const uint32_t b = tc & nsmask, use = tc >> nsshift;
          mbar_wait(sfree0 + 8u * b, (use & 1u) ^ 1u);
          mbar_expect_tx(sfull0 + 8u * b, a.out_buf_bytes);
          const uint32_t dst = o_base + b * a.out_buf_bytes;
          for (int cb = 0; cb < a.ncb; ++cb)
            for (int q = 0; q < a.o_nbox; ++q)
              tma_load_4d(dst + (uint32_t)cb * a.out_block_bytes + (uint32_t)(q * a.o_boxw * a.cbo) * 2u, &a.tmR,
                          sfull0 + 8u * b, cb * a.cbo, un.w0 + q * a.o_boxw, h, un.n);
        }
      }
    }
  } else {
    // ---------------- epilogue warps: TMEM -> registers -> (staging smem -> TMA store | global)
    const int q = warp & 3;                     // TMEM lane quadrant this warp may access
    const int grp = (warp - kEpiWarp0) >> 2;    // 4 warps per quadrant split the (sub-tile, 16-column) work
    const int nchunk = a.N >> 4;
    const uint32_t row_bytes = (uint32_t)a.cbo * 2u;
    const uint32_t nsmask = (uint32_t)NS - 1u, nsshift = NS == 4 ? 2u : (NS == 2 ? 1u : 0u);
    const int staged = a.staged, has_res = a.has_res, act = a.ep.act, cout = a.cout, dual = a.dual;
    const uint32_t swz = a.out_swz_mask;
    const bool wide = a.ncb > 1;                 // several column blocks of cbo (16 / 32 / 64) channels
    const int cbo_shift = a.cbo == 64 ? 6 : (a.cbo == 32 ? 5 : 4);
    // this thread's work items: (sub-tile m, 16-column chunk) pairs dealt round-robin to the 4 warps of a quadrant
    int nitems = 0, it_c0[4] = {0, 0, 0, 0};
    uint32_t it_toff[4] = {0, 0, 0, 0}, it_soff[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    {
      int m = 0, chn = grp;
      while (chn >= nchunk) { chn -= nchunk; ++m; }
      while (m < MT && nitems < 4) {
        const int c0 = chn << 4;
        it_c0[nitems] = c0;
        it_toff[nitems] = (uint32_t)(m * a.N + c0);
        const int R = m * kTileM + q * 32 + lane;
        for (int h = 0; h < 2; ++h) {
          const int cb8 = c0 + 8 * h;
          const int blk = wide ? (cb8 >> cbo_shift) : 0;
          const int cin_blk = wide ? (cb8 & (a.cbo - 1)) : cb8;
          uint32_t off = (uint32_t)R * row_bytes + (uint32_t)cin_blk * 2u;
          off ^= ((off >> 7) & swz) << 4;
          it_soff[nitems][h] = (uint32_t)blk * a.out_block_bytes + off;
        }
        ++nitems;
        chn += kEpiThreads / 128;
        while (chn >= nchunk) { chn -= nchunk; ++m; }
      }
    }
    uint32_t tc = 0, acc = 0, aph = 0;
    UnitIter<MODE> un;
    if (u_begin < u_end) un.init(a, u_begin);
    for (int u = u_begin; u < u_end; u += u_step, un.next(a)) {
      int th0 = un.h0;
      for (int i = 0; i < un.len; ++i, ++tc, th0 += un.hstep) {
        const uint32_t b = staged ? (tc & nsmask) : 0u;
        const uint32_t use = staged ? (tc >> nsshift) : 0u;
        const uint32_t obuf = o_base + b * a.out_buf_bytes;
        if (staged) {
          if (has_res)
            mbar_wait(sfull0 + 8u * b, use & 1u);           // residual tile landed
          else
            mbar_wait(sfree0 + 8u * b, (use & 1u) ^ 1u);    // previous store out of this buffer drained
        }
        mbar_wait(tfull0 + 8u * acc, aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols;
        if (staged) {
          // ---- staged path: every thread's (sub-tile, 16-column) items are fixed for the whole kernel, so their
          // TMEM offsets, swizzled staging offsets and parameter addresses were computed once (it_*)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < nitems) {
              uint32_t r[16];
              tmem_ld16(taddr + it_toff[k], r);
              tmem_ld_wait();
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int cb8 = it_c0[k] + 8 * h;
                if (cb8 < cout) {
                  float f[8];
                  const uint32_t pa = prm_base + 4u * (uint32_t)cb8;
                  const float4 s0 = lds_f4(pa), s1 = lds_f4(pa + 16u);
                  const float4 h0 = lds_f4(pa + 1024u), h1 = lds_f4(pa + 1040u);
                  f[0] = fmaf(__uint_as_float(r[8 * h + 0]), s0.x, h0.x);
                  f[1] = fmaf(__uint_as_float(r[8 * h + 1]), s0.y, h0.y);
                  f[2] = fmaf(__uint_as_float(r[8 * h + 2]), s0.z, h0.z);
                  f[3] = fmaf(__uint_as_float(r[8 * h + 3]), s0.w, h0.w);
                  f[4] = fmaf(__uint_as_float(r[8 * h + 4]), s1.x, h1.x);
                  f[5] = fmaf(__uint_as_float(r[8 * h + 5]), s1.y, h1.y);
                  f[6] = fmaf(__uint_as_float(r[8 * h + 6]), s1.z, h1.z);
                  f[7] = fmaf(__uint_as_float(r[8 * h + 7]), s1.w, h1.w);
                  const uint32_t saddr = obuf + it_soff[k][h];
                  if (has_res) {
                    if (a.ep.pre_act) {   // ext = act(BN(conv)) before "main + ext" (ENet bottlenecks)
                      if (act == ESN_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                      } else if (act == ESN_ACT_PRELU) {
                        const float4 a0 = lds_f4(pa + 2048u), a1 = lds_f4(pa + 2064u);
                        const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = f[j] >= 0.f ? f[j] : f[j] * al[j];
                      }
                    }
                    float g[8];
                    bf16x8_to_float(lds128(saddr), g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += g[j];
                  }
                  if (act == ESN_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                  } else if (act == ESN_ACT_PRELU) {
                    const float4 a0 = lds_f4(pa + 2048u), a1 = lds_f4(pa + 2064u);
                    const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = f[j] >= 0.f ? f[j] : f[j] * al[j];
                  }
                  if (dual) {
                    // second stage on the value as stored (bf16): the same numbers a separate affine pass over y would read
                    const uint4 pk = float_to_bf16x8(f);
                    if (dual == 2) sts128(saddr, pk);
                    float g[8];
                    bf16x8_to_float(pk, g);
                    const uint32_t pb = prm2_base + 4u * (uint32_t)cb8;
                    const float4 s0 = lds_f4(pb), s1 = lds_f4(pb + 16u);
                    const float4 h0 = lds_f4(pb + 1024u), h1 = lds_f4(pb + 1040u);
                    g[0] = fmaf(g[0], s0.x, h0.x); g[1] = fmaf(g[1], s0.y, h0.y);
                    g[2] = fmaf(g[2], s0.z, h0.z); g[3] = fmaf(g[3], s0.w, h0.w);
                    g[4] = fmaf(g[4], s1.x, h1.x); g[5] = fmaf(g[5], s1.y, h1.y);
                    g[6] = fmaf(g[6], s1.z, h1.z); g[7] = fmaf(g[7], s1.w, h1.w);
                    if (a.act2 == ESN_ACT_RELU) {
#pragma unroll
                      for (int j = 0; j < 8; ++j) g[j] = fmaxf(g[j], 0.f);
                    } else if (a.act2 == ESN_ACT_PRELU) {
                      const float4 a0 = lds_f4(pb + 2048u), a1 = lds_f4(pb + 2064u);
                      const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                      for (int j = 0; j < 8; ++j) g[j] = g[j] >= 0.f ? g[j] : g[j] * al[j];
                    }
                    sts128(dual == 2 ? saddr + o2_delta : saddr, float_to_bf16x8(g));
                  } else {
                    sts128(saddr, float_to_bf16x8(f));
                  }
                }
              }
            }
          }
        } else {
          // work items (m, chunk) are dealt round-robin to the warps of a quadrant
          int m = 0, chn = grp;
          while (chn >= nchunk) { chn -= nchunk; ++m; }
          while (m < MT) {
            const int c0 = chn << 4;
            uint32_t r[16];
            tmem_ld16(taddr + (uint32_t)(m * a.N + c0), r);
            tmem_ld_wait();
            if (c0 < cout) {
              const int R = m * kTileM + q * 32 + lane;   // row of the (MT*128)-row tile
  #pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int cb8 = c0 + 8 * h;
                if (cb8 < cout) {
                  float f[8];
                  {
                    const uint32_t pa = prm_base + 4u * (uint32_t)cb8;
                    const float4 s0 = lds_f4(pa), s1 = lds_f4(pa + 16u);
                    const float4 h0 = lds_f4(pa + 1024u), h1 = lds_f4(pa + 1040u);
                    f[0] = fmaf(__uint_as_float(r[8 * h + 0]), s0.x, h0.x);
                    f[1] = fmaf(__uint_as_float(r[8 * h + 1]), s0.y, h0.y);
                    f[2] = fmaf(__uint_as_float(r[8 * h + 2]), s0.z, h0.z);
                    f[3] = fmaf(__uint_as_float(r[8 * h + 3]), s0.w, h0.w);
                    f[4] = fmaf(__uint_as_float(r[8 * h + 4]), s1.x, h1.x);
                    f[5] = fmaf(__uint_as_float(r[8 * h + 5]), s1.y, h1.y);
                    f[6] = fmaf(__uint_as_float(r[8 * h + 6]), s1.z, h1.z);
                    f[7] = fmaf(__uint_as_float(r[8 * h + 7]), s1.w, h1.w);
                  }
                  if (staged) {
                    const int blk = wide ? (cb8 >> cbo_shift) : 0;
                    const int cin_blk = wide ? (cb8 & (a.cbo - 1)) : cb8;
                    uint32_t off = (uint32_t)R * row_bytes + (uint32_t)cin_blk * 2u;
                    off ^= ((off >> 7) & swz) << 4;
                    const uint32_t saddr = obuf + (uint32_t)blk * a.out_block_bytes + off;
                    if (has_res) {
                      if (a.ep.pre_act) {   // ext = act(BN(conv)) before "main + ext" (ENet bottlenecks)
                        if (act == ESN_ACT_RELU) {
  #pragma unroll
                          for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                        } else if (act == ESN_ACT_PRELU) {
                          const uint32_t pa = prm_base + 2048u + 4u * (uint32_t)cb8;
                          const float4 a0 = lds_f4(pa), a1 = lds_f4(pa + 16u);
                          const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  #pragma unroll
                          for (int j = 0; j < 8; ++j) f[j] = f[j] >= 0.f ? f[j] : f[j] * al[j];
                        }
                      }
                      float g[8];
                      bf16x8_to_float(lds128(saddr), g);
  #pragma unroll
                      for (int j = 0; j < 8; ++j) f[j] += g[j];
                    }
                    if (act == ESN_ACT_RELU) {
  #pragma unroll
                      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                    } else if (act == ESN_ACT_PRELU) {
                      const uint32_t pa = prm_base + 2048u + 4u * (uint32_t)cb8;
                      const float4 a0 = lds_f4(pa), a1 = lds_f4(pa + 16u);
                      const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  #pragma unroll
                      for (int j = 0; j < 8; ++j) f[j] = f[j] >= 0.f ? f[j] : f[j] * al[j];
                    }
                    sts128(saddr, float_to_bf16x8(f));
                  } else {
                    // fallback: per-thread global stores (output channel count not a multiple of 8)
                    const int ri = R / a.bw, rj = R - ri * a.bw;
                    const int gi = th0 + ri, gj = un.w0 + rj;
                    if (gi < a.gh && gj < a.gw) {
                      const size_t opix =
                          ((size_t)un.n * a.Hy + (size_t)(gi * a.sy + a.oy)) * a.Wy + (size_t)(gj * a.sx + a.ox);
                      __nv_bfloat16* yp = a.y + opix * a.y_cs;
                      for (int j = 0; j < 8; ++j) {
                        const int c = cb8 + j;
                        if (c < cout) {
                          float v = f[j];
                          if (a.ep.res && a.ep.pre_act) v = apply_act(v, act, prm[512 + c]);
                          if (a.ep.res)
                            v += __bfloat162float(
                                reinterpret_cast<const __nv_bfloat16*>(a.ep.res)[opix * a.ep.res_cstride + c]);
                          v = apply_act(v, act, prm[512 + c]);
                          yp[c] = __float2bfloat16_rn(v);
                        }
                      }
                    }
                  }
                }
              }
            }
            chn += kEpiThreads / 128;
            while (chn >= nchunk) { chn -= nchunk; ++m; }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);   // accumulators drained (one arrival per warp)
        if (++acc == NA) { acc = 0; aph ^= 1u; }
        if (staged) {
          fence_proxy_async();          // my st.shared writes -> visible to the TMA (async proxy)
          epi_bar_sync();
          if (threadIdx.x == kEpiWarp0 * 32) {
            for (int cb = 0; cb < a.ncb; ++cb) {
              // phase-fused transposed conv: column blocks [0, bpa) are output rows 2i, [bpa, 2 bpa) rows 2i+1
              const bool odd = a.shuf_bpa && cb >= a.shuf_bpa;
              const int cc = (a.shuf_bpa ? (cb - (odd ? a.shuf_bpa : 0)) : cb) * a.cbo;
              for (int qb = 0; qb < a.o_nbox; ++qb)
                tma_store_4d(odd ? &a.tmY2 : &a.tmY, obuf + (uint32_t)cb * a.out_block_bytes + (uint32_t)(qb * a.o_boxw) * row_bytes,
                             cc, un.w0 + qb * a.o_boxw, th0, un.n);
              if (dual == 2)
                for (int qb = 0; qb < a.o_nbox; ++qb)
                  tma_store_4d(&a.tmZ, obuf + o2_delta + (uint32_t)cb * a.out_block_bytes + (uint32_t)(qb * a.o_boxw) * row_bytes,
                               cc, un.w0 + qb * a.o_boxw, th0, un.n);
            }
            tma_store_commit();
            // A buffer is handed back as soon as its store has read it -- the store of the PREVIOUS tile, so this thread
            // never waits on the one it just issued -- and not just in time for the tile that reuses it: the residual
            // producer then runs NS-2 tiles ahead of the epilogue instead of issuing each residual load (an HBM round
            // trip) when the epilogue is already waiting for it.
            if (NS == 1) {
              tma_store_wait_read<0>();
              mbar_arrive(sfree0);
            } else {
              tma_store_wait_read<1>();
              if (tc >= 1u) mbar_arrive(sfree0 + 8u * ((tc - 1u) & nsmask));
            }
          }
        }
      }
    }
    if (staged && threadIdx.x == kEpiWarp0 * 32) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();   // the issuer role ran on one lane: reconverge before the .sync.aligned instruction
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
int floordiv2(int t) { return (t >= 0) ? t / 2 : -((-t + 1) / 2); }

struct DeviceLimits {
  int sms = 0;
  int max_smem = 0;
};
const DeviceLimits& limits() {
  if (esn_dry_run()) {      // nominal B200: 148 SMs, 227 KB opt-in shared memory minus the kernel's static part
    static const DeviceLimits nominal = {148, 232448 - 1024};
    return nominal;
  }
  static DeviceLimits ls[kEsnMaxDevices];
  static std::once_flag once[kEsnMaxDevices];
  const int dev = esn_current_device();
  std::call_once(once[dev], [dev] {
    DeviceLimits& l = ls[dev];
    cudaDeviceGetAttribute(&l.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&l.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, conv_umma_kernel<64, MODE_GENERIC>) == cudaSuccess) l.max_smem -= (int)fa.sharedSizeBytes;
#define ESN_SET_SMEM(KBv, Mv) cudaFuncSetAttribute(conv_umma_kernel<KBv, Mv>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.max_smem)
    ESN_SET_SMEM(16, MODE_GENERIC); ESN_SET_SMEM(16, MODE_HREUSE); ESN_SET_SMEM(16, MODE_VREUSE);
    ESN_SET_SMEM(32, MODE_GENERIC); ESN_SET_SMEM(32, MODE_HREUSE); ESN_SET_SMEM(32, MODE_VREUSE);
    ESN_SET_SMEM(64, MODE_GENERIC); ESN_SET_SMEM(64, MODE_HREUSE); ESN_SET_SMEM(64, MODE_VREUSE);
#undef ESN_SET_SMEM
  });
  return ls[dev];
}

}  // namespace

namespace {
// dual: nullptr, or the second epilogue stage (esn_conv2d_umma_dual)
int conv_umma_impl(const EsnConv* p, const EsnConvDual* dual, void* stream) {
  if (!p || !p->w) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  // store_y == 0: only the second-stage value is stored, so every output map is built over y2
  const EsnTensor& y = (dual && !dual->store_y) ? dual->y2 : p->y;
  if (!esn_valid_nhwc(x) || !esn_valid_nhwc(y)) return ESN_ERR_BAD_ARG;
  if (dual) {
    const EsnTensor& z = dual->y2;
    if (!esn_valid_nhwc(z) || z.dtype != ESN_BF16 || z.n != p->y.n || z.h != p->y.h || z.w != p->y.w || z.c != p->y.c)
      return ESN_ERR_BAD_SHAPE;
    if (z.c_stride % 8 || ((uintptr_t)z.ptr % 16)) return ESN_ERR_ALIGN;
    if (dual->act2 != ESN_ACT_NONE && dual->act2 != ESN_ACT_RELU && dual->act2 != ESN_ACT_PRELU) return ESN_ERR_BAD_ARG;
    if (dual->act2 == ESN_ACT_PRELU && !dual->alpha2) return ESN_ERR_BAD_ARG;
    if (p->transposed) return ESN_ERR_UNSUPPORTED;
  }
  if (x.dtype != ESN_BF16 || y.dtype != ESN_BF16 || p->groups != 1) return ESN_ERR_UNSUPPORTED;
  const EsnTensor& res = p->ep.residual;
  if (res.ptr && res.dtype != ESN_BF16) return ESN_ERR_UNSUPPORTED;
  int rc = esn_check_epilogue(p->ep, y);
  if (rc) return rc;
  const bool fused = p->transposed == 2;   // phase-fused ConvTranspose2d(3, s2, p1, op1): 2x2 taps, 4*C outputs, pixel-shuffle store
  const bool tr = p->transposed == 1;
  const int Cin = x.c, Cout = fused ? 4 * y.c : y.c, N = p->cout_pad;
  if (N % 16 || N < 16 || N > 256 || N < Cout) return ESN_ERR_UNSUPPORTED;
  int KB;
  if (Cin == 16 || Cin == 32 || Cin == 64)
    KB = Cin;
  else if (Cin % 64 == 0)
    KB = 64;
  else if (Cin % 32 == 0)
    KB = 32;       // 96 / 160-channel inputs (Fast-SCNN's 64 -> 96 -> 128 bottlenecks, FastSCNN.py:134): 64-byte swizzled K blocks
  else if (Cin % 16 == 0)
    KB = 16;       // 48 / 80 / 112 ...: 32-byte swizzled K blocks (Fast-SCNN's 32 -> 48 -> 64 learning-to-downsample stage)
  else
    return ESN_ERR_UNSUPPORTED;
  const int nkb = Cin / KB;
  if (x.c_stride % 8 || y.c_stride % 8 || ((uintptr_t)x.ptr % 16) || ((uintptr_t)y.ptr % 16) ||
      ((uintptr_t)p->w % 16))
    return ESN_ERR_ALIGN;
  if (res.ptr && (res.c_stride % 8 || ((uintptr_t)res.ptr % 16))) return ESN_ERR_ALIGN;
  if (x.n != y.n) return ESN_ERR_BAD_SHAPE;
  const int ntaps_all = p->kh * p->kw;
  if (ntaps_all > kMaxTaps) return ESN_ERR_UNSUPPORTED;
  if (p->stride != 1 && p->stride != 2) return ESN_ERR_UNSUPPORTED;
  if (tr && (p->stride != 2 || p->dil_h != 1 || p->dil_w != 1)) return ESN_ERR_UNSUPPORTED;
  if (fused) {
    if (p->kh != 2 || p->kw != 2 || p->stride != 1 || p->pad_h || p->pad_w || p->dil_h != 1 || p->dil_w != 1 || res.ptr ||
        y.h != 2 * x.h || y.w != 2 * x.w || y.c_stride != y.c || N != Cout || (y.c != 8 && y.c != 16 && y.c != 32 && y.c != 64))
      return ESN_ERR_UNSUPPORTED;
  } else if (!tr) {
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

  // iteration grid: output positions (conv) or input positions (one transposed-conv phase)
  const int gh = (tr || fused) ? x.h : y.h, gw = (tr || fused) ? x.w : y.w;
  a.gh = gh;
  a.gw = gw;

  // M sub-tiles per pipeline stage: keep every stage ~16 KB so small-C layers amortise the
  // per-tile barrier round trips over 4x (C=16) / 2x (C=32) more pixels
  int MT = 64 / KB;
  while (MT > 1 && MT * N > 256) MT >>= 1;

  // ---- mode: tap reuse needs row tiles (one image row segment per tile)
  static const bool no_reuse = getenv("ESN_UMMA_NOREUSE") != nullptr;
  const bool rowable = !no_reuse && gw >= 128 && !tr && p->stride == 1;
  a.mode = MODE_GENERIC;
  if (rowable && p->kh == 1 && p->kw >= 2 && (p->kw - 1) * p->dil_w <= 256)
    a.mode = MODE_HREUSE;
  else if (rowable && p->kh >= 2 && nkb == 1 && (p->kw - 1) * p->dil_w <= 64 && !getenv("ESN_UMMA_NOHV")) {
    // k x 1, and k_h x k_w (3x3): row ring x shifted windows, every input row loaded once -- if the ring (k_h + 1
    // row windows) fits next to the resident weights and one staging tile; otherwise one box per tap (generic)
    int mt = MT;
    while (mt > 1 && mt * kTileM > gw) mt >>= 1;
    const uint32_t stage = ((uint32_t)(mt * kTileM + (p->kw - 1) * p->dil_w) * (uint32_t)(KB * 2) + 1023u) & ~1023u;
    const uint32_t wreg = ((uint32_t)ntaps_all * N * KB * 2 + 1023u) & ~1023u;
    const bool stg = (Cout % 8 == 0) && (Cout <= 64 || Cout % 64 == 0);
    const uint32_t need = wreg + (uint32_t)(p->kh + 1) * stage + (stg ? (uint32_t)mt * kTileM * Cout * 2 : 0u) + 3072u + 512u + 1024u;
    if (need <= (uint32_t)lim.max_smem) a.mode = MODE_VREUSE;
  }
  if (a.mode == MODE_GENERIC && rowable && p->kh >= 2 && p->kw >= 2 && (p->kw - 1) * p->dil_w <= 256 &&
      !getenv("ESN_UMMA_NOHROWS")) {
    // k_h x k_w convs whose row ring does not fit (two K blocks, or weights that fill the shared memory): row tiles with one
    // window per (tap row, K block); the k_w taps of a row are shifted reads of it -- k_w times less L2 -> shared-memory
    // traffic than one box per tap.  Needs two windows, one staging tile and the weights.
    int mt = MT;
    while (mt > 1 && mt * kTileM > gw) mt >>= 1;
    const uint32_t stage = ((uint32_t)(mt * kTileM + (p->kw - 1) * p->dil_w) * (uint32_t)(KB * 2) + 1023u) & ~1023u;
    const uint32_t wreg = ((uint32_t)ntaps_all * nkb * N * KB * 2 + 1023u) & ~1023u;
    const bool stg = (Cout % 8 == 0) && (Cout <= 64 || Cout % 64 == 0);
    const uint32_t need = wreg + 3u * stage + (stg ? (uint32_t)mt * kTileM * Cout * 2 : 0u) + 3072u + 512u + 1024u +
                          (dual ? 4096u + (dual->store_y ? (uint32_t)mt * kTileM * Cout * 2 : 0u) : 0u);
    if (need <= (uint32_t)lim.max_smem) a.mode = MODE_HREUSE;
  }
  if (a.mode != MODE_GENERIC) {
    if (a.mode == MODE_HREUSE && p->kh == 1 && KB == 64 && nkb == 1 && N <= 64) MT = 2;   // 256-pixel row tiles for C=64
    while (MT > 1 && MT * kTileM > gw) MT >>= 1;
    a.bw = MT * kTileM;
    a.bh = 1;
    a.a_boxw = a.bw > 256 ? 256 : a.bw;
    a.a_nbox = a.bw / a.a_boxw;
  } else {
    int bw;
    if (gw > 128 && MT >= 2) {
      bw = 256;
    } else if (gw > 64) {
      bw = 128;
    } else {
      bw = 8;
      while (bw < gw) bw <<= 1;
    }
    a.bw = bw;
    a.bh = MT * kTileM / bw;
    if (a.bh > 256) return ESN_ERR_UNSUPPORTED;
    a.a_boxw = a.bw;
    a.a_nbox = 1;
  }
  a.o_boxw = a.a_boxw;
  a.o_nbox = a.a_nbox;
  a.MT = MT;
  const int rows = MT * kTileM;
  a.wblock_bytes = N * row_bytes;
  a.tiles_w = esn_cdiv(gw, a.bw);
  a.tiles_h = esn_cdiv(gh, a.bh);
  if (a.mode == MODE_HREUSE) {
    a.hs_d = p->dil_w;
    a.hs_pad = p->pad_w;
    a.hr_kh = p->kh;
    a.hr_dh = p->dil_h;
    a.hr_ph = p->pad_h;
    const int extra = (p->kw - 1) * p->dil_w;
    a.load_bytes = (uint32_t)(a.bw + extra) * row_bytes;
    a.stage_bytes = (a.load_bytes + 1023u) & ~1023u;
    a.nunits = x.n * a.tiles_w * a.tiles_h;
  } else if (a.mode == MODE_VREUSE) {
    a.vr_d = p->dil_h;
    a.vr_pad = p->pad_h;
    const int cnt = esn_cdiv(gh, a.vr_d);
    int L = 32;
    while (L > 4 && (long long)x.n * a.tiles_w * a.vr_d * esn_cdiv(cnt, L) < 3LL * lim.sms) L >>= 1;
    a.vr_L = L;
    a.vr_nseg = esn_cdiv(cnt, L);
    a.vr_cnt = cnt;
    a.vr_rem = gh - (cnt - 1) * a.vr_d;
    a.vr_kh = p->kh;
    a.vr_kw = p->kw;
    a.hs_d = p->dil_w;
    a.hs_pad = p->pad_w;
    a.load_bytes = (uint32_t)(rows + (p->kw - 1) * p->dil_w) * row_bytes;
    a.stage_bytes = (a.load_bytes + 1023u) & ~1023u;
    a.nunits = x.n * a.tiles_w * a.vr_d * a.vr_nseg;
  } else {
    a.load_bytes = (uint32_t)rows * row_bytes;
    a.stage_bytes = a.load_bytes;
    a.nunits = x.n * a.tiles_w * a.tiles_h;
  }
  a.y = reinterpret_cast<__nv_bfloat16*>(y.ptr);
  a.Hy = y.h;
  a.Wy = y.w;
  a.y_cs = y.c_stride;
  a.ep = make_epi(p->ep);

  // staged epilogue (swizzled smem tile + TMA store, residual prefetched by TMA into the same tile)
  a.staged = (Cout % 8 == 0) && (Cout <= 64 || Cout % 64 == 0);
  a.has_res = res.ptr != nullptr;
  if (dual) {
    if (!a.staged) return ESN_ERR_UNSUPPORTED;     // the second stage lives in the staged epilogue only
    a.dual = dual->store_y ? 2 : 1;
    a.act2 = dual->act2;
    a.scale2 = dual->scale2;
    a.shift2 = dual->shift2;
    a.alpha2 = dual->alpha2;
  }
  CUtensorMapSwizzle oswz = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (a.staged) {
    a.cbo = Cout <= 64 ? Cout : 64;
    if (fused) a.cbo = 2 * y.c <= 64 ? 2 * y.c : 64;      // one column block never straddles the two output-row parities
    a.ncb = Cout / a.cbo;
    a.shuf_bpa = fused ? (2 * y.c) / a.cbo : 0;
    a.out_block_bytes = (uint32_t)rows * a.cbo * 2;
    a.out_buf_bytes = a.ncb * a.out_block_bytes;
    const int rb = a.cbo * 2;
    if (rb == 128) { oswz = CU_TENSOR_MAP_SWIZZLE_128B; a.out_swz_mask = 7; }
    else if (rb == 64) { oswz = CU_TENSOR_MAP_SWIZZLE_64B; a.out_swz_mask = 3; }
    else if (rb == 32) { oswz = CU_TENSOR_MAP_SWIZZLE_32B; a.out_swz_mask = 1; }
    else { a.out_swz_mask = 0; }
    a.NS = a.out_buf_bytes <= 16384 ? 4 : 2;
  } else {
    if (fused) return ESN_ERR_UNSUPPORTED;
    a.cbo = 8;  // unused
    a.NS = 0;
  }

  // ---- activation tensor maps (5-D): main box, and the (k-1)*d-pixel tail box of the hreuse window
  for (int which = 0; which < ((a.mode == MODE_HREUSE || (a.mode == MODE_VREUSE && p->kw > 1)) ? 2 : 1); ++which) {
    const cuuint64_t cs = (cuuint64_t)x.c_stride;
    cuuint64_t dims[5], strides[4];
    if (!tr && p->stride == 2) {
      dims[0] = 2 * cs; dims[1] = x.w / 2; dims[2] = 2; dims[3] = x.h / 2; dims[4] = x.n;
      strides[0] = 2 * cs * 2; strides[1] = (cuuint64_t)x.w * cs * 2; strides[2] = 2 * (cuuint64_t)x.w * cs * 2;
      strides[3] = (cuuint64_t)x.h * x.w * cs * 2;
    } else {
      dims[0] = (cuuint64_t)x.c; dims[1] = x.w; dims[2] = 1; dims[3] = x.h; dims[4] = x.n;
      strides[0] = cs * 2; strides[1] = (cuuint64_t)x.w * cs * 2; strides[2] = (cuuint64_t)x.w * cs * 2;
      strides[3] = (cuuint64_t)x.h * x.w * cs * 2;
    }
    const cuuint32_t bwid = which ? (cuuint32_t)((p->kw - 1) * p->dil_w) : (cuuint32_t)a.a_boxw;
    const cuuint32_t box[5] = {(cuuint32_t)KB, bwid, 1, (cuuint32_t)a.bh, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (encode(which ? &a.tmAh : &a.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.ptr, dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
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
  const int nphase = tr ? 4 : 1;
  for (int ph = 0; ph < nphase; ++ph) {
    const int pa = ph >> 1, pb = ph & 1;
    int nt = 0;
    for (int r = 0; r < p->kh; ++r)
      for (int s = 0; s < p->kw; ++s) {
        int dy, dx, par = 0, coff = 0;
        if (tr) {
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
    a.sy = tr ? 2 : 1;
    a.sx = a.sy;
    a.oy = tr ? pa : 0;
    a.ox = tr ? pb : 0;

    if (a.staged && fused) {
      // output seen as (2C, W, H, N) per output-row parity: y[2i+a][2j+b][c] = D[i][j][(a, b, c)]
      for (int par2 = 0; par2 < 2; ++par2) {
        const cuuint64_t C2 = 2 * (cuuint64_t)y.c;
        const cuuint64_t dims[4] = {C2, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)x.n};
        const cuuint64_t strides[3] = {C2 * 2, 2 * (cuuint64_t)y.w * y.c * 2, (cuuint64_t)y.h * y.w * y.c * 2};
        const cuuint32_t box[4] = {(cuuint32_t)a.cbo, (cuuint32_t)a.o_boxw, (cuuint32_t)a.bh, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        void* bp = reinterpret_cast<uint8_t*>(y.ptr) + (size_t)par2 * y.w * y.c * 2;
        if (encode(par2 ? &a.tmY2 : &a.tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, bp, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, oswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return ESN_ERR_CUDA;
      }
    } else if (a.staged) {  // output / residual (/ second output) maps over this phase's (strided) output positions
      for (int which = 0; which < 3; ++which) {
        if ((which == 1 && !a.has_res) || (which == 2 && a.dual != 2)) continue;
        const EsnTensor& t = which == 2 ? dual->y2 : (which ? res : y);
        const cuuint64_t cs = (cuuint64_t)t.c_stride;
        const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)x.n};
        const cuuint64_t strides[3] = {cs * 2 * a.sx, (cuuint64_t)y.w * cs * 2 * a.sy, (cuuint64_t)y.h * y.w * cs * 2};
        const cuuint32_t box[4] = {(cuuint32_t)a.cbo, (cuuint32_t)a.o_boxw, (cuuint32_t)a.bh, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        void* bp = reinterpret_cast<uint8_t*>(t.ptr) + ((size_t)a.oy * y.w + a.ox) * cs * 2;
        if (encode(which == 2 ? &a.tmZ : (which ? &a.tmR : &a.tmY), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, bp, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, oswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          return ESN_ERR_CUDA;
      }
    }

    const uint32_t wbytes = (uint32_t)nt * nkb * a.wblock_bytes;
    a.w_region_bytes = (wbytes + 1023u) & ~1023u;
    // shared memory plan: resident weights + A ring + staging + params + barriers (+1 KB alignment slack)
    // trade staging buffers for A stages until the load pipeline is deep enough to cover ~2 tiles
    const int min_stages = a.mode == MODE_VREUSE ? p->kh + 1 : 2;
    const int per_tile = a.mode == MODE_VREUSE ? 1 : (a.mode == MODE_HREUSE ? p->kh * nkb : nt * nkb);
    int want = a.mode == MODE_VREUSE ? p->kh + 3 : (a.mode == MODE_HREUSE ? per_tile + 1 : 2 * per_tile + 1);
    if (want > 8) want = 8;
    // the residual tile is prefetched into the staging buffer its output leaves from, so with a residual the number of
    // buffers is the number of residual loads in flight: 2 left a 256-pixel tile waiting a full HBM round trip per tile
    const char* ns_env = getenv("ESN_UMMA_NS");            // A/B: "2" restores the old plan
    const bool ns4 = a.has_res && !(ns_env && ns_env[0] == '2');
    int ns = a.staged ? ((a.out_buf_bytes <= 16384 || ns4) ? 4 : 2) : 0;
    // second epilogue stage: its parameter block (+ alignment) and, with both outputs stored, a second staging set
    const auto dual_extra = [&](int nsb) -> uint32_t {
      return a.dual ? 1024u + 3072u + (a.dual == 2 ? (uint32_t)nsb * a.out_buf_bytes : 0u) : 0u;
    };
    for (;;) {
      const uint32_t fixed = a.w_region_bytes + (uint32_t)ns * a.out_buf_bytes + 3072u + 512u + 1024u + dual_extra(ns);
      int stages = fixed < (uint32_t)lim.max_smem ? (int)(((uint32_t)lim.max_smem - fixed) / a.stage_bytes) : 0;
      if (stages > 8) stages = 8;
      if (stages >= want || ns <= 1) {
        if (stages < min_stages) return ESN_ERR_UNSUPPORTED;
        a.stages = stages;
        a.NS = ns;
        break;
      }
      ns >>= 1;
    }
    int na = 4;
    while (na > 2 && na * MT * N > 512) na >>= 1;
    if (na * MT * N > 512) return ESN_ERR_UNSUPPORTED;
    a.NA = na;
    uint32_t cols = 32;
    while (cols < (uint32_t)(na * MT * N)) cols <<= 1;
    a.tmem_cols = cols;
    const size_t smem = a.w_region_bytes + (size_t)a.NS * a.out_buf_bytes + 3072u + 512u + 1024u +
                        (size_t)a.stages * a.stage_bytes + dual_extra(a.NS);
    int grid = lim.sms;
    if (grid > a.nunits) grid = a.nunits;
    {
      // measured on one box (profiles/r02_issue_ab.txt): the one-thread loop is 1.1-1.45x faster for N <= 64 and equal for
      // N = 128; ESN_UMMA_ISSUE=warp selects the old form for A/B runs (read per call)
      const char* force = getenv("ESN_UMMA_ISSUE");
      a.warp_issue = force ? (force[0] == 'w') : 0;
    }
    a.g_dw = grid % a.tiles_w;
    a.g_dh = (grid / a.tiles_w) % a.tiles_h;
    a.g_dn = grid / (a.tiles_w * a.tiles_h);
    if (esn_dry_run()) continue;      // planned and accepted; nothing is launched
#define ESN_LAUNCH(KBv, Mv) conv_umma_kernel<KBv, Mv><<<grid, kThreads, smem, st>>>(a)
#define ESN_LAUNCH_KB(KBv)                                                   \
  do {                                                                       \
    if (a.mode == MODE_HREUSE) ESN_LAUNCH(KBv, MODE_HREUSE);                 \
    else if (a.mode == MODE_VREUSE) ESN_LAUNCH(KBv, MODE_VREUSE);            \
    else ESN_LAUNCH(KBv, MODE_GENERIC);                                      \
  } while (0)
    if (KB == 64) ESN_LAUNCH_KB(64);
    else if (KB == 32) ESN_LAUNCH_KB(32);
    else ESN_LAUNCH_KB(16);
#undef ESN_LAUNCH_KB
#undef ESN_LAUNCH
    ESN_CHECK_LAUNCH();
  }
  return ESN_OK;
}
}  // namespace

extern "C" int esn_conv2d_umma(const EsnConv* p, void* stream) { return conv_umma_impl(p, nullptr, stream); }

extern "C" int esn_conv2d_umma_dual(const EsnConvDual* p, void* stream) {
  if (!p) return ESN_ERR_BAD_ARG;
  return conv_umma_impl(&p->conv, p, stream);
}
