// tcgen05 weight gradient of a dense stride-1 convolution (bf16 operands, fp32 accumulation in TMEM), all taps of a
// filter row (or of the whole filter) from ONE pass over the operands:
//     dW[r][s][ci][co] = sum_p  X[p + (r d_h - pad_h, s d_w - pad_w)][ci] * dY[p][co]
// i.e. D(M = (tap, ci), N = co) = A(M x K) . B(K x N) with the pixel index as the GEMM K dimension.  Both operands are
// "MN-major" exactly as they lie in NHWC memory (channels contiguous, one pixel per 32/64/128-byte row), so a TMA box
// [rows][pixels][channels] with the matching swizzle IS the canonical MN-major UMMA layout (SBO = 8 pixel rows) and no
// transposition happens anywhere.  The horizontal taps are descriptor start addresses shifted by whole pixel rows
// (legal for any row shift); for Cin <= 64 the 128 M rows of one MMA are 128/Cin taps at once (the
// "leading byte offset" between M blocks is one dilated pixel step, so the blocks overlap in shared memory).
// Replaces the weight branch of aten::convolution_backward for DABNet / ERFNet / Fast-SCNN / ESPNetv2 dense convs
// (train.py:353 loss.backward()).  HBM-bound: |x| + |dy| read once from DRAM; the kh-fold re-read of x is L2 traffic.
//
// One persistent CTA per SM (x tap-row groups when the accumulators would not fit 512 TMEM columns):
//   warp 0  TMA producer: per unit (BH x BW output pixels) one dY tile + one X tile per tap row -> smem ring
//   warp 1  MMA issuer: BH*BW/16 K steps x nacc accumulators per stage; the accumulators stay in TMEM over ALL units
//   warps 2-5  epilogue, once: tcgen05.ld -> fp32 red.add into dW[tap][ci][co]
#include "esn_umma_ptx.cuh"

namespace {

constexpr int kWguThreads = 192;
constexpr int kMaxAcc = 16;

struct alignas(64) WguArgs {
  CUtensorMap tmX;   // x : (C, W, H, N), box (cbx, XW, BH, 1)
  CUtensorMap tmG;   // dy: (C, W, H, N), box (cbg, BW, BH, 1)
  float* dw;
  int Cin, Cout, N, kh, kw;
  int nunits, tiles_w, tiles_h;
  int BH, BW, XW;
  int pad_h, pad_w, dil_h;
  int stages, tmem_cols;
  int nxbox, ngbox, nrt, nacc, small, blk;
  uint32_t xbox_alloc, xtile_alloc, gbox_bytes, gtile_bytes, stage_bytes, load_bytes;
  uint32_t xrow_bytes, grow_bytes;
  uint32_t adesc_hi, bdesc_hi, a_lbo, b_lbo, idesc;
  int loop_s, loop_jb;          // MMA issue loops per filter row: taps / tap groups x channel-box pairs
  uint32_t step_s, step_jb;     // their A start-address steps (bytes)
  int acc_rt[kMaxAcc];          // X row tile (tap row relative to the CTA's first)
  uint32_t acc_off[kMaxAcc];    // byte offset of the A start inside that tile
  int acc_s0[kMaxAcc], acc_ci0[kMaxAcc];
};

__device__ __forceinline__ void umma_mn(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kWguThreads, 1) wgrad_umma_kernel(const __grid_constant__ WguArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int S = a.stages;
  const uint32_t bar_base = base + (uint32_t)S * a.stage_bytes + 4096u;   // 4 KB slack: junk M blocks read past the last tile
  const uint32_t full0 = bar_base, empty0 = bar_base + 8u * S, tfull = bar_base + 16u * S, tmem_slot = tfull + 8u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmX);
    tma_prefetch_desc(&a.tmG);
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    mbar_init(tfull, 1);
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

  const int u_begin = (int)(((long long)a.nunits * blockIdx.x) / gridDim.x);
  const int u_end = (int)(((long long)a.nunits * (blockIdx.x + 1)) / gridDim.x);
  const int r0 = blockIdx.y * a.nrt;   // first tap row of this CTA

  if (warp == 0) {
    if (u_begin < u_end) {
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0;
      for (int u = u_begin; u < u_end; ++u) {
        const int tw = u % a.tiles_w, t1 = u / a.tiles_w;
        const int th = t1 % a.tiles_h, n = t1 / a.tiles_h;
        const int w0 = tw * a.BW, h0 = th * a.BH;
        mbar_wait(empty0 + 8u * s, ph ^ 1u);
        if (leader) {
          const uint32_t bar = full0 + 8u * s;
          mbar_expect_tx(bar, a.load_bytes);
          const uint32_t dst = base + (uint32_t)s * a.stage_bytes;
          for (int jb = 0; jb < a.ngbox; ++jb) tma_load_4d(dst + (uint32_t)jb * a.gbox_bytes, &a.tmG, bar, jb * 64, w0, h0, n);
          for (int rt = 0; rt < a.nrt; ++rt)
            for (int jb = 0; jb < a.nxbox; ++jb)
              tma_load_4d(dst + a.gtile_bytes + (uint32_t)rt * a.xtile_alloc + (uint32_t)jb * a.xbox_alloc, &a.tmX, bar, jb * 64,
                          w0 - a.pad_w, h0 - a.pad_h + (r0 + rt) * a.dil_h, n);
        }
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (u_begin < u_end && elect_one()) {
      // MMA issuer: runtime loops whose descriptor words advance with 32-bit adds (no table look-ups, no predicated-off
      // issue slots: ncu showed the first two versions issue-bound at ~300 cycles per MMA), run by ONE elected thread
      // (a warp-uniform loop with the elected lane branching around each tcgen05.mma still cost ~130 cycles per
      // instruction: see esn_umma.cu).  Accumulator order = (filter row, tap / tap group, channel-box pair), as in the
      // host's table.
      constexpr bool leader = true;
      int s = 0;
      uint32_t ph = 0, accum = 0;
      const int ksteps = a.BW >> 4, BH = a.BH, n_rt = a.nrt, n_s = a.loop_s, n_jb = a.loop_jb;
      const uint32_t a_hi = a.adesc_hi, b_hi = a.bdesc_hi, idesc = a.idesc, N = (uint32_t)a.N;
      const uint32_t a_lbo_bits = ((a.a_lbo >> 4) & 0x3FFFu) << 16, b_lbo_bits = ((a.b_lbo >> 4) & 0x3FFFu) << 16;
      const uint32_t kstep_a = a.xrow_bytes, kstep_b = a.grow_bytes;                       // 16 pixel rows, in 16-byte units
      const uint32_t row_a = ((uint32_t)a.XW * a.xrow_bytes) >> 4, row_b = ((uint32_t)a.BW * a.grow_bytes) >> 4;
      const uint32_t step_rt = a.xtile_alloc >> 4, step_s = a.step_s >> 4, step_jb = a.step_jb >> 4;
      const uint32_t stage16 = a.stage_bytes >> 4, gt16 = a.gtile_bytes >> 4;
      const uint32_t base16 = (base & 0x3FFFFu) >> 4;
      for (int u = u_begin; u < u_end; ++u) {
        mbar_wait(full0 + 8u * s, ph);
        tc_fence_after();
        const uint32_t g_lo = (base16 + (uint32_t)s * stage16) | b_lbo_bits;
        const uint32_t x_lo = (base16 + (uint32_t)s * stage16 + gt16) | a_lbo_bits;
        for (int hh = 0; hh < BH; ++hh) {
          uint32_t a_k = x_lo + (uint32_t)hh * row_a, b_k = g_lo + (uint32_t)hh * row_b;
          for (int kk = 0; kk < ksteps; ++kk, a_k += kstep_a, b_k += kstep_b) {
            uint32_t d = tmem_base, a_rt = a_k;
            for (int rt = 0; rt < n_rt; ++rt, a_rt += step_rt) {
              uint32_t a_s = a_rt;
              for (int ts = 0; ts < n_s; ++ts, a_s += step_s) {
                uint32_t a_j = a_s;
                for (int jb = 0; jb < n_jb; ++jb, a_j += step_jb, d += N)
                  if (leader) umma_mn(d, a_j, a_hi, b_k, b_hi, idesc, accum);
              }
            }
            accum = 1u;
          }
        }
        if (leader) umma_commit(empty0 + 8u * s);   // the stage is free once these MMAs have read it
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(tfull);
    }
  } else {
    if (u_begin < u_end) {
      mbar_wait(tfull, 0);
      tc_fence_after();
      const int quad = warp & 3;            // a warp may only touch its own TMEM lane quadrant
      const int m = quad * 32 + lane;
      const bool vec4 = (a.Cout & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.dw) & 15) == 0);
      for (int q = 0; q < a.nacc; ++q) {
        const int sft = a.small ? m / a.blk : 0;
        const int s = a.acc_s0[q] + sft;
        const int ci = a.acc_ci0[q] + (a.small ? m - sft * a.blk : m);
        const bool valid = s < a.kw && ci < a.Cin;
        float* dst = a.dw + ((size_t)((r0 + a.acc_rt[q]) * a.kw + s) * a.Cin + ci) * a.Cout;
        for (int c0 = 0; c0 < a.N; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(q * a.N + c0), v);
          tmem_ld_wait();
          if (valid) {
            if (vec4 && c0 + 16 <= a.Cout) {
              // 16-byte reductions: 4 L2 operations per 16 channels instead of 16
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(__uint_as_float(v[j])),
                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float f = __uint_as_float(v[j]);
                if (c0 + j < a.Cout && f != 0.f) atomicAdd(dst + c0 + j, f);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();   // the issuer role ran on one lane: reconverge before the .sync.aligned instruction
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

struct WguLimits {
  int sms = 0, max_smem = 0;
};
const WguLimits& wgu_limits() {
  static WguLimits ls[kEsnMaxDevices];
  static std::once_flag once[kEsnMaxDevices];
  const int dev = esn_current_device();
  std::call_once(once[dev], [dev] {
    WguLimits& l = ls[dev];
    cudaDeviceGetAttribute(&l.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&l.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, wgrad_umma_kernel) == cudaSuccess) l.max_smem -= (int)fa.sharedSizeBytes;
    cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, l.max_smem);
  });
  return ls[dev];
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
inline uint32_t layout_of(uint32_t row_bytes) { return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u); }
inline CUtensorMapSwizzle swizzle_of(uint32_t row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace

// plans the launch; false when the shape is not taken by this kernel
static bool wgu_plan(const EsnConv* p, WguArgs& a, int& ngroups, bool with_maps) {
  static const bool off = getenv("ESN_WGRAD_NO_UMMA") != nullptr;
  if (off) return false;
  const EsnTensor& x = p->x;
  const EsnTensor& dy = p->y;
  if (x.layout != ESN_NHWC || x.dtype != ESN_BF16 || dy.dtype != ESN_BF16 || p->groups != 1 || p->stride != 1) return false;
  if (p->kh > 3 || p->kw > 3 || x.c < 8 || dy.c < 8 || dy.c > 256) return false;
  if (x.c_stride % 8 || dy.c_stride % 8 || ((uintptr_t)x.ptr & 15) || ((uintptr_t)dy.ptr & 15)) return false;
  if ((long long)dy.n * dy.h * dy.w < 1024) return false;      // tiny problems: launch-bound either way
  EncodeTiledFn encode = get_encode();
  const WguLimits& lim = wgu_limits();
  if (!encode || lim.sms <= 0) return false;

  memset(&a, 0, sizeof(a));
  a.dw = reinterpret_cast<float*>(const_cast<void*>(p->w));
  a.Cin = x.c; a.Cout = dy.c; a.kh = p->kh; a.kw = p->kw;
  a.pad_h = p->pad_h; a.pad_w = p->pad_w; a.dil_h = p->dil_h;
  const int cbx = x.c <= 16 ? 16 : (x.c <= 32 ? 32 : 64);
  const int cbg = dy.c <= 16 ? 16 : (dy.c <= 32 ? 32 : 64);
  a.nxbox = (x.c + 63) / 64;
  a.ngbox = (dy.c + 63) / 64;
  a.N = dy.c <= 64 ? cbg : (dy.c + 15) / 16 * 16;
  a.xrow_bytes = cbx * 2; a.grow_bytes = cbg * 2;
  a.small = a.nxbox == 1;
  a.blk = cbx;
  const int nblk = 128 / cbx;                                  // taps per MMA in the small case
  const int nmma_s = (p->kw + nblk - 1) / nblk;
  const int acc_per_row = a.small ? nmma_s : p->kw * ((a.nxbox + 1) / 2);
  a.BW = dy.w >= 128 ? 128 : (dy.w + 15) / 16 * 16;
  a.XW = a.BW + (p->kw - 1) * p->dil_w;
  if (a.XW > 256) return false;
  const int junk_px = a.small ? (nmma_s * nblk - p->kw) * p->dil_w : 0;
  bool planned = false;
  for (int attempt = 0; attempt < 2 && !planned; ++attempt) {
    a.nrt = attempt == 0 ? p->kh : 1;               // all filter rows in one CTA, else one CTA group per filter row
    if (attempt == 1 && p->kh == 1) break;
    if (acc_per_row * a.nrt * a.N > 512 || acc_per_row * a.nrt > kMaxAcc) continue;
    a.BH = 256 / a.BW > dy.h ? dy.h : 256 / a.BW;
    if (a.BH < 1) a.BH = 1;
    for (;; --a.BH) {
      a.xbox_alloc = round_up((uint32_t)(a.BH * a.XW + junk_px) * a.xrow_bytes, 1024);
      a.xtile_alloc = a.xbox_alloc * (a.small ? 1 : (a.nxbox + 1) / 2 * 2);
      a.gbox_bytes = round_up((uint32_t)(a.BH * a.BW) * a.grow_bytes, 1024);
      a.gtile_bytes = a.gbox_bytes * a.ngbox;
      a.stage_bytes = a.gtile_bytes + a.nrt * a.xtile_alloc;
      a.stages = (int)((lim.max_smem - 1024 - 4096 - 256) / (long long)a.stage_bytes);
      if (a.stages >= 3 || a.BH == 1) break;
    }
    planned = a.stages >= 2;
  }
  if (!planned) return false;
  if (a.stages > 8) a.stages = 8;
  ngroups = p->kh / a.nrt;
  a.nacc = acc_per_row * a.nrt;
  a.tmem_cols = 32;
  while (a.tmem_cols < a.nacc * a.N) a.tmem_cols *= 2;
  a.load_bytes = (uint32_t)(a.ngbox * a.BH * a.BW) * a.grow_bytes + (uint32_t)(a.nrt * a.nxbox * a.BH * a.XW) * a.xrow_bytes;
  a.tiles_w = (dy.w + a.BW - 1) / a.BW;
  a.tiles_h = (dy.h + a.BH - 1) / a.BH;
  a.nunits = dy.n * a.tiles_h * a.tiles_w;

  a.adesc_hi = ((8u * a.xrow_bytes) >> 4) | (1u << 14) | (layout_of(a.xrow_bytes) << 29);
  a.bdesc_hi = ((8u * a.grow_bytes) >> 4) | (1u << 14) | (layout_of(a.grow_bytes) << 29);
  a.a_lbo = a.small ? (uint32_t)p->dil_w * a.xrow_bytes : a.xbox_alloc;
  a.b_lbo = a.gbox_bytes;
  if ((a.a_lbo >> 4) > 0x3FFFu) return false;
  a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.N >> 3) << 17) | ((128u >> 4) << 24);
  a.loop_s = a.small ? nmma_s : p->kw;
  a.loop_jb = a.small ? 1 : (a.nxbox + 1) / 2;
  a.step_s = (uint32_t)((a.small ? nblk : 1) * p->dil_w) * a.xrow_bytes;
  a.step_jb = 2 * a.xbox_alloc;
  int q = 0;
  for (int rt = 0; rt < a.nrt; ++rt) {
    if (a.small) {
      for (int j = 0; j < nmma_s; ++j, ++q) {
        a.acc_rt[q] = rt; a.acc_off[q] = (uint32_t)(j * nblk * p->dil_w) * a.xrow_bytes; a.acc_s0[q] = j * nblk; a.acc_ci0[q] = 0;
      }
    } else {
      for (int s = 0; s < p->kw; ++s)
        for (int jb = 0; jb < (a.nxbox + 1) / 2; ++jb, ++q) {
          a.acc_rt[q] = rt; a.acc_off[q] = (uint32_t)(2 * jb) * a.xbox_alloc + (uint32_t)(s * p->dil_w) * a.xrow_bytes;
          a.acc_s0[q] = s; a.acc_ci0[q] = jb * 128;
        }
    }
  }

  if (!with_maps) return true;
  {
    const cuuint64_t cs = (cuuint64_t)x.c_stride;
    const cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
    const cuuint64_t strides[3] = {cs * 2, (cuuint64_t)x.w * cs * 2, (cuuint64_t)x.h * x.w * cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)cbx, (cuuint32_t)a.XW, (cuuint32_t)a.BH, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(&a.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x.ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(a.xrow_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  {
    const cuuint64_t cs = (cuuint64_t)dy.c_stride;
    const cuuint64_t dims[4] = {(cuuint64_t)dy.c, (cuuint64_t)dy.w, (cuuint64_t)dy.h, (cuuint64_t)dy.n};
    const cuuint64_t strides[3] = {cs * 2, (cuuint64_t)dy.w * cs * 2, (cuuint64_t)dy.h * dy.w * cs * 2};
    const cuuint32_t box[4] = {(cuuint32_t)cbg, (cuuint32_t)a.BW, (cuuint32_t)a.BH, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (encode(&a.tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy.ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(a.grow_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

extern "C" int esn_wgrad_umma_supported(const EsnConv* p) {
  if (!p || !p->w || !p->x.ptr || !p->y.ptr) return 0;
  WguArgs a;
  int ngroups = 1;
  return wgu_plan(p, a, ngroups, false) ? 1 : 0;
}

// called by esn_conv2d_wgrad for bf16 x / bf16 dy dense stride-1 convs; returns false when the shape is not taken
bool esn_wgrad_umma_try(const EsnConv* p, void* stream, int* rc) {
  WguArgs a;
  int ngroups = 1;
  if (!wgu_plan(p, a, ngroups, true)) return false;
  const WguLimits& lim = wgu_limits();
  int gx = lim.sms / ngroups;
  if (gx < 1) gx = 1;
  if (gx > a.nunits) gx = a.nunits;
  const size_t smem = (size_t)a.stages * a.stage_bytes + 1024 + 4096 + 256;
  wgrad_umma_kernel<<<dim3(gx, ngroups), kWguThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  g_esn_launches.fetch_add(1, std::memory_order_relaxed);
  *rc = (cudaPeekAtLastError() == cudaSuccess) ? ESN_OK : ESN_ERR_CUDA;
  if (*rc != ESN_OK) cudaGetLastError();
  return true;
}
