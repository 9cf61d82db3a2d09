// Training-path kernels (round 1: correct first, CUDA cores; the tensor-core wgrad is next):
//   * per-channel batch statistics (sum, sum of squares) for train-mode BatchNorm
//   * BatchNorm finalize (scale/shift for the apply pass, running-stat update as torch does it)
//   * BN + activation backward: reduction pass (sum dz, sum dz*xhat, sum for d alpha) and apply pass
//   * weight gradient of dense / depthwise convolutions
//   * max-pool 2x2 backward, bilinear (align_corners=False) backward
// All reductions: fp32 inside a CTA (warp shuffle -> smem), fp64 atomics across CTAs.
#include "esn_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ void ldv4(const T* p, bool vec, int c, int C, float* v) {
  if (vec) {
    const float4 t = ld4<T>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (c + j < C) ? ld1<T>(p + j) : 0.f;
  }
}
template <typename T>
__device__ __forceinline__ void stv4(T* p, bool vec, int c, int C, const float* v) {
  if (vec) {
    st4<T>(p, make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < C) st1<T>(p + j, v[j]);
  }
}

// ---------------------------------------------------------------- channel statistics
// block = 256 threads = (256/CG) pixel lanes x CG channel groups of 4; grid.x pixel chunks, grid.y channel blocks
constexpr int kStatThreads = 256;

template <typename T, int NQ>
__global__ void __launch_bounds__(kStatThreads) channel_stats_kernel(const T* __restrict__ x, long long M, int C, int cs,
                                                                     double* __restrict__ sums, long long px_per_cta) {
  // NQ quantities per channel: 1 -> sum; 2 -> sum, sum of squares
  __shared__ float red[NQ][kStatThreads][4];
  const int CG = min((C + 3) / 4, 64);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * 64 + cg) * 4;
  const bool vec = (cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(x) % (4 * sizeof(T))) == 0);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  const long long p0 = blockIdx.x * px_per_cta, p1 = min(M, p0 + px_per_cta);
  if (pl < lanes && c < C) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      float v[4];
      ldv4<T>(x + p * cs + c, vec, c, C, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] += v[j];
        if (NQ > 1) q[j] += v[j] * v[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][threadIdx.x][j] = s[j];
    if (NQ > 1) red[NQ - 1][threadIdx.x][j] = q[j];
  }
  __syncthreads();
  if (pl == 0 && c < C) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] += red[0][l * CG + cg][j];
        if (NQ > 1) q[j] += red[NQ - 1][l * CG + cg][j];
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < C) {
        atomicAdd(sums + c + j, (double)s[j]);
        if (NQ > 1) atomicAdd(sums + C + c + j, (double)q[j]);
      }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, float* scale, float* shift, float* mean, float* invstd, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[c] / count;
  double var = sums[C + c] / count - m * m;   // biased variance normalises (torch semantics)
  if (var < 0) var = 0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * is;
  scale[c] = sc;
  shift[c] = b - (float)m * sc;
  mean[c] = (float)m;
  invstd[c] = is;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---------------------------------------------------------------- BN + activation backward
struct BnBwdArgs {
  const void* x;
  const void* dy;
  void* dx;
  const void* extra;
  long long M;
  int C, x_cs, dy_cs, dx_cs, extra_cs, act, train_stats;
  const float *scale, *shift, *alpha, *mean, *invstd;
  double* sums;  // [3][C]: sum dz, sum dz*xhat, sum dy*z*[z<0]
  float *dgamma, *dbeta, *dalpha;
  long long px_per_cta;
};

__device__ __forceinline__ float act_grad(float z, float dy, int act, float alpha) {
  if (act == ESN_ACT_RELU) return z > 0.f ? dy : 0.f;
  if (act == ESN_ACT_PRELU) return z >= 0.f ? dy : dy * alpha;
  return dy;
}

template <typename TX, typename TG>
__global__ void __launch_bounds__(kStatThreads) bn_act_bwd_reduce_kernel(const BnBwdArgs a) {
  __shared__ float red[3][kStatThreads][4];
  const int C = a.C;
  const int CG = min((C + 3) / 4, 64);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * 64 + cg) * 4;
  const TX* x = reinterpret_cast<const TX*>(a.x);
  const TG* dy = reinterpret_cast<const TG*>(a.dy);
  const bool vx = (a.x_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(x) % (4 * sizeof(TX))) == 0);
  const bool vg = (a.dy_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(dy) % (4 * sizeof(TG))) == 0);
  float sc[4], sh[4], al[4], mu[4], is[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = min(c + j, C - 1);
    sc[j] = a.scale ? a.scale[cc] : 1.f;
    sh[j] = a.shift ? a.shift[cc] : 0.f;
    al[j] = (a.act == ESN_ACT_PRELU) ? a.alpha[cc] : 0.f;
    mu[j] = a.mean ? a.mean[cc] : 0.f;
    is[j] = a.invstd ? a.invstd[cc] : 1.f;
  }
  float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(a.M, p0 + a.px_per_cta);
  if (pl < lanes && c < C) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      float xv[4], gv[4];
      ldv4<TX>(x + p * a.x_cs + c, vx, c, C, xv);
      ldv4<TG>(dy + p * a.dy_cs + c, vg, c, C, gv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z = fmaf(xv[j], sc[j], sh[j]);
        const float dz = act_grad(z, gv[j], a.act, al[j]);
        s0[j] += dz;
        s1[j] += dz * (xv[j] - mu[j]) * is[j];
        s2[j] += (z < 0.f) ? gv[j] * z : 0.f;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][threadIdx.x][j] = s0[j];
    red[1][threadIdx.x][j] = s1[j];
    red[2][threadIdx.x][j] = s2[j];
  }
  __syncthreads();
  if (pl == 0 && c < C) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s0[j] += red[0][l * CG + cg][j];
        s1[j] += red[1][l * CG + cg][j];
        s2[j] += red[2][l * CG + cg][j];
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < C) {
        atomicAdd(a.sums + c + j, (double)s0[j]);
        atomicAdd(a.sums + C + c + j, (double)s1[j]);
        atomicAdd(a.sums + 2 * C + c + j, (double)s2[j]);
      }
  }
}

template <typename TX, typename TG, typename TD>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const BnBwdArgs a) {
  const int C = a.C;
  const int ncg = (C + 3) / 4;
  const long long total = a.M * ncg;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // parameter gradients: written once by the first threads of the grid
  if (idx < C) {
    const int c = (int)idx;
    if (a.dbeta) a.dbeta[c] = (float)a.sums[c];
    if (a.dgamma) a.dgamma[c] = (float)a.sums[C + c];
    if (a.dalpha && a.act == ESN_ACT_PRELU) a.dalpha[c] = (float)a.sums[2 * C + c];
  }
  if (idx >= total) return;
  const int c = (int)(idx % ncg) * 4;
  const long long p = idx / ncg;
  const TX* x = reinterpret_cast<const TX*>(a.x) + p * a.x_cs + c;
  const TG* dy = reinterpret_cast<const TG*>(a.dy) + p * a.dy_cs + c;
  TD* dx = reinterpret_cast<TD*>(a.dx) + p * a.dx_cs + c;
  const bool vx = (a.x_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(a.x) % (4 * sizeof(TX))) == 0);
  const bool vg = (a.dy_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(a.dy) % (4 * sizeof(TG))) == 0);
  const bool vd = (a.dx_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(a.dx) % (4 * sizeof(TD))) == 0);
  float xv[4], gv[4], ev[4] = {0, 0, 0, 0}, out[4];
  ldv4<TX>(x, vx, c, C, xv);
  ldv4<TG>(dy, vg, c, C, gv);
  if (a.extra) {
    const TD* e = reinterpret_cast<const TD*>(a.extra) + p * a.extra_cs + c;
    const bool ve = (a.extra_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(a.extra) % (4 * sizeof(TD))) == 0);
    ldv4<TD>(e, ve, c, C, ev);
  }
  const float invM = (float)(1.0 / (double)a.M);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = min(c + j, C - 1);
    const float sc = a.scale ? a.scale[cc] : 1.f, sh = a.shift ? a.shift[cc] : 0.f;
    const float al = (a.act == ESN_ACT_PRELU) ? a.alpha[cc] : 0.f;
    const float z = fmaf(xv[j], sc, sh);
    const float dz = act_grad(z, gv[j], a.act, al);
    float g;
    if (a.train_stats) {
      const float xhat = (xv[j] - a.mean[cc]) * a.invstd[cc];
      g = sc * (dz - (float)a.sums[cc] * invM - xhat * (float)a.sums[C + cc] * invM);
    } else {
      g = sc * dz;
    }
    out[j] = g + ev[j];
  }
  stv4<TD>(dx, vd, c, C, out);
}

// ---------------------------------------------------------------- 16-byte (8 x bf16) variants
// Same decomposition -- a thread owns one 8-channel group and walks a pixel chunk, channel groups fastest across
// the threads so a warp reads whole pixels -- with 128-bit accesses and the per-channel constants in registers.
// Channel counts that are not a multiple of 8 (35, 131, 259 ...) read the full vector (it stays inside the pixel
// stride) and simply never publish / store the lanes past C.
constexpr int kV8Groups = 32;     // channel groups per block (256 channels per blockIdx.y)

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* f) { bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p)), f); }

template <int NQ>
__global__ void __launch_bounds__(kStatThreads) channel_stats_v8_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, int cs,
                                                                        double* __restrict__ sums, long long px_per_cta) {
  __shared__ float red[NQ][kStatThreads][9];
  const int ng = (C + 7) / 8;
  const int CG = min(ng - (int)blockIdx.y * kV8Groups, kV8Groups);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * kV8Groups + cg) * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const long long p0 = blockIdx.x * px_per_cta, p1 = min(M, p0 + px_per_cta);
  if (pl < lanes) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      float v[8];
      ld8(x + p * cs + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        if (NQ > 1) q[j] = fmaf(v[j], v[j], q[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = s[j];
    if (NQ > 1) red[NQ - 1][threadIdx.x][j] = q[j];
  }
  __syncthreads();
  // thread t < CG*8 publishes channel (t / 8 group, t % 8 lane): sum over the pixel lanes
  if ((int)threadIdx.x < CG * 8) {
    const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
    const int cc = (blockIdx.y * kV8Groups + g) * 8 + j;
    if (cc < C) {
      float a0 = 0.f, a1 = 0.f;
      for (int l = 0; l < lanes; ++l) {
        a0 += red[0][l * CG + g][j];
        if (NQ > 1) a1 += red[NQ - 1][l * CG + g][j];
      }
      atomicAdd(sums + cc, (double)a0);
      if (NQ > 1) atomicAdd(sums + C + cc, (double)a1);
    }
  }
}

__global__ void __launch_bounds__(kStatThreads) bn_act_bwd_reduce_v8_kernel(const BnBwdArgs a) {
  __shared__ float red[3][kStatThreads][9];
  const int C = a.C;
  const int ng = (C + 7) / 8;
  const int CG = min(ng - (int)blockIdx.y * kV8Groups, kV8Groups);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * kV8Groups + cg) * 8;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a.x);
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(a.dy);
  float sc[8], sh[8], al[8], mu[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cc = min(c + j, C - 1);
    sc[j] = a.scale ? a.scale[cc] : 1.f;
    sh[j] = a.shift ? a.shift[cc] : 0.f;
    al[j] = (a.act == ESN_ACT_PRELU) ? a.alpha[cc] : 0.f;
    mu[j] = a.mean ? a.mean[cc] : 0.f;
    is[j] = a.invstd ? a.invstd[cc] : 1.f;
  }
  float s0[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = s2[j] = 0.f;
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(a.M, p0 + a.px_per_cta);
  if (pl < lanes) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      float xv[8], gv[8];
      ld8(x + p * a.x_cs + c, xv);
      ld8(dy + p * a.dy_cs + c, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(xv[j], sc[j], sh[j]);
        const float dz = act_grad(z, gv[j], a.act, al[j]);
        s0[j] += dz;
        s1[j] = fmaf(dz, (xv[j] - mu[j]) * is[j], s1[j]);
        s2[j] += (z < 0.f) ? gv[j] * z : 0.f;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = s0[j];
    red[1][threadIdx.x][j] = s1[j];
    red[2][threadIdx.x][j] = s2[j];
  }
  __syncthreads();
  if ((int)threadIdx.x < CG * 8) {
    const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
    const int cc = (blockIdx.y * kV8Groups + g) * 8 + j;
    if (cc < C) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      for (int l = 0; l < lanes; ++l) {
        a0 += red[0][l * CG + g][j];
        a1 += red[1][l * CG + g][j];
        a2 += red[2][l * CG + g][j];
      }
      atomicAdd(a.sums + cc, (double)a0);
      atomicAdd(a.sums + C + cc, (double)a1);
      atomicAdd(a.sums + 2 * C + cc, (double)a2);
    }
  }
}

__global__ void __launch_bounds__(kStatThreads) bn_act_bwd_apply_v8_kernel(const BnBwdArgs a) {
  const int C = a.C;
  const int ng = (C + 7) / 8;
  const int CG = min(ng - (int)blockIdx.y * kV8Groups, kV8Groups);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * kV8Groups + cg) * 8;
  // parameter gradients: written once by the first pixel block
  if (blockIdx.x == 0 && (int)threadIdx.x < CG * 8) {
    const int cc = (blockIdx.y * kV8Groups) * 8 + threadIdx.x;
    if (cc < C) {
      if (a.dbeta) a.dbeta[cc] = (float)a.sums[cc];
      if (a.dgamma) a.dgamma[cc] = (float)a.sums[C + cc];
      if (a.dalpha && a.act == ESN_ACT_PRELU) a.dalpha[cc] = (float)a.sums[2 * C + cc];
    }
  }
  if (pl >= lanes) return;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a.x);
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(a.dy);
  const __nv_bfloat16* ex = reinterpret_cast<const __nv_bfloat16*>(a.extra);
  __nv_bfloat16* dx = reinterpret_cast<__nv_bfloat16*>(a.dx);
  const float invM = (float)(1.0 / (double)a.M);
  float sc[8], sh[8], al[8], k0[8], k1[8], mu[8];   // g = sc*dz - k0 - (x - mu)*k1
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cc = min(c + j, C - 1);
    sc[j] = a.scale ? a.scale[cc] : 1.f;
    sh[j] = a.shift ? a.shift[cc] : 0.f;
    al[j] = (a.act == ESN_ACT_PRELU) ? a.alpha[cc] : 0.f;
    if (a.train_stats) {
      mu[j] = a.mean[cc];
      k0[j] = sc[j] * (float)a.sums[cc] * invM;
      k1[j] = sc[j] * a.invstd[cc] * (float)a.sums[C + cc] * invM;
    } else {
      mu[j] = k0[j] = k1[j] = 0.f;
    }
  }
  const bool full = c + 8 <= C;
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(a.M, p0 + a.px_per_cta);
  for (long long p = p0 + pl; p < p1; p += lanes) {
    float xv[8], gv[8], out[8];
    ld8(x + p * a.x_cs + c, xv);
    ld8(dy + p * a.dy_cs + c, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(xv[j], sc[j], sh[j]);
      const float dz = act_grad(z, gv[j], a.act, al[j]);
      out[j] = sc[j] * dz - k0[j] - (xv[j] - mu[j]) * k1[j];
    }
    if (ex) {
      float ev[8];
      ld8(ex + p * a.extra_cs + c, ev);
#pragma unroll
      for (int j = 0; j < 8; ++j) out[j] += ev[j];
    }
    __nv_bfloat16* o = dx + p * a.dx_cs + c;
    if (full) {
      *reinterpret_cast<uint4*>(o) = float_to_bf16x8(out);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < C) o[j] = __float2bfloat16_rn(out[j]);
    }
  }
}

static inline bool v8_ok(const void* p, int cs) { return p && cs % 8 == 0 && (reinterpret_cast<uintptr_t>(p) % 16) == 0; }

// ---------------------------------------------------------------- dense conv weight gradient
// dW[tap][ci][co] += sum_p X[p + delta_tap, ci] * dY[p, co]; block computes a 64(ci) x 64(co) tile for one
// tap over a chunk of output pixels; thread = 4x4 register tile; smem stages of 16 pixels.
struct WgradArgs {
  const void* x;
  const void* dy;
  float* dw;
  int N, Hi, Wi, Cin, x_cs, x_nchw;
  int Ho, Wo, Cout, dy_cs;
  int kh, kw, stride, pad_h, pad_w, dil_h, dil_w;
  long long px_per_cta;
};

constexpr int kWgPix = 16;

// Network stem (NCHW fp32 image, Cin = 3, 3x3, Cout <= 32): lane = output channel, 27 accumulators per lane.  A warp
// walks output pixels: the 27 image values of a pixel are warp-uniform (broadcast) loads, dY is one coalesced load.
template <typename TG>
__global__ void __launch_bounds__(256) wgrad_stem_kernel(const WgradArgs a) {
  __shared__ float red[8][27][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(M, p0 + a.px_per_cta);
  const float* x = reinterpret_cast<const float*>(a.x);
  const TG* dy = reinterpret_cast<const TG*>(a.dy);
  const size_t plane = (size_t)a.Hi * a.Wi;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  const bool act = lane < a.Cout;
  for (long long p = p0 + warp; p < p1; p += 8) {
    const int wo = (int)(p % a.Wo);
    const int ho = (int)((p / a.Wo) % a.Ho);
    const int n = (int)(p / ((long long)a.Wo * a.Ho));
    const float g = act ? ld1<TG>(dy + (size_t)p * a.dy_cs + lane) : 0.f;
    const float* xn = x + (size_t)n * 3 * plane;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hi = ho * a.stride - a.pad_h + r * a.dil_h;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int wi = wo * a.stride - a.pad_w + q * a.dil_w;
        if (hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) {
          const size_t o = (size_t)hi * a.Wi + wi;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) acc[(r * 3 + q) * 3 + ci] = fmaf(__ldg(xn + ci * plane + o), g, acc[(r * 3 + q) * 3 + ci]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t) red[warp][t][lane] = acc[t];
  __syncthreads();
  for (int e = threadIdx.x; e < 27 * 32; e += 256) {
    const int t = e >> 5, co = e & 31;
    if (co < a.Cout) {
      float v = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) v += red[wv][t][co];
      atomicAdd(a.dw + (size_t)t * a.Cout + co, v);     // t = (tap*3 + ci): the [tap][Cin][Cout] layout
    }
  }
}

// Tiny dense convs (taps*Cin*Cout <= 128: ESPNetv2's 8->6 grouped slices and 3->3 input-reinforcement conv):
// one thread per weight element, walking a pixel strip; the few channels of a pixel are L1 broadcasts.
template <typename TX, typename TG>
__global__ void __launch_bounds__(128) wgrad_small_kernel(const WgradArgs a) {
  const int P = a.kh * a.kw * a.Cin * a.Cout;
  const int t = threadIdx.x;
  if (t >= P) return;
  const int co = t % a.Cout;
  const int ci = (t / a.Cout) % a.Cin;
  const int tap = t / (a.Cout * a.Cin);
  const int r = tap / a.kw, q = tap - r * a.kw;
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(M, p0 + a.px_per_cta);
  if (p0 >= p1) return;
  int wo = (int)(p0 % a.Wo);
  int ho = (int)((p0 / a.Wo) % a.Ho);
  int n = (int)(p0 / ((long long)a.Wo * a.Ho));
  const TX* x = reinterpret_cast<const TX*>(a.x);
  const TG* dy = reinterpret_cast<const TG*>(a.dy) + co;
  float acc = 0.f;
  for (long long p = p0; p < p1; ++p) {
    const int hi = ho * a.stride - a.pad_h + r * a.dil_h, wi = wo * a.stride - a.pad_w + q * a.dil_w;
    if (hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) {
      const float xv = a.x_nchw ? ld1<TX>(x + (((size_t)n * a.Cin + ci) * a.Hi + hi) * a.Wi + wi)
                                : ld1<TX>(x + (((size_t)n * a.Hi + hi) * a.Wi + wi) * a.x_cs + ci);
      acc = fmaf(xv, ld1<TG>(dy + (size_t)p * a.dy_cs), acc);
    }
    if (++wo == a.Wo) { wo = 0; if (++ho == a.Ho) { ho = 0; ++n; } }
  }
  atomicAdd(a.dw + t, acc);      // t == (tap*Cin + ci)*Cout + co: the [tap][Cin][Cout] layout
}

template <typename TX, typename TG>
__global__ void __launch_bounds__(256) wgrad_dense_kernel(const WgradArgs a) {
  __shared__ float xs[kWgPix][64 + 1];
  __shared__ float gs[kWgPix][64 + 1];
  const int tap = blockIdx.y;
  const int r = tap / a.kw, s = tap % a.kw;
  const int nci = (a.Cin + 63) / 64, nco = (a.Cout + 63) / 64;
  const int ci0 = (blockIdx.z % nci) * 64, co0 = (blockIdx.z / nci) * 64;
  (void)nco;
  const int ti = threadIdx.x / 16, tj = threadIdx.x % 16;   // thread tile: ci = ti*4.., co = tj*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(M, p0 + a.px_per_cta);
  const TX* x = reinterpret_cast<const TX*>(a.x);
  const TG* dy = reinterpret_cast<const TG*>(a.dy);
  for (long long pb = p0; pb < p1; pb += kWgPix) {
    // stage kWgPix pixels x 64 channels of X (shifted by the tap) and of dY
    for (int e = threadIdx.x; e < kWgPix * 64; e += 256) {
      const int pp = e / 64, ch = e % 64;
      const long long p = pb + pp;
      float xv = 0.f, gv = 0.f;
      if (p < p1) {
        const int wo = (int)(p % a.Wo);
        const int ho = (int)((p / a.Wo) % a.Ho);
        const int n = (int)(p / ((long long)a.Wo * a.Ho));
        const int hi = ho * a.stride - a.pad_h + r * a.dil_h, wi = wo * a.stride - a.pad_w + s * a.dil_w;
        if (co0 + ch < a.Cout) gv = ld1<TG>(dy + p * a.dy_cs + co0 + ch);
        if (ci0 + ch < a.Cin && hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) {
          if (a.x_nchw)
            xv = ld1<TX>(x + ((size_t)((size_t)n * a.Cin + ci0 + ch) * a.Hi + hi) * a.Wi + wi);
          else
            xv = ld1<TX>(x + ((size_t)((size_t)n * a.Hi + hi) * a.Wi + wi) * a.x_cs + ci0 + ch);
        }
      }
      xs[pp][ch] = xv;
      gs[pp][ch] = gv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < kWgPix; ++pp) {
      float xv[4], gv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xv[i] = xs[pp][ti * 4 + i];
        gv[i] = gs[pp][tj * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + ti * 4 + i, co = co0 + tj * 4 + j;
      if (ci < a.Cin && co < a.Cout && acc[i][j] != 0.f)
        atomicAdd(a.dw + ((size_t)tap * a.Cin + ci) * a.Cout + co, acc[i][j]);
    }
}

// depthwise: dW[tap][c] += sum_p X[p+delta, c] * dY[p, c]
template <typename TX, typename TG>
__global__ void __launch_bounds__(kStatThreads) wgrad_dw_kernel(const WgradArgs a) {
  __shared__ float red[kStatThreads][4];
  const int tap = blockIdx.y;
  const int r = tap / a.kw, s = tap % a.kw;
  const int C = a.Cout;
  const int CG = min((C + 3) / 4, 64);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.z * 64 + cg) * 4;
  const TX* x = reinterpret_cast<const TX*>(a.x);
  const TG* dy = reinterpret_cast<const TG*>(a.dy);
  const bool vx = (a.x_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(x) % (4 * sizeof(TX))) == 0);
  const bool vg = (a.dy_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(dy) % (4 * sizeof(TG))) == 0);
  float acc[4] = {0, 0, 0, 0};
  const long long M = (long long)a.N * a.Ho * a.Wo;
  const long long p0 = blockIdx.x * a.px_per_cta, p1 = min(M, p0 + a.px_per_cta);
  if (pl < lanes && c < C) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      const int wo = (int)(p % a.Wo);
      const int ho = (int)((p / a.Wo) % a.Ho);
      const int n = (int)(p / ((long long)a.Wo * a.Ho));
      const int hi = ho * a.stride - a.pad_h + r * a.dil_h, wi = wo * a.stride - a.pad_w + s * a.dil_w;
      if (hi < 0 || hi >= a.Hi || wi < 0 || wi >= a.Wi) continue;
      float xv[4], gv[4];
      ldv4<TX>(x + ((size_t)((size_t)n * a.Hi + hi) * a.Wi + wi) * a.x_cs + c, vx, c, C, xv);
      ldv4<TG>(dy + p * a.dy_cs + c, vg, c, C, gv);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv[j], gv[j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (pl == 0 && c < C) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += red[l * CG + cg][j];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < C) atomicAdd(a.dw + (size_t)tap * C + c + j, acc[j]);
  }
}

// depthwise, ALL taps in one pass: a thread owns 4 channels and a lane of pixels, loads dY once per pixel and the
// KH*KW (L1-resident) neighbours of x, keeps KH*KW*4 accumulators.  Rows are walked without per-pixel div/mod.
// One read of |x| + |dy| from DRAM instead of KH*KW reads (Fast-SCNN / ESPNetv2 / CGNet 3x3, DABNet / EESP 3x1, 1x3).
template <typename TX, typename TG, int KH, int KW>
__global__ void __launch_bounds__(kStatThreads) wgrad_dw_all_kernel(const WgradArgs a, const int rows_per_cta) {
  constexpr int T = KH * KW;
  __shared__ float red[kStatThreads][4];
  const int C = a.Cout;
  const int CG = min((C + 3) / 4, 64);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.z * 64 + cg) * 4;
  const TX* x = reinterpret_cast<const TX*>(a.x);
  const TG* dy = reinterpret_cast<const TG*>(a.dy);
  const bool vx = (a.x_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(x) % (4 * sizeof(TX))) == 0);
  const bool vg = (a.dy_cs % 4 == 0) && (c + 4 <= C) && ((reinterpret_cast<uintptr_t>(dy) % (4 * sizeof(TG))) == 0);
  float acc[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  const int R = a.N * a.Ho;
  const int row0 = blockIdx.x * rows_per_cta, row1 = min(R, row0 + rows_per_cta);
  if (KW == 3 && a.stride == 1 && a.dil_w == 1) {
    // sliding window: a thread walks a contiguous segment of the row and keeps the KH x 3 window of x in registers,
    // so a step costs KH + 1 loads (the new right-hand column and dY) instead of 3 KH + 1
    const int L = (a.Wo + lanes - 1) / lanes;
    const int wb = pl * L, we = min(a.Wo, wb + L);
    if (pl < lanes && c < C && wb < we) {
      for (int row = row0; row < row1; ++row) {
        const int n = row / a.Ho, ho = row - n * a.Ho;
        const TG* grow = dy + (size_t)row * a.Wo * a.dy_cs + c;
        const TX* xn = x + (size_t)n * a.Hi * a.Wi * a.x_cs + c;
        const TX* xr[KH];
        bool rv[KH];
#pragma unroll
        for (int r = 0; r < KH; ++r) {
          const int hi = ho - a.pad_h + r * a.dil_h;
          rv[r] = hi >= 0 && hi < a.Hi;
          xr[r] = xn + (size_t)(rv[r] ? hi : 0) * a.Wi * a.x_cs;
        }
        float win[KH][3][4];
#pragma unroll
        for (int r = 0; r < KH; ++r)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int wi = wb - a.pad_w + q;
#pragma unroll
            for (int j = 0; j < 4; ++j) win[r][q][j] = 0.f;
            if (rv[r] && wi >= 0 && wi < a.Wi) ldv4<TX>(xr[r] + (size_t)wi * a.x_cs, vx, c, C, win[r][q]);
          }
        for (int wo = wb; wo < we; ++wo) {
          float gv[4];
          ldv4<TG>(grow + (size_t)wo * a.dy_cs, vg, c, C, gv);
          const int wi = wo - a.pad_w + 2;
          const bool cv = wi >= 0 && wi < a.Wi;
#pragma unroll
          for (int r = 0; r < KH; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) win[r][2][j] = 0.f;
            if (rv[r] && cv) ldv4<TX>(xr[r] + (size_t)wi * a.x_cs, vx, c, C, win[r][2]);
          }
#pragma unroll
          for (int r = 0; r < KH; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[r * 3 + q][j] = fmaf(win[r][q][j], gv[j], acc[r * 3 + q][j]);
#pragma unroll
          for (int r = 0; r < KH; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              win[r][0][j] = win[r][1][j];
              win[r][1][j] = win[r][2][j];
            }
        }
      }
    }
  } else if (pl < lanes && c < C) {
    for (int row = row0; row < row1; ++row) {
      const int n = row / a.Ho, ho = row - n * a.Ho;
      const TG* grow = dy + (size_t)row * a.Wo * a.dy_cs + c;
      const TX* xn = x + (size_t)n * a.Hi * a.Wi * a.x_cs + c;
      const int hi0 = ho * a.stride - a.pad_h;
      for (int wo = pl; wo < a.Wo; wo += lanes) {
        float gv[4];
        ldv4<TG>(grow + (size_t)wo * a.dy_cs, vg, c, C, gv);
        const int wi0 = wo * a.stride - a.pad_w;
#pragma unroll
        for (int r = 0; r < KH; ++r) {
          const int hi = hi0 + r * a.dil_h;
          if (hi < 0 || hi >= a.Hi) continue;
#pragma unroll
          for (int q = 0; q < KW; ++q) {
            const int wi = wi0 + q * a.dil_w;
            if (wi < 0 || wi >= a.Wi) continue;
            float xv[4];
            ldv4<TX>(xn + ((size_t)hi * a.Wi + wi) * a.x_cs, vx, c, C, xv);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r * KW + q][j] = fmaf(xv[j], gv[j], acc[r * KW + q][j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = acc[t][j];
    __syncthreads();
    if (pl == 0 && c < C) {
      float v[4] = {acc[t][0], acc[t][1], acc[t][2], acc[t][3]};
      for (int l = 1; l < lanes; ++l)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += red[l * CG + cg][j];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < C && v[j] != 0.f) atomicAdd(a.dw + (size_t)t * C + c + j, v[j]);
    }
  }
}

// depthwise, stride 1, dilation 1 -- column-strip variant (default for these): a CTA owns a strip of the image
// (lanes x LSEG columns, all its channels) and walks DOWN the rows, so the KH input rows of an output row are the rows
// it loaded one and two steps ago: they stay in registers (each x element is loaded from L2/DRAM once per strip, ncu
// showed the row-major variant re-reading x KH times from DRAM and waiting on 4 dependent loads per pixel).  A thread
// = 4 channels x LSEG consecutive pixels: LSEG + 2 + LSEG independent 8/16-byte loads per row step.
template <typename T> struct Raw4;
template <> struct Raw4<float> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ float4 f4() const { return v; }
};
template <> struct Raw4<__nv_bfloat16> {
  uint2 v;
  __device__ __forceinline__ void zero() { v = make_uint2(0u, 0u); }
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ __forceinline__ float4 f4() const {
    return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                       __uint_as_float(v.y & 0xffff0000u));
  }
};

template <typename TX, typename TG, int KH, int KW, int LSEG, int S = 1>
__global__ void __launch_bounds__(kStatThreads) wgrad_dw_strip_kernel(const WgradArgs a, const int rows_per_chunk, const int nstrips,
                                                                      const int nchunks) {
  // S = conv stride (1, or 2 for the strided depthwise convs of Fast-SCNN / ESPNetv2): an output row then advances the input
  // window by S rows (KH - S rows stay in registers) and output column p reads input columns p*S .. p*S + KW - 1
  constexpr int T = KH * KW, XC = (LSEG - 1) * S + KW;
  __shared__ float red[kStatThreads][4];
  const int C = a.Cout;
  const int CG = min(C / 4, 64);
  const int lanes = kStatThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.z * 64 + cg) * 4;
  int b = blockIdx.x;
  const int chunk = b % nchunks; b /= nchunks;
  const int strip = b % nstrips;
  const int n = b / nstrips;
  const int wb = (strip * lanes + pl) * LSEG;                 // first output column of this thread
  const int h_begin = chunk * rows_per_chunk, h_end = min(a.Ho, h_begin + rows_per_chunk);
  float acc[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  if (pl < lanes && c < C && wb < a.Wo && h_begin < h_end) {
    const TX* xn = reinterpret_cast<const TX*>(a.x) + (size_t)n * a.Hi * a.Wi * a.x_cs + c;
    const TG* gn = reinterpret_cast<const TG*>(a.dy) + (size_t)n * a.Ho * a.Wo * a.dy_cs + c;
    Raw4<TX> R[KH][XC];
    auto load_row = [&](Raw4<TX>* dst, const int hi) {
      const bool rv = hi >= 0 && hi < a.Hi;
      const TX* xr = xn + (size_t)(rv ? hi : 0) * a.Wi * a.x_cs;
#pragma unroll
      for (int j = 0; j < XC; ++j) {
        const int wi = wb * S - a.pad_w + j;
        dst[j].zero();
        if (rv && wi >= 0 && wi < a.Wi) dst[j].load(xr + (size_t)wi * a.x_cs);
      }
    };
#pragma unroll
    for (int r = 0; r + S < KH; ++r) load_row(R[r], h_begin * S - a.pad_h + r);
    for (int ho = h_begin; ho < h_end; ++ho) {
#pragma unroll
      for (int r = (KH - S > 0 ? KH - S : 0); r < KH; ++r) load_row(R[r], ho * S - a.pad_h + r);
      Raw4<TG> G[LSEG];
      const TG* gr = gn + (size_t)ho * a.Wo * a.dy_cs;
#pragma unroll
      for (int p = 0; p < LSEG; ++p) {
        G[p].zero();
        if (wb + p < a.Wo) G[p].load(gr + (size_t)(wb + p) * a.dy_cs);
      }
#pragma unroll
      for (int p = 0; p < LSEG; ++p) {
        const float4 g = G[p].f4();
#pragma unroll
        for (int r = 0; r < KH; ++r)
#pragma unroll
          for (int q = 0; q < KW; ++q) {
            const float4 xv = R[r][p * S + q].f4();
            acc[r * KW + q][0] = fmaf(xv.x, g.x, acc[r * KW + q][0]);
            acc[r * KW + q][1] = fmaf(xv.y, g.y, acc[r * KW + q][1]);
            acc[r * KW + q][2] = fmaf(xv.z, g.z, acc[r * KW + q][2]);
            acc[r * KW + q][3] = fmaf(xv.w, g.w, acc[r * KW + q][3]);
          }
      }
#pragma unroll
      for (int r = 0; r + S < KH; ++r)
#pragma unroll
        for (int j = 0; j < XC; ++j) R[r][j] = R[r + S][j];
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = acc[t][j];
    __syncthreads();
    if (pl == 0 && c < C) {
      float v[4] = {acc[t][0], acc[t][1], acc[t][2], acc[t][3]};
      for (int l = 1; l < lanes; ++l)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += red[l * CG + cg][j];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (v[j] != 0.f) atomicAdd(a.dw + (size_t)t * C + c + j, v[j]);
    }
  }
}

// Network stem, second version (3x3, stride 2, Cin = 3 NCHW fp32 image, Cout <= 32): a CTA stages the 3 x 3 input row
// segments of a strip of 128 output pixels in shared memory with coalesced loads; lane = output channel; a warp takes
// two neighbouring output pixels at a time, whose 5 input columns per (ci, row) are ONE 16-byte + one 4-byte broadcast
// shared-memory read (the first version issued 27 dependent global broadcast loads per pixel).
constexpr int kStemTW = 128;
template <typename TG>
__global__ void __launch_bounds__(256) wgrad_stem2_kernel(const WgradArgs a, const int units_per_cta) {
  __shared__ __align__(16) float xs[9][2 * kStemTW + 8];   // [ci*3 + r][input column - wi0]
  __shared__ float red[8][27][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* x = reinterpret_cast<const float*>(a.x);
  const TG* dy = reinterpret_cast<const TG*>(a.dy);
  const size_t plane = (size_t)a.Hi * a.Wi;
  const int tiles_w = (a.Wo + kStemTW - 1) / kStemTW;
  const int nunits = a.N * a.Ho * tiles_w;
  const int u0 = blockIdx.x * units_per_cta, u1 = min(nunits, u0 + units_per_cta);
  const bool act = lane < a.Cout;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  for (int u = u0; u < u1; ++u) {
    const int tw = u % tiles_w, row = u / tiles_w;
    const int n = row / a.Ho, ho = row - n * a.Ho;
    const int wo0 = tw * kStemTW, npx = min(kStemTW, a.Wo - wo0);
    const int wi0 = wo0 * 2 - a.pad_w, hi0 = ho * 2 - a.pad_h;
    __syncthreads();                                   // the previous unit's readers are done
    for (int e = threadIdx.x; e < 9 * (2 * kStemTW + 8); e += 256) {
      const int rr = e / (2 * kStemTW + 8), col = e - rr * (2 * kStemTW + 8);
      const int ci = rr / 3, r = rr - ci * 3;
      const int hi = hi0 + r, wi = wi0 + col;
      float v = 0.f;
      if (hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) v = __ldg(x + ((size_t)n * 3 + ci) * plane + (size_t)hi * a.Wi + wi);
      xs[rr][col] = v;
    }
    __syncthreads();
    const TG* grow = dy + ((size_t)row * a.Wo + wo0) * a.dy_cs + lane;
    for (int j = warp * 2; j < npx; j += 16) {         // pixels j, j+1 of the strip
      const float g0 = act ? ld1<TG>(grow + (size_t)j * a.dy_cs) : 0.f;
      const float g1 = (act && j + 1 < npx) ? ld1<TG>(grow + (size_t)(j + 1) * a.dy_cs) : 0.f;
#pragma unroll
      for (int rr = 0; rr < 9; ++rr) {                 // rr = ci*3 + r
        const float4 v = *reinterpret_cast<const float4*>(&xs[rr][2 * j]);
        const float v4 = xs[rr][2 * j + 4];
        const int ci = rr / 3, r = rr - ci * 3;
        acc[(r * 3 + 0) * 3 + ci] = fmaf(v.x, g0, acc[(r * 3 + 0) * 3 + ci]);
        acc[(r * 3 + 1) * 3 + ci] = fmaf(v.y, g0, acc[(r * 3 + 1) * 3 + ci]);
        acc[(r * 3 + 2) * 3 + ci] = fmaf(v.z, g0, acc[(r * 3 + 2) * 3 + ci]);
        acc[(r * 3 + 0) * 3 + ci] = fmaf(v.z, g1, acc[(r * 3 + 0) * 3 + ci]);
        acc[(r * 3 + 1) * 3 + ci] = fmaf(v.w, g1, acc[(r * 3 + 1) * 3 + ci]);
        acc[(r * 3 + 2) * 3 + ci] = fmaf(v4, g1, acc[(r * 3 + 2) * 3 + ci]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t) red[warp][t][lane] = acc[t];
  __syncthreads();
  for (int e = threadIdx.x; e < 27 * 32; e += 256) {
    const int t = e >> 5, co = e & 31;
    if (co < a.Cout) {
      float v = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) v += red[wv][t][co];
      if (v != 0.f) atomicAdd(a.dw + (size_t)t * a.Cout + co, v);     // t = (tap*3 + ci): the [tap][Cin][Cout] layout
    }
  }
}

// ---------------------------------------------------------------- max-pool 2x2 backward (gather form)
template <typename TX, typename TG>
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const TX* __restrict__ x, const TG* __restrict__ dy,
                                                             TG* __restrict__ dx, int N, int Hi, int Wi, int C, int x_cs,
                                                             int dy_cs, int dx_cs, int accumulate) {
  const long long total = (long long)N * Hi * Wi * C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long p = idx / C;
  const int wi = (int)(p % Wi), hi = (int)((p / Wi) % Hi), n = (int)(p / ((long long)Wi * Hi));
  const int Ho = Hi / 2, Wo = Wi / 2;
  const int ho = hi / 2, wo = wi / 2;
  float g = 0.f;
  if (ho < Ho && wo < Wo) {
    // first maximum in raster order wins (torch max_pool2d_with_indices)
    float best = -INFINITY;
    int bi = 0;
    for (int k = 0; k < 4; ++k) {
      const float v = ld1<TX>(x + ((size_t)((size_t)n * Hi + 2 * ho + (k >> 1)) * Wi + 2 * wo + (k & 1)) * x_cs + c);
      if (v > best) { best = v; bi = k; }
    }
    if (bi == ((hi & 1) << 1 | (wi & 1))) g = ld1<TG>(dy + ((size_t)((size_t)n * Ho + ho) * Wo + wo) * dy_cs + c);
  }
  TG* o = dx + p * dx_cs + c;
  st1<TG>(o, accumulate ? ld1<TG>(o) + g : g);
}

// 16-byte version (bf16, even H and W): a thread owns one pooled pixel x 8 channels -- the four input vectors of its window,
// the (possibly unaligned: the pooled channels sit at channel 29 of DABNet's down-sampler output) gradient values, and the
// four gradient vectors it writes.  The per-element kernel above spent 0.34 ms on DABNet's 8 x 256 x 512 x 35 tensor.
template <bool DYVEC>
__global__ void __launch_bounds__(256) maxpool2x2_bwd_v8_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                                __nv_bfloat16* __restrict__ dx, long long total, int Ho, int Wo, int C,
                                                                int x_cs, int dy_cs, int dx_cs, int accumulate) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ng = (C + 7) / 8;
  const int c = (int)(idx % ng) * 8;
  const long long pp = idx / ng;
  const int wo = (int)(pp % Wo), ho = (int)((pp / Wo) % Ho);
  const long long n = pp / ((long long)Wo * Ho);
  const int Wi = 2 * Wo;
  const size_t p00 = ((size_t)(n * 2 * Ho + 2 * ho) * Wi + 2 * wo);
  float xv[4][8], gv[8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + (p00 + (size_t)(k >> 1) * Wi + (k & 1)) * x_cs + c)), xv[k]);
  const __nv_bfloat16* g = dy + (size_t)pp * dy_cs + c;
  if (DYVEC) {
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(g)), gv);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) gv[j] = (c + j < C) ? __bfloat162float(g[j]) : 0.f;
  }
  float out[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float best = xv[0][j];            // first maximum in raster order wins (torch max_pool2d_with_indices)
    int bi = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (xv[k][j] > best) { best = xv[k][j]; bi = k; }
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k][j] = (bi == k) ? gv[j] : 0.f;
  }
  const bool full = c + 8 <= C;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat16* o = dx + (p00 + (size_t)(k >> 1) * Wi + (k & 1)) * dx_cs + c;
    if (accumulate) {
      float prev[8];
      bf16x8_to_float(*reinterpret_cast<const uint4*>(o), prev);
#pragma unroll
      for (int j = 0; j < 8; ++j) out[k][j] += prev[j];
    }
    if (full) {
      *reinterpret_cast<uint4*>(o) = float_to_bf16x8(out[k]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < C) o[j] = __float2bfloat16_rn(out[k][j]);
    }
  }
}

// ---------------------------------------------------------------- bilinear backward (align_corners=False)
// d low[n,h,w,c] = sum over output pixels of d logits * weight; gather over the <= (2*ceil(1/s)+1)^2 window
template <typename TL, typename TO>
__global__ void __launch_bounds__(128) bilinear_bwd_kernel(const TL* __restrict__ dl, TO* __restrict__ dlow, int N, int C,
                                                           int Hi, int Wi, int Ho, int Wo, int low_cs, float sh, float sw,
                                                           float gscale) {
  const long long total = (long long)N * C * Hi * Wi;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = (int)(idx % Wi);
  const int h = (int)((idx / Wi) % Hi);
  const int c = (int)((idx / ((long long)Wi * Hi)) % C);
  const int n = (int)(idx / ((long long)Wi * Hi * C));
  // The weight of output pixel (ho, wo) on source pixel (h, w) is a product of two hat functions of the clamped source
  // coordinates, max(0, 1 - |fh - h|) * max(0, 1 - |fw - w|) -- the same numbers as (1 - frac, frac) on (floor, floor + 1),
  // including the clamps at both borders -- so only the ~2/s output rows / columns inside the hat are visited (round 1 walked
  // a (2/s + 5)^2 window and rebuilt both index pairs per element: 0.36 ms of DABNet's training step).
  const float rh = 1.f / sh, rw = 1.f / sw;
  const int ho0 = max((int)ceilf(((float)h - 0.5f) * rh - 0.5f) - 1, 0), ho1 = min((int)floorf(((float)h + 1.5f) * rh - 0.5f) + 1, Ho - 1);
  const int wo0 = max((int)ceilf(((float)w - 0.5f) * rw - 0.5f) - 1, 0), wo1 = min((int)floorf(((float)w + 1.5f) * rw - 0.5f) + 1, Wo - 1);
  const TL* plane = dl + ((size_t)n * C + c) * Ho * Wo;
  const float hmax = (float)(Hi - 1), wmax = (float)(Wi - 1);
  float acc = 0.f;
  for (int ho = ho0; ho <= ho1; ++ho) {
    const float fh = fminf(fmaxf(sh * ((float)ho + 0.5f) - 0.5f, 0.f), hmax);
    const float wh = 1.f - fabsf(fh - (float)h);
    if (wh <= 0.f) continue;
    const TL* row = plane + (size_t)ho * Wo;
    float rowacc = 0.f;
    for (int wo = wo0; wo <= wo1; ++wo) {
      const float fw = fminf(fmaxf(sw * ((float)wo + 0.5f) - 0.5f, 0.f), wmax);
      const float ww = fmaxf(1.f - fabsf(fw - (float)w), 0.f);
      rowacc = fmaf(ww, ld1<TL>(row + wo), rowacc);
    }
    acc = fmaf(wh, rowacc, acc);
  }
  st1<TO>(dlow + ((size_t)((size_t)n * Hi + h) * Wi + w) * low_cs + c, acc * gscale);
}

inline long long pick_chunk(long long M, int other_ctas) {
  // ~4 CTAs per SM in total
  long long want = (4LL * 148 + other_ctas - 1) / other_ctas;
  if (want < 1) want = 1;
  long long chunk = (M + want - 1) / want;
  if (chunk < 256) chunk = 256;
  return chunk;
}

}  // namespace

extern "C" int esn_channel_stats(const EsnTensor* x, double* sums, int32_t with_squares, void* stream) {
  if (!x || !sums || !esn_valid_nhwc(*x)) return ESN_ERR_BAD_ARG;
  const long long M = (long long)x->n * x->h * x->w;
  const int cblocks = esn_cdiv(esn_cdiv(x->c, 4), 64);
  const long long chunk = pick_chunk(M, cblocks);
  dim3 grid(esn_cdiv(M, chunk), cblocks);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x->dtype == ESN_BF16 && v8_ok(x->ptr, x->c_stride)) {
    const int cb = esn_cdiv(esn_cdiv(x->c, 8), kV8Groups);
    const long long ch = pick_chunk(M, cb);
    dim3 g8(esn_cdiv(M, ch), cb);
    if (with_squares) channel_stats_v8_kernel<2><<<g8, kStatThreads, 0, st>>>((const __nv_bfloat16*)x->ptr, M, x->c, x->c_stride, sums, ch);
    else channel_stats_v8_kernel<1><<<g8, kStatThreads, 0, st>>>((const __nv_bfloat16*)x->ptr, M, x->c, x->c_stride, sums, ch);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  if (x->dtype == ESN_F32) {
    if (with_squares) channel_stats_kernel<float, 2><<<grid, kStatThreads, 0, st>>>((const float*)x->ptr, M, x->c, x->c_stride, sums, chunk);
    else channel_stats_kernel<float, 1><<<grid, kStatThreads, 0, st>>>((const float*)x->ptr, M, x->c, x->c_stride, sums, chunk);
  } else {
    if (with_squares) channel_stats_kernel<__nv_bfloat16, 2><<<grid, kStatThreads, 0, st>>>((const __nv_bfloat16*)x->ptr, M, x->c, x->c_stride, sums, chunk);
    else channel_stats_kernel<__nv_bfloat16, 1><<<grid, kStatThreads, 0, st>>>((const __nv_bfloat16*)x->ptr, M, x->c, x->c_stride, sums, chunk);
  }
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_bn_finalize(const EsnBnFinalize* p, void* stream) {
  if (!p || !p->sums || !p->scale || !p->shift || !p->mean || !p->invstd || p->channels < 1 || p->count < 1)
    return ESN_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  bn_finalize_kernel<<<esn_cdiv(p->channels, 128), 128, 0, st>>>(p->sums, (double)p->count, p->gamma, p->beta, p->eps,
                                                                 p->momentum, p->running_mean, p->running_var, p->scale,
                                                                 p->shift, p->mean, p->invstd, p->channels);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

static int fill_bn_bwd(const EsnBnBwd* p, BnBwdArgs& a) {
  if (!p || !esn_valid_nhwc(p->x) || !esn_valid_nhwc(p->dy) || !p->sums) return ESN_ERR_BAD_ARG;
  if (p->x.n != p->dy.n || p->x.h != p->dy.h || p->x.w != p->dy.w || p->x.c != p->dy.c) return ESN_ERR_BAD_SHAPE;
  if (p->act == ESN_ACT_PRELU && !p->alpha) return ESN_ERR_BAD_ARG;
  if (p->train_stats && (!p->mean || !p->invstd)) return ESN_ERR_BAD_ARG;
  a.x = p->x.ptr;
  a.dy = p->dy.ptr;
  a.dx = p->dx.ptr;
  a.extra = p->extra.ptr;
  a.M = (long long)p->x.n * p->x.h * p->x.w;
  a.C = p->x.c;
  a.x_cs = p->x.c_stride;
  a.dy_cs = p->dy.c_stride;
  a.dx_cs = p->dx.c_stride;
  a.extra_cs = p->extra.c_stride;
  a.act = p->act;
  a.train_stats = p->train_stats;
  a.scale = p->scale;
  a.shift = p->shift;
  a.alpha = p->alpha;
  a.mean = p->mean;
  a.invstd = p->invstd;
  a.sums = p->sums;
  a.dgamma = p->dgamma;
  a.dbeta = p->dbeta;
  a.dalpha = p->dalpha;
  return ESN_OK;
}

extern "C" int esn_bn_act_bwd_reduce(const EsnBnBwd* p, void* stream) {
  BnBwdArgs a;
  int rc = fill_bn_bwd(p, a);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool xf = p->x.dtype == ESN_F32, gf = p->dy.dtype == ESN_F32;
  if (!xf && !gf && v8_ok(a.x, a.x_cs) && v8_ok(a.dy, a.dy_cs)) {
    const int cb = esn_cdiv(esn_cdiv(a.C, 8), kV8Groups);
    a.px_per_cta = pick_chunk(a.M, cb);
    dim3 g8(esn_cdiv(a.M, a.px_per_cta), cb);
    bn_act_bwd_reduce_v8_kernel<<<g8, kStatThreads, 0, st>>>(a);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  const int cblocks = esn_cdiv(esn_cdiv(a.C, 4), 64);
  a.px_per_cta = pick_chunk(a.M, cblocks);
  dim3 grid(esn_cdiv(a.M, a.px_per_cta), cblocks);
  if (xf && gf) bn_act_bwd_reduce_kernel<float, float><<<grid, kStatThreads, 0, st>>>(a);
  else if (xf) bn_act_bwd_reduce_kernel<float, __nv_bfloat16><<<grid, kStatThreads, 0, st>>>(a);
  else if (gf) bn_act_bwd_reduce_kernel<__nv_bfloat16, float><<<grid, kStatThreads, 0, st>>>(a);
  else bn_act_bwd_reduce_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kStatThreads, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_bn_act_bwd_apply(const EsnBnBwd* p, void* stream) {
  BnBwdArgs a;
  int rc = fill_bn_bwd(p, a);
  if (rc) return rc;
  if (!esn_valid_nhwc(p->dx) || p->dx.c != p->x.c || p->dx.dtype != p->dy.dtype) return ESN_ERR_BAD_ARG;
  if (p->extra.ptr && (!esn_valid_nhwc(p->extra) || p->extra.dtype != p->dx.dtype)) return ESN_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool xf = p->x.dtype == ESN_F32, gf = p->dy.dtype == ESN_F32;
  if (!xf && !gf && v8_ok(a.x, a.x_cs) && v8_ok(a.dy, a.dy_cs) && v8_ok(a.dx, a.dx_cs) && (!a.extra || v8_ok(a.extra, a.extra_cs))) {
    const int cb = esn_cdiv(esn_cdiv(a.C, 8), kV8Groups);
    long long want = (8LL * 148 + cb - 1) / cb;
    a.px_per_cta = (a.M + want - 1) / want;
    if (a.px_per_cta < 64) a.px_per_cta = 64;
    dim3 g8(esn_cdiv(a.M, a.px_per_cta), cb);
    bn_act_bwd_apply_v8_kernel<<<g8, kStatThreads, 0, st>>>(a);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
  const long long total = a.M * ((a.C + 3) / 4);
  const int grid = esn_cdiv(total > a.C ? total : a.C, 256);
  if (xf && gf) bn_act_bwd_apply_kernel<float, float, float><<<grid, 256, 0, st>>>(a);
  else if (xf) bn_act_bwd_apply_kernel<float, __nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else if (gf) bn_act_bwd_apply_kernel<__nv_bfloat16, float, float><<<grid, 256, 0, st>>>(a);
  else bn_act_bwd_apply_kernel<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

bool esn_wgrad_mma_try(const EsnConv* p, void* stream, int* rc);   // esn_wgrad_mma.cu
bool esn_bilinear_bwd_rows_try(const EsnTensor* dy, const EsnTensor* dx, int align_corners, int accumulate, float gscale,
                               void* stream);                      // esn_train3.cu
bool esn_wgrad_umma_try(const EsnConv* p, void* stream, int* rc);  // esn_wgrad_umma.cu (tcgen05, stride 1)
bool esn_wgrad_rows_try(const EsnConv* p, void* stream, int* rc);  // esn_wgrad_rows.cu (mma.sync, dense 3x3, all taps per pass)
bool esn_wgrad_taps3_try(const EsnConv* p, void* stream, int* rc); // esn_wgrad_taps3.cu (mma.sync, dense 3x1 / 1x3)

extern "C" int esn_conv2d_wgrad(const EsnConv* p, void* stream) {
  // p->x: forward input, p->y: gradient of the conv output, p->w: fp32 dW accumulator
  // [kh*kw][Cin/groups][Cout] (same layout as the direct kernel's weights), accumulated with atomics.
  if (!p || !p->w || !esn_valid_nhwc(p->y)) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->x;
  const EsnTensor& dy = p->y;
  const bool nchw = x.layout == ESN_NCHW;
  if (nchw ? (!x.ptr || x.dtype != ESN_F32) : !esn_valid_nhwc(x)) return ESN_ERR_BAD_ARG;
  if (p->transposed) return ESN_ERR_UNSUPPORTED;
  const bool dw = p->groups != 1;
  if (dw && (p->groups != x.c || x.c != dy.c || nchw)) return ESN_ERR_UNSUPPORTED;
  const int eh = (x.h + 2 * p->pad_h - p->dil_h * (p->kh - 1) - 1) / p->stride + 1;
  const int ew = (x.w + 2 * p->pad_w - p->dil_w * (p->kw - 1) - 1) / p->stride + 1;
  if (eh != dy.h || ew != dy.w || x.n != dy.n) return ESN_ERR_BAD_SHAPE;
  {
    int rc = ESN_OK;   // bf16 dense convs: tensor-core path
    if (!dw && !nchw && esn_wgrad_rows_try(p, stream, &rc)) return rc;
    if (!dw && !nchw && esn_wgrad_taps3_try(p, stream, &rc)) return rc;
    if (!dw && !nchw && esn_wgrad_umma_try(p, stream, &rc)) return rc;
    if (!dw && !nchw && esn_wgrad_mma_try(p, stream, &rc)) return rc;
  }
  WgradArgs a;
  a.x = x.ptr;
  a.dy = dy.ptr;
  a.dw = reinterpret_cast<float*>(const_cast<void*>(p->w));
  a.N = x.n; a.Hi = x.h; a.Wi = x.w; a.Cin = x.c; a.x_cs = nchw ? 0 : x.c_stride; a.x_nchw = nchw;
  a.Ho = dy.h; a.Wo = dy.w; a.Cout = dy.c; a.dy_cs = dy.c_stride;
  a.kh = p->kh; a.kw = p->kw; a.stride = p->stride; a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.dil_h = p->dil_h; a.dil_w = p->dil_w;
  const long long M = (long long)dy.n * dy.h * dy.w;
  const int taps = p->kh * p->kw;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool xf = x.dtype == ESN_F32, gf = dy.dtype == ESN_F32;
  static const bool dw_old = getenv("ESN_WGRAD_DW_PER_TAP") != nullptr, stem_old = getenv("ESN_WGRAD_STEM_V1") != nullptr;
  static const bool dw_nostrip = getenv("ESN_WGRAD_DW_NOSTRIP") != nullptr;
  const bool dw_k = (p->kh == 3 && p->kw == 3) || (p->kh == 3 && p->kw == 1) || (p->kh == 1 && p->kw == 3);
  const bool s2_33 = p->stride == 2 && p->kh == 3 && p->kw == 3;
  if (dw && !dw_old && !dw_nostrip && dw_k && (p->stride == 1 || s2_33) && p->dil_h == 1 && p->dil_w == 1 && dy.c % 4 == 0 &&
      x.c_stride % 4 == 0 && dy.c_stride % 4 == 0 && ((uintptr_t)x.ptr % (xf ? 16 : 8)) == 0 &&
      ((uintptr_t)dy.ptr % (gf ? 16 : 8)) == 0) {
    constexpr int LSEG = 4;
    const int cblocks = esn_cdiv(dy.c / 4, 64);
    const int CG = dy.c / 4 < 64 ? dy.c / 4 : 64;
    const int lanes = kStatThreads / CG;
    const int nstrips = esn_cdiv(dy.w, lanes * LSEG);
    const int base_ctas = dy.n * nstrips * cblocks;
    int nchunks = (148 * 4 + base_ctas - 1) / base_ctas;       // >= 4 CTAs per SM worth of work, chunks of >= 16 rows
    if (nchunks > dy.h / 16) nchunks = dy.h / 16;
    if (nchunks < 1) nchunks = 1;
    const int rpc = esn_cdiv(dy.h, nchunks);
    nchunks = esn_cdiv(dy.h, rpc);
    dim3 grid(dy.n * nstrips * nchunks, 1, cblocks);
#define ESN_DWSTRIP(KH, KW)                                                                                                     \
    do {                                                                                                                         \
      if (xf && gf) wgrad_dw_strip_kernel<float, float, KH, KW, LSEG><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks);  \
      else if (xf) wgrad_dw_strip_kernel<float, __nv_bfloat16, KH, KW, LSEG><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks); \
      else if (gf) wgrad_dw_strip_kernel<__nv_bfloat16, float, KH, KW, LSEG><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks); \
      else wgrad_dw_strip_kernel<__nv_bfloat16, __nv_bfloat16, KH, KW, LSEG><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks); \
    } while (0)
    if (s2_33) {
      if (xf && gf) wgrad_dw_strip_kernel<float, float, 3, 3, LSEG, 2><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks);
      else if (xf) wgrad_dw_strip_kernel<float, __nv_bfloat16, 3, 3, LSEG, 2><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks);
      else if (gf) wgrad_dw_strip_kernel<__nv_bfloat16, float, 3, 3, LSEG, 2><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks);
      else wgrad_dw_strip_kernel<__nv_bfloat16, __nv_bfloat16, 3, 3, LSEG, 2><<<grid, kStatThreads, 0, st>>>(a, rpc, nstrips, nchunks);
    } else if (p->kh == 3 && p->kw == 3) ESN_DWSTRIP(3, 3);
    else if (p->kh == 3) ESN_DWSTRIP(3, 1);
    else ESN_DWSTRIP(1, 3);
#undef ESN_DWSTRIP
  } else if (dw && !dw_old && dw_k) {
    const int cblocks = esn_cdiv(esn_cdiv(dy.c, 4), 64);
    const int R = dy.n * dy.h;
    int want = (148 * 8) / cblocks;                     // CTAs along the row axis (8 x 256 threads per SM)
    if (want < 1) want = 1;
    int rpc = esn_cdiv(R, want);
    if (rpc < 1) rpc = 1;
    dim3 grid(esn_cdiv(R, rpc), 1, cblocks);
#define ESN_DWALL(KH, KW)                                                                                         \
    do {                                                                                                           \
      if (xf && gf) wgrad_dw_all_kernel<float, float, KH, KW><<<grid, kStatThreads, 0, st>>>(a, rpc);              \
      else if (xf) wgrad_dw_all_kernel<float, __nv_bfloat16, KH, KW><<<grid, kStatThreads, 0, st>>>(a, rpc);       \
      else if (gf) wgrad_dw_all_kernel<__nv_bfloat16, float, KH, KW><<<grid, kStatThreads, 0, st>>>(a, rpc);       \
      else wgrad_dw_all_kernel<__nv_bfloat16, __nv_bfloat16, KH, KW><<<grid, kStatThreads, 0, st>>>(a, rpc);       \
    } while (0)
    if (p->kh == 3 && p->kw == 3) ESN_DWALL(3, 3);
    else if (p->kh == 3) ESN_DWALL(3, 1);
    else ESN_DWALL(1, 3);
#undef ESN_DWALL
  } else if (dw) {
    const int cblocks = esn_cdiv(esn_cdiv(dy.c, 4), 64);
    a.px_per_cta = pick_chunk(M, taps * cblocks);
    dim3 grid(esn_cdiv(M, a.px_per_cta), taps, cblocks);
    if (xf && gf) wgrad_dw_kernel<float, float><<<grid, kStatThreads, 0, st>>>(a);
    else if (xf) wgrad_dw_kernel<float, __nv_bfloat16><<<grid, kStatThreads, 0, st>>>(a);
    else if (gf) wgrad_dw_kernel<__nv_bfloat16, float><<<grid, kStatThreads, 0, st>>>(a);
    else wgrad_dw_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kStatThreads, 0, st>>>(a);
  } else if (nchw && x.c == 3 && p->kh == 3 && p->kw == 3 && dy.c <= 32 && p->stride == 2 && p->dil_h == 1 && p->dil_w == 1 &&
             !stem_old) {
    const int nunits = dy.n * dy.h * esn_cdiv(dy.w, kStemTW);
    int upc = esn_cdiv(nunits, 148 * 4);
    if (upc < 1) upc = 1;
    const int grid = esn_cdiv(nunits, upc);
    if (gf) wgrad_stem2_kernel<float><<<grid, 256, 0, st>>>(a, upc);
    else wgrad_stem2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, upc);
  } else if (nchw && x.c == 3 && p->kh == 3 && p->kw == 3 && dy.c <= 32) {
    a.px_per_cta = (M + 148 * 8 - 1) / (148 * 8);
    if (a.px_per_cta < 256) a.px_per_cta = 256;
    const int grid = esn_cdiv(M, a.px_per_cta);
    if (gf) wgrad_stem_kernel<float><<<grid, 256, 0, st>>>(a);
    else wgrad_stem_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
  } else if (taps * x.c * dy.c <= 128) {
    a.px_per_cta = (M + 148 * 16 - 1) / (148 * 16);
    if (a.px_per_cta < 64) a.px_per_cta = 64;
    const int grid = esn_cdiv(M, a.px_per_cta);
    if (xf && gf) wgrad_small_kernel<float, float><<<grid, 128, 0, st>>>(a);
    else if (xf) wgrad_small_kernel<float, __nv_bfloat16><<<grid, 128, 0, st>>>(a);
    else if (gf) wgrad_small_kernel<__nv_bfloat16, float><<<grid, 128, 0, st>>>(a);
    else wgrad_small_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 128, 0, st>>>(a);
  } else {
    const int tiles = esn_cdiv(x.c, 64) * esn_cdiv(dy.c, 64);
    a.px_per_cta = pick_chunk(M, taps * tiles);
    a.px_per_cta = (a.px_per_cta + kWgPix - 1) / kWgPix * kWgPix;
    dim3 grid(esn_cdiv(M, a.px_per_cta), taps, tiles);
    if (xf && gf) wgrad_dense_kernel<float, float><<<grid, 256, 0, st>>>(a);
    else if (xf) wgrad_dense_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
    else if (gf) wgrad_dense_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(a);
    else wgrad_dense_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  }
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_maxpool2x2_bwd(const EsnTensor* x, const EsnTensor* dy, const EsnTensor* dx, int32_t accumulate,
                                  void* stream) {
  if (!x || !dy || !dx || !esn_valid_nhwc(*x) || !esn_valid_nhwc(*dy) || !esn_valid_nhwc(*dx)) return ESN_ERR_BAD_ARG;
  if (dy->h != x->h / 2 || dy->w != x->w / 2 || dy->c != x->c || dx->h != x->h || dx->w != x->w || dx->c != x->c ||
      dx->dtype != dy->dtype)
    return ESN_ERR_BAD_SHAPE;
  const long long total = (long long)x->n * x->h * x->w * x->c;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = esn_cdiv(total, 256);
  const bool xf = x->dtype == ESN_F32, gf = dy->dtype == ESN_F32;
  if (!xf && !gf && v8_ok(x->ptr, x->c_stride) && v8_ok(dx->ptr, dx->c_stride) && !((x->h | x->w) & 1)) {
    const long long groups = (long long)dy->n * dy->h * dy->w * ((x->c + 7) / 8);
    const int g8 = esn_cdiv(groups, 256);
    if (v8_ok(dy->ptr, dy->c_stride))
      maxpool2x2_bwd_v8_kernel<true><<<g8, 256, 0, st>>>((const __nv_bfloat16*)x->ptr, (const __nv_bfloat16*)dy->ptr, (__nv_bfloat16*)dx->ptr,
                                                         groups, dy->h, dy->w, x->c, x->c_stride, dy->c_stride, dx->c_stride, accumulate);
    else
      maxpool2x2_bwd_v8_kernel<false><<<g8, 256, 0, st>>>((const __nv_bfloat16*)x->ptr, (const __nv_bfloat16*)dy->ptr, (__nv_bfloat16*)dx->ptr,
                                                          groups, dy->h, dy->w, x->c, x->c_stride, dy->c_stride, dx->c_stride, accumulate);
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
#define ESN_MPB(TX, TG)                                                                                              \
  maxpool2x2_bwd_kernel<TX, TG><<<grid, 256, 0, st>>>((const TX*)x->ptr, (const TG*)dy->ptr, (TG*)dx->ptr, x->n, x->h, \
                                                      x->w, x->c, x->c_stride, dy->c_stride, dx->c_stride, accumulate)
  if (xf && gf) ESN_MPB(float, float);
  else if (xf) ESN_MPB(float, __nv_bfloat16);
  else if (gf) ESN_MPB(__nv_bfloat16, float);
  else ESN_MPB(__nv_bfloat16, __nv_bfloat16);
#undef ESN_MPB
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}

extern "C" int esn_bilinear_bwd(const EsnTensor* dlogits, const EsnTensor* dlow, float gscale, void* stream) {
  if (!dlogits || !dlow || !dlogits->ptr || !esn_valid_nhwc(*dlow)) return ESN_ERR_BAD_ARG;
  if (dlogits->layout != ESN_NCHW || dlogits->n != dlow->n || dlogits->c != dlow->c) return ESN_ERR_BAD_SHAPE;
  const long long total = (long long)dlow->n * dlow->c * dlow->h * dlow->w;
  const int grid = esn_cdiv(total, 128);
  const float sh = (float)dlow->h / (float)dlogits->h, sw = (float)dlow->w / (float)dlogits->w;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool lf = dlogits->dtype == ESN_F32, of = dlow->dtype == ESN_F32;
  if (esn_bilinear_bwd_rows_try(dlogits, dlow, 0, 0, gscale, stream)) {
    ESN_CHECK_LAUNCH();
    return ESN_OK;
  }
#define ESN_BLB(TL, TO)                                                                                            \
  bilinear_bwd_kernel<TL, TO><<<grid, 128, 0, st>>>((const TL*)dlogits->ptr, (TO*)dlow->ptr, dlow->n, dlow->c,      \
                                                    dlow->h, dlow->w, dlogits->h, dlogits->w, dlow->c_stride, sh, sw, \
                                                    gscale)
  if (lf && of) ESN_BLB(float, float);
  else if (lf) ESN_BLB(float, __nv_bfloat16);
  else if (of) ESN_BLB(__nv_bfloat16, float);
  else ESN_BLB(__nv_bfloat16, __nv_bfloat16);
#undef ESN_BLB
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
