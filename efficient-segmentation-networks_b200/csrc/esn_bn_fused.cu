// Train-mode BatchNorm + activation in ONE launch per direction (round 2).
//
// Round 1 ran a BatchNorm layer as three launches forward (esn_channel_stats -> esn_bn_finalize -> esn_affine_act) and two
// backward (esn_bn_act_bwd_reduce -> esn_bn_act_bwd_apply).  DABNet's training step has 71 such layers on 8-67 MB tensors:
// 355 of its 570 launches, each 5-30 us long, each reading its operands from DRAM again.  Here a layer is one CO-RESIDENT
// grid (cooperative launch, ~4 CTAs per SM) with one grid-wide barrier in the middle:
//
//   forward : [per-channel sum / sum of squares over the CTA's pixel chunk, fp64 atomics into one of 8 replicas]  -- barrier --
//             [every CTA derives scale / shift from the finished sums (the x == 0 CTAs also store scale / shift / mean / invstd
//              and update the running statistics), then normalises + activates THE SAME pixel chunk, which is still in L2]
//   backward: [sum dz, sum dz*xhat, sum dy*z*[z<0] over the chunk]  -- barrier --
//             [dx = scale*(dz - mean dz - xhat * mean dz*xhat) (+ extra) over the same chunk; parameter gradients stored]
//
// so x (and dy) cross the DRAM interface once and the layer costs one launch.  bf16 NHWC, 16-byte accesses; the entry points
// return ESN_ERR_UNSUPPORTED for anything else and the caller keeps the multi-launch path (fp32 parity runs).
// Semantics are those of esn_train.cu's kernels (torch: the biased batch variance normalises, running_var gets the unbiased
// one): replaces aten::native_batch_norm(training=True) + _prelu_kernel / threshold and their backward
// (ERFNet.py:21,38,45,107; DABNet.py:41; train.py:351-356).
#include "esn_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBlockChannels = 256;                  // channels per blockIdx.y
constexpr int kRep = ESN_BN_FUSED_REPLICAS;          // accumulator replicas: CTA x adds into replica x % kRep, so one address
                                                     // takes grid/kRep atomics instead of grid (444 same-address fp64 atomics
                                                     // cost 5.5 us of a 17 us layer, measured)

template <int VEC> struct Pack;
template <> struct Pack<8> { using T = uint4; };
template <> struct Pack<4> { using T = uint2; };

__device__ __forceinline__ void unpack(const uint4& r, float* f) { bf16x8_to_float(r, f); }
__device__ __forceinline__ void unpack(const uint2& r, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void pack(const float* f, uint4& r) { r = float_to_bf16x8(f); }
__device__ __forceinline__ void pack(const float* f, uint2& r) {
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
  h[0] = __floats2bfloat162_rn(f[0], f[1]);
  h[1] = __floats2bfloat162_rn(f[2], f[3]);
}
__device__ __forceinline__ uint4 zero_of(uint4) { return make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ uint2 zero_of(uint2) { return make_uint2(0, 0); }

// All CTAs of the grid are resident (cooperative launch), `bar` was zero when the kernel started.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < expected);
  }
  __syncthreads();
}

__device__ __forceinline__ float act_grad(float z, float dy, int act, float alpha) {
  if (act == ESN_ACT_RELU) return z > 0.f ? dy : 0.f;
  if (act == ESN_ACT_PRELU) return z >= 0.f ? dy : dy * alpha;
  return dy;
}

// sum over the kRep replicas of quantity q of channel c (after the barrier: L2 loads, never the non-coherent path)
__device__ __forceinline__ double total_of(const double* sums, int nq, int C, int q, int c) {
  double t = 0.0;
#pragma unroll
  for (int r = 0; r < kRep; ++r) t += __ldcg(sums + ((size_t)r * nq + q) * C + c);
  return t;
}

struct FwdArgs {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  long long M, px_per_cta;
  int C, x_cs, y_cs, act;
  double* sums;          // [kRep][2][C], zero on entry
  unsigned int* bar;     // zero on entry
  double count;
  const float *gamma, *beta, *alpha;
  float eps, momentum;
  float *running_mean, *running_var, *scale, *shift, *mean, *invstd;
};

// A thread owns VEC channels and every `lanes`-th pixel of the CTA's chunk; U independent loads in flight, tail predicated
// (no serial remainder loop: a layer on an 8 MB tensor is 1-2 rounds per phase, so every dependent round trip counts).
template <int VEC, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) bn_act_train_fwd_kernel(const FwdArgs a) {
  using P = typename Pack<VEC>::T;
  constexpr int G = kBlockChannels / VEC;
  __shared__ float red[2][kThreads][VEC + 1];
  __shared__ float s_sc[kBlockChannels], s_sh[kBlockChannels];
  const int C = a.C;
  const int ng = (C + VEC - 1) / VEC;
  const int CG = min(ng - (int)blockIdx.y * G, G);
  const int lanes = kThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * G + cg) * VEC;
  const long long p0 = blockIdx.x * a.px_per_cta;
  const int npx = (int)(min(a.M, p0 + a.px_per_cta) - p0);      // pixels of this CTA's chunk; 32-bit offsets inside it
  const __nv_bfloat16* xb = a.x + p0 * a.x_cs + c;
  const bool active = pl < lanes;

  // ---- phase 1: statistics of the chunk
  float s[VEC], q[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.f;
  if (active) {
    for (int p = pl; p < npx; p += U * lanes) {
      P raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pp = p + u * lanes;
        raw[u] = pp < npx ? __ldg(reinterpret_cast<const P*>(xb + pp * a.x_cs)) : zero_of(P());
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[VEC];
        unpack(raw[u], v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          s[j] += v[j];
          q[j] = fmaf(v[j], v[j], q[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    red[0][threadIdx.x][j] = s[j];
    red[1][threadIdx.x][j] = q[j];
  }
  __syncthreads();
  const bool publisher = (int)threadIdx.x < CG * VEC;
  const int pg = threadIdx.x / VEC, pj = threadIdx.x % VEC;
  const int pc = blockIdx.y * kBlockChannels + threadIdx.x;       // the channel this thread publishes / finalizes
  if (publisher && pc < C) {
    float a0 = 0.f, a1 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a0 += red[0][l * CG + pg][pj];
      a1 += red[1][l * CG + pg][pj];
    }
    double* dst = a.sums + (size_t)(blockIdx.x % kRep) * 2 * C;
    atomicAdd(dst + pc, (double)a0);
    atomicAdd(dst + C + pc, (double)a1);
  }

  grid_barrier(a.bar, gridDim.x * gridDim.y);

  // ---- finalize (every CTA for its own channels; the first pixel chunk stores the layer's record)
  if (publisher) {
    float sc = 0.f, sh = 0.f;
    if (pc < C) {
      const double m = total_of(a.sums, 2, C, 0, pc) / a.count;
      double var = total_of(a.sums, 2, C, 1, pc) / a.count - m * m;   // biased variance normalises (torch semantics)
      if (var < 0) var = 0;
      const float is = (float)(1.0 / sqrt(var + (double)a.eps));
      const float g = a.gamma ? a.gamma[pc] : 1.f, b = a.beta ? a.beta[pc] : 0.f;
      sc = g * is;
      sh = b - (float)m * sc;
      if (blockIdx.x == 0) {
        a.scale[pc] = sc;
        a.shift[pc] = sh;
        a.mean[pc] = (float)m;
        a.invstd[pc] = is;
        if (a.running_mean) a.running_mean[pc] = (1.f - a.momentum) * a.running_mean[pc] + a.momentum * (float)m;
        if (a.running_var) {
          const double unbiased = a.count > 1 ? var * a.count / (a.count - 1) : var;
          a.running_var[pc] = (1.f - a.momentum) * a.running_var[pc] + a.momentum * (float)unbiased;
        }
      }
    }
    s_sc[threadIdx.x] = sc;
    s_sh[threadIdx.x] = sh;
  }
  __syncthreads();
  if (!active) return;

  // ---- phase 2: normalise + activate the same chunk (L2-resident)
  float sc[VEC], sh[VEC], al[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sc[j] = s_sc[cg * VEC + j];
    sh[j] = s_sh[cg * VEC + j];
    al[j] = (a.act == ESN_ACT_PRELU) ? a.alpha[min(c + j, C - 1)] : 0.f;
  }
  const bool full = c + VEC <= C;
  __nv_bfloat16* yb = a.y + p0 * a.y_cs + c;
  const int act = a.act;
  for (int p = pl; p < npx; p += U * lanes) {
    P raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      raw[u] = pp < npx ? __ldg(reinterpret_cast<const P*>(xb + pp * a.x_cs)) : zero_of(P());
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < npx) {
        float v[VEC];
        unpack(raw[u], v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = apply_act(fmaf(v[j], sc[j], sh[j]), act, al[j]);
        __nv_bfloat16* o = yb + pp * a.y_cs;
        if (full) {
          P r;
          pack(v, r);
          *reinterpret_cast<P*>(o) = r;
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (c + j < C) o[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
  }
}

struct BwdArgs {
  const __nv_bfloat16 *x, *dy, *extra;
  __nv_bfloat16* dx;
  long long M, px_per_cta;
  int C, x_cs, dy_cs, dx_cs, extra_cs, act;
  const float *scale, *shift, *alpha, *mean, *invstd;
  double* sums;          // [kRep][3][C], zero on entry
  unsigned int* bar;     // zero on entry
  float *dgamma, *dbeta, *dalpha;
};

template <int VEC, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) bn_act_bwd_fused_kernel(const BwdArgs a) {
  using P = typename Pack<VEC>::T;
  constexpr int G = kBlockChannels / VEC;
  __shared__ float red[3][kThreads][VEC + 1];
  __shared__ float s_k0[kBlockChannels], s_k1[kBlockChannels];
  const int C = a.C;
  const int ng = (C + VEC - 1) / VEC;
  const int CG = min(ng - (int)blockIdx.y * G, G);
  const int lanes = kThreads / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int c = (blockIdx.y * G + cg) * VEC;
  const long long p0 = blockIdx.x * a.px_per_cta;
  const int npx = (int)(min(a.M, p0 + a.px_per_cta) - p0);
  const __nv_bfloat16* xb = a.x + p0 * a.x_cs + c;
  const __nv_bfloat16* gb = a.dy + p0 * a.dy_cs + c;
  const int act = a.act;
  const bool active = pl < lanes;
  float sc[VEC], sh[VEC], al[VEC], mu[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int cc = min(c + j, C - 1);
    sc[j] = a.scale[cc];
    sh[j] = a.shift[cc];
    al[j] = (act == ESN_ACT_PRELU) ? a.alpha[cc] : 0.f;
    mu[j] = a.mean[cc];
  }

  // ---- phase 1: sum dz, sum dz*(x - mean), sum dy*z*[z<0] over the chunk
  float s0[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s0[j] = s1[j] = s2[j] = 0.f;
  if (active) {
    for (int p = pl; p < npx; p += U * lanes) {
      P rx[U], rg[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pp = p + u * lanes;
        const bool ok = pp < npx;
        rx[u] = ok ? __ldg(reinterpret_cast<const P*>(xb + pp * a.x_cs)) : zero_of(P());
        rg[u] = ok ? __ldg(reinterpret_cast<const P*>(gb + pp * a.dy_cs)) : zero_of(P());
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float xv[VEC], gv[VEC];
        unpack(rx[u], xv);
        unpack(rg[u], gv);                      // a predicated-off pixel has dy = 0: it adds nothing to any sum
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float z = fmaf(xv[j], sc[j], sh[j]);
          const float dz = act_grad(z, gv[j], act, al[j]);
          s0[j] += dz;
          s1[j] = fmaf(dz, xv[j] - mu[j], s1[j]);
          s2[j] += (z < 0.f) ? gv[j] * z : 0.f;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    red[0][threadIdx.x][j] = s0[j];
    red[1][threadIdx.x][j] = s1[j];
    red[2][threadIdx.x][j] = s2[j];
  }
  __syncthreads();
  const bool publisher = (int)threadIdx.x < CG * VEC;
  const int pg = threadIdx.x / VEC, pj = threadIdx.x % VEC;
  const int pc = blockIdx.y * kBlockChannels + threadIdx.x;
  if (publisher && pc < C) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a0 += red[0][l * CG + pg][pj];
      a1 += red[1][l * CG + pg][pj];
      a2 += red[2][l * CG + pg][pj];
    }
    double* dst = a.sums + (size_t)(blockIdx.x % kRep) * 3 * C;
    atomicAdd(dst + pc, (double)a0);
    atomicAdd(dst + C + pc, (double)a1 * (double)a.invstd[pc]);      // = sum dz * xhat
    atomicAdd(dst + 2 * C + pc, (double)a2);
  }

  grid_barrier(a.bar, gridDim.x * gridDim.y);

  // per-channel constants of the apply pass; parameter gradients stored once, by the first pixel chunk
  if (publisher) {
    float k0 = 0.f, k1 = 0.f;
    if (pc < C) {
      const double t0 = total_of(a.sums, 3, C, 0, pc), t1 = total_of(a.sums, 3, C, 1, pc);
      const float invM = (float)(1.0 / (double)a.M);
      const float scp = a.scale[pc];
      k0 = scp * (float)t0 * invM;
      k1 = scp * a.invstd[pc] * (float)t1 * invM;
      if (blockIdx.x == 0) {
        if (a.dbeta) a.dbeta[pc] = (float)t0;
        if (a.dgamma) a.dgamma[pc] = (float)t1;
        if (a.dalpha && act == ESN_ACT_PRELU) a.dalpha[pc] = (float)total_of(a.sums, 3, C, 2, pc);
      }
    }
    s_k0[threadIdx.x] = k0;
    s_k1[threadIdx.x] = k1;
  }
  __syncthreads();
  if (!active) return;

  // ---- phase 2: dx over the same chunk:  dx = sc*dz - k0 - (x - mu)*k1 (+ extra) = sc*dz - x*k1 + (mu*k1 - k0) (+ extra)
  float k1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    k1[j] = -s_k1[cg * VEC + j];
    mu[j] = fmaf(mu[j], s_k1[cg * VEC + j], -s_k0[cg * VEC + j]);      // mu now holds the constant term
  }
  const bool full = c + VEC <= C;
  __nv_bfloat16* db = a.dx + p0 * a.dx_cs + c;
  const __nv_bfloat16* eb = a.extra ? a.extra + p0 * a.extra_cs + c : nullptr;
  for (int p = pl; p < npx; p += U * lanes) {
    P rx[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      const bool ok = pp < npx;
      rx[u] = ok ? __ldg(reinterpret_cast<const P*>(xb + pp * a.x_cs)) : zero_of(P());
      rg[u] = ok ? __ldg(reinterpret_cast<const P*>(gb + pp * a.dy_cs)) : zero_of(P());
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < npx) {
        float xv[VEC], gv[VEC], out[VEC];
        unpack(rx[u], xv);
        unpack(rg[u], gv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float z = fmaf(xv[j], sc[j], sh[j]);
          const float dz = act_grad(z, gv[j], act, al[j]);
          out[j] = fmaf(sc[j], dz, fmaf(k1[j], xv[j], mu[j]));
        }
        if (eb) {      // a second consumer's gradient (few layers); may alias dx: coherent load
          float ev[VEC];
          unpack(*reinterpret_cast<const P*>(eb + pp * a.extra_cs), ev);
#pragma unroll
          for (int j = 0; j < VEC; ++j) out[j] += ev[j];
        }
        __nv_bfloat16* o = db + pp * a.dx_cs;
        if (full) {
          P r;
          pack(out, r);
          *reinterpret_cast<P*>(o) = r;
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (c + j < C) o[j] = __float2bfloat16_rn(out[j]);
        }
      }
    }
  }
}

// Backward of a bare activation (conv bias + ReLU, no BatchNorm, no PReLU slope to learn): dx = dy * act'(x*scale + shift)
// * scale (+ extra).  No sums, no barrier: one streaming pass with kU independent 16-byte loads per operand in flight (the
// general apply kernel of esn_train.cu keeps one: 29 us on ERFNet's 17 MB tensors, 51 launches per step).
template <int U>
__global__ void __launch_bounds__(kThreads) act_bwd_kernel(const BwdArgs a) {
  const long long nvec = a.M * (a.C / 8);
  const long long stride = (long long)gridDim.x * kThreads;
  const int ng = a.C / 8;
  for (long long i0 = blockIdx.x * (long long)kThreads + threadIdx.x; i0 < nvec; i0 += stride * U) {
    uint4 rx[U], rg[U];
    long long pix[U];
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      const bool ok = i < nvec;
      pix[u] = ok ? i / ng : 0;
      c[u] = ok ? (int)(i % ng) * 8 : 0;
      rx[u] = ok ? __ldg(reinterpret_cast<const uint4*>(a.x + pix[u] * a.x_cs + c[u])) : make_uint4(0, 0, 0, 0);
      rg[u] = ok ? __ldg(reinterpret_cast<const uint4*>(a.dy + pix[u] * a.dy_cs + c[u])) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= nvec) continue;
      float xv[8], gv[8], out[8];
      bf16x8_to_float(rx[u], xv);
      bf16x8_to_float(rg[u], gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sc = a.scale ? a.scale[c[u] + j] : 1.f, sh = a.shift ? a.shift[c[u] + j] : 0.f;
        out[j] = sc * act_grad(fmaf(xv[j], sc, sh), gv[j], a.act, 0.f);
      }
      if (a.extra) {
        float ev[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(a.extra + pix[u] * a.extra_cs + c[u]), ev);
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] += ev[j];
      }
      *reinterpret_cast<uint4*>(a.dx + pix[u] * a.dx_cs + c[u]) = float_to_bf16x8(out);
    }
  }
}

inline bool v8(const void* p, int cs) { return p && cs % 8 == 0 && (reinterpret_cast<uintptr_t>(p) % 16) == 0; }

// CTAs that fit on the device at once (per kernel variant, per device; computed on first use)
template <typename K>
int resident_ctas(K kernel, int* cache) {
  const int dev = esn_current_device();
  if (cache[dev] == 0) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || per_sm < 1 || sms < 1) {
      cudaGetLastError();
      cache[dev] = -1;
    } else {
      cache[dev] = per_sm * sms;
    }
  }
  return cache[dev];
}

// pixel chunk so that (chunks x channel blocks) <= the co-resident capacity; >= 256 pixels per CTA, the same floor as the
// multi-launch kernels (esn_train.cu: pick_chunk), so that small layers sum their statistics in exactly the same order on both
// paths: train-mode BatchNorm over a handful of values (Fast-SCNN's 1x1 pyramid level at batch 2 is a sign function of the
// difference of two numbers) turns a last-bit difference of the mean into a different network output
inline long long plan_chunk(long long M, int cblocks, int capacity, int* gx) {
  long long want = capacity / cblocks;
  if (want < 1) want = 1;
  long long chunk = (M + want - 1) / want;
  if (chunk < 256) chunk = 256;
  *gx = esn_cdiv(M, chunk);
  return chunk;
}

template <typename K, typename A>
int launch_coop(K kernel, dim3 grid, const A& args, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args);
  g_esn_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (getenv("ESN_DEBUG")) fprintf(stderr, "esn: %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e));
    return ESN_ERR_CUDA;
  }
  return ESN_OK;
}

constexpr int kVariants = 4;
int g_fwd_capacity[kVariants][kEsnMaxDevices], g_bwd_capacity[kVariants][kEsnMaxDevices];

template <typename K>
int run_fwd(K kernel, int variant, FwdArgs& a, cudaStream_t st) {
  const int capacity = resident_ctas(kernel, g_fwd_capacity[variant]);
  const int cb = esn_cdiv(a.C, kBlockChannels);
  if (capacity < 1 || cb > capacity) return ESN_ERR_UNSUPPORTED;
  int gx = 1;
  a.px_per_cta = plan_chunk(a.M, cb, capacity, &gx);
  if (a.px_per_cta * (long long)max(a.x_cs, a.y_cs) >= (1LL << 31)) return ESN_ERR_UNSUPPORTED;
  return launch_coop(kernel, dim3(gx, cb), a, st);
}
template <typename K>
int run_bwd(K kernel, int variant, BwdArgs& a, cudaStream_t st) {
  const int capacity = resident_ctas(kernel, g_bwd_capacity[variant]);
  const int cb = esn_cdiv(a.C, kBlockChannels);
  if (capacity < 1 || cb > capacity) return ESN_ERR_UNSUPPORTED;
  int gx = 1;
  a.px_per_cta = plan_chunk(a.M, cb, capacity, &gx);
  if (a.px_per_cta * (long long)max(max(a.x_cs, a.dy_cs), max(a.dx_cs, a.extra_cs)) >= (1LL << 31)) return ESN_ERR_UNSUPPORTED;
  return launch_coop(kernel, dim3(gx, cb), a, st);
}

int variant_of(const char* env, int dflt) {
  const char* v = getenv(env);          // tuning switch (tools/bench_bn.py); the default is the measured best
  const int k = v ? atoi(v) : dflt;
  return (k < 0 || k >= kVariants) ? dflt : k;
}

}  // namespace

extern "C" int esn_bn_act_train_fwd(const EsnBnTrainFwd* p, void* stream) {
  if (!p || !esn_valid_nhwc(p->x) || !esn_valid_nhwc(p->y) || !p->fin.sums || !p->barrier || !p->fin.scale || !p->fin.shift ||
      !p->fin.mean || !p->fin.invstd || p->fin.count < 1)
    return ESN_ERR_BAD_ARG;
  const EsnTensor &x = p->x, &y = p->y;
  if (x.n != y.n || x.h != y.h || x.w != y.w || x.c != y.c || p->fin.channels != x.c) return ESN_ERR_BAD_SHAPE;
  if (p->act == ESN_ACT_PRELU && !p->alpha) return ESN_ERR_BAD_ARG;
  if (x.dtype != ESN_BF16 || y.dtype != ESN_BF16 || !v8(x.ptr, x.c_stride) || !v8(y.ptr, y.c_stride) || x.ptr == y.ptr)
    return ESN_ERR_UNSUPPORTED;
  FwdArgs a;
  a.x = (const __nv_bfloat16*)x.ptr;
  a.y = (__nv_bfloat16*)y.ptr;
  a.M = (long long)x.n * x.h * x.w;
  a.C = x.c;
  a.x_cs = x.c_stride;
  a.y_cs = y.c_stride;
  a.act = p->act;
  a.sums = const_cast<double*>(p->fin.sums);
  a.bar = p->barrier;
  a.count = (double)p->fin.count;
  a.gamma = p->fin.gamma;
  a.beta = p->fin.beta;
  a.alpha = p->alpha;
  a.eps = p->fin.eps;
  a.momentum = p->fin.momentum;
  a.running_mean = p->fin.running_mean;
  a.running_var = p->fin.running_var;
  a.scale = p->fin.scale;
  a.shift = p->fin.shift;
  a.mean = p->fin.mean;
  a.invstd = p->fin.invstd;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const int variant = variant_of("ESN_BN_FWD_VARIANT", 0);
  switch (variant) {
    // measured equal within 5 % on B200 (11.2-11.6 us on an 8.4 MB tensor, 39-45 us on 67 MB); <8,4,3> is the fastest overall
    case 1: return run_fwd(bn_act_train_fwd_kernel<4, 8, 4>, 1, a, st);
    case 2: return run_fwd(bn_act_train_fwd_kernel<8, 8, 2>, 2, a, st);
    case 3: return run_fwd(bn_act_train_fwd_kernel<8, 2, 4>, 3, a, st);
    default: return run_fwd(bn_act_train_fwd_kernel<8, 4, 3>, 0, a, st);
  }
}

extern "C" int esn_bn_act_bwd_fused(const EsnBnBwd* p, uint32_t* barrier, void* stream) {
  if (!p || !barrier || !esn_valid_nhwc(p->x) || !esn_valid_nhwc(p->dy) || !esn_valid_nhwc(p->dx) || !p->sums)
    return ESN_ERR_BAD_ARG;
  if (p->x.n != p->dy.n || p->x.h != p->dy.h || p->x.w != p->dy.w || p->x.c != p->dy.c || p->dx.c != p->x.c ||
      p->dx.n != p->x.n || p->dx.h != p->x.h || p->dx.w != p->x.w)
    return ESN_ERR_BAD_SHAPE;
  if (p->act == ESN_ACT_PRELU && !p->alpha) return ESN_ERR_BAD_ARG;
  if (p->extra.ptr && !esn_valid_nhwc(p->extra)) return ESN_ERR_BAD_ARG;
  if (!p->train_stats || !p->scale || !p->shift || !p->mean || !p->invstd) return ESN_ERR_UNSUPPORTED;
  if (p->x.dtype != ESN_BF16 || p->dy.dtype != ESN_BF16 || p->dx.dtype != ESN_BF16 ||
      (p->extra.ptr && p->extra.dtype != ESN_BF16) || !v8(p->x.ptr, p->x.c_stride) || !v8(p->dy.ptr, p->dy.c_stride) ||
      !v8(p->dx.ptr, p->dx.c_stride) || (p->extra.ptr && !v8(p->extra.ptr, p->extra.c_stride)) || p->dx.ptr == p->dy.ptr ||
      p->dx.ptr == p->x.ptr)
    return ESN_ERR_UNSUPPORTED;
  BwdArgs a;
  a.x = (const __nv_bfloat16*)p->x.ptr;
  a.dy = (const __nv_bfloat16*)p->dy.ptr;
  a.extra = (const __nv_bfloat16*)p->extra.ptr;
  a.dx = (__nv_bfloat16*)p->dx.ptr;
  a.M = (long long)p->x.n * p->x.h * p->x.w;
  a.C = p->x.c;
  a.x_cs = p->x.c_stride;
  a.dy_cs = p->dy.c_stride;
  a.dx_cs = p->dx.c_stride;
  a.extra_cs = p->extra.c_stride;
  a.act = p->act;
  a.scale = p->scale;
  a.shift = p->shift;
  a.alpha = p->alpha;
  a.mean = p->mean;
  a.invstd = p->invstd;
  a.sums = p->sums;
  a.bar = barrier;
  a.dgamma = p->dgamma;
  a.dbeta = p->dbeta;
  a.dalpha = p->dalpha;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const int variant = variant_of("ESN_BN_BWD_VARIANT", 0);
  switch (variant) {
    // measured on B200 (tools/bench_bn.py, 8.4 / 16.8 / 33.6 / 67 MB tensors, us per layer): <8,4,2> 16.7 / 26.4 / 44.3 / 86.8,
    // <4,4,4> 19.5 / 31.2 / 52.8 / 97.6, <8,2,3> 27.9 / 44.8 / 84.7 / 161, <4,8,3> 36.4 / 56.3 / 102 / 194
    case 1: return run_bwd(bn_act_bwd_fused_kernel<4, 4, 4>, 1, a, st);
    case 2: return run_bwd(bn_act_bwd_fused_kernel<8, 2, 3>, 2, a, st);
    case 3: return run_bwd(bn_act_bwd_fused_kernel<8, 6, 2>, 3, a, st);
    default: return run_bwd(bn_act_bwd_fused_kernel<8, 4, 2>, 0, a, st);
  }
}

extern "C" int esn_act_bwd(const EsnBnBwd* p, void* stream) {
  if (!p || !esn_valid_nhwc(p->x) || !esn_valid_nhwc(p->dy) || !esn_valid_nhwc(p->dx)) return ESN_ERR_BAD_ARG;
  if (p->x.n != p->dy.n || p->x.h != p->dy.h || p->x.w != p->dy.w || p->x.c != p->dy.c || p->dx.c != p->x.c ||
      p->dx.n != p->x.n || p->dx.h != p->x.h || p->dx.w != p->x.w)
    return ESN_ERR_BAD_SHAPE;
  if (p->extra.ptr && !esn_valid_nhwc(p->extra)) return ESN_ERR_BAD_ARG;
  if (p->train_stats || p->act == ESN_ACT_PRELU || p->x.c % 8) return ESN_ERR_UNSUPPORTED;
  if (p->x.dtype != ESN_BF16 || p->dy.dtype != ESN_BF16 || p->dx.dtype != ESN_BF16 ||
      (p->extra.ptr && p->extra.dtype != ESN_BF16) || !v8(p->x.ptr, p->x.c_stride) || !v8(p->dy.ptr, p->dy.c_stride) ||
      !v8(p->dx.ptr, p->dx.c_stride) || (p->extra.ptr && !v8(p->extra.ptr, p->extra.c_stride)))
    return ESN_ERR_UNSUPPORTED;
  BwdArgs a = {};
  a.x = (const __nv_bfloat16*)p->x.ptr;
  a.dy = (const __nv_bfloat16*)p->dy.ptr;
  a.extra = (const __nv_bfloat16*)p->extra.ptr;
  a.dx = (__nv_bfloat16*)p->dx.ptr;
  a.M = (long long)p->x.n * p->x.h * p->x.w;
  a.C = p->x.c;
  a.x_cs = p->x.c_stride;
  a.dy_cs = p->dy.c_stride;
  a.dx_cs = p->dx.c_stride;
  a.extra_cs = p->extra.c_stride;
  a.act = p->act;
  a.scale = p->scale;
  a.shift = p->shift;
  constexpr int U = 4;
  const long long nvec = a.M * (a.C / 8);
  long long ctas = (nvec + (long long)kThreads * U - 1) / ((long long)kThreads * U);
  if (ctas > 148 * 8) ctas = 148 * 8;
  if (ctas < 1) ctas = 1;
  act_bwd_kernel<U><<<(unsigned)ctas, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
