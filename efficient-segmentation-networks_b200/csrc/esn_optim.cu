// Optimizer step of the training iteration (train.py:355 `optimizer.step()` on the torch.optim.Adam that train.py:212-215
// builds): Adam over EVERY parameter tensor of a param group in ONE launch.
//
// A lightweight segmentation net is ~400 parameter tensors of 16 ... 200 k elements (DABNet: 0.75 M parameters).  A
// multi-tensor-apply optimizer packs tensor addresses into the 4 KB kernel-argument space -- 8 launches for DABNet, each a few
// dozen CTAs that loop serially over up to 64 k elements: 0.38 ms of a 7.7 ms step, at the very end of the critical path
// (after the gradient all-reduce).  Here the address table lives in device memory (written once per set of gradient
// buffers; inside a replayed CUDA graph the addresses never change), a CTA owns one 2048-element chunk of one tensor
// (blocks[] maps CTA -> (tensor, chunk)), every thread has its eight elements in flight at once, and the step counter is
// advanced by the last CTA to finish -- so the whole update is one launch with no serial loop.
#include "esn_common.cuh"

namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 2048;      // elements per CTA (ESN_ADAM_CHUNK in esn.h)

// b = (beta1, 1 - beta1, beta2, 1 - beta2): the complements are rounded from the double values, not formed as 1 - float(beta)
// (1 - 0.999f is off by 1.3e-5 relative, which the second moment would carry)
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const float4 b, float wd, float step_size,
                                      float bc2_sqrt, float eps) {
  if (wd != 0.f) g = fmaf(p, wd, g);                               // L2 penalty folded into the gradient (torch.optim.Adam)
  m = fmaf(b.x, m, b.y * g);
  v = fmaf(b.z, v, b.w * g * g);
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p -= step_size * m / denom;
}

__global__ void __launch_bounds__(kAdamThreads) adam_table_kernel(const EsnAdamTensor* __restrict__ tab,
                                                                  const int2* __restrict__ blocks,
                                                                  const float* __restrict__ lr, float* step,
                                                                  unsigned int* done, const double beta1, const double beta2,
                                                                  const float eps, const float wd) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const double t = (double)*reinterpret_cast<volatile float*>(step) + 1.0;  // this update's step number
    s_c[0] = (float)((double)__ldg(lr) / (1.0 - pow(beta1, t)));             // lr / bias_correction1
    s_c[1] = (float)sqrt(1.0 - pow(beta2, t));                               // sqrt(bias_correction2)
  }
  const float4 bt = make_float4((float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2));
  const int2 b = __ldg(blocks + blockIdx.x);
  const EsnAdamTensor T = tab[b.x];
  __syncthreads();
  const float step_size = s_c[0], bc2_sqrt = s_c[1];
  const long long base = (long long)b.y * kAdamChunk;
  const long long left = T.n - base;
  float* __restrict__ p = T.p + base;
  const float* __restrict__ g = T.g + base;
  float* __restrict__ m = T.m + base;
  float* __restrict__ v = T.v + base;
  const bool vec = left >= kAdamChunk && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                           reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    // whole chunk, 16-byte aligned: two float4 per array per thread, all eight loads issued before the first use
    float4 P[2], G[2], M[2], V[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * kAdamThreads;
      P[k] = reinterpret_cast<const float4*>(p)[i];
      G[k] = __ldg(reinterpret_cast<const float4*>(g) + i);
      M[k] = reinterpret_cast<const float4*>(m)[i];
      V[k] = reinterpret_cast<const float4*>(v)[i];
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * kAdamThreads;
      adam1(P[k].x, G[k].x, M[k].x, V[k].x, bt, wd, step_size, bc2_sqrt, eps);
      adam1(P[k].y, G[k].y, M[k].y, V[k].y, bt, wd, step_size, bc2_sqrt, eps);
      adam1(P[k].z, G[k].z, M[k].z, V[k].z, bt, wd, step_size, bc2_sqrt, eps);
      adam1(P[k].w, G[k].w, M[k].w, V[k].w, bt, wd, step_size, bc2_sqrt, eps);
      reinterpret_cast<float4*>(p)[i] = P[k];
      reinterpret_cast<float4*>(m)[i] = M[k];
      reinterpret_cast<float4*>(v)[i] = V[k];
    }
  } else {
    const int cnt = (int)(left < kAdamChunk ? left : kAdamChunk);
    float P[8], G[8], M[8], V[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = threadIdx.x + k * kAdamThreads;
      if (i < cnt) { P[k] = p[i]; G[k] = __ldg(g + i); M[k] = m[i]; V[k] = v[i]; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = threadIdx.x + k * kAdamThreads;
      if (i < cnt) {
        adam1(P[k], G[k], M[k], V[k], bt, wd, step_size, bc2_sqrt, eps);
        p[i] = P[k]; m[i] = M[k]; v[i] = V[k];
      }
    }
  }
  // every CTA read *step before it arrives here; the last one to arrive advances it and re-arms the counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *reinterpret_cast<volatile float*>(step) = *reinterpret_cast<volatile float*>(step) + 1.f;
      *done = 0u;
      __threadfence();
    }
  }
}

}  // namespace

extern "C" int32_t esn_adam_chunk(void) { return kAdamChunk; }

extern "C" int esn_adam_step(const EsnAdamTensor* table, const int32_t* blocks, int32_t n_blocks, const float* lr, float* step,
                             uint32_t* done, double beta1, double beta2, double eps, double weight_decay, void* stream) {
  if (!table || !blocks || !lr || !step || !done) return ESN_ERR_BAD_ARG;
  if (n_blocks < 0) return ESN_ERR_BAD_SHAPE;
  if (n_blocks == 0) return ESN_OK;
  if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && weight_decay >= 0.0)) return ESN_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(table) % 8) || (reinterpret_cast<uintptr_t>(blocks) % 8)) return ESN_ERR_ALIGN;
  adam_table_kernel<<<(unsigned)n_blocks, kAdamThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, reinterpret_cast<const int2*>(blocks), lr, step, done, beta1, beta2, (float)eps, (float)weight_decay);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
