// TEST INFRASTRUCTURE ONLY: just enough of the CUDA execution model to run a simple kernel's SOURCE on the CPU.
// One CTA at a time; every CUDA thread of the CTA is a real pthread, __syncthreads() is a pthread barrier, __shared__
// variables are function-local statics (one instance, visible to all threads of the running CTA).  Good for checking
// indexing, tails, alignment branches and barrier placement (a misplaced barrier deadlocks or races here as well);
// says nothing about memory coalescing or speed.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <vector>

struct shim_dim3 { unsigned x, y, z; };
static thread_local shim_dim3 threadIdx;
static thread_local shim_dim3 blockIdx;
static shim_dim3 blockDim, gridDim;
static pthread_barrier_t shim_barrier;

#define __global__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __forceinline__ inline
#define __device__

static inline void __syncthreads() { pthread_barrier_wait(&shim_barrier); }
struct uint4 { uint32_t x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
template <typename T> static inline T __ldg(const T* p) { return *p; }

template <typename F>
static void shim_launch(unsigned grid, unsigned block, F body) {
  gridDim = {grid, 1, 1};
  blockDim = {block, 1, 1};
  struct Arg { F* body; unsigned tid, bid; };
  for (unsigned b = 0; b < grid; ++b) {
    pthread_barrier_init(&shim_barrier, nullptr, block);
    std::vector<pthread_t> th(block);
    std::vector<Arg> args(block);
    for (unsigned t = 0; t < block; ++t) {
      args[t] = Arg{&body, t, b};
      pthread_create(&th[t], nullptr, [](void* p) -> void* {
        Arg* a = static_cast<Arg*>(p);
        threadIdx = {a->tid, 0, 0};
        blockIdx = {a->bid, 0, 0};
        (*a->body)();
        return nullptr;
      }, &args[t]);
    }
    for (unsigned t = 0; t < block; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&shim_barrier);
  }
}

// element access helpers of esn_common.cuh (float only on the CPU; the bf16 specialisations differ in the conversion only)
template <typename T> static inline float ld1(const T* p) { return static_cast<float>(*p); }
template <typename T> static inline void st1(T* p, float v) { *p = static_cast<T>(v); }
