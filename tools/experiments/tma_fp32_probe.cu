// Stand-alone probe (not part of the library): does ONE 4-D TMA box load of an fp32 NCHW image complete its mbarrier
// for a given (box width, start column, start row)?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe
// tma_fp32_probe.cu ; run: ./tma_probe BOXW C0 C1 [W] [H]   (one configuration per process: a fault is sticky).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct alignas(64) Args {
  CUtensorMap tm;
  float* out;
  int* flag;
  int c0, c1, bytes, nfloat;
};

__global__ void probe(const __grid_constant__ Args a) {
  extern __shared__ unsigned char raw[];
  const unsigned rawa = (unsigned)__cvta_generic_to_shared(raw);
  const unsigned base = (rawa + 127u) & ~127u;
  const unsigned bar = base + 32768u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(a.bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(base), "l"(reinterpret_cast<unsigned long long>(&a.tm)), "r"(bar), "r"(a.c0), "r"(a.c1), "r"(0), "r"(0)
        : "memory");
    int ok = 0;
    const long long t0 = clock64();
    while (clock64() - t0 < 200000000LL) {
      unsigned p;
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                   : "=r"(p) : "r"(bar) : "memory");
      if (p) { ok = 1; break; }
    }
    *a.flag = ok;
    const float* s = reinterpret_cast<const float*>(raw + (base - rawa));
    if (ok) for (int i = 0; i < a.nfloat && i < 2048; ++i) a.out[i] = s[i];
  }
}

int main(int argc, char** argv) {
  const int boxw = argc > 1 ? atoi(argv[1]) : 136, c0 = argc > 2 ? atoi(argv[2]) : -1, c1 = argc > 3 ? atoi(argv[3]) : -1;
  const int W = argc > 4 ? atoi(argv[4]) : 44, H = argc > 5 ? atoi(argv[5]) : 24;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode\n"); return 2; }
  std::vector<float> h((size_t)3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *x, *out; int* flag;
  cudaMalloc(&x, h.size() * 4); cudaMalloc(&out, 2048 * 4); cudaMalloc(&flag, 4);
  cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(flag, 0xff, 4); cudaMemset(out, 0, 2048 * 4);
  Args a; memset(&a, 0, sizeof(a));
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, 1};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)boxw, 3, 3, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = ((EncodeTiledFn)fn)(&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("boxw %d c0 %d c1 %d W %d H %d: encode rc %d", boxw, c0, c1, W, H, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  a.out = out; a.flag = flag; a.c0 = c0; a.c1 = c1; a.bytes = boxw * 9 * 4; a.nfloat = boxw * 9;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  probe<<<1, 32, 33000>>>(a);
  const cudaError_t e = cudaDeviceSynchronize();
  int f = -1; float v[4] = {0, 0, 0, 0};
  if (e == cudaSuccess) { cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost); cudaMemcpy(v, out, 16, cudaMemcpyDeviceToHost); }
  // element (c=0, row c1+1, col c0+1) is the first in-bounds element when c0 = c1 = -1
  float probe_v = 0; if (e == cudaSuccess && f == 1) cudaMemcpy(&probe_v, out + boxw + 1, 4, cudaMemcpyDeviceToHost);
  printf("  sync: %s  completed %d  s[0..3] = %g %g %g %g  s[boxw+1] = %g\n", cudaGetErrorString(e), f, v[0], v[1], v[2], v[3], probe_v);
  return 0;
}
