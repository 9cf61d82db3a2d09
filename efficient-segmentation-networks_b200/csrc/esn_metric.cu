// Confusion matrix of a batch of predicted masks against the labels, on the device.
// Replaces the reference's per-pixel Python loop ConfusionMatrix.generateM (utils/metric/metric.py:68-76, called through
// get_iou from test.py:90 / train.py:404): M[gt, pred] += 1 for every pixel whose label is a class (gt < nclass; the
// ignore label 255 and anything else >= nclass is skipped).  Integer work, HBM-bound: 1 byte of mask + 1 or 8 bytes of
// label per pixel.  Per-warp shared-memory histograms, equal bins of a warp merged with __match_any_sync (label maps
// are piecewise constant, so most of a warp hits one bin), one 64-bit global atomic per non-zero bin per CTA.
#include "esn_common.cuh"

namespace {

constexpr int kCmThreads = 256;
constexpr int kCmWarps = kCmThreads / 32;

template <typename TG>
__global__ void __launch_bounds__(kCmThreads) confusion_kernel(const uint8_t* __restrict__ pred, const TG* __restrict__ gt,
                                                               const long long n, const int nclass,
                                                               unsigned long long* __restrict__ M) {
  extern __shared__ unsigned int cm_hist[];        // [kCmWarps][nclass * nclass]
  const int bins = nclass * nclass;
  for (int i = threadIdx.x; i < kCmWarps * bins; i += kCmThreads) cm_hist[i] = 0u;
  __syncthreads();
  unsigned int* mine = cm_hist + (threadIdx.x >> 5) * bins;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * kCmThreads;
  // all lanes of a warp run the same number of iterations (the tail is predicated), so the warp-wide match is safe
  const long long iters = (n + stride - 1) / stride;
  long long i = (long long)blockIdx.x * kCmThreads + threadIdx.x;
  for (long long it = 0; it < iters; ++it, i += stride) {
    int bin = -1;
    if (i < n) {
      const long long g = (long long)gt[i];
      const int p = pred[i];
      if (g >= 0 && g < nclass && p < nclass) bin = (int)g * nclass + p;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(mine + bin, (unsigned)__popc(peers));
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += kCmThreads) {
    unsigned long long s = 0;
#pragma unroll
    for (int w = 0; w < kCmWarps; ++w) s += cm_hist[w * bins + b];
    if (s) atomicAdd(M + b, s);
  }
}

}  // namespace

extern "C" int esn_confusion_matrix(const uint8_t* pred, const void* gt, int32_t gt_is_int64, int64_t n_pixels, int32_t nclass,
                                    uint64_t* M, void* stream) {
  if (!pred || !gt || !M || n_pixels < 0) return ESN_ERR_BAD_ARG;
  if (nclass < 1 || nclass > 32) return ESN_ERR_UNSUPPORTED;
  if (gt_is_int64 && ((uintptr_t)gt & 7)) return ESN_ERR_ALIGN;
  if ((uintptr_t)M & 7) return ESN_ERR_ALIGN;
  if (n_pixels == 0) return ESN_OK;
  const size_t smem = (size_t)kCmWarps * nclass * nclass * sizeof(unsigned int);     // <= 32 KB
  long long grid = (n_pixels + kCmThreads - 1) / kCmThreads;
  if (grid > 148 * 8) grid = 148 * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long* m = reinterpret_cast<unsigned long long*>(M);
  if (gt_is_int64)
    confusion_kernel<long long><<<(unsigned)grid, kCmThreads, smem, st>>>(pred, reinterpret_cast<const long long*>(gt), n_pixels, nclass, m);
  else
    confusion_kernel<uint8_t><<<(unsigned)grid, kCmThreads, smem, st>>>(pred, reinterpret_cast<const uint8_t*>(gt), n_pixels, nclass, m);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
