// OHEM threshold selection on the device (ProbOhemCrossEntropy2d, utils/losses/loss.py:199-203 of the reference).
// The reference sorts the per-pixel probabilities of the labelled class on the GPU and reads the min_kept-th smallest back
// to the host (argsort + Python comparisons: two host synchronisations per training step).  Here the k-th smallest value
// is found exactly, without sorting and without leaving the device: the values are non-negative floats, so the numeric
// order is the order of their bit patterns and a radix select over (12, 12, 8)-bit digits needs three histogram passes over
// the 4-byte-per-pixel array (HBM-bound integer work: 3 reads of n*4 bytes; n = 4 M pixels for 8 x 512 x 1024) and three
// single-CTA scans of <= 4096 bins.
#include "esn_common.cuh"

namespace {

constexpr int kBins = 4096;
struct OhemState {              // lives at the start of the workspace (zeroed by the caller)
  unsigned int hist[kBins];
  unsigned int prefix;          // bits of the k-th value fixed so far (high bits)
  unsigned int done;            // 1 = no selection needed (min_kept > num_valid)
  unsigned long long k;         // rank still to find inside the current prefix (1-based)
};

// pass p: digit = bits [shift, shift + width) of the values whose higher bits equal state->prefix
__global__ void __launch_bounds__(256) ohem_hist_kernel(const float* __restrict__ prob, long long n, OhemState* st, int shift,
                                                        int width, int pass) {
  __shared__ unsigned int sh[kBins];
  if (st->done) return;
  const int bins = 1 << width;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned int prefix = st->prefix;
  const unsigned int hi_mask = pass == 0 ? 0u : ~((1u << (shift + width)) - 1u);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned int u = __float_as_uint(__ldg(prob + i));
    if ((u & hi_mask) == prefix) atomicAdd(&sh[(u >> shift) & (bins - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (sh[i]) atomicAdd(&st->hist[i], sh[i]);
}

// one CTA: find the bin in which the cumulative count reaches k, fix its bits, clear the histogram for the next pass;
// the last pass writes the threshold
__global__ void __launch_bounds__(1024) ohem_scan_kernel(OhemState* st, int shift, int width, int last, float thresh, float* out) {
  __shared__ unsigned long long part[1024];
  __shared__ int found_bin;
  __shared__ unsigned long long found_before;
  if (st->done) {
    if (last && threadIdx.x == 0) *out = INFINITY;
    return;
  }
  const int bins = 1 << width;
  const int per = (bins + 1023) / 1024;
  const int lo = threadIdx.x * per;
  unsigned long long s = 0;
  for (int i = lo; i < lo + per && i < bins; ++i) s += st->hist[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; ++t) {
      const unsigned long long v = part[t];
      part[t] = run;          // exclusive prefix of the thread's chunk
      run += v;
    }
    found_bin = -1;
  }
  __syncthreads();
  const unsigned long long k = st->k;
  unsigned long long run = part[threadIdx.x];
  for (int i = lo; i < lo + per && i < bins; ++i) {
    const unsigned long long c = st->hist[i];
    if (run < k && k <= run + c) {      // exactly one bin satisfies this
      found_bin = i;
      found_before = run;
    }
    run += c;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) st->hist[i] = 0;
  if (threadIdx.x == 0) {
    const unsigned int prefix = st->prefix | ((unsigned int)found_bin << shift);
    st->prefix = prefix;
    st->k = k - found_before;
    if (last) *out = fmaxf(thresh, __uint_as_float(prefix));
  }
}

__global__ void ohem_init_kernel(OhemState* st, long long n, long long min_kept, const float* num_valid) {
  // loss.py:199: `if self.min_kept > num_valid` -> nothing is filtered; loss.py:207: index[min(len(index), min_kept) - 1]
  // (loss.py:201: with min_kept <= 0 the reference applies no mask at all -- kept_mask sits inside `if self.min_kept > 0`)
  const bool skip = (double)min_kept > (double)*num_valid || *num_valid <= 0.f || min_kept <= 0;
  st->done = skip ? 1u : 0u;
  st->prefix = 0u;
  st->k = (unsigned long long)(min_kept < n ? min_kept : n);
}

}  // namespace

extern "C" int64_t esn_ohem_workspace_bytes(void) { return (int64_t)sizeof(OhemState); }

extern "C" int esn_ohem_threshold(const float* prob, int64_t n, int64_t min_kept, float thresh, const float* num_valid, float* out,
                                  void* workspace, void* stream) {
  if (!prob || !num_valid || !out || !workspace || n <= 0) return ESN_ERR_BAD_ARG;
  if (((uintptr_t)workspace % 8) || ((uintptr_t)prob % 4)) return ESN_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  OhemState* st = reinterpret_cast<OhemState*>(workspace);
  ohem_init_kernel<<<1, 1, 0, s>>>(st, (long long)n, (long long)min_kept, num_valid);
  ESN_CHECK_LAUNCH();
  int grid = esn_cdiv(n, 256 * 16);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  const int shifts[3] = {20, 8, 0}, widths[3] = {12, 12, 8};
  for (int p = 0; p < 3; ++p) {
    ohem_hist_kernel<<<grid, 256, 0, s>>>(prob, (long long)n, st, shifts[p], widths[p], p);
    ESN_CHECK_LAUNCH();
    ohem_scan_kernel<<<1, 1024, 0, s>>>(st, shifts[p], widths[p], p == 2, thresh, out);
    ESN_CHECK_LAUNCH();
  }
  return ESN_OK;
}
