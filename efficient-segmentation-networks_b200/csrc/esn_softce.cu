// Label-smoothing cross-entropy: CrossEntropyLoss2dLabelSmooth, utils/losses/loss.py:56-86 of the reference
// (`--use_label_smoothing` in train.py): soft targets (1 - eps) * onehot + eps / C fed to nn.CrossEntropyLoss(weight,
// reduction='mean'), i.e. the mean over ALL pixels of -sum_c w_c t_c log softmax(x)_c.  Forward sum and gradient in one
// pass over the fp32 NCHW logits each; HBM-bound (C floats read, + C written for the gradient, per pixel).
#include "esn_common.cuh"

#include "esn_softce_kernel.cuh"

extern "C" int esn_soft_ce(const EsnSoftCE* p, void* stream) {
  if (!p || !p->logits.ptr || !p->target || !p->sum) return ESN_ERR_BAD_ARG;
  const EsnTensor& x = p->logits;
  if (x.dtype != ESN_F32 || x.layout != ESN_NCHW) return ESN_ERR_UNSUPPORTED;
  if (x.n <= 0 || x.c <= 0 || x.h <= 0 || x.w <= 0) return ESN_ERR_BAD_SHAPE;
  if (!(p->epsilon >= 0.f && p->epsilon <= 1.f)) return ESN_ERR_BAD_ARG;
  float* dx = nullptr;
  if (p->dlogits.ptr) {
    const EsnTensor& d = p->dlogits;
    if (d.dtype != ESN_F32 || d.layout != ESN_NCHW) return ESN_ERR_UNSUPPORTED;
    if (d.n != x.n || d.c != x.c || d.h != x.h || d.w != x.w) return ESN_ERR_BAD_SHAPE;
    dx = static_cast<float*>(d.ptr);
  }
  const long long hw = (long long)x.h * x.w, npix = hw * x.n;
  long long grid = (npix + kSoftCeThreads - 1) / kSoftCeThreads;
  if (grid > 148 * 8) grid = 148 * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  soft_ce_kernel<<<(unsigned)grid, kSoftCeThreads, 0, st>>>(static_cast<const float*>(x.ptr),
                                                           reinterpret_cast<const long long*>(p->target), p->weight, p->sum, dx,
                                                           p->gout, x.c, hw, npix, p->epsilon, p->ignore_label, p->grad_scale);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
