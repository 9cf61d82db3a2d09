// Per-pixel gate with a per-image bias: y = g * x + b, g one channel (N,1,H,W), x (N,C,H,W), b (N,C,1,1).
// LEDNet's attention pyramid closes with it (APNModule.forward, model/LEDNet.py:279-281 of the reference:
// `x = torch.mul(x, mid); x = x + b1`, where b1 is the global-pooling branch upsampled from 1x1 with
// align_corners=True, i.e. a constant per image and class).  Elementwise, HBM-bound: C reads + C writes (+1) per pixel
// on the 1/8-resolution class scores; one launch instead of broadcast-multiply, upsample and add.
#include "esn_common.cuh"

#include "esn_gate_kernel.cuh"

extern "C" int esn_gate_bcast(const EsnTensor* g, const EsnTensor* x, const EsnTensor* b, const EsnTensor* y, void* stream) {
  if (!g || !x || !y || !esn_valid_nhwc(*g) || !esn_valid_nhwc(*x) || !esn_valid_nhwc(*y)) return ESN_ERR_BAD_ARG;
  if (b && b->ptr && !esn_valid_nhwc(*b)) return ESN_ERR_BAD_ARG;
  const bool has_b = b && b->ptr;
  if (g->c != 1 || g->n != x->n || g->h != x->h || g->w != x->w) return ESN_ERR_BAD_SHAPE;
  if (y->n != x->n || y->h != x->h || y->w != x->w || y->c != x->c) return ESN_ERR_BAD_SHAPE;
  if (has_b && (b->n != x->n || b->c != x->c || b->h != 1 || b->w != 1)) return ESN_ERR_BAD_SHAPE;
  // the gate may be fp32 next to bf16 scores (fp32 pyramid, bf16 class scores); everything else shares x's type
  if (y->dtype != x->dtype || (has_b && b->dtype != x->dtype)) return ESN_ERR_UNSUPPORTED;
  if (g->dtype != x->dtype && !(g->dtype == ESN_F32 && x->dtype == ESN_BF16)) return ESN_ERR_UNSUPPORTED;
  const long long npix = (long long)x->n * x->h * x->w;
  const long long total = npix * x->c;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int hw = x->h * x->w;
  if (x->dtype == ESN_F32)
    gate_bcast_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)g->ptr, g->c_stride, (const float*)x->ptr, x->c_stride,
                                                            has_b ? (const float*)b->ptr : nullptr, has_b ? b->c_stride : 0,
                                                            (float*)y->ptr, y->c_stride, npix, hw, x->c);
  else if (g->dtype == ESN_F32)
    gate_bcast_kernel<__nv_bfloat16, float><<<(unsigned)grid, 256, 0, st>>>(
        (const float*)g->ptr, g->c_stride, (const __nv_bfloat16*)x->ptr, x->c_stride,
        has_b ? (const __nv_bfloat16*)b->ptr : nullptr, has_b ? b->c_stride : 0, (__nv_bfloat16*)y->ptr, y->c_stride, npix, hw, x->c);
  else
    gate_bcast_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>(
        (const __nv_bfloat16*)g->ptr, g->c_stride, (const __nv_bfloat16*)x->ptr, x->c_stride,
        has_b ? (const __nv_bfloat16*)b->ptr : nullptr, has_b ? b->c_stride : 0, (__nv_bfloat16*)y->ptr, y->c_stride, npix, hw, x->c);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
