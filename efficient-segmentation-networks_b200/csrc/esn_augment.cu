// Training-time augmentation of the reference's CityscapesDataSet on the device: see esn_augment_kernel.cuh.
#include "esn_common.cuh"

namespace {
__device__ __forceinline__ double aug_dmul(double a, double b) { return __dmul_rn(a, b); }     // never contracted into an FMA:
__device__ __forceinline__ double aug_dsub(double a, double b) { return __dsub_rn(a, b); }     // cv2 rounds each step
__device__ __forceinline__ int aug_rint(float v) { return __float2int_rn(v); }
__device__ __forceinline__ int aug_floorf(float v) { return __float2int_rd(v); }
__device__ __forceinline__ double aug_floord(double v) { return floor(v); }
}  // namespace

#include "esn_augment_kernel.cuh"

static_assert(sizeof(EsnAugItem) == sizeof(AugItem), "EsnAugItem (include/esn.h) and the kernel's AugItem must match");

extern "C" int32_t esn_augment_max_batch(void) { return kAugMaxBatch; }

extern "C" int esn_augment_u8(const EsnAugItem* items, int32_t n, int32_t crop_h, int32_t crop_w, const float* mean3,
                              int32_t ignore_label, float* out_img, int64_t* out_label, void* stream) {
  if (!items || !mean3 || !out_img || !out_label || n < 0 || crop_h <= 0 || crop_w <= 0) return ESN_ERR_BAD_ARG;
  if (n > kAugMaxBatch) return ESN_ERR_UNSUPPORTED;
  if (n == 0) return ESN_OK;
  AugArgs a;
  for (int i = 0; i < n; ++i) {
    const EsnAugItem& s = items[i];
    if (!s.img || !s.label || s.h <= 0 || s.w <= 0 || s.rh <= 0 || s.rw <= 0 || !(s.scale > 0.0)) return ESN_ERR_BAD_ARG;
    // the crop window must lie inside the padded resized image (the reference draws the offsets that way)
    const int ph = s.rh > crop_h ? s.rh : crop_h, pw = s.rw > crop_w ? s.rw : crop_w;
    if (s.h_off < 0 || s.w_off < 0 || s.h_off + crop_h > ph || s.w_off + crop_w > pw) return ESN_ERR_BAD_SHAPE;
    AugItem& d = a.it[i];
    d.img = s.img; d.label = s.label; d.h = s.h; d.w = s.w; d.rh = s.rh; d.rw = s.rw; d.scale = s.scale;
    d.h_off = s.h_off; d.w_off = s.w_off; d.flip = s.flip; d.do_scale = s.do_scale;
  }
  a.n = n; a.crop_h = crop_h; a.crop_w = crop_w; a.ignore_label = ignore_label;
  a.mean[0] = mean3[0]; a.mean[1] = mean3[1]; a.mean[2] = mean3[2];
  a.out_img = out_img;
  a.out_label = reinterpret_cast<long long*>(out_label);
  dim3 grid((unsigned)esn_cdiv((long long)crop_h * crop_w, 256), (unsigned)n);
  augment_u8_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ESN_CHECK_LAUNCH();
  return ESN_OK;
}
