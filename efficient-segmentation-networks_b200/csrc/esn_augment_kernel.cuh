// Device source of esn_augment_u8 (csrc/esn_augment.cu), free of CUDA headers so the CPU test suite can compile it with g++
// behind tests/cuda_cpu_shim.h.  aug_dmul / aug_dsub (unfused double multiply / subtract), aug_rint (float -> int, half to
// even) and aug_floorf come from the including file.
//
// One thread = one pixel of one output crop.  Everything the reference's CityscapesDataSet.__getitem__ does after decoding
// (dataset/cityscapes.py:67-104) happens here in one pass over the OUTPUT: random scale (cv2.resize INTER_LINEAR on the uint8
// image -- OpenCV's 11-bit fixed-point arithmetic, reproduced bit for bit -- and INTER_NEAREST on the label map), mean
// subtraction in fp32, BGR -> RGB, zero / ignore padding up to the crop size, the crop offset, CHW, the mirror.  The resized
// image is never materialised: each output pixel reads its 2 x 2 source neighbourhood.  HBM-bound byte work: 20 bytes written
// per output pixel (3 fp32 planes + int64 label), <= 16 bytes gathered.
#pragma once
#include <stdint.h>

namespace {

struct AugItem {
  const uint8_t* img;     // (h, w, 3) uint8, BGR as cv2.imread returns it
  const uint8_t* label;   // (h, w) uint8
  int32_t h, w;           // decoded size
  int32_t rh, rw;         // size after cv2.resize: cvRound(h * f), cvRound(w * f)
  double scale;           // 1 / f (cv2: scale_x = 1. / inv_scale_x)
  int32_t h_off, w_off;   // crop offset inside the padded resized image
  int32_t flip;           // 1 = mirror columns (image[:, :, ::-1] after the crop)
  int32_t do_scale;       // 0 = the scale step is skipped (scale=False)
};

constexpr int kAugMaxBatch = 24;
struct AugArgs {
  AugItem it[kAugMaxBatch];
  int32_t n, crop_h, crop_w, ignore_label;
  float mean[3];          // in the image's channel order (BGR)
  float* out_img;         // (n, 3, crop_h, crop_w) fp32, RGB planes
  long long* out_label;   // (n, crop_h, crop_w) int64
};

__device__ __forceinline__ void aug_pixel(const AugArgs& a, int b, int y, int x) {
  const AugItem& it = a.it[b];
  const int xs = it.flip ? (a.crop_w - 1 - x) : x;
  const int Y = it.h_off + y, X = it.w_off + xs;
  const long long plane = (long long)a.crop_h * a.crop_w;
  float* op = a.out_img + (long long)b * 3 * plane + (long long)y * a.crop_w + x;
  long long* lp = a.out_label + (long long)b * plane + (long long)y * a.crop_w + x;
  if (Y >= it.rh || X >= it.rw) {          // cv2.copyMakeBorder: 0.0 on the mean-subtracted image, ignore_label on the labels
    op[0] = 0.f;
    op[plane] = 0.f;
    op[2 * plane] = 0.f;
    *lp = a.ignore_label;
    return;
  }
  int v[3], lab;
  if (!it.do_scale) {
    const uint8_t* p = it.img + ((long long)Y * it.w + X) * 3;
    v[0] = p[0];
    v[1] = p[1];
    v[2] = p[2];
    lab = it.label[(long long)Y * it.w + X];
  } else {
    // resize.cpp: fx = (float)((dx + 0.5) * scale_x - 0.5); sx = cvFloor(fx); fx -= sx; the fraction is reset at the left /
    // right border, the ROW indices are clipped instead; weights = saturate_cast<short>(w * 2048)
    float fx = (float)aug_dsub(aug_dmul((double)X + 0.5, it.scale), 0.5);
    int sx = aug_floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= it.w - 1) { fx = 0.f; sx = it.w - 1; }
    const int a0 = aug_rint((1.f - fx) * 2048.f), a1 = aug_rint(fx * 2048.f);
    const int x1 = sx + 1 < it.w ? sx + 1 : it.w - 1;
    float fy = (float)aug_dsub(aug_dmul((double)Y + 0.5, it.scale), 0.5);
    const int sy = aug_floorf(fy);
    fy -= (float)sy;
    const int b0 = aug_rint((1.f - fy) * 2048.f), b1 = aug_rint(fy * 2048.f);
    const int y0 = sy < 0 ? 0 : (sy > it.h - 1 ? it.h - 1 : sy);
    const int y1 = sy + 1 < 0 ? 0 : (sy + 1 > it.h - 1 ? it.h - 1 : sy + 1);
    const uint8_t* r0 = it.img + (long long)y0 * it.w * 3;
    const uint8_t* r1 = it.img + (long long)y1 * it.w * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s0 = r0[sx * 3 + c] * a0 + r0[x1 * 3 + c] * a1;       // HResizeLinear<uchar, int, short>
      const int s1 = r1[sx * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
      v[c] = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;      // VResizeLinear, FixedPtCast<int, uchar, 22>
    }
    // INTER_NEAREST: sx = min(cvFloor(dx * scale_x), w - 1)
    int ly = (int)aug_floord(aug_dmul((double)Y, it.scale)), lx = (int)aug_floord(aug_dmul((double)X, it.scale));
    ly = ly < it.h - 1 ? ly : it.h - 1;
    lx = lx < it.w - 1 ? lx : it.w - 1;
    lab = it.label[(long long)ly * it.w + lx];
  }
  // image = float32(image) - mean (BGR), then [:, :, ::-1] (RGB), then CHW
  op[0] = (float)v[2] - a.mean[2];
  op[plane] = (float)v[1] - a.mean[1];
  op[2 * plane] = (float)v[0] - a.mean[0];
  *lp = lab;
}

__global__ void __launch_bounds__(256) augment_u8_kernel(const AugArgs a) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)a.crop_h * a.crop_w) return;
  aug_pixel(a, b, (int)(i / a.crop_w), (int)(i % a.crop_w));
}

}  // namespace
