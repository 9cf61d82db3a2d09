"""CPU restatement of the reference's image pre-processing tail (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The dataset classes of the reference finish every sample the same way (dataset/cityscapes.py:74-78 train, :164-170
val, :208-214 test; the CamVid twins are identical): the decoded uint8 BGR image becomes float32, the per-channel mean
(BGR order; float32 in dataset/inform/cityscapes_inform.pkl) is subtracted, the channel axis is reversed to RGB and
moved to the front.  Pinned on tests/golden/pipeline.npz (tools/make_golden_pipeline.py: the unmodified
CityscapesValDataSet / CityscapesTestDataSet run on PNG files written by the generator).
"""
import numpy as np

CITYSCAPES_MEAN_BGR = np.array([72.3924, 82.90902, 73.158325], dtype=np.float32)   # cityscapes_inform.pkl['mean']


def image_to_input(image_bgr_u8, mean_bgr=CITYSCAPES_MEAN_BGR):
    """(H,W,3) uint8 BGR -> (3,H,W) float32 RGB, mean-subtracted (float32 arithmetic, as numpy does for an fp32 mean)."""
    x = image_bgr_u8.astype(np.float32) - np.asarray(mean_bgr, dtype=np.float32).reshape(1, 1, 3)
    return np.ascontiguousarray(x[:, :, ::-1].transpose(2, 0, 1))


def batch_to_input(images_bgr_u8, mean_bgr=CITYSCAPES_MEAN_BGR):
    """(N,H,W,3) uint8 -> (N,3,H,W) float32: what the DataLoader's default collate stacks."""
    return np.stack([image_to_input(im, mean_bgr) for im in images_bgr_u8])


# --------------------------------------------------------------------------- training-time augmentation
# CityscapesDataSet.__getitem__ (dataset/cityscapes.py:58-106): random scale (cv2.resize, INTER_LINEAR image / INTER_NEAREST
# label), mean subtraction, BGR -> RGB, zero / ignore padding up to the crop size, random crop, CHW, random mirror.
# cv2 is not in /root/reference; its resize arithmetic (opencv imgproc/resize.cpp, fixed point for 8-bit images) is restated
# below and pinned bit-exactly against cv2 4.13 itself in tests/test_pipeline_cpu.py, and against the reference class on PNG
# files in tests/golden/augment.npz (tools/make_golden_augment.py).
SCALES = (0.75, 1.0, 1.25, 1.5, 1.75, 2.0)          # cityscapes.py:68


def _cv_round(v):
    return int(np.rint(v))                           # cvRound: half to even


def resized_size(h, w, f):
    """dsize of cv2.resize(src, None, fx=f, fy=f): saturate_cast<int>(size * f)."""
    return _cv_round(h * f), _cv_round(w * f)


def _linear_coeffs(dn, sn, scale, clamp):
    """Source index and the two 11-bit fixed-point weights per destination index (resize.cpp: fx = (float)((dx + 0.5) *
    scale - 0.5) in double then float; weights = saturate_cast<short>(w * 2048)).  Horizontally the fraction is reset at the
    borders (`clamp`); vertically the row INDICES are clipped instead and the weights stay."""
    ofs = np.zeros(dn, np.int64)
    a = np.zeros((dn, 2), np.int32)
    for d in range(dn):
        fx = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(fx))
        fx = np.float32(fx - np.float32(s))
        if clamp:
            if s < 0:
                fx, s = np.float32(0), 0
            if s >= sn - 1:
                fx, s = np.float32(0), sn - 1
        ofs[d] = s
        a[d, 0] = int(np.rint(np.float32(np.float32(1.0) - fx) * np.float32(2048)))
        a[d, 1] = int(np.rint(fx * np.float32(2048)))
    return ofs, a


def resize_linear_u8(src, f):
    """cv2.resize(src, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR) for uint8 (H,W) or (H,W,C), bit-exact."""
    h, w = src.shape[:2]
    dh, dw = resized_size(h, w, f)
    scale = 1.0 / f
    xo, xa = _linear_coeffs(dw, w, scale, True)
    yo, ya = _linear_coeffs(dh, h, scale, False)
    s = src.astype(np.int32)
    x1 = np.minimum(xo + 1, w - 1)
    tail = (None,) * (s.ndim - 2)
    hbuf = s[:, xo] * xa[:, 0][(None, slice(None)) + tail] + s[:, x1] * xa[:, 1][(None, slice(None)) + tail]
    s0, s1 = hbuf[np.clip(yo, 0, h - 1)], hbuf[np.clip(yo + 1, 0, h - 1)]
    b0 = ya[:, 0][(slice(None), None) + tail]
    b1 = ya[:, 1][(slice(None), None) + tail]
    return ((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def resize_nearest_u8(src, f):
    """cv2.resize(src, None, fx=f, fy=f, interpolation=cv2.INTER_NEAREST): sx = min(floor(dx / f), w - 1)."""
    h, w = src.shape[:2]
    dh, dw = resized_size(h, w, f)
    scale = 1.0 / f
    sy = np.minimum(np.floor(np.arange(dh) * scale).astype(np.int64), h - 1)
    sx = np.minimum(np.floor(np.arange(dw) * scale).astype(np.int64), w - 1)
    return src[sy][:, sx]


def train_item(image_bgr_u8, label_u8, f_scale, h_off, w_off, flip, crop_hw, mean_bgr=CITYSCAPES_MEAN_BGR, ignore_label=255):
    """One sample of CityscapesDataSet with the random draws given: -> ((3,crop_h,crop_w) float32 RGB, (crop_h,crop_w) float32).
    f_scale None = no scaling (scale=False); flip in {+1, -1} (cityscapes.py:100)."""
    crop_h, crop_w = crop_hw
    image, label = image_bgr_u8, label_u8
    if f_scale is not None:
        image, label = resize_linear_u8(image, f_scale), resize_nearest_u8(label, f_scale)
    image = image.astype(np.float32) - np.asarray(mean_bgr, dtype=np.float32).reshape(1, 1, 3)
    image = image[:, :, ::-1]
    img_h, img_w = label.shape
    pad_h, pad_w = max(crop_h - img_h, 0), max(crop_w - img_w, 0)
    if pad_h > 0 or pad_w > 0:
        image = np.pad(image, ((0, pad_h), (0, pad_w), (0, 0)), constant_values=0.0)
        label = np.pad(label, ((0, pad_h), (0, pad_w)), constant_values=ignore_label)
    image = image[h_off:h_off + crop_h, w_off:w_off + crop_w].transpose(2, 0, 1)
    label = label[h_off:h_off + crop_h, w_off:w_off + crop_w].astype(np.float32)
    return np.ascontiguousarray(image[:, :, ::flip]), np.ascontiguousarray(label[:, ::flip])


def draw_train_params(img_hw, crop_hw, scale=True, mirror=True):
    """The reference's random draws in its own order (cityscapes.py:69, 92-93, 100): random.randint(0, 5),
    random.randint(0, H' - crop_h), random.randint(0, W' - crop_w), np.random.choice(2) -- H', W' the padded resized size."""
    import random
    f = SCALES[random.randint(0, 5)] if scale else None
    h, w = resized_size(img_hw[0], img_hw[1], f) if f is not None else img_hw
    h, w = max(h, crop_hw[0]), max(w, crop_hw[1])
    h_off = random.randint(0, h - crop_hw[0])
    w_off = random.randint(0, w - crop_hw[1])
    flip = int(np.random.choice(2) * 2 - 1) if mirror else 1
    return f, h_off, w_off, flip
