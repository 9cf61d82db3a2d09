"""Training-time augmentation of the reference's CityscapesDataSet (dataset/cityscapes.py:27-106) on the device.

The reference decodes an image on a DataLoader worker and then, still on the CPU, rescales it with cv2 by a random factor,
subtracts the mean, pads, crops and mirrors it; the finished fp32 crop (12 bytes per pixel) is what crosses PCIe.  Here the
decoded uint8 image and label map cross PCIe (4 bytes per pixel of the ORIGINAL image) and ONE kernel launch per batch
(esn_augment_u8) produces the fp32 NCHW crops and int64 label crops, bit-identical to the reference class for the same
random draws -- cv2's fixed-point INTER_LINEAR arithmetic included.

    aug = CityscapesTrainAugment(crop_size=(512, 1024), mean=mean_bgr)          # same arguments as the reference class
    images, labels = aug(list_of_uint8_HxWx3_cuda_tensors, list_of_uint8_HxW_cuda_tensors)

The random draws are made on the host with Python's `random` and `numpy.random`, in the reference's order
(cityscapes.py:69, 92-93, 100), so seeding both reproduces the reference's samples.
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _lib as L
from . import ops

SCALES = (0.75, 1.0, 1.25, 1.5, 1.75, 2.0)      # cityscapes.py:68


def _cv_round(v):
    return int(np.rint(v))                       # cvRound (dsize of cv2.resize): half to even


def draw_params(img_hw, crop_hw, scale=True, mirror=True):
    """(f_scale or None, h_off, w_off, flip in {+1, -1}) drawn as CityscapesDataSet.__getitem__ draws them."""
    f = SCALES[random.randint(0, 5)] if scale else None
    h, w = (_cv_round(img_hw[0] * f), _cv_round(img_hw[1] * f)) if f is not None else img_hw
    ph, pw = max(h, crop_hw[0]), max(w, crop_hw[1])
    h_off = random.randint(0, ph - crop_hw[0])
    w_off = random.randint(0, pw - crop_hw[1])
    flip = int(np.random.choice(2) * 2 - 1) if mirror else 1
    return f, h_off, w_off, flip


class CityscapesTrainAugment:
    def __init__(self, crop_size=(512, 1024), mean=(128, 128, 128), scale=True, mirror=True, ignore_label=255):
        self.crop_h, self.crop_w = crop_size
        self.mean = [float(v) for v in np.asarray(mean, dtype=np.float32)]      # BGR, the pickle's fp32 values
        self.scale, self.is_mirror, self.ignore_label = scale, mirror, ignore_label

    def __call__(self, images, labels, params=None):
        """images: uint8 CUDA tensors (H, W, 3) in cv2's BGR order (sizes may differ), labels: uint8 (H, W); params: optional
        list of (f_scale, h_off, w_off, flip) per image, drawn here when omitted.  Returns (fp32 (N, 3, crop_h, crop_w),
        int64 (N, crop_h, crop_w))."""
        n = len(images)
        if n != len(labels):
            raise ValueError("one label map per image")
        if n == 0:
            raise ValueError("empty batch")
        dev = images[0].device
        ops.require_cuda(images[0], "CityscapesTrainAugment")
        out = torch.empty((n, 3, self.crop_h, self.crop_w), dtype=torch.float32, device=dev)
        lab = torch.empty((n, self.crop_h, self.crop_w), dtype=torch.int64, device=dev)
        if params is None:
            params = [draw_params(tuple(im.shape[:2]), (self.crop_h, self.crop_w), self.scale, self.is_mirror) for im in images]
        mean = (C.c_float * 3)(*self.mean)
        step = int(L.lib.esn_augment_max_batch())
        keep = []
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            items = (L.EsnAugItem * (hi - lo))()
            for k in range(lo, hi):
                im, lb = images[k], labels[k]
                if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or lb.dtype != torch.uint8 or lb.shape != im.shape[:2]:
                    raise TypeError("expected a uint8 (H, W, 3) image and a uint8 (H, W) label map")
                ops.require_cuda(lb, "CityscapesTrainAugment")
                im, lb = im.contiguous(), lb.contiguous()
                keep += [im, lb]
                f, h_off, w_off, flip = params[k]
                h, w = im.shape[:2]
                it = items[k - lo]
                it.img, it.label, it.h, it.w = im.data_ptr(), lb.data_ptr(), h, w
                it.rh, it.rw = (_cv_round(h * f), _cv_round(w * f)) if f is not None else (h, w)
                it.scale = 1.0 / f if f is not None else 1.0
                it.h_off, it.w_off, it.flip, it.do_scale = h_off, w_off, int(flip < 0), int(f is not None)
            ops._call(L.lib.esn_augment_u8, "esn_augment_u8",
                      (items, hi - lo, self.crop_h, self.crop_w, mean, self.ignore_label, C.c_void_p(out[lo:hi].data_ptr()),
                       C.c_void_p(lab[lo:hi].data_ptr())), (hi - lo) * self.crop_h * self.crop_w * 24)
        return out, lab
