#!/usr/bin/env python
"""Golden vectors for the input pipeline from the UNMODIFIED reference dataset classes (needs /root/reference and cv2;
run in the build container): seeded uint8 BGR images are written as lossless PNGs in the Cityscapes directory layout,
read back through CityscapesValDataSet / CityscapesTestDataSet.__getitem__ with the mean of the shipped pickle, and the
resulting float32 CHW arrays stored next to the raw images -> tests/golden/pipeline.npz."""
import os
import pickle
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)
from dataset.cityscapes import CityscapesTestDataSet, CityscapesValDataSet  # noqa: E402  (the reference's own classes)

mean = pickle.load(open(os.path.join(REF, "dataset/inform/cityscapes_inform.pkl"), "rb"))["mean"]
rng = np.random.RandomState(1234)
out = {"mean": np.asarray(mean)}
with tempfile.TemporaryDirectory() as root:
    lines = []
    for i, (h, w) in enumerate(((24, 40), (17, 23), (32, 64))):       # one size with h*w % 4 != 0
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        img[0, 0] = (0, 0, 0)
        img[0, 1] = (255, 255, 255)
        lab = rng.randint(0, 19, (h, w)).astype(np.uint8)
        rel_i = "leftImg8bit/val/city/img%d_leftImg8bit.png" % i
        rel_l = "gtFine/val/city/img%d_gtFine_labelTrainIds.png" % i
        for rel, arr in ((rel_i, img), (rel_l, lab)):
            os.makedirs(os.path.dirname(os.path.join(root, rel)), exist_ok=True)
            assert cv2.imwrite(os.path.join(root, rel), arr)
        lines.append("%s %s" % (rel_i, rel_l))
        out["image%d" % i] = img
    lst = os.path.join(root, "list.txt")
    open(lst, "w").write("\n".join(lines) + "\n")
    val = CityscapesValDataSet(root, lst, f_scale=1, mean=mean)
    test = CityscapesTestDataSet(root, lst, mean=mean)
    for i in range(len(lines)):
        x, _, size, _ = val[i]
        xt, _, _ = test[i]
        assert x.dtype == np.float32 and np.array_equal(x, xt)
        out["input%d" % i] = x
np.savez_compressed(os.path.join(ROOT, "tests/golden/pipeline.npz"), **out)
print("pipeline golden written:", {k: v.shape for k, v in out.items()}, "mean", mean, mean.dtype)
