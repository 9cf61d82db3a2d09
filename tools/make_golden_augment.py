#!/usr/bin/env python
"""Golden vectors for the training-time augmentation from the UNMODIFIED reference class CityscapesDataSet
(dataset/cityscapes.py:27-106; needs /root/reference and cv2, run in the build container): seeded uint8 BGR images and label
maps are written as lossless PNGs, the class is read with `random` / `np.random` seeded per sample, and its outputs are stored
next to the raw inputs and the seeds -> tests/golden/augment.npz.  Twelve samples cover every scale factor, padding in one /
both / no direction, and both mirror states."""
import os
import pickle
import random
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)
from dataset.cityscapes import CityscapesDataSet  # noqa: E402  (the reference's own class)

mean = pickle.load(open(os.path.join(REF, "dataset/inform/cityscapes_inform.pkl"), "rb"))["mean"]
rng = np.random.RandomState(4321)
out = {"mean": np.asarray(mean)}
sizes = ((40, 72), (33, 47), (64, 128))
crops = ((32, 64), (48, 48), (40, 96))
with tempfile.TemporaryDirectory() as root:
    lines = []
    for i, (h, w) in enumerate(sizes):
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        lab = rng.randint(0, 19, (h, w)).astype(np.uint8)
        lab[rng.rand(h, w) < 0.05] = 255
        rel_i, rel_l = "leftImg8bit/train/c/i%d_leftImg8bit.png" % i, "gtFine/train/c/i%d_gtFine_labelTrainIds.png" % i
        for rel, arr in ((rel_i, img), (rel_l, lab)):
            os.makedirs(os.path.dirname(os.path.join(root, rel)), exist_ok=True)
            assert cv2.imwrite(os.path.join(root, rel), arr)
        lines.append("%s %s" % (rel_i, rel_l))
        out["image%d" % i], out["label%d" % i] = img, lab
    lst = os.path.join(root, "list.txt")
    open(lst, "w").write("\n".join(lines) + "\n")
    k = 0
    for ci, crop in enumerate(crops):
        ds = CityscapesDataSet(root, lst, crop_size=crop, mean=mean, scale=True, mirror=True, ignore_label=255)
        for i in range(len(sizes)):
            for seed in (11 + k, 101 + k):
                random.seed(seed)
                np.random.seed(seed)
                x, y, size, _ = ds[i]
                out["case%d" % k] = np.array([i, crop[0], crop[1], seed])
                out["x%d" % k], out["y%d" % k] = x, y
                k += 1
    ds = CityscapesDataSet(root, lst, crop_size=(32, 32), mean=mean, scale=False, mirror=False, ignore_label=255)
    random.seed(5)
    np.random.seed(5)
    x, y, _, _ = ds[0]
    out["noscale_x"], out["noscale_y"] = x, y
out["n_cases"] = np.array([k])
np.savez_compressed(os.path.join(ROOT, "tests/golden/augment.npz"), **out)
print("augment golden written:", k, "cases")
