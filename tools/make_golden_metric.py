#!/usr/bin/env python
"""Golden vectors for the evaluation metric from the UNMODIFIED reference class (needs /root/reference; run in the
build container): seeded Cityscapes-like labels (piecewise-constant regions, 8 % ignore = 255) and noisy predictions ->
ConfusionMatrix.generateM + jaccard -> tests/golden/metric.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from utils.metric.metric import ConfusionMatrix  # noqa: E402  (the reference's own class)

rng = np.random.RandomState(1234)
nclass, H, W = 19, 48, 96
gts, preds = [], []
conf = ConfusionMatrix(nclass)
for img in range(3):
    coarse = rng.randint(0, nclass, (H // 8, W // 8))
    gt = np.kron(coarse, np.ones((8, 8), dtype=np.int64))
    gt[rng.rand(H, W) < 0.08] = 255
    pred = np.where(rng.rand(H, W) < 0.7, np.minimum(gt, nclass - 1), rng.randint(0, nclass, (H, W))).astype(np.uint8)
    gt[gt == 17] = 255              # a class that never occurs: exercises the non-zero-diagonal rule of jaccard()
    pred[pred == 17] = 3
    gts.append(gt.astype(np.int64))
    preds.append(pred)
    conf.addM(conf.generateM([gt.flatten(), pred.flatten()]))
aveJ, j_list, M = conf.jaccard()
np.savez_compressed(os.path.join(ROOT, "tests/golden/metric.npz"), gt=np.stack(gts), pred=np.stack(preds), M=M,
                    meanIoU=np.float64(aveJ), per_class=np.asarray(j_list, dtype=np.float64), nclass=np.int64(nclass))
print("meanIoU", aveJ, "classes with IoU", len(j_list), "pixels counted", int(M.sum()))
