"""Segmentation metrics on the device: the confusion matrix of the (fused-argmax) predicted masks against the labels is
accumulated by `esn_confusion_matrix`; the 19 x 19 result is the only thing that crosses PCIe.

Replaces the reference's evaluation tail (`test.py:79-90`, `train.py:398-404`): D2H copy of the fp32 logits, numpy argmax,
`data_list.append([gt.flatten(), output.flatten()])` and the per-pixel Python loop of `ConfusionMatrix.generateM`
(`utils/metric/metric.py:68-76`) in a multiprocessing pool.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import ops


class ConfusionMatrix:
    """Same arithmetic and `jaccard()` return value as `utils/metric/metric.py:ConfusionMatrix`; `add_batch` takes CUDA
    tensors (pred: uint8 masks from `model.predict_mask`, gt: uint8 or int64 labels, any shape with equal numel)."""

    def __init__(self, nclass, classes=None, ignore_label=255, device="cuda"):
        self.nclass = nclass
        self.classes = classes
        self.ignore_label = ignore_label
        self._m = torch.zeros(nclass * nclass, dtype=torch.int64, device=device)

    def add_batch(self, pred, gt):
        ops.require_cuda(pred, "ConfusionMatrix.add_batch")
        ops.require_cuda(gt, "ConfusionMatrix.add_batch")
        if pred.dtype != torch.uint8:
            raise TypeError("pred must be a uint8 mask")
        if gt.dtype not in (torch.uint8, torch.int64):
            raise TypeError("gt must be uint8 or int64")
        if pred.numel() != gt.numel():
            raise ValueError("pred and gt differ in size")
        pred, gt = pred.contiguous(), gt.contiguous()
        L.check(L.lib.esn_confusion_matrix(C.c_void_p(pred.data_ptr()), C.c_void_p(gt.data_ptr()), int(gt.dtype == torch.int64),
                                           pred.numel(), self.nclass, C.c_void_p(self._m.data_ptr()), ops.stream()),
                "esn_confusion_matrix")

    @property
    def M(self):
        """The matrix as the reference keeps it: float64 [nclass, nclass], rows = ground truth, columns = prediction."""
        return self._m.view(self.nclass, self.nclass).cpu().numpy().astype(np.float64)

    def jaccard(self):
        return jaccard_from_matrix(self.M)


def jaccard_from_matrix(M):
    """`ConfusionMatrix.jaccard` (utils/metric/metric.py:58-66): per-class IoU over the classes with a non-zero diagonal,
    their mean, and the matrix."""
    per_class = []
    for i in range(M.shape[0]):
        if not M[i, i] == 0:
            per_class.append(M[i, i] / (np.sum(M[i, :]) + np.sum(M[:, i]) - M[i, i]))
    return np.sum(per_class) / len(per_class), per_class, M
