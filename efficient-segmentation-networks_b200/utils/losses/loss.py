"""CrossEntropyLoss2d -- drop-in for the reference's utils/losses/loss.py:15-32
(nn.CrossEntropyLoss(weight, ignore_index, reduction='mean') on (N,C,H,W) logits / (N,H,W) int64
targets), computed by the fused esn_weighted_ce kernel (forward sums and backward gradient).
Under torch.distributed the weighted mean is taken over the GLOBAL batch."""
import torch
import torch.nn as nn

from esn import ops
from esn import train as T

__all__ = ["CrossEntropyLoss2d", "FocalLoss2d"]


class CrossEntropyLoss2d(nn.Module):
    def __init__(self, weight=None, ignore_label=255, reduction='mean', distributed=True):
        super().__init__()
        self.distributed = distributed   # all-reduce the two loss sums over the default process group
        if reduction != 'mean':
            raise NotImplementedError("only reduction='mean' (the reference's default) is on the hot path")
        self.ignore_label = ignore_label
        if weight is not None:
            self.register_buffer("weight", torch.as_tensor(weight, dtype=torch.float32))
        else:
            self.weight = None

    def forward(self, output, target):
        ops.require_cuda(output, "CrossEntropyLoss2d")
        w = self.weight
        if w is not None and w.device != output.device:
            w = w.to(output.device)
        return T.cross_entropy(output, target.long(), w, self.ignore_label, self.distributed)


class FocalLoss2d(nn.Module):
    """Drop-in for the reference's FocalLoss2d (utils/losses/loss.py:96-127; `--use_focal` in train.py).  The reference
    passes the MEAN-reduced weighted cross-entropy -- a scalar -- through the focal factor (loss.py:122-124), so the loss
    is alpha * (1 - exp(-L))^gamma * L with L the value of the fused weighted-CE kernel; the three scalar operations run
    on the device tensor (no host sync) and autograd hands their derivative to the CE kernel's backward as its upstream
    gradient, which it folds into the d-logits it writes.  size_average: mean / sum of a scalar are the identity."""

    def __init__(self, alpha=0.5, gamma=2, weight=None, ignore_index=255, size_average=True):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma
        self.weight = weight
        self.ignore_index = ignore_index
        self.size_average = size_average
        self.ce_fn = CrossEntropyLoss2d(weight=weight, ignore_label=ignore_index)

    def forward(self, output, target):
        if target.dim() == 4:          # (N,1,H,W) labels, as the reference accepts
            target = target[:, 0]
        ce = self.ce_fn(output, target)
        return self.alpha * (1.0 - torch.exp(-ce)) ** self.gamma * ce
