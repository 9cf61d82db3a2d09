"""CrossEntropyLoss2d -- drop-in for the reference's utils/losses/loss.py:15-32
(nn.CrossEntropyLoss(weight, ignore_index, reduction='mean') on (N,C,H,W) logits / (N,H,W) int64
targets), computed by the fused esn_weighted_ce kernel (forward sums and backward gradient).
Under torch.distributed the weighted mean is taken over the GLOBAL batch."""
import torch
import torch.nn as nn

from esn import train as T

__all__ = ["CrossEntropyLoss2d"]


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
        if not output.is_cuda:
            raise RuntimeError("CrossEntropyLoss2d: tensors must live on a CUDA device; this framework has no CPU path")
        w = self.weight
        if w is not None and w.device != output.device:
            w = w.to(output.device)
        return T.cross_entropy(output, target.long(), w, self.ignore_label, self.distributed)
