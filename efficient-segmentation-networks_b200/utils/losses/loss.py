"""CrossEntropyLoss2d -- drop-in for the reference's utils/losses/loss.py:15-32
(nn.CrossEntropyLoss(weight, ignore_index, reduction='mean') on (N,C,H,W) logits / (N,H,W) int64
targets), computed by the fused esn_weighted_ce kernel (forward sums and backward gradient).
Under torch.distributed the weighted mean is taken over the GLOBAL batch."""
import os

import torch
import torch.nn as nn

from esn import ops
from esn import train as T

__all__ = ["CrossEntropyLoss2d", "FocalLoss2d", "ProbOhemCrossEntropy2d"]


def _fused_ce(output, target, w, ignore_label, distributed, reduction, keep_thresh=None):
    ops.require_cuda(output, "CrossEntropyLoss2d")
    if w is not None and w.device != output.device:
        w = w.to(output.device)
    target = target.long()
    if os.environ.get("ESN_CHECK_LABELS") == "1":
        bad = (target != ignore_label) & ((target < 0) | (target >= output.shape[1]))
        if bool(bad.any()):
            raise IndexError("Target %d is out of bounds." % int(target[bad][0]))
    if distributed is None:
        from esn import parallel
        distributed = parallel.is_active()
    return T.cross_entropy(output, target, w, ignore_label, distributed, reduction, keep_thresh)


def fused_head_spec(criterion, output_device, target, classes):
    """What a model's fused head + loss (e.g. DABNet.fused_loss -> esn_bilinear_ce) needs from a CrossEntropyLoss2d:
    (int64 target, class weights on the device or None, ignore label, reduction, distributed) -- the same label check and the
    same resolution of `distributed=None` as the module's own forward.  None for any other criterion (no fused form)."""
    if type(criterion) is not CrossEntropyLoss2d:
        return None
    w = criterion.weight
    if w is not None and w.device != output_device:
        w = w.to(output_device)
    target = target.long()
    if os.environ.get("ESN_CHECK_LABELS") == "1":
        bad = (target != criterion.ignore_label) & ((target < 0) | (target >= classes))
        if bool(bad.any()):
            raise IndexError("Target %d is out of bounds." % int(target[bad][0]))
    distributed = criterion.distributed
    if distributed is None:
        from esn import parallel
        distributed = parallel.is_active()
    return target, w, criterion.ignore_label, criterion.reduction, distributed


class CrossEntropyLoss2d(nn.Module):
    """Same constructor as the reference (weight, ignore_label, reduction) and the same `state_dict` key
    (`nll_loss.weight`: the reference keeps an nn.CrossEntropyLoss child named nll_loss, loss.py:23; here that child only
    holds the class weights, the arithmetic is the fused kernel).

    distributed: None (default) = all-reduce the two loss sums over the default process group ONLY when the model was wrapped
    by esn.parallel.data_parallel in this process, whose gradient buckets are SUM-reduced -- each rank's loss is then already
    divided by the GLOBAL sum of class weights, so summed gradients equal the reference's gathered-batch gradient.
    Under stock DistributedDataParallel (mean-reduced gradients) or any other process-group use the loss stays local, as
    nn.CrossEntropyLoss is.  True / False force it.

    reduction: 'mean' (reference default) or 'sum'; 'none' (a per-pixel loss map) is not on the hot path and raises.
    Labels outside [0, classes) other than ignore_label are an error in torch; the kernel treats them as ignored -- set
    ESN_CHECK_LABELS=1 to validate them (one host sync per call) while debugging a dataset."""

    def __init__(self, weight=None, ignore_label=255, reduction='mean', distributed=None):
        super().__init__()
        self.distributed = distributed
        if reduction not in ('mean', 'sum'):
            raise NotImplementedError("reduction=%r is not on the hot path (only 'mean', the reference's default, and 'sum')"
                                      % (reduction,))
        self.reduction = reduction
        self.ignore_label = ignore_label
        self.nll_loss = nn.CrossEntropyLoss(weight=None if weight is None else torch.as_tensor(weight, dtype=torch.float32),
                                            ignore_index=ignore_label, reduction=reduction)

    @property
    def weight(self):
        return self.nll_loss.weight

    def forward(self, output, target):
        return _fused_ce(output, target, self.weight, self.ignore_label, self.distributed, self.reduction)


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
        # the reference's child and state_dict key (`ce_fn.weight`, loss.py:104); here it only holds the class weights
        self.ce_fn = nn.CrossEntropyLoss(weight=None if weight is None else torch.as_tensor(weight, dtype=torch.float32),
                                         ignore_index=ignore_index)

    def forward(self, output, target):
        if target.dim() == 4:          # (N,1,H,W) labels, as the reference accepts
            target = target[:, 0]
        ce = _fused_ce(output, target, self.ce_fn.weight, self.ignore_index, None, "mean")
        return self.alpha * (1.0 - torch.exp(-ce)) ** self.gamma * ce


class ProbOhemCrossEntropy2d(nn.Module):
    """Drop-in for the reference's ProbOhemCrossEntropy2d (utils/losses/loss.py:163-216; what train.py:147-149 selects for
    Cityscapes with --use_ohem): online hard example mining.  Same constructor, same `criterion` child (state_dict key
    `criterion.weight` with use_weight=True, the class-balance weights of loss.py:173-177).

    Three device steps, no host synchronisation (the reference argsorts all pixels and compares on the host twice per step):
    1. esn_weighted_ce writes the softmax probability of the labelled class per pixel (mask_prob) and counts the valid pixels;
    2. esn_ohem_threshold finds max(thresh, min_kept-th smallest mask_prob) by an exact radix select (+inf when
       min_kept > number of valid pixels, where the reference filters nothing);
    3. the fused weighted-CE kernel (forward sums, and the gradient in backward) skips pixels above that threshold.
    The threshold carries no gradient, as in the reference (masks are index / comparison results).  down_ratio is stored and
    unused, as in the reference.  Under esn.parallel the threshold is per rank (each rank mines its own shard)."""

    def __init__(self, ignore_label=255, reduction='mean', thresh=0.6, min_kept=256, down_ratio=1, use_weight=False):
        super().__init__()
        if reduction not in ('mean', 'sum'):
            raise NotImplementedError("reduction=%r is not on the hot path" % (reduction,))
        self.ignore_label = ignore_label
        self.thresh = float(thresh)
        self.min_kept = int(min_kept)
        self.down_ratio = down_ratio
        self.reduction = reduction
        weight = None
        if use_weight:
            weight = torch.FloatTensor([0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023, 0.9539,
                                        0.9843, 1.1116, 0.9037, 1.0865, 1.0955, 1.0865, 1.1529, 1.0507])      # loss.py:173-177
        self.criterion = nn.CrossEntropyLoss(reduction=reduction, weight=weight, ignore_index=ignore_label)

    def forward(self, pred, target):
        ops.require_cuda(pred, "ProbOhemCrossEntropy2d")
        target = target.long().contiguous()
        lg = pred.detach().contiguous()
        prob = torch.empty(target.shape, dtype=torch.float32, device=pred.device)
        sums, _ = ops.weighted_ce(lg, target, None, self.ignore_label, prob_out=prob)      # sums[1] = number of valid pixels
        thr = ops.ohem_threshold(prob, self.min_kept, self.thresh, sums[1:2])
        return _fused_ce(pred, target, self.criterion.weight, self.ignore_label, None, self.reduction, keep_thresh=thr)
