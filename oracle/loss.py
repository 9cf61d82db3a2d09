"""CPU restatement of the weighted cross-entropy head (TEST INFRASTRUCTURE).

Follows `CrossEntropyLoss2d` (utils/losses/loss.py:15-32), i.e.
`nn.CrossEntropyLoss(weight, ignore_index, reduction='mean')`:
    loss = sum_i w[y_i] * (-log softmax(x_i)[y_i]) / sum_i w[y_i]   over y_i != ignore
written out explicitly (not by calling F.cross_entropy) so the kernel's
arithmetic has an independent statement; `tests/test_oracle_golden.py` pins it
against the reference class.
"""
import torch


def weighted_ce(logits, target, weight=None, ignore_label=255):
    """logits (N,C,H,W) float, target (N,H,W) int64 -> (loss, sum_wl, sum_w)."""
    n, c, h, w = logits.shape
    x = logits.permute(0, 2, 3, 1).reshape(-1, c)
    y = target.reshape(-1)
    valid = y != ignore_label
    ys = torch.where(valid, y, torch.zeros_like(y))
    m = x.max(dim=1, keepdim=True).values
    lse = m.squeeze(1) + torch.log(torch.exp(x - m).sum(dim=1))
    nll = lse - x.gather(1, ys.view(-1, 1)).squeeze(1)
    wv = torch.ones(c, dtype=x.dtype) if weight is None else weight.to(x.dtype)
    wi = wv[ys] * valid.to(x.dtype)
    sum_wl = (wi * nll).sum()
    sum_w = wi.sum()
    return sum_wl / sum_w, sum_wl, sum_w


def weighted_ce_grad(logits, target, weight=None, ignore_label=255):
    """d loss / d logits for the mean-reduced weighted CE (closed form)."""
    n, c, h, w = logits.shape
    x = logits.permute(0, 2, 3, 1).reshape(-1, c)
    y = target.reshape(-1)
    valid = y != ignore_label
    ys = torch.where(valid, y, torch.zeros_like(y))
    p = torch.softmax(x, dim=1)
    onehot = torch.zeros_like(p).scatter_(1, ys.view(-1, 1), 1.0)
    wv = torch.ones(c, dtype=x.dtype) if weight is None else weight.to(x.dtype)
    wi = wv[ys] * valid.to(x.dtype)
    g = (p - onehot) * (wi / wi.sum()).view(-1, 1)
    return g.view(n, h, w, c).permute(0, 3, 1, 2).contiguous()


def focal(logits, target, weight=None, ignore_label=255, alpha=0.5, gamma=2):
    """FocalLoss2d (utils/losses/loss.py:96-127).  The reference feeds the MEAN-reduced weighted cross-entropy (a scalar,
    :122) through the focal factor, so the loss is a scalar function of the CE value:  alpha * (1 - e^-L)^gamma * L.
    Returns (loss, dloss/dlogits)."""
    ce, _, _ = weighted_ce(logits, target, weight, ignore_label)
    pt = torch.exp(-ce)
    loss = alpha * (1 - pt) ** gamma * ce
    dloss_dce = alpha * ((1 - pt) ** gamma + ce * gamma * (1 - pt) ** (gamma - 1) * pt)
    return loss, dloss_dce * weighted_ce_grad(logits, target, weight, ignore_label)


def ohem(logits, target, weight=None, ignore_label=255, thresh=0.6, min_kept=256):
    """ProbOhemCrossEntropy2d.forward (utils/losses/loss.py:187-213): online hard example mining.  mask_prob = softmax
    probability of the labelled class (1 for ignored pixels); if min_kept > number of valid pixels nothing is filtered;
    otherwise (and only if min_kept > 0) the threshold is max(thresh, min_kept-th smallest mask_prob) and pixels with
    mask_prob above it become ignored; the loss is the mean-reduced (weighted) cross-entropy of what is left.
    Returns (loss, dloss/dlogits, threshold or None)."""
    n, c, h, w = logits.shape
    y = target.reshape(-1)
    valid = y != ignore_label
    ys = torch.where(valid, y, torch.zeros_like(y))
    num_valid = int(valid.sum())
    prob = torch.softmax(logits, dim=1).transpose(0, 1).reshape(c, -1)
    threshold = None
    if min_kept > num_valid:
        pass
    elif num_valid > 0:
        mask_prob = prob[ys, torch.arange(len(ys))].masked_fill(~valid, 1.0)
        threshold = float(thresh)
        if min_kept > 0:
            kth = torch.sort(mask_prob).values[min(len(mask_prob), min_kept) - 1]
            if kth > thresh:
                threshold = float(kth)
            valid = valid & (mask_prob <= threshold)
    tgt = torch.where(valid, y, torch.full_like(y, ignore_label)).view(n, h, w)
    loss, _, _ = weighted_ce(logits, tgt, weight, ignore_label)
    return loss, weighted_ce_grad(logits, tgt, weight, ignore_label), threshold


# ProbOhemCrossEntropy2d(use_weight=True) class-balance weights, utils/losses/loss.py:173-177
OHEM_CLASS_WEIGHTS = (0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023, 0.9539, 0.9843, 1.1116,
                      0.9037, 1.0865, 1.0955, 1.0865, 1.1529, 1.0507)
