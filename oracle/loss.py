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
