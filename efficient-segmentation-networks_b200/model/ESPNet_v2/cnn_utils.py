"""ESPNetv2 building blocks on B200 kernels -- drop-in for the reference's model/ESPNet_v2/cnn_utils.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/ESPNet_v2/cnn_utils.py:11-190.  Every block folds its eval-mode BatchNorm and PReLU into
the epilogue of the conv that produces it; grouped 1x1 convs run as one dense conv per group over channel
slices of the NHWC buffers (tcgen05 when the per-group K is 16/32/64/128), depthwise convs use the vectorised
NHWC stencil.
"""
import torch
import torch.nn as nn

from esn import ops
from esn._lib import ACT_NONE, ACT_PRELU
from esn.prep import PrepMixin

__all__ = ["PSPModule", "CBR", "BR", "CB", "C", "CDilated", "CDilatedB"]


def _no_train(mod):
    if mod.training:
        raise NotImplementedError(
            "%s: training-mode kernels are not built yet for this model; call .eval(). "
            "There is no eager-PyTorch fallback." % type(mod).__name__)


def fold(conv, bn=None, prelu=None, device=None, **kw):
    """ConvPrep of conv -> eval BN -> PReLU."""
    s, b = ops.bn_affine(bn, device) if bn is not None else (None, None)
    if prelu is not None:
        return ops.ConvPrep(conv, s, b, ACT_PRELU, prelu.weight, device=device, **kw)
    return ops.ConvPrep(conv, s, b, ACT_NONE, device=device, **kw)


class _ConvBlock(PrepMixin, nn.Module):
    """conv (+bn) (+act): the forward shared by CBR / CB / C / CDilated / CDilatedB."""

    def _build_prep(self, device):
        return fold(self.conv, getattr(self, "bn", None), getattr(self, "act", None), device)

    def forward(self, input, out=None, residual=None):
        _no_train(self)
        c = self.conv
        if (input.shape[1] == 3 and input.dtype == torch.float32 and input.is_contiguous() and not ops.is_nhwc(input)
                and c.kernel_size == (3, 3) and c.stride == (2, 2) and c.padding == (1, 1) and c.groups == 1
                and c.out_channels % 4 == 0 and c.out_channels <= 32 and residual is None
                and not ((input.shape[2] | input.shape[3]) & 1)):
            ops.require_cuda(input, type(self).__name__)
            prep = self.prep(input.device)
            n, _, h, w = input.shape
            if out is None:
                out = ops.new_act(n, prep.cout, h // 2, w // 2, ops.compute_dtype(input), input.device)
            return ops.stem_conv3x3s2(input, prep.w_direct, prep.cout, 0, out, prep.scale, prep.shift, prep.alpha, prep.act)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device), out=out, residual=residual)


class CBR(_ConvBlock):
    def __init__(self, nIn, nOut, kSize, stride=1, groups=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, kSize, stride=stride, padding=padding, bias=False, groups=groups)
        self.bn = nn.BatchNorm2d(nOut)
        self.act = nn.PReLU(nOut)


class BR(PrepMixin, nn.Module):
    def __init__(self, nOut):
        super().__init__()
        self.bn = nn.BatchNorm2d(nOut)
        self.act = nn.PReLU(nOut)

    def _build_prep(self, device):
        s, b = ops.bn_affine(self.bn, device)
        return s, b, self.act.weight.detach().float().to(device).contiguous()

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        s, b, a = self.prep(x.device)
        return ops.affine_act(x, s, b, a, ACT_PRELU, out=out)


class CB(_ConvBlock):
    def __init__(self, nIn, nOut, kSize, stride=1, groups=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, kSize, stride=stride, padding=padding, bias=False, groups=groups)
        self.bn = nn.BatchNorm2d(nOut)


class C(_ConvBlock):
    def __init__(self, nIn, nOut, kSize, stride=1, groups=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, kSize, stride=stride, padding=padding, bias=False, groups=groups)


class CDilated(_ConvBlock):
    def __init__(self, nIn, nOut, kSize, stride=1, d=1, groups=1):
        super().__init__()
        padding = int((kSize - 1) / 2) * d
        self.conv = nn.Conv2d(nIn, nOut, kSize, stride=stride, padding=padding, bias=False, dilation=d, groups=groups)


class CDilatedB(_ConvBlock):
    def __init__(self, nIn, nOut, kSize, stride=1, d=1, groups=1):
        super().__init__()
        padding = int((kSize - 1) / 2) * d
        self.conv = nn.Conv2d(nIn, nOut, kSize, stride=stride, padding=padding, bias=False, dilation=d, groups=groups)
        self.bn = nn.BatchNorm2d(nOut)


class PSPModule(nn.Module):
    def __init__(self, features, out_features=1024, sizes=(1, 2, 4, 8)):
        super().__init__()
        self.stages = nn.ModuleList([C(features, features, 3, 1, groups=features) for size in sizes])
        self.project = CBR(features * (len(sizes) + 1), out_features, 1, 1)

    def forward(self, feats, cat=None, out=None):
        """`cat`: optional pre-allocated (N, 5*C, H, W) buffer whose first C channels already hold feats."""
        _no_train(self)
        x = ops.as_act(feats)
        n, c, h, w = x.shape
        if cat is None:
            cat = ops.new_act(n, c * (len(self.stages) + 1), h, w, x.dtype, x.device)
            ops.affine_act(x, None, None, None, ACT_NONE, out=cat[:, :c])
        f = x
        for i, stage in enumerate(self.stages):
            fh, fw = (f.shape[2] - 1) // 2 + 1, (f.shape[3] - 1) // 2 + 1
            f = ops.avgpool3x3s2(f, ops.new_act(n, c, fh, fw, x.dtype, x.device))
            ops.bilinear(stage(f), h, w, True, out=cat[:, (i + 1) * c:(i + 2) * c])
        return self.project(cat, out=out)
