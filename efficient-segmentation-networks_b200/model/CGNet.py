"""CGNet on B200 kernels -- drop-in for the reference's model/CGNet.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/CGNet.py:13-367.  Inference launch plan per ContextGuidedBlock (CGNet.py:230-260),
6 launches instead of ~17 ATen kernels:
  1x1 conv + BN + PReLU (tcgen05) -> depthwise 3x3 "loc" and dilated depthwise 3x3 "sur", each writing its
  half of the joint buffer with its slice of BN + PReLU in the epilogue (the torch.cat never happens)
  -> per-image global average pool -> gate MLP (two Linear layers, ReLU, Sigmoid) -> x * gate + input.
ContextGuidedBlock_Down (CGNet.py:193-227) adds the strided 3x3 (tensor cores, over the zero-padded concat
width) and the 1x1 "reduce".  Stem, input injection, concat BNPReLU and the classifier + bilinear head are
shared with DABNet's plan.
"""
import torch
import torch.nn as nn

from esn import ops
from esn._lib import ACT_NONE, ACT_PRELU
from esn.prep import PrepMixin, weights_generation

__all__ = ["CGNet"]


def _no_train(mod):
    if mod.training:
        raise NotImplementedError(
            "%s: training-mode kernels are not built yet for this model; call .eval(). "
            "There is no eager-PyTorch fallback." % type(mod).__name__)


def _affine(bn, act, device):
    s, b = ops.bn_affine(bn, device)
    return s.contiguous(), b.contiguous(), act.weight.detach().to(device=device, dtype=torch.float32).contiguous()


def _tc_channels(c):
    for v in (16, 32, 64):
        if c <= v:
            return v
    return (c + 63) // 64 * 64


class ConvBNPReLU(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False)
        self.bn = nn.BatchNorm2d(nOut, eps=1e-03)
        self.act = nn.PReLU(nOut)

    def _build_prep(self, device):
        s, b, a = _affine(self.bn, self.act, device)
        cin = self.conv.in_channels
        padded = None
        if cin > 8 and _tc_channels(cin) != cin:
            padded = ops.ConvPrep(self.conv, s, b, ACT_PRELU, a, device=device, cin_pad=_tc_channels(cin))
        return ops.ConvPrep(self.conv, s, b, ACT_PRELU, a, device=device), padded

    def forward(self, input, out=None):
        _no_train(self)
        c = self.conv
        if (input.shape[1] == 3 and input.dtype == torch.float32 and input.is_contiguous() and not ops.is_nhwc(input)
                and c.kernel_size == (3, 3) and c.stride == (2, 2) and c.out_channels % 4 == 0 and c.out_channels <= 32):
            ops.require_cuda(input, "ConvBNPReLU")
            prep, _ = self.prep(input.device)
            if out is None:
                n, _, h, w = input.shape
                ho, wo = prep.out_hw(h, w)
                out = ops.new_act(n, prep.cout, ho, wo, ops.compute_dtype(input), input.device)
            return ops.stem_conv3x3s2(input, prep.w_direct, prep.cout, 0, out, prep.scale, prep.shift, prep.alpha, prep.act)
        x = ops.as_act(input)
        prep, padded = self.prep(x.device)
        if padded is not None and x.dtype == torch.bfloat16 and x.stride(3) >= padded.cin:
            return ops.conv2d(ops.widen(x, padded.cin), padded, out=out)    # zero-padded concat buffer on tensor cores
        return ops.conv2d(x, prep, out=out)


class BNPReLU(PrepMixin, nn.Module):
    def __init__(self, nOut):
        super().__init__()
        self.bn = nn.BatchNorm2d(nOut, eps=1e-03)
        self.act = nn.PReLU(nOut)

    def _build_prep(self, device):
        s, b, a = _affine(self.bn, self.act, device)
        z = torch.zeros((-s.numel()) % 8, device=device)
        return s, b, a, torch.cat([s, z]), torch.cat([b, z]), torch.cat([a, z])

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        s, b, a, sp, bp, ap = self.prep(x.device)
        if x.shape[1] % 4 and out is x and x.stride(3) >= sp.numel():
            xw = ops.widen(x, sp.numel())         # vector path over the zero-padded width of a concat buffer
            ops.affine_act(xw, sp, bp, ap, ACT_PRELU, out=xw)
            return x
        return ops.affine_act(x, s, b, a, ACT_PRELU, out=out)


class ConvBN(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False)
        self.bn = nn.BatchNorm2d(nOut, eps=1e-03)

    def _build_prep(self, device):
        s, b = ops.bn_affine(self.bn, device)
        return ops.ConvPrep(self.conv, s, b, ACT_NONE, device=device)

    def forward(self, input):
        _no_train(self)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device))


class _PlainConv(PrepMixin, nn.Module):
    """conv only (Conv / ChannelWiseConv / DilatedConv / ChannelWiseDilatedConv of the reference)."""

    def _build_prep(self, device):
        return ops.ConvPrep(self.conv, device=device)

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device), out=out)


class Conv(_PlainConv):
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False)


class ChannelWiseConv(_PlainConv):
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), groups=nIn, bias=False)


class DilatedConv(_PlainConv):
    def __init__(self, nIn, nOut, kSize, stride=1, d=1):
        super().__init__()
        padding = int((kSize - 1) / 2) * d
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False, dilation=d)


class ChannelWiseDilatedConv(_PlainConv):
    def __init__(self, nIn, nOut, kSize, stride=1, d=1):
        super().__init__()
        padding = int((kSize - 1) / 2) * d
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), groups=nIn, bias=False,
                              dilation=d)


class FGlo(PrepMixin, nn.Module):
    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel), nn.Sigmoid())

    def _build_prep(self, device):
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        return f(self.fc[0].weight), f(self.fc[0].bias), f(self.fc[2].weight), f(self.fc[2].bias)

    def forward(self, x, out=None, residual=None):
        _no_train(self)
        x = ops.as_act(x)
        w1, b1, w2, b2 = self.prep(x.device)
        return ops.fglo_gate(x, w1, b1, w2, b2, out=out, residual=residual)


def _joint(block, y, bn, act, device):
    """[F_loc(y), F_sur(y)] -> BN -> PReLU, each depthwise conv writing its half with its BN/PReLU slice."""
    n, c, h, w = y.shape
    key = (str(device), weights_generation(), tuple((t.data_ptr(), t._version) for t in list(bn.parameters()) + list(bn.buffers()) + [act.weight]),
           block.F_loc.conv.weight._version, block.F_sur.conv.weight._version, bn.eps)
    cached = block.__dict__.get("_esn_joint")
    if cached is None or cached[0] != key:
        s, b, a = _affine(bn, act, device)
        loc = ops.ConvPrep(block.F_loc.conv, s[:c], b[:c], ACT_PRELU, a[:c], device=device)
        sur = ops.ConvPrep(block.F_sur.conv, s[c:], b[c:], ACT_PRELU, a[c:], device=device)
        cached = (key, (loc, sur))
        block.__dict__["_esn_joint"] = cached
    loc, sur = cached[1]
    j = ops.new_act(n, 2 * c, h, w, y.dtype, y.device)
    ops.conv2d(y, loc, out=j[:, :c])
    ops.conv2d(y, sur, out=j[:, c:])
    return j


class ContextGuidedBlock_Down(nn.Module):
    def __init__(self, nIn, nOut, dilation_rate=2, reduction=16):
        super().__init__()
        self.conv1x1 = ConvBNPReLU(nIn, nOut, 3, 2)
        self.F_loc = ChannelWiseConv(nOut, nOut, 3, 1)
        self.F_sur = ChannelWiseDilatedConv(nOut, nOut, 3, 1, dilation_rate)
        self.bn = nn.BatchNorm2d(2 * nOut, eps=1e-3)
        self.act = nn.PReLU(2 * nOut)
        self.reduce = Conv(2 * nOut, nOut, 1, 1)
        self.F_glo = FGlo(nOut, reduction)

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        y = self.conv1x1(x)
        j = _joint(self, y, self.bn, self.act, x.device)
        r = self.reduce(j)
        return self.F_glo(r, out=out)


class ContextGuidedBlock(nn.Module):
    def __init__(self, nIn, nOut, dilation_rate=2, reduction=16, add=True):
        super().__init__()
        n = int(nOut / 2)
        self.conv1x1 = ConvBNPReLU(nIn, n, 1, 1)
        self.F_loc = ChannelWiseConv(n, n, 3, 1)
        self.F_sur = ChannelWiseDilatedConv(n, n, 3, 1, dilation_rate)
        self.bn_prelu = BNPReLU(nOut)
        self.add = add
        self.F_glo = FGlo(nOut, reduction)

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        y = self.conv1x1(x)
        j = _joint(self, y, self.bn_prelu.bn, self.bn_prelu.act, x.device)
        return self.F_glo(j, out=out, residual=x if self.add else None)


class InputInjection(nn.Module):
    def __init__(self, downsamplingRatio):
        super().__init__()
        self.pool = nn.ModuleList()
        for i in range(0, downsamplingRatio):
            self.pool.append(nn.AvgPool2d(3, stride=2, padding=1))

    def forward(self, input):
        ops.require_cuda(input, "InputInjection")
        x = input
        for _ in self.pool:
            n, c, h, w = x.shape
            y = ops.new_act(n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, torch.float32, x.device, c_alloc=4)
            x = ops.avgpool3x3s2(x if (ops.is_nhwc(x) or x.is_contiguous()) else x.contiguous(), y)
        return x


class CGNet(PrepMixin, nn.Module):
    def __init__(self, classes=19, M=3, N=21, dropout_flag=False):
        super().__init__()
        self.level1_0 = ConvBNPReLU(3, 32, 3, 2)
        self.level1_1 = ConvBNPReLU(32, 32, 3, 1)
        self.level1_2 = ConvBNPReLU(32, 32, 3, 1)
        self.sample1 = InputInjection(1)
        self.sample2 = InputInjection(2)
        self.b1 = BNPReLU(32 + 3)
        self.level2_0 = ContextGuidedBlock_Down(32 + 3, 64, dilation_rate=2, reduction=8)
        self.level2 = nn.ModuleList()
        for i in range(0, M - 1):
            self.level2.append(ContextGuidedBlock(64, 64, dilation_rate=2, reduction=8))
        self.bn_prelu_2 = BNPReLU(128 + 3)
        self.level3_0 = ContextGuidedBlock_Down(128 + 3, 128, dilation_rate=4, reduction=16)
        self.level3 = nn.ModuleList()
        for i in range(0, N - 1):
            self.level3.append(ContextGuidedBlock(128, 128, dilation_rate=4, reduction=16))
        self.bn_prelu_3 = BNPReLU(256)
        if dropout_flag:
            self.classifier = nn.Sequential(nn.Dropout2d(0.1, False), Conv(256, classes, 1, 1))
        else:
            self.classifier = nn.Sequential(Conv(256, classes, 1, 1))
        for m in self.modules():        # the reference re-initialises every Conv2d (CGNet.py:316-325)
            if m.__class__.__name__.find('Conv2d') != -1:
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    m.bias.data.zero_()

    def _build_prep(self, device):
        conv = self.classifier[-1].conv
        return ops.ConvPrep(conv, device=device, cout_pad=32), conv.out_channels

    def _scores(self, input):
        ops.require_cuda(input, "CGNet")
        _no_train(self)
        if input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        dt = ops.compute_dtype(input)
        dev = input.device
        n, _, h, w = input.shape
        inp1 = self.sample1(input)
        inp2 = self.sample1(inp1)
        cat0 = ops.new_act(n, 35, inp1.shape[2], inp1.shape[3], dt, dev, c_alloc=64, zero=True)
        y = self.level1_0(input)
        y = self.level1_1(y)
        self.level1_2(y, out=cat0[:, :32])
        ops.affine_act(inp1, None, None, None, ACT_NONE, out=cat0[:, 32:35])
        self.b1(cat0, out=cat0)

        cat1 = ops.new_act(n, 131, inp2.shape[2], inp2.shape[3], dt, dev, c_alloc=192, zero=True)
        y10 = self.level2_0(cat0, out=cat1[:, 64:128])
        y = y10
        for i, layer in enumerate(self.level2):
            y = layer(y, out=cat1[:, 0:64] if i == len(self.level2) - 1 else None)
        ops.affine_act(inp2, None, None, None, ACT_NONE, out=cat1[:, 128:131])
        self.bn_prelu_2(cat1, out=cat1)

        h3, w3 = (cat1.shape[2] - 1) // 2 + 1, (cat1.shape[3] - 1) // 2 + 1
        cat2 = ops.new_act(n, 256, h3, w3, dt, dev)
        y20 = self.level3_0(cat1, out=cat2[:, 0:128])
        y = y20
        for i, layer in enumerate(self.level3):
            y = layer(y, out=cat2[:, 128:256] if i == len(self.level3) - 1 else None)
        self.bn_prelu_3(cat2, out=cat2)

        cls, classes = self.prep(dev)
        if dt == torch.bfloat16:
            scores = ops.new_act(n, classes, h3, w3, torch.bfloat16, dev, c_alloc=32)
            ops.conv2d(cat2, cls, out=ops.widen(scores, 32))
        else:
            scores = ops.new_act(n, classes, h3, w3, torch.float32, dev, c_alloc=32)
            self.classifier[-1](cat2, out=scores)
        return scores, (h, w), dt

    def fused_loss(self, input, target, criterion):
        """criterion(self(input), target) (train.py:351-352) with the bilinear head (CGNet.py:332), CrossEntropyLoss2d and both
        their backward passes as ONE launch (esn_bilinear_ce); esn.graph.GraphedTrainStep calls this.  Falls back to the
        two-module form for other criteria and in eval mode."""
        from esn import train as T
        from model._cgnet_train import cgnet_train_forward
        return T.fused_bilinear_loss(self, cgnet_train_forward, input, target, criterion, self.classifier[-1].conv.out_channels)

    def forward(self, input):
        if self.training:
            # batch-statistics BatchNorm and the recorded backward (esn/train.py); one autograd node for the net
            from esn import train as T
            from model._cgnet_train import cgnet_train_forward
            return T.run_network(self, lambda inp: cgnet_train_forward(self, inp), input)
        scores, (h, w), dt = self._scores(input)
        ldt = torch.bfloat16 if dt == torch.bfloat16 else torch.float32
        return ops.head_bilinear(scores, scores.shape[1], h, w, True, False, ldt)[0]

    @torch.no_grad()
    def predict_mask(self, input, with_logits=False):
        scores, (h, w), dt = self._scores(input)
        ldt = torch.bfloat16 if dt == torch.bfloat16 else torch.float32
        logits, mask = ops.head_bilinear(scores, scores.shape[1], h, w, with_logits, True, ldt)
        return (logits, mask) if with_logits else mask
