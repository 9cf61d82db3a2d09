"""Fast-SCNN on B200 kernels -- drop-in for the reference's model/FastSCNN.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/FastSCNN.py:15-235.  Launch plan: every conv carries its BatchNorm (+ReLU, +shortcut)
in the epilogue; 1x1 convs (the bulk of the MACs: K = 64...768) run on the tcgen05 kernel -- channel counts
that are not tensor-core friendly (48, 96) live in buffers padded to 64 / 128 with exact-zero tails produced
by zero weight rows, expansions wider than 256 outputs run as 256-channel N slices; depthwise 3x3 convs use
the vectorised NHWC stencil; pyramid pooling = adaptive-average-pool + tiny 1x1 + bilinear
(align_corners=True) written straight into the concat buffer; the final x8 bilinear (align_corners=True) is
fused with the NCHW logits store / argmax.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (kept for parity with the reference module's namespace)

from esn import ops
from esn._lib import ACT_NONE, ACT_RELU
from esn.prep import PrepMixin

__all__ = ["FastSCNN"]

_STEM_PAD0 = 256


def _no_train(mod):
    if mod.training:
        raise NotImplementedError(
            "%s: training-mode kernels are not built yet for this model; call .eval(). "
            "There is no eager-PyTorch fallback." % type(mod).__name__)


def _tc(c):
    """Channel count the tensor-core conv accepts as Cin: 16/32/64 or a multiple of 64."""
    for v in (16, 32, 64):
        if c <= v:
            return v
    return (c + 63) // 64 * 64


def _conv_prep(conv, bn, act, device):
    """conv (+bias) -> eval BN -> act, with Cin / Cout zero-padded to tensor-core friendly widths when
    the conv is dense (the padded output channels evaluate to exactly 0)."""
    s, b = ops.bn_affine(bn, device) if bn is not None else (None, None)
    dense = conv.groups == 1 and conv.in_channels > 8
    cin_pad = _tc(conv.in_channels) if dense and _tc(conv.in_channels) != conv.in_channels else None
    cout_pad = _tc(conv.out_channels) if (conv.groups == 1 and conv.out_channels % 16 == 0
                                          and _tc(conv.out_channels) != conv.out_channels) else None
    plain = ops.ConvPrep(conv, s, b, act, device=device)
    padded = ops.ConvPrep(conv, s, b, act, device=device, cin_pad=cin_pad, cout_pad=cout_pad) if (cin_pad or cout_pad) else None
    return plain, padded


def _run_conv(x, preps, residual=None, out=None):
    """Pick the padded (tensor-core) variant when x is bf16 and wide enough; output buffers are always
    allocated at the padded width so the next 1x1 conv can read zero tails."""
    plain, padded = preps
    n, c, h, w = x.shape
    ho, wo = plain.out_hw(h, w)
    width = _tc(plain.cout) if plain.groups == 1 and plain.cout % 16 == 0 else plain.cout
    if out is None:
        out = ops.new_act(n, plain.cout, ho, wo, x.dtype, x.device, c_alloc=max(width, plain.cout), zero=padded is None and width != plain.cout)
    if padded is not None and x.dtype == torch.bfloat16 and x.stride(3) >= padded.cin and out.stride(3) >= padded.cout:
        res = None if residual is None else ops.widen(residual, padded.cout)
        ops.conv2d(ops.widen(x, padded.cin), padded, out=ops.widen(out, padded.cout), residual=res)
        return out
    if padded is not None and out.stride(3) > plain.cout:
        ops.widen(out, out.stride(3))[:, plain.cout:].zero_()     # exact path: keep the padded tail zero
    ops.conv2d(x, plain, out=out, residual=residual)
    return out


class _ConvBNReLU(PrepMixin, nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, **kwargs):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=False),
                                  nn.BatchNorm2d(out_channels), nn.ReLU(True))

    def _build_prep(self, device):
        return _conv_prep(self.conv[0], self.conv[1], ACT_RELU, device)

    def forward(self, x, out=None):
        _no_train(self)
        c = self.conv[0]
        if (x.shape[1] == 3 and x.dtype == torch.float32 and x.is_contiguous() and not ops.is_nhwc(x)
                and c.kernel_size == (3, 3) and c.stride == (2, 2) and c.padding in ((0, 0), (1, 1))
                and c.out_channels % 4 == 0 and c.out_channels <= 32):
            ops.require_cuda(x, "_ConvBNReLU")
            prep, _ = self.prep(x.device)
            n, _, h, w = x.shape
            ho, wo = prep.out_hw(h, w)
            y = ops.new_act(n, prep.cout, ho, wo, ops.compute_dtype(x), x.device)
            return ops.stem_conv3x3s2(x, prep.w_direct, prep.cout, _STEM_PAD0 if c.padding == (0, 0) else 0, y,
                                      prep.scale, prep.shift, None, ACT_RELU)
        x = ops.as_act(x)
        return _run_conv(x, self.prep(x.device), out=out)


class _DSConv(PrepMixin, nn.Module):
    def __init__(self, dw_channels, out_channels, stride=1, **kwargs):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(dw_channels, dw_channels, 3, stride, 1, groups=dw_channels, bias=False), nn.BatchNorm2d(dw_channels),
            nn.ReLU(True), nn.Conv2d(dw_channels, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels), nn.ReLU(True))

    def _build_prep(self, device):
        return _conv_prep(self.conv[0], self.conv[1], ACT_RELU, device), _conv_prep(self.conv[3], self.conv[4], ACT_RELU, device)

    def forward(self, x):
        _no_train(self)
        x = ops.as_act(x)
        dw, pw = self.prep(x.device)
        y = _dw(x, dw[0])
        return _run_conv(y, pw)


def _dw(x, prep):
    """Depthwise conv over the logical channels; the output keeps the input's padded pixel stride with a zero tail."""
    n, c, h, w = x.shape
    ho, wo = prep.out_hw(h, w)
    ca = x.stride(3)
    y = ops.new_act(n, c, ho, wo, x.dtype, x.device, c_alloc=ca, zero=ca != c)
    return ops.conv2d(x, prep, out=y)


class _DWConv(PrepMixin, nn.Module):
    def __init__(self, dw_channels, out_channels, stride=1, **kwargs):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(dw_channels, out_channels, 3, stride, 1, groups=dw_channels, bias=False),
                                  nn.BatchNorm2d(out_channels), nn.ReLU(True))

    def _build_prep(self, device):
        return _conv_prep(self.conv[0], self.conv[1], ACT_RELU, device)

    def forward(self, x):
        _no_train(self)
        x = ops.as_act(x)
        return _dw(x, self.prep(x.device)[0])


class LinearBottleneck(PrepMixin, nn.Module):
    def __init__(self, in_channels, out_channels, t=6, stride=2, **kwargs):
        super().__init__()
        self.use_shortcut = stride == 1 and in_channels == out_channels
        self.block = nn.Sequential(_ConvBNReLU(in_channels, in_channels * t, 1), _DWConv(in_channels * t, in_channels * t, stride),
                                   nn.Conv2d(in_channels * t, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels))

    def _build_prep(self, device):
        return _conv_prep(self.block[2], self.block[3], ACT_NONE, device)

    def forward(self, x, out=None):
        _no_train(self)
        x = ops.as_act(x)
        y = self.block[0](x)
        y = self.block[1](y)
        return _run_conv(y, self.prep(x.device), residual=x if self.use_shortcut else None, out=out)


class PyramidPooling(nn.Module):
    def __init__(self, in_channels, out_channels, **kwargs):
        super().__init__()
        inter_channels = int(in_channels / 4)
        self.conv1 = _ConvBNReLU(in_channels, inter_channels, 1, **kwargs)
        self.conv2 = _ConvBNReLU(in_channels, inter_channels, 1, **kwargs)
        self.conv3 = _ConvBNReLU(in_channels, inter_channels, 1, **kwargs)
        self.conv4 = _ConvBNReLU(in_channels, inter_channels, 1, **kwargs)
        self.out = _ConvBNReLU(in_channels * 2, out_channels, 1)

    def forward(self, x, cat=None):
        """`cat`: optional pre-allocated (N, 2*C, H, W) buffer whose first C channels already hold x."""
        _no_train(self)
        x = ops.as_act(x)
        n, c, h, w = x.shape
        if cat is None:
            cat = ops.new_act(n, 2 * c, h, w, x.dtype, x.device)
            ops.affine_act(x, None, None, None, ACT_NONE, out=cat[:, :c])
        ci = c // 4
        for i, (conv, size) in enumerate(((self.conv1, 1), (self.conv2, 2), (self.conv3, 3), (self.conv4, 6))):
            f = conv(ops.adaptive_avgpool(x, size))
            ops.bilinear(f[:, :ci] if f.shape[1] != ci else f, h, w, True, out=cat[:, c + i * ci:c + (i + 1) * ci])
        return self.out(cat)


class LearningToDownsample(nn.Module):
    def __init__(self, dw_channels1=32, dw_channels2=48, out_channels=64, **kwargs):
        super().__init__()
        self.conv = _ConvBNReLU(3, dw_channels1, 3, 2)
        self.dsconv1 = _DSConv(dw_channels1, dw_channels2, 2)
        self.dsconv2 = _DSConv(dw_channels2, out_channels, 2)

    def forward(self, x):
        return self.dsconv2(self.dsconv1(self.conv(x)))


class GlobalFeatureExtractor(nn.Module):
    def __init__(self, in_channels=64, block_channels=(64, 96, 128), out_channels=128, t=6, num_blocks=(3, 3, 3), **kwargs):
        super().__init__()
        self.bottleneck1 = self._make_layer(LinearBottleneck, in_channels, block_channels[0], num_blocks[0], t, 2)
        self.bottleneck2 = self._make_layer(LinearBottleneck, block_channels[0], block_channels[1], num_blocks[1], t, 2)
        self.bottleneck3 = self._make_layer(LinearBottleneck, block_channels[1], block_channels[2], num_blocks[2], t, 1)
        self.ppm = PyramidPooling(block_channels[2], out_channels)

    def _make_layer(self, block, inplanes, planes, blocks, t=6, stride=1):
        layers = [block(inplanes, planes, t, stride)]
        for i in range(1, blocks):
            layers.append(block(planes, planes, t, 1))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.bottleneck2(self.bottleneck1(x))
        blocks = list(self.bottleneck3)
        for blk in blocks[:-1]:
            x = blk(x)
        # the last bottleneck writes straight into the pyramid-pooling concat buffer
        last = blocks[-1]
        xa = ops.as_act(x)
        n, _, h, w = xa.shape
        c = last.block[3].num_features
        cat = ops.new_act(n, 2 * c, h, w, xa.dtype, xa.device)
        y = last(xa, out=cat[:, :c])
        return self.ppm(y, cat=cat)


class FeatureFusionModule(PrepMixin, nn.Module):
    def __init__(self, highter_in_channels, lower_in_channels, out_channels, scale_factor=4, **kwargs):
        super().__init__()
        self.scale_factor = scale_factor
        self.dwconv = _DWConv(lower_in_channels, out_channels, 1)
        self.conv_lower_res = nn.Sequential(nn.Conv2d(out_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        self.conv_higher_res = nn.Sequential(nn.Conv2d(highter_in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        self.relu = nn.ReLU(True)

    def _build_prep(self, device):
        return (_conv_prep(self.conv_lower_res[0], self.conv_lower_res[1], ACT_NONE, device),
                _conv_prep(self.conv_higher_res[0], self.conv_higher_res[1], ACT_RELU, device))

    def forward(self, higher_res_feature, lower_res_feature):
        _no_train(self)
        hi = ops.as_act(higher_res_feature)
        lo = ops.as_act(lower_res_feature)
        lower_p, higher_p = self.prep(hi.device)
        _, _, h, w = hi.shape
        lo = ops.bilinear(lo, h, w, True)
        lo = self.dwconv(lo)
        lo = _run_conv(lo, lower_p)
        return _run_conv(hi, higher_p, residual=lo)     # relu(higher + lower)


class Classifer(PrepMixin, nn.Module):
    def __init__(self, dw_channels, num_classes, stride=1, **kwargs):
        super().__init__()
        self.dsconv1 = _DSConv(dw_channels, dw_channels, stride)
        self.dsconv2 = _DSConv(dw_channels, dw_channels, stride)
        self.conv = nn.Sequential(nn.Dropout(0.1), nn.Conv2d(dw_channels, num_classes, 1))

    def _build_prep(self, device):
        conv = self.conv[1]
        return (ops.ConvPrep(conv, device=device), ops.ConvPrep(conv, device=device, cout_pad=32)), conv.out_channels

    def forward(self, x):
        _no_train(self)        # eval: Dropout is the identity
        x = self.dsconv2(self.dsconv1(ops.as_act(x)))
        (plain, padded), classes = self.prep(x.device)
        n, _, h, w = x.shape
        scores = ops.new_act(n, classes, h, w, x.dtype, x.device, c_alloc=32)
        if x.dtype == torch.bfloat16:
            ops.conv2d(x, padded, out=ops.widen(scores, 32))
        else:
            ops.conv2d(x, plain, out=scores)
        return scores


class FastSCNN(nn.Module):
    def __init__(self, classes, aux=False, **kwargs):
        super().__init__()
        self.aux = aux
        self.learning_to_downsample = LearningToDownsample(32, 48, 64)
        self.global_feature_extractor = GlobalFeatureExtractor(64, [64, 96, 128], 128, 6, [3, 3, 3])
        self.feature_fusion = FeatureFusionModule(64, 128, 128)
        self.classifier = Classifer(128, classes)
        if self.aux:
            self.auxlayer = nn.Sequential(nn.Conv2d(64, 32, 3, padding=1, bias=False), nn.BatchNorm2d(32), nn.ReLU(True),
                                          nn.Dropout(0.1), nn.Conv2d(32, classes, 1))

    def _scores(self, x):
        ops.require_cuda(x, "FastSCNN")
        _no_train(self)
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        higher = self.learning_to_downsample(x)
        y = self.global_feature_extractor(higher)
        y = self.feature_fusion(higher, y)
        return self.classifier(y), x.shape[2:]

    def fused_loss(self, input, target, criterion):
        """criterion(self(input), target) (train.py:351-352) with the bilinear head (FastSCNN.py:233, align_corners=True),
        CrossEntropyLoss2d and both their backward passes as ONE launch (esn_bilinear_ce): the 2.5 GB of fp32 logits of a
        16 x 1024 x 2048 batch and their gradient are never written.  esn.graph.GraphedTrainStep calls this; other criteria and
        eval mode take the two-module form."""
        if self.aux:
            return criterion(self(input), target)
        from esn import train as T
        from model._fastscnn_train import fastscnn_train_forward
        return T.fused_bilinear_loss(self, fastscnn_train_forward, input, target, criterion, self.classifier.conv[1].out_channels)

    def forward(self, x):
        if self.training:
            # batch-statistics BatchNorm + Dropout + recorded backward (esn/train.py); one autograd node for the net
            from esn import train as T
            from model._fastscnn_train import fastscnn_train_forward
            if self.aux:
                raise NotImplementedError("FastSCNN(aux=True): the auxiliary head is never used by the reference's forward")
            return T.run_network(self, lambda inp: fastscnn_train_forward(self, inp), x)
        scores, (h, w) = self._scores(x)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        return ops.head_bilinear(scores, scores.shape[1], h, w, True, False, ldt, align_corners=True)[0]

    @torch.no_grad()
    def predict_mask(self, x, with_logits=False):
        scores, (h, w) = self._scores(x)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        logits, mask = ops.head_bilinear(scores, scores.shape[1], h, w, with_logits, True, ldt, align_corners=True)
        return (logits, mask) if with_logits else mask
