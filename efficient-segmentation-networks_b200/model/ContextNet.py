"""ContextNet on B200 kernels -- drop-in for the reference's model/ContextNet.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/ContextNet.py:16-226.  The block set is Fast-SCNN's under other names (Custom_Conv ==
_ConvBNReLU, DepthSepConv == _DSConv, DepthConv == _DWConv, LinearBottleneck, FeatureFusionModule, Classifer), so
each class here derives from its model/FastSCNN.py twin and inherits that launch plan: every conv carries its
BatchNorm (+ReLU, +shortcut) in the epilogue, the 1x1 convs run on the tcgen05 kernel over channel-padded buffers,
the depthwise 3x3 convs on the vectorised NHWC stencil, and the final bilinear (align_corners=True) is fused with
the NCHW logits store / argmax.

What ContextNet adds is the quarter-resolution deep branch: the reference resizes the fp32 NCHW image with
``F.interpolate(scale_factor=0.25, mode='bilinear', align_corners=True)`` (ContextNet.py:207).  Here the image is
viewed as N*3 single-channel planes -- for C = 1 the NCHW and NHWC layouts coincide -- and resized by the NHWC
bilinear kernel straight into an fp32 NCHW buffer, which is exactly what the stem kernel of the deep branch reads.
"""
import torch
import torch.nn as nn

from esn import ops
from model import FastSCNN as _fs

__all__ = ["ContextNet"]


class Custom_Conv(_fs._ConvBNReLU):
    pass


class DepthSepConv(_fs._DSConv):
    pass


class DepthConv(_fs._DWConv):
    pass


class LinearBottleneck(_fs.LinearBottleneck):
    def __init__(self, in_channels, out_channels, t=6, stride=2, **kwargs):
        nn.Module.__init__(self)
        self.use_shortcut = stride == 1 and in_channels == out_channels
        self.block = nn.Sequential(Custom_Conv(in_channels, in_channels * t, 1), DepthConv(in_channels * t, in_channels * t, stride),
                                   nn.Conv2d(in_channels * t, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels))


class Shallow_net(nn.Module):
    """Full-resolution branch: 3x3/s2 conv, then three depthwise-separable convs (strides 2, 2, 1) -> 1/8 scale."""

    def __init__(self, dw_channels1=32, dw_channels2=64, out_channels=128, **kwargs):
        super().__init__()
        self.conv = Custom_Conv(3, dw_channels1, 3, 2)
        self.dsconv1 = DepthSepConv(dw_channels1, dw_channels2, 2)
        self.dsconv2 = DepthSepConv(dw_channels2, out_channels, 2)
        self.dsconv3 = DepthSepConv(out_channels, out_channels, 1)

    def forward(self, x):
        return self.dsconv3(self.dsconv2(self.dsconv1(self.conv(x))))


class Deep_net(nn.Module):
    """Quarter-resolution branch: 3x3/s2 conv and six LinearBottleneck stages -> 1/32 scale."""

    def __init__(self, in_channels, block_channels, t, num_blocks, **kwargs):
        super().__init__()
        self.block_channels = block_channels
        self.t = t
        self.num_blocks = num_blocks
        self.conv_ = Custom_Conv(3, in_channels, 3, 2)
        strides = (1, 1, 2, 2, 1, 1)
        cin = in_channels
        for i, (cout, nb, ti, s) in enumerate(zip(block_channels, num_blocks, t, strides)):
            setattr(self, "bottleneck%d" % (i + 1), self._layer(LinearBottleneck, cin, cout, nb, ti, s))
            cin = cout

    def _layer(self, block, in_channels, out_channels, blocks, t, stride):
        layers = [block(in_channels, out_channels, t, stride)]
        for _ in range(1, blocks):
            layers.append(block(out_channels, out_channels, t, 1))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.conv_(x)
        for i in range(1, 7):
            x = getattr(self, "bottleneck%d" % i)(x)
        return x


class FeatureFusionModule(_fs.FeatureFusionModule):
    def __init__(self, highter_in_channels, lower_in_channels, out_channels, scale_factor=4, **kwargs):
        nn.Module.__init__(self)
        self.scale_factor = scale_factor
        self.dwconv = DepthConv(lower_in_channels, out_channels, 1)
        self.conv_lower_res = nn.Sequential(nn.Conv2d(out_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        self.conv_higher_res = nn.Sequential(nn.Conv2d(highter_in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        self.relu = nn.ReLU(True)


class Classifer(_fs.Classifer):
    def __init__(self, dw_channels, num_classes, stride=1, **kwargs):
        nn.Module.__init__(self)
        self.dsconv1 = DepthSepConv(dw_channels, dw_channels, stride)
        self.dsconv2 = DepthSepConv(dw_channels, dw_channels, stride)
        self.conv = nn.Sequential(nn.Dropout(0.1), nn.Conv2d(dw_channels, num_classes, 1))


def quarter_scale_image(x):
    """``F.interpolate(x, scale_factor=0.25, mode='bilinear', align_corners=True)`` of the fp32 NCHW image
    (ContextNet.py:207) -> fp32 NCHW (N, C, H//4, W//4), one launch of the NHWC bilinear kernel over N*C planes."""
    n, c, h, w = x.shape
    ho, wo = h // 4, w // 4          # floor(size * scale_factor), as torch computes the output size
    if ho < 1 or wo < 1:
        raise ValueError("ContextNet: input %dx%d is too small for the quarter-scale branch" % (h, w))
    y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device)
    planes_in = x.view(n * c, h, w, 1).permute(0, 3, 1, 2)        # (N*C, 1, H, W): NHWC-strided view, no copy
    planes_out = y.view(n * c, ho, wo, 1).permute(0, 3, 1, 2)
    ops.bilinear(planes_in, ho, wo, True, out=planes_out)
    return y


class ContextNet(nn.Module):
    def __init__(self, classes, aux=False, **kwargs):
        super().__init__()
        self.aux = aux
        self.spatial_detail = Shallow_net(32, 64, 128)
        self.context_feature_extractor = Deep_net(32, [32, 32, 48, 64, 96, 128], [1, 6, 6, 6, 6, 6], [1, 1, 3, 3, 2, 2])
        self.feature_fusion = FeatureFusionModule(128, 128, 128)
        self.classifier = Classifer(128, classes)
        if self.aux:
            self.auxlayer = nn.Sequential(nn.Conv2d(128, 32, 3, padding=1, bias=False), nn.BatchNorm2d(32), nn.ReLU(True),
                                          nn.Dropout(0.1), nn.Conv2d(32, classes, 1))

    def _scores(self, x):
        ops.require_cuda(x, "ContextNet")
        if self.training:
            raise NotImplementedError("ContextNet: training-mode kernels are not wired for this model; call .eval(). "
                                      "There is no eager-PyTorch fallback.")
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        higher = self.spatial_detail(x)
        lower = self.context_feature_extractor(quarter_scale_image(x))
        y = self.feature_fusion(higher, lower)
        return self.classifier(y), x.shape[2:]        # the auxiliary head never reaches the output (ContextNet.py:217-224)

    def forward(self, x):
        scores, (h, w) = self._scores(x)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        return ops.head_bilinear(scores, scores.shape[1], h, w, True, False, ldt, align_corners=True)[0]

    @torch.no_grad()
    def predict_mask(self, x, with_logits=False):
        """uint8 (N,H,W) argmax mask computed inside the head kernel (replaces test.py:79-82)."""
        scores, (h, w) = self._scores(x)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        logits, mask = ops.head_bilinear(scores, scores.shape[1], h, w, with_logits, True, ldt, align_corners=True)
        return (logits, mask) if with_logits else mask
