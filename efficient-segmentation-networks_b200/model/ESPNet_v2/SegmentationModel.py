"""ESPNetv2 segmentation network on B200 kernels -- drop-in for the reference's
model/ESPNet_v2/SegmentationModel.py.

Same class name, constructor signature and attribute names (identical ``state_dict`` keys) as
/root/reference/model/ESPNet_v2/SegmentationModel.py:20-77.  The decoder's concats are channel slices of
pre-allocated NHWC buffers (147 -> 192 and 51 -> 64 channels, zero tails against zero weight columns, so the
1x1 projections run on the tcgen05 kernel); the x2 bilinear ladders (align_corners=True) write straight into
those slices and the last one is fused with the NCHW logits store / argmax.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

from esn import ops
from esn._lib import ACT_NONE
from esn.prep import PrepMixin
from model.ESPNet_v2.Model import EESPNet, EESP
from model.ESPNet_v2.cnn_utils import *          # noqa: F401,F403
from model.ESPNet_v2.cnn_utils import _no_train, fold

__all__ = ["EESPNet_Seg"]


def _tc(c):
    for v in (16, 32, 64):
        if c <= v:
            return v
    return (c + 63) // 64 * 64


class EESPNet_Seg(PrepMixin, nn.Module):
    def __init__(self, classes=19, s=2, pretrained=None, gpus=1):
        super().__init__()
        classificationNet = EESPNet(classes=1000, s=s)
        if pretrained:
            if not os.path.isfile(pretrained):
                print('Weight file does not exist. Training without pre-trained weights')
            else:
                sd = torch.load(pretrained, map_location="cpu")
                classificationNet.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in sd.items()})
                print('Model initialized with pretrained weights')
        self.net = classificationNet
        del self.net.classifier
        del self.net.level5
        del self.net.level5_0
        p = 0.1 if s <= 0.5 else 0.2
        c4 = self.net.level4[-1].module_act.num_parameters
        c3 = self.net.level3[-1].module_act.num_parameters
        self.proj_L4_C = CBR(c4, c3, 1, 1)
        pspSize = 2 * c3
        self.pspMod = nn.Sequential(EESP(pspSize, pspSize // 2, stride=1, k=4, r_lim=7),
                                    PSPModule(pspSize // 2, pspSize // 2))
        self.project_l3 = nn.Sequential(nn.Dropout2d(p=p), C(pspSize // 2, classes, 1, 1))
        self.act_l3 = BR(classes)
        self.project_l2 = CBR(self.net.level2_0.act.num_parameters + classes, classes, 1, 1)
        self.project_l1 = nn.Sequential(nn.Dropout2d(p=p), C(self.net.level1.act.num_parameters + classes, classes, 1, 1))

    def hierarchicalUpsample(self, x, factor=3):
        x = ops.as_act(x)
        for i in range(factor):
            x = ops.bilinear(x, 2 * x.shape[2], 2 * x.shape[3], True)
        return x

    def _build_prep(self, device):
        c2in, c1in = self.project_l2.conv.in_channels, self.project_l1[1].conv.in_channels
        l3 = fold(self.project_l3[1].conv, self.act_l3.bn, self.act_l3.act, device)
        l2 = self.project_l2
        l1 = self.project_l1[1]
        return dict(l3=l3, l3p=fold(self.project_l3[1].conv, self.act_l3.bn, self.act_l3.act, device, cout_pad=32),
                    l2=fold(l2.conv, l2.bn, l2.act, device), l2p=fold(l2.conv, l2.bn, l2.act, device, cin_pad=_tc(c2in), cout_pad=32),
                    l1=fold(l1.conv, None, None, device), l1p=fold(l1.conv, None, None, device, cin_pad=_tc(c1in), cout_pad=32))

    def _scores(self, input):
        ops.require_cuda(input, "EESPNet_Seg")
        _no_train(self)        # eval: Dropout2d is the identity
        n, _, h, w = input.shape
        if (h | w) & 15:
            raise ValueError("EESPNet_Seg: input height and width must be multiples of 16, got %dx%d" % (h, w))
        dt, dev = ops.compute_dtype(input), input.device
        P = self.prep(dev)
        tc = dt == torch.bfloat16
        classes = P["l3"].cout
        c1 = self.net.level1.act.num_parameters
        c2 = self.net.level2_0.act.num_parameters
        c3 = self.net.level3[-1].module_act.num_parameters
        w1, w2 = _tc(c1 + classes), _tc(c2 + classes)
        cat1 = ops.new_act(n, c1 + classes, h // 2, w // 2, dt, dev, c_alloc=w1, zero=True)
        cat2 = ops.new_act(n, c2 + classes, h // 4, w // 4, dt, dev, c_alloc=w2, zero=True)
        cat3 = ops.new_act(n, 2 * c3, h // 8, w // 8, dt, dev)
        out_l1, out_l2, out_l3, out_l4 = self.net(input, seg=True,
                                                  bufs=dict(l1=cat1[:, :c1], l2=cat2[:, :c2], l3=cat3[:, :c3]))
        l4p = self.proj_L4_C(out_l4)
        ops.bilinear(l4p, h // 8, w // 8, True, out=cat3[:, c3:])
        psp_cat = ops.new_act(n, 5 * c3, h // 8, w // 8, dt, dev)
        m3 = self.pspMod[0](cat3, out=psp_cat[:, :c3])
        m3 = self.pspMod[1](m3, cat=psp_cat)
        s3 = ops.new_act(n, classes, h // 8, w // 8, dt, dev, c_alloc=32)
        if tc:
            ops.conv2d(m3, P["l3p"], out=ops.widen(s3, 32))
        else:
            ops.conv2d(m3, P["l3"], out=s3)
        ops.bilinear(s3, h // 4, w // 4, True, out=cat2[:, c2:c2 + classes])
        m2 = ops.new_act(n, classes, h // 4, w // 4, dt, dev, c_alloc=32)
        if tc:
            ops.conv2d(ops.widen(cat2, w2), P["l2p"], out=ops.widen(m2, 32))
        else:
            ops.conv2d(cat2, P["l2"], out=m2)
        ops.bilinear(m2, h // 2, w // 2, True, out=cat1[:, c1:c1 + classes])
        m1 = ops.new_act(n, classes, h // 2, w // 2, dt, dev, c_alloc=32)
        if tc:
            ops.conv2d(ops.widen(cat1, w1), P["l1p"], out=ops.widen(m1, 32))
        else:
            ops.conv2d(cat1, P["l1"], out=m1)
        return m1, (h, w)

    # No fused_loss here: the scores are at 1/2 resolution, and esn_bilinear_ce maps 8 lanes to the rows of a source cell -- at an
    # up-sampling factor of 2 a quarter of them work and the fused close takes as long as the launches it replaces (measured:
    # 160.8 against 161.4 ms per step).  It pays from a factor of 4 (DABNet, CGNet, Fast-SCNN: 8).

    def forward(self, input):
        if self.training:
            # batch-statistics BatchNorm + Dropout2d + recorded backward (esn/train.py); one autograd node for the net
            from esn import train as T
            from model.ESPNet_v2._train import espnetv2_train_forward
            return T.run_network(self, lambda inp: espnetv2_train_forward(self, inp), input)
        scores, (h, w) = self._scores(input)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        return ops.head_bilinear(scores, scores.shape[1], h, w, True, False, ldt, align_corners=True)[0]

    @torch.no_grad()
    def predict_mask(self, input, with_logits=False):
        scores, (h, w) = self._scores(input)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        logits, mask = ops.head_bilinear(scores, scores.shape[1], h, w, with_logits, True, ldt, align_corners=True)
        return (logits, mask) if with_logits else mask
