"""ESPNetv2 encoder (EESP units) on B200 kernels -- drop-in for the reference's model/ESPNet_v2/Model.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/ESPNet_v2/Model.py:15-287.  Launch plan of one EESP unit: grouped 1x1 reduce (+BN+PReLU)
-> k depthwise dilated 3x3 convs writing channel slices of the concat buffer, each taking the previous slice as
its residual operand (the hierarchical feature fusion) -> BN+PReLU over the concat -> grouped 1x1 expand with BN,
the block input as residual and the unit's PReLU in the epilogue.  A DownSampler writes the average-pooled
input and the strided EESP output into the two halves of one buffer and finishes with the input-reinforcement
1x1 conv, which takes that buffer as its residual and applies the PReLU.
"""
import math

import torch
import torch.nn as nn
from torch.nn import init

from esn import ops
from esn._lib import ACT_NONE, ACT_PRELU
from esn.prep import PrepMixin
from model.ESPNet_v2.cnn_utils import *          # noqa: F401,F403  (same namespace as the reference module)
from model.ESPNet_v2.cnn_utils import _no_train, fold

__all__ = ["EESP", "DownSampler", "EESPNet"]

config_inp_reinf = 3


class ImagePyramid:
    """The RGB image average-pooled (3x3, stride 2, pad 1) level by level, shared by the DownSamplers."""

    def __init__(self, image):
        self.levels = [image]

    def at_height(self, h):
        """First level below full resolution whose height is h (Model.py:136-142 pools until the heights match)."""
        i = 1
        while True:
            if i == len(self.levels):
                x = self.levels[-1]
                n, c, hh, ww = x.shape
                if hh <= 1 and ww <= 1:
                    raise ValueError("input reinforcement: no pyramid level of height %d" % h)
                y = ops.new_act(n, c, (hh - 1) // 2 + 1, (ww - 1) // 2 + 1, torch.float32, x.device, c_alloc=4)
                self.levels.append(ops.avgpool3x3s2(x if (ops.is_nhwc(x) or x.is_contiguous()) else x.contiguous(), y))
            if self.levels[i].shape[2] == h:
                return self.levels[i]
            i += 1


class EESP(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, stride=1, k=4, r_lim=7, down_method='esp'):
        super().__init__()
        self.stride = stride
        n = int(nOut / k)
        n1 = nOut - (k - 1) * n
        assert down_method in ['avg', 'esp'], 'One of these is suppported (avg or esp)'
        assert n == n1, "n(={}) and n1(={}) should be equal for Depth-wise Convolution ".format(n, n1)
        self.proj_1x1 = CBR(nIn, n, 1, stride=1, groups=k)
        map_receptive_ksize = {3: 1, 5: 2, 7: 3, 9: 4, 11: 5, 13: 6, 15: 7, 17: 8}
        self.k_sizes = list()
        for i in range(k):
            ksize = int(3 + 2 * i)
            ksize = ksize if ksize <= r_lim else 3
            self.k_sizes.append(ksize)
        self.k_sizes.sort()
        self.spp_dw = nn.ModuleList()
        for i in range(k):
            d_rate = map_receptive_ksize[self.k_sizes[i]]
            self.spp_dw.append(CDilated(n, n, kSize=3, stride=stride, groups=n, d=d_rate))
        self.conv_1x1_exp = CB(nOut, nOut, 1, 1, groups=k)
        self.br_after_cat = BR(nOut)
        self.module_act = nn.PReLU(nOut)
        self.downAvg = True if down_method == 'avg' else False

    def _build_prep(self, device):
        e = self.conv_1x1_exp
        return dict(proj=self.proj_1x1.prep(device), dw=[m.prep(device) for m in self.spp_dw],
                    br=self.br_after_cat.prep(device), exp_lin=fold(e.conv, e.bn, None, device),
                    exp_act=fold(e.conv, e.bn, self.module_act, device))

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        P = self.prep(x.device)
        o1 = ops.conv2d(x, P["proj"])
        nb, n, h, w = o1.shape
        ho, wo = P["dw"][0].out_hw(h, w)
        k = len(P["dw"])
        cat = ops.new_act(nb, n * k, ho, wo, o1.dtype, o1.device)
        for i, dw in enumerate(P["dw"]):
            ops.conv2d(o1, dw, out=cat[:, i * n:(i + 1) * n], residual=cat[:, (i - 1) * n:i * n] if i else None)
        s, b, a = P["br"]
        ops.affine_act(cat, s, b, a, ACT_PRELU, out=cat)
        if self.stride == 2 and self.downAvg:
            return ops.conv2d(cat, P["exp_lin"], out=out)
        same = cat.shape == x.shape
        return ops.conv2d(cat, P["exp_act"], out=out, residual=x if same else None)


class DownSampler(PrepMixin, nn.Module):
    def __init__(self, nin, nout, k=4, r_lim=9, reinf=True):
        super().__init__()
        nout_new = nout - nin
        self.eesp = EESP(nin, nout_new, stride=2, k=k, r_lim=r_lim, down_method='avg')
        self.avg = nn.AvgPool2d(kernel_size=3, padding=1, stride=2)
        if reinf:
            self.inp_reinf = nn.Sequential(CBR(config_inp_reinf, config_inp_reinf, 3, 1),
                                           CB(config_inp_reinf, nout, 1, 1))
        self.act = nn.PReLU(nout)

    def _build_prep(self, device):
        alpha = self.act.weight.detach().float().to(device).contiguous()
        reinf = None
        if hasattr(self, "inp_reinf"):
            cb = self.inp_reinf[1]
            reinf = fold(cb.conv, cb.bn, self.act, device)
        return alpha, reinf

    def forward(self, input, input2=None, out=None):
        _no_train(self)
        x = ops.as_act(input)
        alpha, reinf = self.prep(x.device)
        nb, cin, h, w = x.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        cout = self.act.num_parameters
        fused_act = input2 is None
        cat = out if (out is not None and fused_act) else ops.new_act(nb, cout, ho, wo, x.dtype, x.device)
        ops.avgpool3x3s2(x, cat[:, :cin], None, None, alpha[:cin].contiguous() if fused_act else None,
                         ACT_PRELU if fused_act else ACT_NONE)
        self.eesp(x, out=cat[:, cin:])
        if fused_act:
            ops.affine_act(cat[:, cin:], None, None, alpha[cin:].contiguous(), ACT_PRELU, out=cat[:, cin:])
            return cat
        pyr = input2 if isinstance(input2, ImagePyramid) else ImagePyramid(input2)
        img = self.inp_reinf[0](pyr.at_height(ho))
        return ops.conv2d(img, reinf, out=out, residual=cat)


class EESPNet(nn.Module):
    def __init__(self, classes=19, s=1):
        super().__init__()
        reps = [0, 3, 7, 3]
        channels = 3
        r_lim = [13, 11, 9, 7, 5]
        K = [4] * len(r_lim)
        base = 32
        config_len = 5
        config = [base] * config_len
        base_s = 0
        for i in range(config_len):
            if i == 0:
                base_s = int(base * s)
                base_s = math.ceil(base_s / K[0]) * K[0]
                config[i] = base if base_s > base else base_s
            else:
                config[i] = base_s * pow(2, i)
        if s <= 1.5:
            config.append(1024)
        elif s in [1.5, 2]:
            config.append(1280)
        else:
            raise ValueError('Configuration not supported')
        self.input_reinforcement = True
        self.level1 = CBR(channels, config[0], 3, 2)
        self.level2_0 = DownSampler(config[0], config[1], k=K[0], r_lim=r_lim[0], reinf=self.input_reinforcement)
        self.level3_0 = DownSampler(config[1], config[2], k=K[1], r_lim=r_lim[1], reinf=self.input_reinforcement)
        self.level3 = nn.ModuleList()
        for i in range(reps[1]):
            self.level3.append(EESP(config[2], config[2], stride=1, k=K[2], r_lim=r_lim[2]))
        self.level4_0 = DownSampler(config[2], config[3], k=K[2], r_lim=r_lim[2], reinf=self.input_reinforcement)
        self.level4 = nn.ModuleList()
        for i in range(reps[2]):
            self.level4.append(EESP(config[3], config[3], stride=1, k=K[3], r_lim=r_lim[3]))
        self.level5_0 = DownSampler(config[3], config[4], k=K[3], r_lim=r_lim[3])
        self.level5 = nn.ModuleList()
        for i in range(reps[3]):
            self.level5.append(EESP(config[4], config[4], stride=1, k=K[4], r_lim=r_lim[4]))
        self.level5.append(CBR(config[4], config[4], 3, 1, groups=config[4]))
        self.level5.append(CBR(config[4], config[5], 1, 1, groups=K[4]))
        self.classifier = nn.Linear(config[5], classes)
        self.init_params()

    def init_params(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    init.constant_(m.bias, 0)

    def forward(self, input, p=0.2, seg=True, bufs=None):
        """`bufs`: optional dict of pre-allocated concat buffers {'l1','l2','l3'} the level outputs are written into."""
        _no_train(self)
        ops.require_cuda(input, "EESPNet")
        if not seg:
            raise NotImplementedError("EESPNet: the ImageNet classification head is outside the segmentation path")
        if input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        bufs = bufs or {}
        out_l1 = self.level1(input, out=bufs.get("l1"))
        pyr = ImagePyramid(input) if self.input_reinforcement else None
        out_l2 = self.level2_0(out_l1, pyr, out=bufs.get("l2"))
        out_l3 = self.level3_0(out_l2, pyr)
        for i, layer in enumerate(self.level3):
            out_l3 = layer(out_l3, out=bufs.get("l3") if i == len(self.level3) - 1 else None)
        out_l4 = self.level4_0(out_l3, pyr)
        for layer in self.level4:
            out_l4 = layer(out_l4)
        return out_l1, out_l2, out_l3, out_l4
