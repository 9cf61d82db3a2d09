"""ENet on B200 kernels -- drop-in for the reference's model/ENet.py.

Same class names, constructor signatures and attribute names as /root/reference/model/ENet.py:14-432,
including the reference's aliasing: one `activation` module instance is shared by every call site of a
block (ENet.py:53-89), so ``state_dict`` carries the same aliased keys.  Inference launch plan:

* InitialBlock (ENet.py:14-44): conv3x3/s2 || MaxPool2d(3,2,1) -> BN -> act in ONE stem kernel pass
* RegularBottleneck (ENet.py:46-100): 1x1 -> {3x3 (dilated) | 5x1,1x5} -> 1x1, every conv with BN+act in
  its epilogue; the last conv also applies "act(main + ext)" (epilogue flag ACT_BEFORE_RESIDUAL):
  3-4 launches instead of 11-14 ATen kernels
* DownsamplingBottleneck (ENet.py:102-197): max-pool with int32 indices; the zero channel padding of the
  main branch is never materialised (the add runs on the first Cin channels only)
* UpsamplingBottleneck (ENet.py:199-272): MaxUnpool2d as a deterministic gather fused with "+ ext, act"
* head: ConvTranspose2d(16, classes, 3, s2) on the tensor cores (Cout padded to 32) + identity head
  kernel writing NCHW logits and/or the argmax mask.
"""
import os

import torch
import torch.nn as nn

from esn import ops
from esn._lib import ACT_NONE, ACT_PRELU, ACT_RELU
from esn.prep import PrepMixin

__all__ = ["ENet"]

_PRE = 1   # ESN_EP_ACT_BEFORE_RESIDUAL


def _no_train(mod):
    if mod.training:
        raise NotImplementedError(
            "%s: training-mode kernels are not built yet for this model; call .eval(). "
            "There is no eager-PyTorch fallback." % type(mod).__name__)


def _act_of(activation, channels, device):
    """(act code, per-channel alpha) of the block's shared activation (nn.PReLU() has ONE alpha)."""
    if isinstance(activation, nn.PReLU):
        a = activation.weight.detach().to(device=device, dtype=torch.float32)
        return ACT_PRELU, a.expand(channels).contiguous() if a.numel() == 1 else a.contiguous()
    return ACT_RELU, None


def _cba(conv, bn, activation, device, pre_residual=False, cout_pad=None):
    """ConvPrep of conv -> eval BN -> activation."""
    s, b = ops.bn_affine(bn, device)
    act, alpha = _act_of(activation, conv.out_channels, device)
    prep = ops.ConvPrep(conv, s, b, act, alpha, device=device, cout_pad=cout_pad)
    prep.ep_flags = _PRE if pre_residual else 0
    return prep


FUSED_HEAD = os.environ.get("ESN_ENET_FUSED_HEAD", "1") != "0"      # transposed conv + argmax as one tensor-core launch
FUSED_BNECK4 = os.environ.get("ESN_ENET_FUSED_BNECK4", "1") != "0"  # the 16-channel bottleneck (4 internal channels) in one launch

class InitialBlock(PrepMixin, nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, padding=0, bias=False, relu=True):
        super().__init__()
        activation = nn.ReLU() if relu else nn.PReLU()
        self.main_branch = nn.Conv2d(in_channels, out_channels - 3, kernel_size=kernel_size, stride=2,
                                     padding=padding, bias=bias)
        self.ext_branch = nn.MaxPool2d(kernel_size, stride=2, padding=padding)
        self.batch_norm = nn.BatchNorm2d(out_channels)
        self.out_prelu = activation

    def _build_prep(self, device):
        s, b = ops.bn_affine(self.batch_norm, device)
        act, alpha = _act_of(self.out_prelu, self.batch_norm.num_features, device)
        conv = ops.ConvPrep(self.main_branch, device=device)
        if self.main_branch.bias is not None:
            b = b.clone()
            b[:conv.cout] += conv.shift * s[:conv.cout]
        return conv, s.contiguous(), b.contiguous(), act, alpha

    def forward(self, input):
        _no_train(self)
        ops.require_cuda(input, "InitialBlock")
        c = self.main_branch
        if not (input.shape[1] == 3 and c.kernel_size == (3, 3) and c.padding == (1, 1) and input.dtype == torch.float32):
            raise NotImplementedError("InitialBlock is built for the 3-channel fp32 network input with a 3x3/s2/p1 conv")
        x = input.contiguous()
        conv, s, b, act, alpha = self.prep(x.device)
        n, _, h, w = x.shape
        y = ops.new_act(n, conv.cout + 3, (h - 1) // 2 + 1, (w - 1) // 2 + 1, ops.compute_dtype(x), x.device)
        return ops.stem_conv3x3s2(x, conv.w_direct, conv.cout, 2, y, s, b, alpha, act)


class RegularBottleneck(PrepMixin, nn.Module):
    def __init__(self, channels, internal_ratio=4, kernel_size=3, padding=0, dilation=1, asymmetric=False,
                 dropout_prob=0., bias=False, relu=True):
        super().__init__()
        internal_channels = channels // internal_ratio
        activation = nn.ReLU() if relu else nn.PReLU()
        self.ext_conv1 = nn.Sequential(
            nn.Conv2d(channels, internal_channels, kernel_size=1, stride=1, bias=bias),
            nn.BatchNorm2d(internal_channels), activation)
        if asymmetric:
            self.ext_conv2 = nn.Sequential(
                nn.Conv2d(internal_channels, internal_channels, kernel_size=(kernel_size, 1), stride=1,
                          padding=(padding, 0), dilation=dilation, bias=bias),
                nn.BatchNorm2d(internal_channels), activation,
                nn.Conv2d(internal_channels, internal_channels, kernel_size=(1, kernel_size), stride=1,
                          padding=(0, padding), dilation=dilation, bias=bias),
                nn.BatchNorm2d(internal_channels), activation)
        else:
            self.ext_conv2 = nn.Sequential(
                nn.Conv2d(internal_channels, internal_channels, kernel_size=kernel_size, stride=1, padding=padding,
                          dilation=dilation, bias=bias),
                nn.BatchNorm2d(internal_channels), activation)
        self.ext_conv3 = nn.Sequential(
            nn.Conv2d(internal_channels, channels, kernel_size=1, stride=1, bias=bias),
            nn.BatchNorm2d(channels), activation)
        self.ext_regu1 = nn.Dropout2d(p=dropout_prob)
        self.out_prelu = activation

    def _build_prep(self, device):
        a = self.out_prelu
        preps = [_cba(self.ext_conv1[0], self.ext_conv1[1], a, device)]
        for i in range(0, len(self.ext_conv2), 3):
            preps.append(_cba(self.ext_conv2[i], self.ext_conv2[i + 1], a, device))
        preps.append(_cba(self.ext_conv3[0], self.ext_conv3[1], a, device, pre_residual=True))
        return preps

    def forward(self, input):
        _no_train(self)   # eval: Dropout2d is the identity
        x = ops.as_act(input)
        preps = self.prep(x.device)
        if (FUSED_BNECK4 and len(preps) == 3 and x.dtype == torch.bfloat16 and x.shape[1] == 16 and preps[0].cout == 4
                and (preps[1].kh, preps[1].kw, preps[1].stride) == (3, 3, 1) and preps[1].dil_h == preps[1].dil_w <= 4
                and (preps[1].pad_h, preps[1].pad_w) == (preps[1].dil_h, preps[1].dil_w)
                and x.stride(3) % 8 == 0 and x.data_ptr() % 16 == 0):
            # four internal channels (regular5_1): the whole block in one launch, the intermediates in shared memory
            n, c, h, w = x.shape
            y = ops.new_act(n, c, h, w, x.dtype, x.device)
            q = ops.L.EsnBneck4()
            q.x, q.y = ops.tdesc(x), ops.tdesc(y)
            q.w1, q.w2, q.w3 = (p.w_direct.data_ptr() for p in preps)
            for i, p in enumerate(preps, 1):
                setattr(q, "scale%d" % i, p.scale.data_ptr())
                setattr(q, "shift%d" % i, p.shift.data_ptr())
                setattr(q, "alpha%d" % i, p.alpha.data_ptr() if p.alpha is not None else None)
            q.dilation, q.act = preps[1].dil_h, preps[0].act
            ops._call(ops.L.lib.esn_bottleneck4, "esn_bottleneck4", (ops.C.byref(q),), 2 * ops._nbytes(x),
                      2 * n * h * w * (64 + 144 + 64))
            return y
        y = x
        for p in preps[:-1]:
            y = ops.conv2d(y, p)
        return ops.conv2d(y, preps[-1], residual=x)      # act(main + act(BN(conv)))


class DownsamplingBottleneck(PrepMixin, nn.Module):
    def __init__(self, in_channels, out_channels, internal_ratio=4, kernel_size=3, padding=0, return_indices=False,
                 dropout_prob=0., bias=False, relu=True):
        super().__init__()
        self.return_indices = return_indices
        internal_channels = in_channels // internal_ratio
        activation = nn.ReLU() if relu else nn.PReLU()
        self.main_max1 = nn.MaxPool2d(kernel_size, stride=2, padding=padding, return_indices=return_indices)
        self.ext_conv1 = nn.Sequential(
            nn.Conv2d(in_channels, internal_channels, kernel_size=2, stride=2, bias=bias),
            nn.BatchNorm2d(internal_channels), activation)
        self.ext_conv2 = nn.Sequential(
            nn.Conv2d(internal_channels, internal_channels, kernel_size=kernel_size, stride=1, padding=padding,
                      bias=bias), nn.BatchNorm2d(internal_channels), activation)
        self.ext_conv3 = nn.Sequential(
            nn.Conv2d(internal_channels, out_channels, kernel_size=1, stride=1, bias=bias),
            nn.BatchNorm2d(out_channels), activation)
        self.ext_regul = nn.Dropout2d(p=dropout_prob)
        self.out_prelu = activation

    def _build_prep(self, device):
        a = self.out_prelu
        act, alpha = _act_of(a, self.ext_conv3[0].out_channels, device)
        return ([_cba(self.ext_conv1[0], self.ext_conv1[1], a, device), _cba(self.ext_conv2[0], self.ext_conv2[1], a, device),
                 _cba(self.ext_conv3[0], self.ext_conv3[1], a, device)], act, alpha)

    def forward(self, x):
        _no_train(self)
        if self.main_max1.kernel_size != 3 or self.main_max1.padding != 1:
            raise NotImplementedError("DownsamplingBottleneck pooling kernel is built for MaxPool2d(3, 2, 1)")
        x = ops.as_act(x)
        preps, act, alpha = self.prep(x.device)
        main, max_indices = ops.maxpool3x3s2_idx(x)
        ext = x
        for p in preps:
            ext = ops.conv2d(ext, p)
        # out = act(cat(main, zeros) + ext): the add only touches the first Cin channels
        cin, cout = main.shape[1], ext.shape[1]
        ops.affine_act(ext[:, :cin], None, None, alpha[:cin].contiguous() if alpha is not None else None, act,
                       out=ext[:, :cin], residual=main)
        ops.affine_act(ext[:, cin:], None, None, alpha[cin:].contiguous() if alpha is not None else None, act,
                       out=ext[:, cin:])
        return ext, max_indices


class UpsamplingBottleneck(PrepMixin, nn.Module):
    def __init__(self, in_channels, out_channels, internal_ratio=4, kernel_size=3, padding=0, dropout_prob=0.,
                 bias=False, relu=True):
        super().__init__()
        internal_channels = in_channels // internal_ratio
        activation = nn.ReLU() if relu else nn.PReLU()
        self.main_conv1 = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=bias),
                                        nn.BatchNorm2d(out_channels))
        self.main_unpool1 = nn.MaxUnpool2d(kernel_size=2)
        self.ext_conv1 = nn.Sequential(nn.Conv2d(in_channels, internal_channels, kernel_size=1, bias=bias),
                                       nn.BatchNorm2d(internal_channels), activation)
        self.ext_conv2 = nn.Sequential(
            nn.ConvTranspose2d(internal_channels, internal_channels, kernel_size=kernel_size, stride=2, padding=padding,
                               output_padding=1, bias=bias), nn.BatchNorm2d(internal_channels), activation)
        self.ext_conv3 = nn.Sequential(nn.Conv2d(internal_channels, out_channels, kernel_size=1, bias=bias),
                                       nn.BatchNorm2d(out_channels), activation)
        self.ext_regul = nn.Dropout2d(p=dropout_prob)
        self.out_prelu = activation

    def _build_prep(self, device):
        a = self.out_prelu
        s, b = ops.bn_affine(self.main_conv1[1], device)
        main = ops.ConvPrep(self.main_conv1[0], s, b, ACT_NONE, device=device)
        act, alpha = _act_of(a, self.ext_conv3[0].out_channels, device)
        return (main, [_cba(self.ext_conv1[0], self.ext_conv1[1], a, device), _cba(self.ext_conv2[0], self.ext_conv2[1], a, device),
                       _cba(self.ext_conv3[0], self.ext_conv3[1], a, device)], act, alpha)

    def forward(self, x, max_indices):
        _no_train(self)
        x = ops.as_act(x)
        main_p, preps, act, alpha = self.prep(x.device)
        main = ops.conv2d(x, main_p)
        ext = x
        for p in preps:
            ext = ops.conv2d(ext, p)
        return ops.max_unpool2x2(main, max_indices, ext=ext, act=act, alpha=alpha)    # act(unpool(main) + ext)


class ENet(PrepMixin, nn.Module):
    def __init__(self, classes, encoder_relu=False, decoder_relu=True):
        super().__init__()
        self.name = 'BaseLine_ENet_trans'
        self.initial_block = InitialBlock(3, 16, kernel_size=3, padding=1, relu=encoder_relu)
        self.downsample1_0 = DownsamplingBottleneck(16, 64, padding=1, return_indices=True, dropout_prob=0.01, relu=encoder_relu)
        self.regular1_1 = RegularBottleneck(64, padding=1, dropout_prob=0.01, relu=encoder_relu)
        self.regular1_2 = RegularBottleneck(64, padding=1, dropout_prob=0.01, relu=encoder_relu)
        self.regular1_3 = RegularBottleneck(64, padding=1, dropout_prob=0.01, relu=encoder_relu)
        self.regular1_4 = RegularBottleneck(64, padding=1, dropout_prob=0.01, relu=encoder_relu)
        self.downsample2_0 = DownsamplingBottleneck(64, 128, padding=1, return_indices=True, dropout_prob=0.1, relu=encoder_relu)
        self.regular2_1 = RegularBottleneck(128, padding=1, dropout_prob=0.1, relu=encoder_relu)
        self.dilated2_2 = RegularBottleneck(128, dilation=2, padding=2, dropout_prob=0.1, relu=encoder_relu)
        self.asymmetric2_3 = RegularBottleneck(128, kernel_size=5, padding=2, asymmetric=True, dropout_prob=0.1, relu=encoder_relu)
        self.dilated2_4 = RegularBottleneck(128, dilation=4, padding=4, dropout_prob=0.1, relu=encoder_relu)
        self.regular2_5 = RegularBottleneck(128, padding=1, dropout_prob=0.1, relu=encoder_relu)
        self.dilated2_6 = RegularBottleneck(128, dilation=8, padding=8, dropout_prob=0.1, relu=encoder_relu)
        self.asymmetric2_7 = RegularBottleneck(128, kernel_size=5, asymmetric=True, padding=2, dropout_prob=0.1, relu=encoder_relu)
        self.dilated2_8 = RegularBottleneck(128, dilation=16, padding=16, dropout_prob=0.1, relu=encoder_relu)
        self.regular3_0 = RegularBottleneck(128, padding=1, dropout_prob=0.1, relu=encoder_relu)
        self.dilated3_1 = RegularBottleneck(128, dilation=2, padding=2, dropout_prob=0.1, relu=encoder_relu)
        self.asymmetric3_2 = RegularBottleneck(128, kernel_size=5, padding=2, asymmetric=True, dropout_prob=0.1, relu=encoder_relu)
        self.dilated3_3 = RegularBottleneck(128, dilation=4, padding=4, dropout_prob=0.1, relu=encoder_relu)
        self.regular3_4 = RegularBottleneck(128, padding=1, dropout_prob=0.1, relu=encoder_relu)
        self.dilated3_5 = RegularBottleneck(128, dilation=8, padding=8, dropout_prob=0.1, relu=encoder_relu)
        self.asymmetric3_6 = RegularBottleneck(128, kernel_size=5, asymmetric=True, padding=2, dropout_prob=0.1, relu=encoder_relu)
        self.dilated3_7 = RegularBottleneck(128, dilation=16, padding=16, dropout_prob=0.1, relu=encoder_relu)
        self.upsample4_0 = UpsamplingBottleneck(128, 64, padding=1, dropout_prob=0.1, relu=decoder_relu)
        self.regular4_1 = RegularBottleneck(64, padding=1, dropout_prob=0.1, relu=decoder_relu)
        self.regular4_2 = RegularBottleneck(64, padding=1, dropout_prob=0.1, relu=decoder_relu)
        self.upsample5_0 = UpsamplingBottleneck(64, 16, padding=1, dropout_prob=0.1, relu=decoder_relu)
        self.regular5_1 = RegularBottleneck(16, padding=1, dropout_prob=0.1, relu=decoder_relu)
        self.transposed_conv = nn.ConvTranspose2d(16, classes, kernel_size=3, stride=2, padding=1, output_padding=1, bias=False)
        self.project_layer = nn.Conv2d(128, classes, 1, bias=False)     # unused by forward (ENet.py:384, 414)

    def _build_prep(self, device):
        classes = self.transposed_conv.out_channels
        tc = self.transposed_conv
        # bf16 mask-only path: transposed conv + argmax in one launch on the tensor cores (esn_head_convt3x3s2_mask)
        frags = None
        if (tc.in_channels == 16 and classes <= 24 and tc.kernel_size == (3, 3) and tc.stride == (2, 2) and tc.padding == (1, 1)
                and tc.output_padding == (1, 1) and tc.dilation == (1, 1) and tc.groups == 1):
            frags = (ops.pack_convt3x3s2_frags(tc.weight.to(device), classes),
                     None if tc.bias is None else tc.bias.detach().to(device=device, dtype=torch.float32).contiguous())
        return ops.ConvPrep(tc, device=device, cout_pad=(classes + 7) // 8 * 8), classes, frags

    def _features(self, x):
        ops.require_cuda(x, "ENet")
        if self.training:
            _no_train(self)
        x = self.initial_block(x)
        x, i1 = self.downsample1_0(x)
        for m in (self.regular1_1, self.regular1_2, self.regular1_3, self.regular1_4):
            x = m(x)
        x, i2 = self.downsample2_0(x)
        for m in (self.regular2_1, self.dilated2_2, self.asymmetric2_3, self.dilated2_4, self.regular2_5, self.dilated2_6,
                  self.asymmetric2_7, self.dilated2_8, self.regular3_0, self.dilated3_1, self.asymmetric3_2, self.dilated3_3,
                  self.regular3_4, self.dilated3_5, self.asymmetric3_6, self.dilated3_7):
            x = m(x)
        x = self.upsample4_0(x, i2)
        x = self.regular4_2(self.regular4_1(x))
        x = self.upsample5_0(x, i1)
        return self.regular5_1(x)

    def _scores(self, x):
        x = self._features(x)
        head, classes, _ = self.prep(x.device)
        s = ops.conv2d(x, head)                       # (N, classes padded to 24, H, W) NHWC
        return s[:, :classes], classes

    def forward(self, x):
        if self.training:
            # batch-statistics BatchNorm, Dropout2d and the recorded backward (esn/train.py); one autograd node for the net
            from esn import train as T
            from model._enet_train import enet_train_forward
            return T.run_network(self, lambda inp: enet_train_forward(self, inp), x)
        s, classes = self._scores(x)
        n, _, h, w = s.shape
        ldt = torch.bfloat16 if s.dtype == torch.bfloat16 else torch.float32
        return ops.head_bilinear(s, classes, h, w, True, False, ldt)[0]     # same-size "interpolation": NHWC -> NCHW

    @torch.no_grad()
    def predict_mask(self, x, with_logits=False):
        if not with_logits and FUSED_HEAD:
            f = self._features(x)
            _, classes, frags = self.prep(f.device)
            mask = ops.head_convt3x3s2_mask(f, frags[0], frags[1], classes) if frags is not None else None
            if mask is not None:          # bf16 features: the full-resolution scores never leave the registers
                return mask
            head, classes, _ = self.prep(f.device)
            s = ops.conv2d(f, head)[:, :classes]
        else:
            s, classes = self._scores(x)
        n, _, h, w = s.shape
        ldt = torch.bfloat16 if s.dtype == torch.bfloat16 else torch.float32
        logits, mask = ops.head_bilinear(s, classes, h, w, with_logits, True, ldt)
        return (logits, mask) if with_logits else mask
