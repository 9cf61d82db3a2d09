"""ERFNet on B200 kernels -- drop-in for the reference's model/ERFNet.py.

Same class names, constructor signatures, attribute names (hence identical
``state_dict`` keys) as /root/reference/model/ERFNet.py:16-156; the forwards
issue C-ABI kernel calls (esn.ops) instead of ATen ops:

* non_bottleneck_1d (ERFNet.py:30-65): four dense factorized convs, each one
  tcgen05 implicit-GEMM launch with bias / folded BN / ReLU / residual in the
  epilogue  (4 launches instead of 11 ATen kernels)
* DownsamplerBlock (ERFNet.py:16-27): strided conv written straight into its
  slice of the concat buffer with the BN slice + ReLU folded in; max-pool
  branch likewise (2 launches instead of 5)
* UpsamplerBlock (ERFNet.py:103-112): transposed conv as 4 output-parity phases
* Decoder.output_conv (ERFNet.py:128): 2x2 transposed conv fused with the NCHW
  logits store and/or the uint8 argmax mask.
"""
import torch
import torch.nn as nn

from esn import ops
from esn._lib import ACT_NONE, ACT_RELU
from esn.prep import PrepMixin

__all__ = ["ERFNet"]


def _no_train(mod):
    if mod.training:
        raise NotImplementedError(
            "%s: training-mode kernels (batch-stat BN, backward) are not built yet for this model; "
            "call .eval(). There is no eager-PyTorch fallback." % type(mod).__name__)


class DownsamplerBlock(PrepMixin, nn.Module):
    def __init__(self, ninput, noutput):
        super().__init__()
        self.conv = nn.Conv2d(ninput, noutput - ninput, (3, 3), stride=2, padding=1, bias=True)
        self.pool = nn.MaxPool2d(2, stride=2)
        self.bn = nn.BatchNorm2d(noutput, eps=1e-3)

    def _build_prep(self, device):
        scale, shift = ops.bn_affine(self.bn, device)
        nc = self.conv.out_channels
        conv = ops.ConvPrep(self.conv, scale[:nc], shift[:nc], ACT_RELU, device=device)
        ps, pb = scale[nc:].contiguous(), shift[nc:].contiguous()
        return conv, ps, pb, torch.cat([conv.scale, ps]).contiguous(), torch.cat([conv.shift, pb]).contiguous()

    def forward(self, input):
        _no_train(self)
        dtype = ops.compute_dtype(input)
        # the network input (3 channels, NCHW fp32) is read in place; anything else is NHWC
        x = input if (input.shape[1] < 8 and input.is_contiguous() and input.dtype == torch.float32
                      and not ops.is_nhwc(input)) else ops.as_act(input, dtype)
        ops.require_cuda(x, "DownsamplerBlock")
        conv, pscale, pshift, fscale, fshift = self.prep(x.device)
        n, c, h, w = x.shape
        nc = conv.cout
        y = ops.new_act(n, nc + c, h // 2, w // 2, dtype, x.device)
        if c == 3 and not ops.is_nhwc(x) and (h | w) % 2 == 0 and (nc + c) % 4 == 0 and nc + c <= 32:
            # network stem: conv + pool + BN + ReLU in one pass over the NCHW image
            return ops.stem_conv3x3s2(x, conv.w_direct, nc, True, y, fscale, fshift, None, ACT_RELU)
        ops.conv2d(x, conv, out=y[:, :nc])
        ops.maxpool2x2(x, y[:, nc:], pscale, pshift, None, ACT_RELU)
        return y


class non_bottleneck_1d(PrepMixin, nn.Module):
    def __init__(self, chann, dropprob, dilated):
        super().__init__()
        self.conv3x1_1 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1, 0), bias=True)
        self.conv1x3_1 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1), bias=True)
        self.bn1 = nn.BatchNorm2d(chann, eps=1e-03)
        self.conv3x1_2 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1 * dilated, 0), bias=True,
                                   dilation=(dilated, 1))
        self.conv1x3_2 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1 * dilated), bias=True,
                                   dilation=(1, dilated))
        self.bn2 = nn.BatchNorm2d(chann, eps=1e-03)
        self.dropout = nn.Dropout2d(dropprob)

    def _build_prep(self, device):
        s1, b1 = ops.bn_affine(self.bn1, device)
        s2, b2 = ops.bn_affine(self.bn2, device)
        return (ops.ConvPrep(self.conv3x1_1, act=ACT_RELU, device=device),
                ops.ConvPrep(self.conv1x3_1, s1, b1, ACT_RELU, device=device),
                ops.ConvPrep(self.conv3x1_2, act=ACT_RELU, device=device),
                ops.ConvPrep(self.conv1x3_2, s2, b2, ACT_RELU, device=device))

    def forward(self, input):
        _no_train(self)   # eval: Dropout2d is the identity (ERFNet.py:62-63)
        x = ops.as_act(input)
        c1, c2, c3, c4 = self.prep(x.device)
        n, c, h, w = x.shape
        # each factorized pair is one kernel when the shape allows (intermediate row kept in shared memory)
        if x.dtype == torch.bfloat16:
            y = ops.new_act(n, c, h, w, x.dtype, x.device)
            if ops.pair_supported(x, c1, c2, y, None):
                ops.conv_pair(x, c1, c2, out=y)
            else:
                y = ops.conv2d(ops.conv2d(x, c1), c2, out=y)
            out = ops.new_act(n, c, h, w, x.dtype, x.device)
            if ops.pair_supported(y, c3, c4, out, x):
                return ops.conv_pair(y, c3, c4, out=out, residual=x)
            return ops.conv2d(ops.conv2d(y, c3), c4, out=out, residual=x)
        y = ops.conv2d(x, c1)
        y = ops.conv2d(y, c2)
        y = ops.conv2d(y, c3)
        return ops.conv2d(y, c4, residual=x)   # relu(bn2(conv) + input)


class Encoder(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16)
        self.layers = nn.ModuleList()
        self.layers.append(DownsamplerBlock(16, 64))
        for x in range(0, 5):
            self.layers.append(non_bottleneck_1d(64, 0.03, 1))
        self.layers.append(DownsamplerBlock(64, 128))
        for x in range(0, 2):
            self.layers.append(non_bottleneck_1d(128, 0.3, 2))
            self.layers.append(non_bottleneck_1d(128, 0.3, 4))
            self.layers.append(non_bottleneck_1d(128, 0.3, 8))
            self.layers.append(non_bottleneck_1d(128, 0.3, 16))
        # only used in encoder-only mode (ERFNet.py:88-89)
        self.output_conv = nn.Conv2d(128, num_classes, 1, stride=1, padding=0, bias=True)

    def forward(self, input, predict=False):
        output = self.initial_block(input)
        for layer in self.layers:
            output = layer(output)
        if predict:
            prep = ops.ConvPrep(self.output_conv, device=output.device)
            output = ops.to_nchw(ops.conv2d(output, prep), torch.float32)
        return output


class UpsamplerBlock(PrepMixin, nn.Module):
    def __init__(self, ninput, noutput):
        super().__init__()
        self.conv = nn.ConvTranspose2d(ninput, noutput, 3, stride=2, padding=1, output_padding=1, bias=True)
        self.bn = nn.BatchNorm2d(noutput, eps=1e-3)

    def _build_prep(self, device):
        s, b = ops.bn_affine(self.bn, device)
        return ops.ConvPrep(self.conv, s, b, ACT_RELU, device=device)

    def forward(self, input):
        _no_train(self)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device))


class Decoder(PrepMixin, nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(UpsamplerBlock(128, 64))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(UpsamplerBlock(64, 16))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.output_conv = nn.ConvTranspose2d(16, num_classes, 2, stride=2, padding=0, output_padding=0, bias=True)

    def _build_prep(self, device):
        w = self.output_conv.weight.detach().to(device=device, dtype=torch.float32)  # (Cin, classes, 2, 2)
        cin, classes = w.shape[0], w.shape[1]
        packed = torch.zeros((2, 2, cin, 32), dtype=torch.float32, device=device)
        packed[:, :, :, :classes] = w.permute(2, 3, 0, 1)
        # mask-only bf16 path: transposed conv + argmax on the tensor cores (esn_head_convt2x2_mask)
        frags = ops.pack_convt2x2_frags(w, classes) if (cin == 16 and classes <= 24) else None
        return (packed.contiguous(), self.output_conv.bias.detach().to(device=device, dtype=torch.float32).contiguous(), classes,
                frags)

    def features(self, input):
        output = input
        for layer in self.layers:
            output = layer(output)
        return output

    def head(self, feat, want_logits=True, want_mask=False):
        w, b, classes, frags = self.prep(feat.device)
        if want_mask and not want_logits and frags is not None:
            mask = ops.head_convt2x2_mask(feat, frags, b, classes)
            if mask is not None:          # bf16 features: the full-resolution scores never leave the registers
                return None, mask
        ldt = torch.bfloat16 if feat.dtype == torch.bfloat16 else torch.float32
        return ops.head_convt2x2(feat, w, b, classes, want_logits, want_mask, ldt)

    def forward(self, input):
        return self.head(self.features(input))[0]


class ERFNet(nn.Module):
    def __init__(self, classes, encoder=None):
        super().__init__()
        if encoder is None:
            self.encoder = Encoder(classes)
        else:
            self.encoder = encoder
        self.decoder = Decoder(classes)

    def forward(self, input, only_encode=False):
        ops.require_cuda(input, "ERFNet")
        if self.training and not only_encode:
            # batch-statistics BatchNorm + Dropout2d + recorded backward (esn/train.py); one autograd node for the net.
            # encoder.output_conv takes no part (ERFNet.py:88-89: predict=False), so it gets no gradient -- as in the reference.
            from esn import train as T
            from model._erfnet_train import erfnet_train_forward
            return T.run_network(self, lambda inp: erfnet_train_forward(self, inp), input)
        if only_encode:
            return self.encoder.forward(input, predict=True)
        output = self.encoder(input)
        return self.decoder.forward(output)

    @torch.no_grad()
    def predict_mask(self, input, with_logits=False):
        """GPU replacement of the reference's CPU argmax (test.py:79-82): uint8 (N,H,W) mask,
        computed inside the head kernel from the fp32 accumulators."""
        feat = self.decoder.features(self.encoder(input))
        logits, mask = self.decoder.head(feat, want_logits=with_logits, want_mask=True)
        return (logits, mask) if with_logits else mask
