"""DABNet on B200 kernels -- drop-in for the reference's model/DABNet.py.

Same class names, constructor signatures and attribute names (identical
``state_dict`` keys) as /root/reference/model/DABNet.py:16-183.  Launch plan:

* Conv (+BNPReLU) (DABNet.py:16-35): one conv launch, BN + PReLU in the epilogue
* DABModule (DABNet.py:51-83): 4 launches instead of 22 ATen kernels --
  bn_relu_1 (affine+PReLU), conv3x3+BNPReLU (tcgen05), the fused depthwise
  asymmetric pair kernel (both branches, 4 BNPReLUs, add, bn_relu_2), and
  conv1x1 + residual (tcgen05)
* DownSamplingBlock (DABNet.py:86-110): conv and max-pool branches write their
  channel slices of the output with the BNPReLU slice folded in
* concats (DABNet.py:166,171,176) are never materialised by a copy: producers
  write into channel slices of one pre-allocated NHWC buffer, BNPReLU runs in place
* classifier + F.interpolate (DABNet.py:179-181): 1x1 conv at 1/8 resolution and
  one fused bilinear kernel that writes NCHW logits and/or the argmax mask.
"""
import torch
import torch.nn as nn

from esn import ops
from esn._lib import ACT_NONE, ACT_PRELU
from esn.prep import PrepMixin, weights_generation

__all__ = ["DABNet"]


def _no_train(mod):
    if mod.training:
        raise NotImplementedError(
            "%s: training-mode kernels (batch-stat BN, backward) are not built yet for this model; "
            "call .eval(). There is no eager-PyTorch fallback." % type(mod).__name__)


def _bnprelu_affine(bnp, device):
    scale, shift = ops.bn_affine(bnp.bn, device)
    alpha = bnp.acti.weight.detach().to(device=device, dtype=torch.float32).contiguous()
    return scale.contiguous(), shift.contiguous(), alpha


class Conv(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, kSize, stride, padding, dilation=(1, 1), groups=1, bn_acti=False, bias=False):
        super().__init__()
        self.bn_acti = bn_acti
        self.conv = nn.Conv2d(nIn, nOut, kernel_size=kSize, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        if self.bn_acti:
            self.bn_prelu = BNPReLU(nOut)

    def _build_prep(self, device):
        if self.bn_acti:
            s, b, a = _bnprelu_affine(self.bn_prelu, device)
            return ops.ConvPrep(self.conv, s, b, ACT_PRELU, a, device=device)
        return ops.ConvPrep(self.conv, device=device)

    def forward(self, input, out=None, residual=None, then=None, out2=None, keep=True):
        """then = (scale, shift, alpha) of a BNPReLU that consumes this conv's output: computed in the same launch
        (esn_conv2d_umma_dual) into out2; keep=False writes only out2.  Returns out, or (out, out2) with `then`."""
        _no_train(self)
        n_in = input.shape[1]
        if then is not None:
            x = ops.as_act(input)
            return ops.conv2d_then_affine(x, self.prep(x.device), then[0], then[1], then[2], ACT_PRELU, out2, out=out,
                                          residual=residual, store_y=keep)
        x = input if (n_in < 8 and input.is_contiguous() and input.dtype == torch.float32
                      and not ops.is_nhwc(input)) else ops.as_act(input)
        prep = self.prep(x.device)
        if out is None and x.dtype == torch.float32 and ops.compute_dtype(x) == torch.bfloat16:
            n, c, h, w = x.shape
            ho, wo = prep.out_hw(h, w)
            out = ops.new_act(n, prep.cout, ho, wo, torch.bfloat16, x.device)
        c = self.conv
        if (n_in == 3 and x is input and not ops.is_nhwc(x) and residual is None and c.kernel_size == (3, 3)
                and c.stride == (2, 2) and c.padding == (1, 1) and c.dilation == (1, 1) and c.groups == 1
                and prep.cout % 4 == 0 and prep.cout <= 32):
            if out is None:
                n, _, h, w = x.shape
                ho, wo = prep.out_hw(h, w)
                out = ops.new_act(n, prep.cout, ho, wo, ops.compute_dtype(x), x.device)
            return ops.stem_conv3x3s2(x, prep.w_direct, prep.cout, False, out, prep.scale, prep.shift, prep.alpha,
                                      prep.act)
        return ops.conv2d(x, prep, out=out, residual=residual)


class BNPReLU(PrepMixin, nn.Module):
    def __init__(self, nIn):
        super().__init__()
        self.bn = nn.BatchNorm2d(nIn, eps=1e-3)
        self.acti = nn.PReLU(nIn)

    def _build_prep(self, device):
        s, b, a = _bnprelu_affine(self, device)
        pad = (-s.numel()) % 8      # zero scale/shift/alpha tail: padded channels of a concat buffer stay 0
        z = torch.zeros(pad, device=device)
        return s, b, a, torch.cat([s, z]), torch.cat([b, z]), torch.cat([a, z])

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        s, b, a, sp, bp, ap = self.prep(x.device)
        c = x.shape[1]
        if c % 4 and out is x and x.stride(3) >= sp.numel():
            # in-place BNPReLU of a channel-padded concat buffer: run the vector path over the padded width
            xw = ops.widen(x, sp.numel())
            ops.affine_act(xw, sp, bp, ap, ACT_PRELU, out=xw)
            return x
        return ops.affine_act(x, s, b, a, ACT_PRELU, out=out)


class DABModule(PrepMixin, nn.Module):
    def __init__(self, nIn, d=1, kSize=3, dkSize=3):
        super().__init__()
        self.bn_relu_1 = BNPReLU(nIn)
        self.conv3x3 = Conv(nIn, nIn // 2, kSize, 1, padding=1, bn_acti=True)
        self.dconv3x1 = Conv(nIn // 2, nIn // 2, (dkSize, 1), 1, padding=(1, 0), groups=nIn // 2, bn_acti=True)
        self.dconv1x3 = Conv(nIn // 2, nIn // 2, (1, dkSize), 1, padding=(0, 1), groups=nIn // 2, bn_acti=True)
        self.ddconv3x1 = Conv(nIn // 2, nIn // 2, (dkSize, 1), 1, padding=(1 * d, 0), dilation=(d, 1),
                              groups=nIn // 2, bn_acti=True)
        self.ddconv1x3 = Conv(nIn // 2, nIn // 2, (1, dkSize), 1, padding=(0, 1 * d), dilation=(1, d),
                              groups=nIn // 2, bn_acti=True)
        self.bn_relu_2 = BNPReLU(nIn // 2)
        self.conv1x1 = Conv(nIn // 2, nIn, 1, 1, padding=0, bn_acti=False)
        self._d = d
        self._dk = dkSize

    def _build_prep(self, device):
        """[27][C] parameter block of esn_dab_dw_pair (layout: csrc/esn_stencil.cu)."""
        if self._dk != 3:
            raise NotImplementedError("fused depthwise pair kernel is built for dkSize=3")
        rows = []
        for m in (self.dconv3x1, self.dconv1x3, self.ddconv3x1, self.ddconv1x3):
            w = m.conv.weight.detach().to(device=device, dtype=torch.float32)  # (C,1,3,1) or (C,1,1,3)
            rows.append(w.reshape(w.shape[0], 3).t())
        for m in (self.dconv3x1, self.dconv1x3, self.ddconv3x1, self.ddconv1x3):
            rows.append(torch.stack(_bnprelu_affine(m.bn_prelu, device)))
        rows.append(torch.stack(_bnprelu_affine(self.bn_relu_2, device)))
        return torch.cat(rows, 0).contiguous()

    def forward(self, input, out=None, pre=None, then=None, keep=True):
        """pre: bn_relu_1(input) when the producer of `input` already computed it.  then = (scale, shift, alpha) of the
        BNPReLU that reads this module's output (the next module's bn_relu_1, or a concat's bn_prelu slice): computed in
        the conv1x1 + residual launch.  With `then` the result is (output, BNPReLU(output)); keep=False stores only the
        second one, into `out`."""
        _no_train(self)
        x = ops.as_act(input)
        prm = self.prep(x.device)
        y = self.bn_relu_1(x) if pre is None else pre
        y = self.conv3x3(y)
        y = ops.dab_dw_pair(y, prm, self._d)
        if then is None:
            return self.conv1x1(y, out=out, residual=x)     # output + input
        if not keep:
            return self.conv1x1(y, residual=x, then=then, out2=out, keep=False)
        n, c, h, w = x.shape
        return self.conv1x1(y, out=out, residual=x, then=then, out2=ops.new_act(n, c, h, w, x.dtype, x.device))


class DownSamplingBlock(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut):
        super().__init__()
        self.nIn = nIn
        self.nOut = nOut
        if self.nIn < self.nOut:
            nConv = nOut - nIn
        else:
            nConv = nOut
        self.conv3x3 = Conv(nIn, nConv, kSize=3, stride=2, padding=1)
        self.max_pool = nn.MaxPool2d(2, stride=2)
        self.bn_prelu = BNPReLU(nOut)

    @staticmethod
    def _tc_channels(c):
        """Smallest channel count >= c the tcgen05 conv takes as Cin (16/32/64 or a multiple of 64)."""
        for v in (16, 32, 64):
            if c <= v:
                return v
        return (c + 63) // 64 * 64

    def _build_prep(self, device):
        s, b, a = _bnprelu_affine(self.bn_prelu, device)
        nc = self.conv3x3.conv.out_channels
        conv = ops.ConvPrep(self.conv3x3.conv, s[:nc], b[:nc], ACT_PRELU, a[:nc], device=device)
        # tensor-core variant for channel-padded concat inputs: Cin zero-extended, Cout rounded up to 8
        # (the extra output channels land where the max-pool branch writes afterwards)
        conv_tc = ops.ConvPrep(self.conv3x3.conv, s[:nc], b[:nc], ACT_PRELU, a[:nc], device=device,
                               cin_pad=self._tc_channels(self.nIn), cout_pad=(nc + 7) // 8 * 8)
        return conv, conv_tc, s[nc:].contiguous(), b[nc:].contiguous(), a[nc:].contiguous()

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        conv, conv_tc, ps, pb, pa = self.prep(x.device)
        n, c, h, w = x.shape
        if out is None:
            out = ops.new_act(n, self.nOut, h // 2, w // 2, x.dtype, x.device)
        nc = conv.cout
        pooled = self.nIn < self.nOut
        if (x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and x.stride(3) >= conv_tc.cin
                and (conv_tc.cout == nc or (pooled and conv_tc.cout <= self.nOut)) and (h | w) % 2 == 0):
            # x is a zero-padded concat buffer: read it at its padded width on the tensor cores
            ops.conv2d(ops.widen(x, conv_tc.cin), conv_tc, out=out[:, :conv_tc.cout])
        else:
            ops.conv2d(x, conv, out=out[:, :nc])
        if pooled:
            ops.maxpool2x2(x, out[:, nc:], ps, pb, pa, ACT_PRELU)
        return out


class InputInjection(nn.Module):
    def __init__(self, ratio):
        super().__init__()
        self.pool = nn.ModuleList()
        for i in range(0, ratio):
            self.pool.append(nn.AvgPool2d(3, stride=2, padding=1))

    def forward(self, input):
        ops.require_cuda(input, "InputInjection")
        x = input
        for _ in self.pool:
            n, c, h, w = x.shape
            y = ops.new_act(n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, torch.float32, x.device, c_alloc=4)
            x = ops.avgpool3x3s2(x if (ops.is_nhwc(x) or x.is_contiguous()) else x.contiguous(), y)
        return x


class DABNet(nn.Module):
    def __init__(self, classes=19, block_1=3, block_2=6):
        super().__init__()
        self.init_conv = nn.Sequential(
            Conv(3, 32, 3, 2, padding=1, bn_acti=True),
            Conv(32, 32, 3, 1, padding=1, bn_acti=True),
            Conv(32, 32, 3, 1, padding=1, bn_acti=True),
        )
        self.down_1 = InputInjection(1)
        self.down_2 = InputInjection(2)
        self.down_3 = InputInjection(3)
        self.bn_prelu_1 = BNPReLU(32 + 3)
        self.downsample_1 = DownSamplingBlock(32 + 3, 64)
        self.DAB_Block_1 = nn.Sequential()
        for i in range(0, block_1):
            self.DAB_Block_1.add_module("DAB_Module_1_" + str(i), DABModule(64, d=2))
        self.bn_prelu_2 = BNPReLU(128 + 3)
        dilation_block_2 = [4, 4, 8, 8, 16, 16]
        self.downsample_2 = DownSamplingBlock(128 + 3, 128)
        self.DAB_Block_2 = nn.Sequential()
        for i in range(0, block_2):
            self.DAB_Block_2.add_module("DAB_Module_2_" + str(i), DABModule(128, d=dilation_block_2[i]))
        self.bn_prelu_3 = BNPReLU(256 + 3)
        self.classifier = nn.Sequential(Conv(259, classes, 1, 1, padding=0))

    def _scores(self, input):
        """Everything up to the 1/8-resolution class scores (NHWC fp32)."""
        ops.require_cuda(input, "DABNet")
        if input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        dt = ops.compute_dtype(input)
        dev = input.device
        n, _, h, w = input.shape

        # input-injection pyramid, computed once (the reference recomputes it 3x: DABNet.py:160-165)
        d1 = self.down_1(input)
        d2 = self.down_1(d1)
        d3 = self.down_1(d2)

        h1, w1 = d1.shape[2:]
        # concat buffers are written whole (producers' slices + esn_concat_tail, which also zeroes the channel padding):
        # no zero fill
        cat0 = ops.new_act(n, 35, h1, w1, dt, dev, c_alloc=64)
        # The BNPReLU over each concat (DABNet.py:166,171,176) never runs as a pass over the whole buffer: every producer
        # applies its channel slice of it -- in its own epilogue where the raw value has no other reader
        y = self.init_conv[0](input)
        y = self.init_conv[1](y)
        s, b, a = self.bn_prelu_1.prep(dev)[:3]
        self.init_conv[2](y, then=(s[:32], b[:32], a[:32]), out2=cat0[:, :32], keep=False)
        ops.concat_tail(d1, cat0, 32, s[32:], b[32:], a[32:], ACT_PRELU)

        h2, w2 = d2.shape[2:]
        cat1 = ops.new_act(n, 131, h2, w2, dt, dev, c_alloc=192)
        s, b, a = self.bn_prelu_2.prep(dev)[:3]
        y = self.downsample_1(cat0, out=cat1[:, 64:128])
        self._dab_block(list(self.DAB_Block_1), y, cat1[:, 0:64], (s[:64], b[:64], a[:64]))
        ops.affine_act(y, s[64:128], b[64:128], a[64:128], ACT_PRELU, out=y)     # the down-sampler's slice, after its last reader
        ops.concat_tail(d2, cat1, 128, s[128:], b[128:], a[128:], ACT_PRELU)

        h3, w3 = d3.shape[2:]
        cat2 = ops.new_act(n, 259, h3, w3, dt, dev, c_alloc=320)
        s, b, a = self.bn_prelu_3.prep(dev)[:3]
        y = self.downsample_2(cat1, out=cat2[:, 128:256])
        self._dab_block(list(self.DAB_Block_2), y, cat2[:, 0:128], (s[:128], b[:128], a[:128]))
        ops.affine_act(y, s[128:256], b[128:256], a[128:256], ACT_PRELU, out=y)
        ops.concat_tail(d3, cat2, 256, s[256:], b[256:], a[256:], ACT_PRELU)

        classes = self.classifier[0].conv.out_channels
        if dt == torch.bfloat16:
            # 1x1 classifier on the tensor cores over the padded 320-channel view (zero weights on the tail)
            scores = ops.new_act(n, classes, h3, w3, torch.bfloat16, dev, c_alloc=32)
            cls = self._cls_prep(dev)
            ops.conv2d(ops.widen(cat2, cls.cin), cls, out=ops.widen(scores, cls.cout))
        else:
            scores = ops.new_act(n, classes, h3, w3, torch.float32, dev, c_alloc=32)
            self.classifier[0](cat2, out=scores)
        return scores, (h, w), dt

    @staticmethod
    def _dab_block(blocks, y, out, out_affine):
        """A chain of DABModules: each conv1x1 + residual launch also emits the next module's bn_relu_1; the last one
        writes only BNPReLU_concat(output) into its slice `out` of the concat buffer."""
        pre = None
        for i, blk in enumerate(blocks):
            if i == len(blocks) - 1:
                blk(y, out=out, pre=pre, then=out_affine, keep=False)
            else:
                y, pre = blk(y, pre=pre, then=blocks[i + 1].bn_relu_1.prep(y.device)[:3])

    def _cls_prep(self, device):
        m = self.classifier[0]
        key = (str(device), weights_generation(), m.conv.weight.data_ptr(), m.conv.weight._version)
        cached = self.__dict__.get("_esn_cls")
        if cached is None or cached[0] != key:
            cached = (key, ops.ConvPrep(m.conv, device=device, cin_pad=320, cout_pad=32))
            self.__dict__["_esn_cls"] = cached
        return cached[1]

    def forward(self, input):
        if self.training:
            # batch-statistics BatchNorm + recorded backward (esn/train.py); one autograd node for the net
            from esn import train as T
            from model._dabnet_train import dabnet_train_forward
            return T.run_network(self, lambda x: dabnet_train_forward(self, x), input)
        scores, (h, w), dt = self._scores(input)
        ldt = torch.bfloat16 if dt == torch.bfloat16 else torch.float32
        return ops.head_bilinear(scores, scores.shape[1], h, w, True, False, ldt)[0]

    def fused_loss(self, input, target, criterion):
        """criterion(self(input), target) -- train.py:351-352 -- with the bilinear head, CrossEntropyLoss2d and both their
        backward passes as ONE launch (esn_bilinear_ce: the full-resolution logits and their gradient are never written).
        esn.graph.GraphedTrainStep calls this when the model offers it.  Any other criterion or eval mode takes the
        two-module form."""
        from esn import train as T
        from model._dabnet_train import dabnet_train_forward
        return T.fused_bilinear_loss(self, dabnet_train_forward, input, target, criterion, self.classifier[0].conv.out_channels)

    @torch.no_grad()
    def predict_mask(self, input, with_logits=False):
        """uint8 (N,H,W) argmax mask computed inside the upsampling kernel (test.py:79-82 on GPU)."""
        scores, (h, w), dt = self._scores(input)
        ldt = torch.bfloat16 if dt == torch.bfloat16 else torch.float32
        logits, mask = ops.head_bilinear(scores, scores.shape[1], h, w, with_logits, True, ldt)
        return (logits, mask) if with_logits else mask
