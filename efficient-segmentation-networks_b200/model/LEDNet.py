"""LEDNet on B200 kernels -- drop-in for the reference's model/LEDNet.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/LEDNet.py:46-327.

* SS_nbt_module_paper (LEDNet.py:108-186): the two halves of the input are channel slices of one NHWC buffer (no
  split copy); each half runs its four factorized convs (left 3x1 -> 1x3, right 1x3 -> 3x1, the second pair dilated)
  on the tcgen05 kernel (16 / 32 / 64 channels) with bias / BatchNorm / ReLU in the epilogue; the closing conv of each
  half adds its half of the block input and applies the ReLU while writing into its half of the merged buffer (no
  concat, no add pass).  The channel shuffle is ONE 1x1 conv whose weight is the permutation matrix: exact in bf16
  (every output is a single input times 1.0, accumulated in fp32) and on the tensor cores.
* APNModule (LEDNet.py:189-283): the reference's asymmetric-stride pairs -- (k,1) conv with stride (2,1), then (1,k)
  with stride (1,2) -- are run as a stride-1 (k,1) conv followed by a (1,k) conv with stride 2 in both directions:
  the second conv has a single row tap, so it reads exactly the even rows the reference keeps (same outputs, the
  single-channel maps are tiny).  The three-level pyramid is resized by the NHWC bilinear kernel (align_corners=True);
  `x * mid + b1` -- a one-channel gate times the class scores plus the global-pooling branch, which after its
  1x1 -> (h,w) upsampling is a constant per image and class -- is one launch of esn_gate_bcast.
* DownsamplerBlock and the final bilinear (align_corners=True) + argmax: ERFNet's / Fast-SCNN's.
"""
import torch
import torch.nn as nn

from esn import ops
from esn._lib import ACT_NONE, ACT_RELU
from esn.prep import PrepMixin
from model.ESNet import DownsamplerBlock as _EsDownsamplerBlock

__all__ = ["LEDNet"]


def _no_train(mod):
    if mod.training:
        raise NotImplementedError("%s: training-mode kernels are not wired for this model; call .eval(). "
                                  "There is no eager-PyTorch fallback." % type(mod).__name__)


class Conv2dBnRelu(PrepMixin, nn.Module):
    def __init__(self, in_ch, out_ch, kernel_size=3, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size, stride, padding, dilation=dilation, bias=bias),
                                  nn.BatchNorm2d(out_ch, eps=1e-3), nn.ReLU(inplace=True))

    def _build_prep(self, device):
        s, b = ops.bn_affine(self.conv[1], device)
        conv = self.conv[0]
        plain = ops.ConvPrep(conv, s, b, ACT_RELU, device=device)
        # class scores live in 32-channel buffers (zero tail) so the 1x1 runs on the tensor cores
        wide = ops.ConvPrep(conv, s, b, ACT_RELU, device=device, cout_pad=32) if conv.out_channels < 32 else plain
        return plain, wide

    def forward(self, x):
        _no_train(self)
        x = ops.as_act(x)
        plain, wide = self.prep(x.device)
        n, _, h, w = x.shape
        ho, wo = plain.out_hw(h, w)
        if wide is plain:
            return ops.conv2d(x, plain)
        y = ops.new_act(n, plain.cout, ho, wo, x.dtype, x.device, c_alloc=32, zero=x.dtype != torch.bfloat16)
        if x.dtype == torch.bfloat16:
            ops.conv2d(x, wide, out=ops.widen(y, 32))
        else:
            ops.conv2d(x, plain, out=y)
        return y


class DownsamplerBlock(_EsDownsamplerBlock):
    pass


def _shuffle_matrix(c, groups, device):
    """(C, C, 1, 1) permutation weight of Channel_shuffle (LEDNet.py:27-39): out[j*groups + g] = in[g*(C/groups) + j]."""
    w = torch.zeros((c, c, 1, 1), dtype=torch.float32, device=device)
    per = c // groups
    for g in range(groups):
        for j in range(per):
            w[j * groups + g, g * per + j, 0, 0] = 1.0
    return w


class SS_nbt_module_paper(PrepMixin, nn.Module):
    def __init__(self, chann, dropprob, dilated):
        super().__init__()
        oup_inc = chann // 2
        for side in ("l", "r"):
            setattr(self, "conv3x1_1_" + side, nn.Conv2d(oup_inc, oup_inc, (3, 1), stride=1, padding=(1, 0), bias=True))
            setattr(self, "conv1x3_1_" + side, nn.Conv2d(oup_inc, oup_inc, (1, 3), stride=1, padding=(0, 1), bias=True))
            setattr(self, "bn1_" + side, nn.BatchNorm2d(oup_inc, eps=1e-03))
            setattr(self, "conv3x1_2_" + side, nn.Conv2d(oup_inc, oup_inc, (3, 1), stride=1, padding=(1 * dilated, 0), bias=True,
                                                         dilation=(dilated, 1)))
            setattr(self, "conv1x3_2_" + side, nn.Conv2d(oup_inc, oup_inc, (1, 3), stride=1, padding=(0, 1 * dilated), bias=True,
                                                         dilation=(1, dilated)))
            setattr(self, "bn2_" + side, nn.BatchNorm2d(oup_inc, eps=1e-03))
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout2d(dropprob)
        self._chann = chann

    def _build_prep(self, device):
        def P(name, bn=None, act=ACT_RELU):
            s, b = ops.bn_affine(getattr(self, bn), device) if bn else (None, None)
            return ops.ConvPrep(getattr(self, name), s, b, act, device=device)
        # order of execution; the last conv of a side also adds that half of the input and applies the ReLU
        left = (P("conv3x1_1_l"), P("conv1x3_1_l", "bn1_l"), P("conv3x1_2_l"), P("conv1x3_2_l", "bn2_l"))
        right = (P("conv1x3_1_r"), P("conv3x1_1_r", "bn1_r"), P("conv1x3_2_r"), P("conv3x1_2_r", "bn2_r"))
        shuffle = ops.ConvPrep.from_weight(_shuffle_matrix(self._chann, 2, device))
        return left, right, shuffle

    def forward(self, x):
        _no_train(self)            # eval: Dropout2d is the identity
        x = ops.as_act(x)
        left, right, shuffle = self.prep(x.device)
        n, c, h, w = x.shape
        c1 = c // 2
        merged = ops.new_act(n, c, h, w, x.dtype, x.device)
        for convs, lo in ((left, 0), (right, c1)):
            half = x[:, lo:lo + c1]
            t = ops.conv2d(ops.conv2d(ops.conv2d(half, convs[0]), convs[1]), convs[2])
            ops.conv2d(t, convs[3], out=merged[:, lo:lo + c1], residual=half)       # relu(half + bn2(...))
        return ops.conv2d(merged, shuffle)


def _asym_pair(seq, i, k, stride, device):
    """seq[i] = (k,1) conv with stride (stride,1), seq[i+1] = (1,k) conv with stride (1,stride), seq[i+2] = BatchNorm,
    then ReLU.  Returned as (stride-1 (k,1) conv, (1,k) conv with stride `stride` in both directions + BN + ReLU): with a
    single row tap the second conv picks exactly the rows 0, stride, 2*stride, ... that the reference computes."""
    a, b, bn = seq[i], seq[i + 1], seq[i + 2]
    s, sh = ops.bn_affine(bn, device)
    first = ops.ConvPrep.from_weight(a.weight.detach().to(device), stride=1, padding=(k // 2, 0),
                                     bias=a.bias.detach().to(device=device, dtype=torch.float32))
    second = ops.ConvPrep.from_weight(b.weight.detach().to(device), stride=stride, padding=(0, k // 2),
                                      bias=b.bias.detach().to(device=device, dtype=torch.float32), scale=s, shift=sh, act=ACT_RELU)
    return first, second


def _one_channel(n, h, w, like):
    """Single-channel map: pixel stride 1 (for C = 1 the NHWC and NCHW layouts coincide, and ops.tdesc accepts only that).
    ALWAYS fp32: the pyramid is a 128 -> 1 channel reduction with heavy cancellation whose result gates every class score
    (measured on the oracle: an 8e-3 perturbation of the features comes out of the module as 4e-2 ... 9e-2), the maps are
    1/128 of the feature bytes, so bf16 storage would buy nothing and cost a further 2x of error."""
    return ops.new_act(n, 1, h, w, torch.float32, like.device)


def _run_pair(pair, x):
    first, second = pair
    n, _, h, w = x.shape
    t = ops.conv2d(x, first, out=_one_channel(n, *first.out_hw(h, w), x))
    return ops.conv2d(t, second, out=_one_channel(n, *second.out_hw(h, w), x))


def _asym(k, stride=1):
    """The reference's asymmetric pair as modules: (k,1) stride (s,1), (1,k) stride (1,s), BN, ReLU."""
    return [nn.Conv2d(1, 1, kernel_size=(k, 1), stride=(stride, 1), padding=(k // 2, 0), bias=True),
            nn.Conv2d(1, 1, kernel_size=(1, k), stride=(1, stride), padding=(0, k // 2), bias=True),
            nn.BatchNorm2d(1, eps=1e-03), nn.ReLU(inplace=True)]


class APNModule(PrepMixin, nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.branch1 = nn.Sequential(nn.AdaptiveAvgPool2d(1), Conv2dBnRelu(in_ch, out_ch, kernel_size=1, stride=1, padding=0))
        self.mid = nn.Sequential(Conv2dBnRelu(in_ch, out_ch, kernel_size=1, stride=1, padding=0))
        d1 = _asym(7, 2)
        d1[0] = nn.Conv2d(in_ch, 1, kernel_size=(7, 1), stride=(2, 1), padding=(3, 0), bias=True)
        self.down1 = nn.Sequential(*d1)
        self.down2 = nn.Sequential(*_asym(5, 2))
        self.down3 = nn.Sequential(*(_asym(3, 2) + _asym(3, 1)))
        self.conv2 = nn.Sequential(*_asym(5, 1))
        self.conv1 = nn.Sequential(*_asym(7, 1))

    def _build_prep(self, device):
        return {"down1": _asym_pair(self.down1, 0, 7, 2, device), "down2": _asym_pair(self.down2, 0, 5, 2, device),
                "down3a": _asym_pair(self.down3, 0, 3, 2, device), "down3b": _asym_pair(self.down3, 4, 3, 1, device),
                "conv2": _asym_pair(self.conv2, 0, 5, 1, device), "conv1": _asym_pair(self.conv1, 0, 7, 1, device)}

    def forward(self, x):
        _no_train(self)
        x = ops.as_act(x)
        P = self.prep(x.device)
        n, _, h, w = x.shape
        b1 = self.branch1[1](ops.adaptive_avgpool(x, 1))            # (N, classes, 1, 1): constant after its upsampling
        mid = self.mid[0](x)                                        # (N, classes, h, w)
        x1 = _run_pair(P["down1"], x)
        x2 = _run_pair(P["down2"], x1)
        x3 = _run_pair(P["down3b"], _run_pair(P["down3a"], x2))
        h4, w4, h2, w2 = (h + 3) // 4, (w + 3) // 4, (h + 1) // 2, (w + 1) // 2
        x3 = ops.bilinear(x3, h4, w4, True, out=_one_channel(n, h4, w4, x))
        y = ops.affine_act(_run_pair(P["conv2"], x2), None, None, None, ACT_NONE, residual=x3)          # x2 + x3
        y = ops.bilinear(y, h2, w2, True, out=_one_channel(n, h2, w2, x))
        y = ops.affine_act(_run_pair(P["conv1"], x1), None, None, None, ACT_NONE, residual=y)           # x + x1
        y = ops.bilinear(y, h, w, True, out=_one_channel(n, h, w, x))
        return ops.gate_bcast(y, mid, b1)                           # y * mid + b1


class LEDNet(nn.Module):
    def __init__(self, classes):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 32)
        self.layers = nn.ModuleList()
        for _ in range(3):
            self.layers.append(SS_nbt_module_paper(32, 0.03, 1))
        self.layers.append(DownsamplerBlock(32, 64))
        for _ in range(2):
            self.layers.append(SS_nbt_module_paper(64, 0.03, 1))
        self.layers.append(DownsamplerBlock(64, 128))
        for d in (1, 2, 5, 9, 2, 5, 9, 17):
            self.layers.append(SS_nbt_module_paper(128, 0.3, d))
        self.apn = APNModule(in_ch=128, out_ch=classes)

    def _scores(self, input):
        ops.require_cuda(input, "LEDNet")
        _no_train(self)
        # any input size: the DownsamplerBlock pads like the reference's (LEDNet.py:84-88), the pyramid rounds its levels itself
        output = self.initial_block(input)
        for layer in self.layers:
            output = layer(output)
        return self.apn(output), input.shape[2:]

    def forward(self, input):
        scores, (h, w) = self._scores(input)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        return ops.head_bilinear(scores, scores.shape[1], h, w, True, False, ldt, align_corners=True)[0]

    @torch.no_grad()
    def predict_mask(self, input, with_logits=False):
        """uint8 (N,H,W) argmax mask computed inside the head kernel (replaces test.py:79-82)."""
        scores, (h, w) = self._scores(input)
        ldt = torch.bfloat16 if scores.dtype == torch.bfloat16 else torch.float32
        logits, mask = ops.head_bilinear(scores, scores.shape[1], h, w, with_logits, True, ldt, align_corners=True)
        return (logits, mask) if with_logits else mask
