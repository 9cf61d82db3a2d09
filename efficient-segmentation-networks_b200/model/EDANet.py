"""EDANet on B200 kernels -- drop-in for the reference's model/EDANet.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/EDANet.py:18-157.  EDANet is a dense-concatenation net of ERFNet-style asymmetric convs:
every EDAModule computes k = 40 new channels (1x1 -> 3x1 -> 1x3 -> dilated 3x1 -> dilated 1x3, BatchNorm + ReLU
after the 1x1 and after each 1x3) and returns ``cat([new, input])``.

Launch plan:

* The concat is never materialised.  An EDANetBlock owns ONE NHWC buffer holding its final channel count
  (in + L*k, pixel stride padded for the tensor cores, tail zeroed); the producer of the block's input (the preceding
  DownsamplerBlock) writes straight into the buffer's last ``in`` channels, module i reads the channel slice
  ``[(L-i)*k:]`` and writes its 40 channels right in front of it -- so ``cat([new, input])`` is a pointer offset.
* Convs carry bias / BatchNorm / ReLU in the epilogue.  In bf16 the 1x1 reductions (60...410 input channels) and the
  40-channel asymmetric convs run on the tcgen05 kernel over zero-extended weights: the 1x1 reads the slice widened
  to the next multiple of 64 channels (the extra channels are older data or the zero tail, times zero weights), the
  40-channel intermediates live in 64-channel buffers whose tail is written as exact zeros by zero weight rows.
* DownsamplerBlock: strided conv (+ 2x2 max-pool branch when ninput < noutput) with the BatchNorm slice + ReLU folded
  into each branch, written into the consumer's buffer.
* project_layer (1x1 -> classes) + ``F.interpolate(scale_factor=8, bilinear, align_corners=True)`` (EDANet.py:155-156):
  tcgen05 1x1 into a 32-channel score buffer, then the fused bilinear head (NCHW logits and/or uint8 argmax).
"""
import torch
import torch.nn as nn

from esn import ops
from esn._lib import ACT_NONE, ACT_RELU
from esn.prep import PrepMixin, weights_generation

__all__ = ["EDANet"]


def _no_train(mod):
    if mod.training:
        raise NotImplementedError("%s: training-mode kernels are not wired for this model; call .eval(). "
                                  "There is no eager-PyTorch fallback." % type(mod).__name__)


def _tc(c):
    """Channel count the tensor-core conv accepts as Cin: 16/32/64 or a multiple of 64."""
    for v in (16, 32, 64):
        if c <= v:
            return v
    return (c + 63) // 64 * 64


class _Packed:
    """A conv (+BN) (+ReLU) in two packings: exact channels (fp32 / any layout) and, for bf16, weights zero-extended
    to tensor-core friendly channel counts (cin_pad always; cout_pad only when the output buffer is private)."""

    def __init__(self, conv, bn, act, device, pad_out):
        s, b = ops.bn_affine(bn, device) if bn is not None else (None, None)
        self.plain = ops.ConvPrep(conv, s, b, act, device=device)
        cin_pad = _tc(conv.in_channels)
        cout_pad = _tc(conv.out_channels) if pad_out else None
        self.padded = ops.ConvPrep(conv, s, b, act, device=device, cin_pad=cin_pad if cin_pad != conv.in_channels else None,
                                   cout_pad=cout_pad if cout_pad != conv.out_channels else None)
        self.pad_out = pad_out

    def __call__(self, x, out=None):
        """x: NHWC view with the conv's logical channels; when its pixel stride leaves room and it is bf16, the padded
        packing runs on the widened view.  out=None allocates a private buffer (padded stride, zero tail)."""
        plain, padded = self.plain, self.padded
        n, _, h, w = x.shape
        ho, wo = plain.out_hw(h, w)
        room = _room(x)
        use_padded = x.dtype == torch.bfloat16 and room >= padded.cin
        if out is None:
            alloc = _tc(plain.cout) if self.pad_out else plain.cout
            out = ops.new_act(n, plain.cout, ho, wo, x.dtype, x.device, c_alloc=alloc,
                              zero=alloc != plain.cout and not use_padded)
        elif self.pad_out:
            raise ValueError("a conv whose output is zero-extended cannot write into a caller's slice")
        if use_padded:
            ops.conv2d(ops.widen(x, padded.cin), padded, out=ops.widen(out, padded.cout) if padded.cout != plain.cout else out)
        else:
            ops.conv2d(x, plain, out=out)
        return out


def _room(x):
    """Channels readable from x's first channel inside one pixel of its buffer (>= x.shape[1]); the views used here
    always start `shape[1] + tail` channels before the end of the pixel, which _BlockBuffer records on the view."""
    return getattr(x, "_esn_room", x.stride(3) if x.storage_offset() % x.stride(3) == 0 else x.shape[1])


def _slice(buf_full, lo, hi):
    """Channel slice [lo:hi] of a full-width NHWC buffer view, remembering how far it may be widened."""
    v = buf_full[:, lo:hi]
    v._esn_room = buf_full.stride(3) - lo
    return v


class DownsamplerBlock(PrepMixin, nn.Module):
    def __init__(self, ninput, noutput):
        super().__init__()
        self.ninput = ninput
        self.noutput = noutput
        if self.ninput < self.noutput:
            self.conv = nn.Conv2d(ninput, noutput - ninput, kernel_size=3, stride=2, padding=1)
            self.pool = nn.MaxPool2d(2, stride=2)
        else:
            self.conv = nn.Conv2d(ninput, noutput, kernel_size=3, stride=2, padding=1)
        self.bn = nn.BatchNorm2d(noutput)

    def _build_prep(self, device):
        scale, shift = ops.bn_affine(self.bn, device)
        nc = self.conv.out_channels
        plain = ops.ConvPrep(self.conv, scale[:nc], shift[:nc], ACT_RELU, device=device)
        cin_pad = _tc(self.conv.in_channels)
        padded = (ops.ConvPrep(self.conv, scale[:nc], shift[:nc], ACT_RELU, device=device, cin_pad=cin_pad)
                  if self.conv.in_channels > 8 and cin_pad != self.conv.in_channels else plain)
        return plain, padded, scale[nc:].contiguous(), shift[nc:].contiguous()

    def forward(self, x, out=None):
        _no_train(self)
        x = ops.as_act(x)
        if (x.shape[2] | x.shape[3]) & 1:
            raise NotImplementedError("EDANet DownsamplerBlock: odd input sizes are not supported, got %dx%d"
                                      % (x.shape[2], x.shape[3]))
        plain, padded, pscale, pshift = self.prep(x.device)
        n, c, h, w = x.shape
        if out is None:      # pixel stride rounded up to 8 channels with a zero tail, so a consumer can read it widened
            alloc = (self.noutput + 7) // 8 * 8
            out = ops.new_act(n, self.noutput, h // 2, w // 2, x.dtype, x.device, c_alloc=alloc, zero=alloc != self.noutput)
        nc = plain.cout
        if x.dtype == torch.bfloat16 and padded is not plain and _room(x) >= padded.cin:
            ops.conv2d(ops.widen(x, padded.cin), padded, out=out[:, :nc])
        else:
            ops.conv2d(x, plain, out=out[:, :nc])
        if self.ninput < self.noutput:
            ops.maxpool2x2(x, out[:, nc:], pscale, pshift, None, ACT_RELU)
        return out


class EDAModule(PrepMixin, nn.Module):
    def __init__(self, ninput, dilated, k=40, dropprob=0.02):
        super().__init__()
        self.conv1x1 = nn.Conv2d(ninput, k, kernel_size=1)
        self.bn0 = nn.BatchNorm2d(k)
        self.conv3x1_1 = nn.Conv2d(k, k, kernel_size=(3, 1), padding=(1, 0))
        self.conv1x3_1 = nn.Conv2d(k, k, kernel_size=(1, 3), padding=(0, 1))
        self.bn1 = nn.BatchNorm2d(k)
        self.conv3x1_2 = nn.Conv2d(k, k, (3, 1), stride=1, padding=(dilated, 0), dilation=dilated)
        self.conv1x3_2 = nn.Conv2d(k, k, (1, 3), stride=1, padding=(0, dilated), dilation=dilated)
        self.bn2 = nn.BatchNorm2d(k)
        self.dropout = nn.Dropout2d(dropprob)

    def _build_prep(self, device):
        return (_Packed(self.conv1x1, self.bn0, ACT_RELU, device, True),
                _Packed(self.conv3x1_1, None, ACT_NONE, device, True),          # no activation between the 3x1 and the 1x3
                _Packed(self.conv1x3_1, self.bn1, ACT_RELU, device, True),
                _Packed(self.conv3x1_2, None, ACT_NONE, device, True),
                _Packed(self.conv1x3_2, self.bn2, ACT_RELU, device, False))     # writes the 40 new channels of the concat

    def new_channels(self, x, out):
        """The module without its concat: relu(bn2(...)) of x written into `out` (k channels, any NHWC slice)."""
        _no_train(self)            # eval: Dropout2d is the identity
        c0, c1, c2, c3, c4 = self.prep(x.device)
        return c4(c3(c2(c1(c0(x)))), out=out)

    def forward(self, x):
        """Stand-alone use: returns cat([new, x]) in a fresh buffer (inside EDANetBlock no copy is made)."""
        x = ops.as_act(x)
        n, c, h, w = x.shape
        k = self.conv1x1.out_channels
        full = ops.new_act(n, k + c, h, w, x.dtype, x.device)
        ops.affine_act(x, None, None, None, ACT_NONE, out=full[:, k:])
        self.new_channels(full[:, k:], full[:, :k])
        return full


class EDANetBlock(nn.Module):
    def __init__(self, in_channels, num_dense_layer, dilated, growth_rate):
        super().__init__()
        _in_channels = in_channels
        modules = []
        for i in range(num_dense_layer):
            modules.append(EDAModule(_in_channels, dilated[i], growth_rate))
            _in_channels += growth_rate
        self.residual_dense_layers = nn.Sequential(*modules)
        self._in, self._k, self._layers = in_channels, growth_rate, num_dense_layer

    @property
    def out_channels(self):
        return self._in + self._k * self._layers

    def new_buffer(self, n, h, w, dtype, device, consumer_pad=0):
        """Zeroed NHWC buffer for the block's final concat.  The pixel stride leaves room for every widened read: module
        i reads from channel (L-i)*k on, `consumer_pad` is what the block's consumer reads from channel 0."""
        total, k, L = self.out_channels, self._k, self._layers
        need = max([total, consumer_pad] + [(L - i) * k + _tc(self._in + i * k) for i in range(L)])
        alloc = (need + 63) // 64 * 64
        full = ops.new_act(n, alloc, h, w, dtype, device, zero=True)
        return full, _slice(full, L * k, total)          # (whole buffer, the slice the block's input is written to)

    def run(self, full):
        """Fill channels [0, L*k) of `full` (its channels [L*k, total) hold the block's input); returns the concat view."""
        total, k, L = self.out_channels, self._k, self._layers
        for i, mod in enumerate(self.residual_dense_layers):
            lo = (L - i) * k
            mod.new_channels(_slice(full, lo, total), _slice(full, lo - k, lo))
        return _slice(full, 0, total)

    def forward(self, x):
        x = ops.as_act(x)
        n, _, h, w = x.shape
        full, slot = self.new_buffer(n, h, w, x.dtype, x.device)
        ops.affine_act(x, None, None, None, ACT_NONE, out=slot)
        return self.run(full)


class EDANet(nn.Module):
    def __init__(self, classes=19):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(DownsamplerBlock(3, 15))
        self.layers.append(DownsamplerBlock(15, 60))
        self.layers.append(EDANetBlock(60, 5, [1, 1, 1, 2, 2], 40))
        self.layers.append(DownsamplerBlock(260, 130))
        self.layers.append(EDANetBlock(130, 8, [2, 2, 4, 4, 8, 8, 16, 16], 40))
        self.project_layer = nn.Conv2d(450, classes, kernel_size=1)
        self.weights_init()

    def weights_init(self):
        """Same distributions as the reference's initialiser (EDANet.py:137-146): conv weights N(0, 0.02), BatchNorm
        weights N(1, 0.02) with zero bias."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0.0, 0.02)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.normal_(1.0, 0.02)
                m.bias.data.fill_(0)

    def _build_project(self, device):
        conv = self.project_layer
        return (ops.ConvPrep(conv, device=device),
                ops.ConvPrep(conv, device=device, cin_pad=_tc(conv.in_channels), cout_pad=32), conv.out_channels)

    def _scores(self, x):
        ops.require_cuda(x, "EDANet")
        _no_train(self)
        if (x.shape[2] | x.shape[3]) % 8:
            raise NotImplementedError("EDANet: input height and width must be multiples of 8, got %dx%d"
                                      % (x.shape[2], x.shape[3]))
        down1, down2, block1, down3, block2 = self.layers
        plain, padded, classes = self._project_prep(x.device)
        y = down1(x)
        n, _, h, w = y.shape
        full1, slot1 = block1.new_buffer(n, h // 2, w // 2, y.dtype, y.device, consumer_pad=_tc(block1.out_channels))
        down2(y, out=slot1)
        cat1 = block1.run(full1)
        full2, slot2 = block2.new_buffer(n, h // 4, w // 4, y.dtype, y.device, consumer_pad=padded.cin)
        down3(cat1, out=slot2)
        cat2 = block2.run(full2)
        scores = ops.new_act(n, classes, h // 4, w // 4, cat2.dtype, cat2.device, c_alloc=32)
        if cat2.dtype == torch.bfloat16:
            ops.conv2d(ops.widen(cat2, padded.cin), padded, out=ops.widen(scores, 32))
        else:
            ops.conv2d(cat2, plain, out=scores)
        return scores, x.shape[2:]

    def _project_prep(self, device):
        conv = self.project_layer
        sig = (str(device), weights_generation(), conv.weight.data_ptr(), conv.weight._version, conv.bias.data_ptr(), conv.bias._version)
        cached = self.__dict__.get("_esn_project")
        if cached is None or cached[0] != sig:
            with torch.no_grad():
                cached = (sig, self._build_project(device))
            self.__dict__["_esn_project"] = cached
        return cached[1]

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
