"""Train-mode forward of ERFNet on the training kernels (esn/train.py): same graph as the reference's
ERFNet.forward (model/ERFNet.py:16-156) with batch-statistics BatchNorm and Dropout2d, recording the backward on a
tape.  The downsampler concat [conv | max-pool] is two channel slices of one buffer (forward and gradient)."""
import torch

from esn import ops
from esn import train as T
from esn._lib import ACT_NONE, ACT_RELU


def _convT(conv):
    t = conv.__dict__.get("_esn_T")
    if t is None:
        t = T.ConvTransposeT(conv) if isinstance(conv, torch.nn.ConvTranspose2d) else T.ConvT(conv)
        conv.__dict__["_esn_T"] = t
    return t


def _relu(tape, x):
    return T.BNActT(None, ACT_RELU).forward(tape, x)


def _down(tape, m, x, dt, image=False):
    """DownsamplerBlock (ERFNet.py:16-27): cat[conv3x3 s2, maxpool 2x2] -> BN -> ReLU."""
    n, c, h, w = x.t.shape
    nc = m.conv.out_channels
    cat = T.V(ops.new_act(n, nc + c, h // 2, w // 2, dt, x.t.device))
    if image and not ops.is_nhwc(x.t) and not ((h | w) & 1):
        # initial block on the NCHW fp32 image: conv and pool in ONE pass of the stem kernel (raw concat, the conv bias as the
        # shift); only the conv's weight gradient is recorded -- no gradient flows to the image
        ct = _convT(m.conv)
        fwd_prep, _ = ct.preps()
        ones = torch.ones(nc + c, device=x.t.device)
        shift = torch.zeros(nc + c, device=x.t.device)
        shift[:nc] = fwd_prep.shift
        ops.stem_conv3x3s2(x.t, fwd_prep.w_direct, nc, 1, cat.t, ones, shift, None, ACT_NONE)
        ct.forward(tape, x, out=cat.slice(0, nc), need_dx=False, dtype=dt, precomputed=True)
        return T.BNActT(m.bn, ACT_RELU).forward(tape, cat)
    _convT(m.conv).forward(tape, x, out=cat.slice(0, nc), need_dx=not image, dtype=dt)
    T.maxpool2x2(tape, x, cat.slice(nc, nc + c), need_dx=not image)
    return T.BNActT(m.bn, ACT_RELU).forward(tape, cat)


def _nb1d(tape, m, x):
    """non_bottleneck_1d (ERFNet.py:30-65)."""
    y = _relu(tape, _convT(m.conv3x1_1).forward(tape, x))
    y = T.BNActT(m.bn1, ACT_RELU).forward(tape, _convT(m.conv1x3_1).forward(tape, y))
    y = _relu(tape, _convT(m.conv3x1_2).forward(tape, y))
    y = T.BNActT(m.bn2, ACT_NONE).forward(tape, _convT(m.conv1x3_2).forward(tape, y))
    if m.dropout.p != 0:
        return _relu(tape, T.dropout(tape, y, m.dropout.p, per_channel=True, residual=x))     # dropout(y) + x in one pass
    return _relu(tape, T.add(tape, y, x))


def _up(tape, m, x):
    """UpsamplerBlock (ERFNet.py:103-112): ConvTranspose2d(3, s2, p1, op1) -> BN -> ReLU."""
    return T.BNActT(m.bn, ACT_RELU).forward(tape, _convT(m.conv).forward(tape, x))


def erfnet_train_forward(model, input):
    ops.require_cuda(input, "ERFNet")
    if input.dtype != torch.float32 or not input.is_contiguous():
        input = input.float().contiguous()
    n, _, H, W = input.shape
    if (H | W) & 7:
        raise ValueError("ERFNet: input height and width must be multiples of 8, got %dx%d" % (H, W))
    dt = ops.compute_dtype(input)
    tape = T.Tape(model.__dict__.get("_esn_buckets"), device=input.device)
    enc, dec = model.encoder, model.decoder
    y = _down(tape, enc.initial_block, T.V(input), dt, image=True)
    for layer in enc.layers:
        y = _nb1d(tape, layer, y) if hasattr(layer, "conv3x1_1") else _down(tape, layer, y, dt)
    for layer in dec.layers:
        y = _nb1d(tape, layer, y) if hasattr(layer, "conv3x1_1") else _up(tape, layer, y)
    w, b, classes, _ = dec.prep(y.t.device)
    logits, holder = T.convt2x2_logits(tape, _convT(dec.output_conv), y, w, b, classes)
    return logits, tape, holder
