"""Train-mode forward of Fast-SCNN on the training kernels (esn/train.py): same graph as the reference's
FastSCNN.forward (model/FastSCNN.py:204-235) with batch-statistics BatchNorm and Dropout(0.1), recording the
backward on a tape.  The pyramid-pooling concat is a set of channel slices of one buffer (forward and gradient)."""
import torch

from esn import ops
from esn import train as T
from esn._lib import ACT_NONE, ACT_RELU


def _convT(conv):
    t = conv.__dict__.get("_esn_T")
    if t is None:
        t = T.ConvT(conv)
        conv.__dict__["_esn_T"] = t
    return t


def _cbr(tape, seq, x, need_dx=True, dtype=None, out=None, act=ACT_RELU, i=0):
    """seq[i] conv -> seq[i+1] BatchNorm -> act."""
    y = _convT(seq[i]).forward(tape, x, need_dx=need_dx, dtype=dtype)
    return T.BNActT(seq[i + 1], act).forward(tape, y, out=out)


def _dsconv(tape, m, x):
    """_DSConv: dw3x3(s)+BN+ReLU -> 1x1+BN+ReLU (FastSCNN.py:29-45)."""
    return _cbr(tape, m.conv, _cbr(tape, m.conv, x), i=3)


def _bottleneck(tape, m, x):
    """LinearBottleneck (FastSCNN.py:62-82)."""
    y = _cbr(tape, m.block[0].conv, x)
    y = _cbr(tape, m.block[1].conv, y)
    y = _cbr(tape, m.block, y, act=ACT_NONE, i=2)
    return T.add(tape, x, y) if m.use_shortcut else y


def fastscnn_train_forward(model, input, loss=None):
    ops.require_cuda(input, "FastSCNN")
    if input.dtype != torch.float32 or not input.is_contiguous():
        input = input.float().contiguous()
    dt = ops.compute_dtype(input)
    dev = input.device
    n, _, H, W = input.shape
    tape = T.Tape(model.__dict__.get("_esn_buckets"), device=input.device)     # data-parallel gradient buckets, if attached

    ltd = model.learning_to_downsample
    y = _cbr(tape, ltd.conv.conv, T.V(input), need_dx=False, dtype=dt)
    y = _dsconv(tape, ltd.dsconv1, y)
    higher = _dsconv(tape, ltd.dsconv2, y)

    gfe = model.global_feature_extractor
    y = higher
    for stage in (gfe.bottleneck1, gfe.bottleneck2, gfe.bottleneck3):
        for blk in stage:
            y = _bottleneck(tape, blk, y)
    # pyramid pooling (FastSCNN.py:85-112): [x | up(conv(pool_s(x))) for s in 1,2,3,6] -> 1x1
    ppm = gfe.ppm
    _, c, h, w = y.t.shape
    cat = T.V(ops.new_act(n, 2 * c, h, w, dt, dev))
    T.copy_into(tape, y, cat.slice(0, c))
    ci = c // 4
    for i, (conv, size) in enumerate(((ppm.conv1, 1), (ppm.conv2, 2), (ppm.conv3, 3), (ppm.conv4, 6))):
        f = _cbr(tape, conv.conv, T.adaptive_avgpool(tape, y, size))
        T.bilinear(tape, f, h, w, True, out=cat.slice(c + i * ci, c + (i + 1) * ci))
    y = _cbr(tape, ppm.out.conv, cat)

    # feature fusion (FastSCNN.py:157-182)
    ffm = model.feature_fusion
    _, _, hh, wh = higher.t.shape
    lo = T.bilinear(tape, y, hh, wh, True)
    lo = _cbr(tape, ffm.dwconv.conv, lo)
    lo = _cbr(tape, ffm.conv_lower_res, lo, act=ACT_NONE)
    hi = _cbr(tape, ffm.conv_higher_res, higher, act=ACT_NONE)
    y = T.BNActT(None, ACT_RELU).forward(tape, T.add(tape, hi, lo))

    # classifier (FastSCNN.py:185-201) + final bilinear (align_corners=True) to the input size
    cl = model.classifier
    y = _dsconv(tape, cl.dsconv2, _dsconv(tape, cl.dsconv1, y))
    y = T.dropout(tape, y, cl.conv[0].p, per_channel=False, training=True)
    classes = cl.conv[1].out_channels
    scores = T.V(ops.new_act(n, classes, y.t.shape[2], y.t.shape[3], dt, dev, c_alloc=32))
    _convT(cl.conv[1]).forward(tape, y, out=scores)
    # fp32 logits, or the loss sums of the fused close when called from FastSCNN.fused_loss (esn_bilinear_ce)
    return T.bilinear_close(tape, scores, H, W, loss, align_corners=True)
