"""Train-mode forward of CGNet on the training kernels (esn/train.py): same graph as the reference's CGNet.forward
(model/CGNet.py:193-367) with batch-statistics BatchNorm, recording the backward on a tape.  The joint feature
cat[F_loc, F_sur] is one buffer whose halves the two depthwise convs write (forward) and read their gradients from
(backward); the stage concats are channel slices of one buffer as in the inference path; FGlo = esn.train.fglo."""
import torch

from esn import ops
from esn import train as T
from esn._lib import ACT_NONE, ACT_PRELU


def _convT(conv):
    t = conv.__dict__.get("_esn_T")
    if t is None:
        t = T.ConvT(conv)
        conv.__dict__["_esn_T"] = t
    return t


def _bnprelu(tape, bn, act, x, out=None):
    return T.BNActT(bn, ACT_PRELU, act).forward(tape, x, out=out)


def _cbr(tape, m, x, out=None, need_dx=True, dtype=None):
    """ConvBNPReLU (CGNet.py:13-40)."""
    return _bnprelu(tape, m.bn, m.act, _convT(m.conv).forward(tape, x, need_dx=need_dx, dtype=dtype), out=out)


def _joint(tape, blk, y, bn, act):
    """cat[F_loc(y), F_sur(y)] -> BN -> PReLU (CGNet.py:214-219, 250-254)."""
    n, c, h, w = y.t.shape
    joi = T.V(ops.new_act(n, 2 * c, h, w, y.t.dtype, y.t.device))
    _convT(blk.F_loc.conv).forward(tape, y, out=joi.slice(0, c))
    _convT(blk.F_sur.conv).forward(tape, y, out=joi.slice(c, 2 * c))
    return _bnprelu(tape, bn, act, joi)


def _block_down(tape, blk, x, out=None):
    """ContextGuidedBlock_Down (CGNet.py:193-227)."""
    y = _cbr(tape, blk.conv1x1, x)
    j = _joint(tape, blk, y, blk.bn, blk.act)
    r = _convT(blk.reduce.conv).forward(tape, j)
    return T.fglo(tape, blk.F_glo.fc, r, out=out)


def _block(tape, blk, x, out=None):
    """ContextGuidedBlock (CGNet.py:230-260)."""
    y = _cbr(tape, blk.conv1x1, x)
    j = _joint(tape, blk, y, blk.bn_prelu.bn, blk.bn_prelu.act)
    return T.fglo(tape, blk.F_glo.fc, j, out=out, residual=x if blk.add else None)


def cgnet_train_forward(model, input, loss=None):
    ops.require_cuda(input, "CGNet")
    if input.dtype != torch.float32 or not input.is_contiguous():
        input = input.float().contiguous()
    n, _, H, W = input.shape
    dt = ops.compute_dtype(input)
    dev = input.device
    tape = T.Tape(model.__dict__.get("_esn_buckets"), device=input.device)
    m = model
    inp1 = m.sample1(input)            # input-injection pyramid: no parameters, no gradient
    inp2 = m.sample1(inp1)

    def cat_buffer(c, like):
        hh, ww = like.shape[2:]
        return T.V(ops.new_act(n, c, hh, ww, dt, dev, c_alloc=(c + 7) // 8 * 8, zero=True))

    y = _cbr(tape, m.level1_0, T.V(input), need_dx=False, dtype=dt)
    y = _cbr(tape, m.level1_1, y)
    cat0 = cat_buffer(35, inp1)
    _cbr(tape, m.level1_2, y, out=cat0.slice(0, 32))
    ops.affine_act(inp1, None, None, None, ACT_NONE, out=cat0.t[:, 32:35])
    c0 = _bnprelu(tape, m.b1.bn, m.b1.act, cat0)

    cat1 = cat_buffer(131, inp2)
    y10 = _block_down(tape, m.level2_0, c0, out=cat1.slice(64, 128))
    y = y10
    for i, layer in enumerate(m.level2):
        y = _block(tape, layer, y, out=cat1.slice(0, 64) if i == len(m.level2) - 1 else None)
    ops.affine_act(inp2, None, None, None, ACT_NONE, out=cat1.t[:, 128:131])
    c1 = _bnprelu(tape, m.bn_prelu_2.bn, m.bn_prelu_2.act, cat1)

    h3, w3 = (cat1.t.shape[2] - 1) // 2 + 1, (cat1.t.shape[3] - 1) // 2 + 1
    cat2 = T.V(ops.new_act(n, 256, h3, w3, dt, dev))
    y20 = _block_down(tape, m.level3_0, c1, out=cat2.slice(0, 128))
    y = y20
    for i, layer in enumerate(m.level3):
        y = _block(tape, layer, y, out=cat2.slice(128, 256) if i == len(m.level3) - 1 else None)
    c2 = _bnprelu(tape, m.bn_prelu_3.bn, m.bn_prelu_3.act, cat2)

    if len(m.classifier) == 2 and m.classifier[0].p > 0:          # dropout_flag=True: nn.Dropout2d(0.1) (CGNet.py:309-311)
        c2 = T.dropout(tape, c2, m.classifier[0].p, per_channel=True)
    conv = m.classifier[-1].conv
    classes = conv.out_channels
    scores = T.V(ops.new_act(n, classes, h3, w3, dt, dev, c_alloc=32))
    _convT(conv).forward(tape, c2, out=scores)
    # fp32 logits, or the loss sums of the fused close when called from CGNet.fused_loss (esn_bilinear_ce)
    return T.bilinear_close(tape, scores, H, W, loss)
