"""Train-mode forward of ESPNet (v1, ESPNet-A decoder) on the training kernels (esn/train.py): same graph as the
reference's ESPNet.forward (model/ESPNet.py:138-385) with batch-statistics BatchNorm, recording the backward on a tape.

The inference path keeps its activations in tensor-core-friendly physical channel layouts (model/ESPNet.py: _Map); the
training path uses the LOGICAL layouts throughout -- every concat is a buffer whose channel slices are written in place by
their producers (no torch.cat copy), the hierarchical feature fusion sums (ESPNet.py:160-166) are adds written straight into
their concat slice, and a tensor with two consumers (the down-sampled stage input, the level-1 concat) is the concat slice
itself.  Channel counts here are 12 / 16 / 25 / 28 / 19 / 38: those convs run on the direct CUDA kernels in all three
directions (forward, input gradient, weight gradient).
"""
import torch
import torch.nn as nn

from esn import ops
from esn import train as T
from esn._lib import ACT_NONE, ACT_PRELU


def _convT(conv):
    t = conv.__dict__.get("_esn_T")
    if t is None:
        t = T.ConvTransposeT(conv) if isinstance(conv, nn.ConvTranspose2d) else T.ConvT(conv)
        conv.__dict__["_esn_T"] = t
    return t


def _br(tape, bn, act, x, out=None):
    """BN (eps 1e-3) + per-channel PReLU (BR / the tail of CBR, ESPNet.py:17-62)."""
    return T.BNActT(bn, ACT_PRELU, act).forward(tape, x, out=out)


def _new(like, c, h, w, dt):
    n = like.shape[0]
    return T.V(ops.new_act(n, c, h, w, dt, like.device, c_alloc=(c + 7) // 8 * 8, zero=True))


def _five_branches(tape, m, o1, dt):
    """d1 | d2 | d2+d4 | d2+d4+d8 | d2+d4+d8+d16 (ESPNet.py:155-169, 207-221) written into one concat buffer."""
    n1, n = m.d1.conv.out_channels, m.d2.conv.out_channels
    _, _, h, w = o1.t.shape
    cat = _new(o1.t, n1 + 4 * n, h, w, dt)
    _convT(m.d1.conv).forward(tape, o1, out=cat.slice(0, n1))
    prev = _convT(m.d2.conv).forward(tape, o1, out=cat.slice(n1, n1 + n))
    lo = n1 + n
    for name in ("d4", "d8", "d16"):
        d = _convT(getattr(m, name).conv).forward(tape, o1)
        prev = T.add(tape, prev, d, out=cat.slice(lo, lo + n))
        lo += n
    return cat


def _down(tape, m, x, dt, out=None):
    """DownSamplerB (ESPNet.py:138-173)."""
    o1 = _convT(m.c1.conv).forward(tape, x)
    return _br(tape, m.bn, m.act, _five_branches(tape, m, o1, dt), out=out)


def _esp(tape, m, x, dt, out=None):
    """DilatedParllelResidualBlockB (ESPNet.py:176-229): BN(input + combine) -> PReLU when add."""
    o1 = _convT(m.c1.conv).forward(tape, x)
    c = _five_branches(tape, m, o1, dt)
    if m.add:
        c = T.add(tape, x, c)
    return _br(tape, m.bn.bn, m.bn.act, c, out=out)


def espnet_train_forward(model, input):
    ops.require_cuda(input, "ESPNet")
    if input.dtype != torch.float32 or not input.is_contiguous():
        input = input.float().contiguous()
    n, _, H, W = input.shape
    if (H | W) & 7:
        raise ValueError("ESPNet: input height and width must be multiples of 8, got %dx%d" % (H, W))
    dt = ops.compute_dtype(input)
    tape = T.Tape(model.__dict__.get("_esn_buckets"), device=input.device)
    enc = model.encoder
    classes = model.classifier.out_channels
    img = T.V(input)

    # ---- encoder (ESPNet.py:283-318)
    inp1 = enc.sample1(input)                      # avg-pool pyramid of the image: no parameters, no gradient
    inp2 = enc.sample1(inp1)
    c1 = enc.level1.conv.out_channels
    cat0 = _new(input, c1 + 3, H // 2, W // 2, dt)                  # [level1 | inp1]
    # CBR = conv -> BN -> PReLU on the 16 conv channels alone (ESPNet.py:17-40), written into its concat slice
    o0 = _convT(enc.level1.conv).forward(tape, img, need_dx=False, dtype=dt)
    _br(tape, enc.level1.bn, enc.level1.act, o0, out=cat0.slice(0, c1))
    ops.affine_act(inp1, None, None, None, ACT_NONE, out=cat0.t[:, c1:c1 + 3])
    o0_cat = _br(tape, enc.b1.bn, enc.b1.act, cat0)

    w2 = enc.level2_0.bn.num_features
    cat1 = _new(input, 2 * w2 + 3, H // 4, W // 4, dt)              # [output1 | output1_0 | inp2]
    o1_0 = _down(tape, enc.level2_0, o0_cat, dt, out=cat1.slice(w2, 2 * w2))
    o1 = o1_0
    for i, blk in enumerate(enc.level2):
        o1 = _esp(tape, blk, o1, dt, out=cat1.slice(0, w2) if i == len(enc.level2) - 1 else None)
    ops.affine_act(inp2, None, None, None, ACT_NONE, out=cat1.t[:, 2 * w2:2 * w2 + 3])
    o1_cat = _br(tape, enc.b2.bn, enc.b2.act, cat1)

    w3 = enc.level3_0.bn.num_features
    cat2 = _new(input, 2 * w3, H // 8, W // 8, dt)                  # [output2_0 | output2]
    o2_0 = _down(tape, enc.level3_0, o1_cat, dt, out=cat2.slice(0, w3))
    o2 = o2_0
    for i, blk in enumerate(enc.level3):
        o2 = _esp(tape, blk, o2, dt, out=cat2.slice(w3, 2 * w3) if i == len(enc.level3) - 1 else None)
    o2_cat = _br(tape, enc.b3.bn, enc.b3.act, cat2)

    # ---- decoder (ESPNet.py:372-385)
    s = T.BNActT(model.br, ACT_NONE).forward(tape, _convT(enc.classifier.conv).forward(tape, o2_cat))
    cat_d = _new(input, 2 * classes, H // 4, W // 4, dt)            # [output1_C | output2_c]
    _convT(model.up_l3[0]).forward(tape, s, out=cat_d.slice(classes, 2 * classes))
    _convT(model.level3_C.conv).forward(tape, o1_cat, out=cat_d.slice(0, classes))
    comb = _br(tape, model.combine_l2_l3[0].bn, model.combine_l2_l3[0].act, cat_d)
    comb = _esp(tape, model.combine_l2_l3[1], comb, dt)
    up = _convT(model.up_l2[0]).forward(tape, comb)
    cat_e = _new(input, classes + c1 + 3, H // 2, W // 2, dt)       # [comb | output0_cat] (ESPNet.py:382)
    _br(tape, model.up_l2[1].bn, model.up_l2[1].act, up, out=cat_e.slice(0, classes))
    T.copy_into(tape, o0_cat, cat_e.slice(classes, classes + c1 + 3))
    y = _convT(model.conv.conv).forward(tape, cat_e)
    y = _br(tape, model.conv.bn, model.conv.act, y)
    scores = T.V(ops.new_act(n, classes, H, W, dt, input.device, c_alloc=(classes + 7) // 8 * 8))
    _convT(model.classifier).forward(tape, y, out=scores)
    logits, holder = T.bilinear_logits(tape, scores, H, W, torch.float32)     # same size: NHWC -> NCHW fp32 logits
    return logits, tape, holder
