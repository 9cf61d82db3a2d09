"""Train-mode forward of ESPNetv2 (EESPNet_Seg) on the training kernels (esn/train.py): same graph as the
reference's EESPNet_Seg.forward / EESPNet.forward(seg=True) (model/ESPNet_v2/SegmentationModel.py:59-77,
Model.py:236-287) with batch-statistics BatchNorm and Dropout2d, recording the backward on a tape.  Concats are
channel slices of one buffer (forward and gradient); grouped 1x1 convs are one dense conv per group."""
import torch

from esn import ops
from esn import train as T
from esn._lib import ACT_NONE, ACT_PRELU


def _convT(conv):
    t = conv.__dict__.get("_esn_T")
    if t is None:
        grouped = conv.groups > 1 and not (conv.groups == conv.in_channels == conv.out_channels)
        t = T.GroupedConvT(conv) if grouped else T.ConvT(conv)
        conv.__dict__["_esn_T"] = t
    return t


def _conv_padded(tape, conv, x, cin_pad, classes):
    """1x1 conv onto `classes` scores over a zero-padded input width, on the tensor cores in all three directions (forward,
    input gradient, weight gradient): x lives in a buffer cin_pad channels wide with a zero tail, the result in a 32-channel
    one (zero weight rows) of which the first `classes` are the scores.  The CUDA-core kernel ran these odd-channel convs
    (147 -> 19, 51 -> 19 and their input gradients at 1/4 and 1/2 resolution) at 1.8-3.9 TFLOP/s: 24 of 161 ms per step."""
    t = conv.__dict__.get("_esn_T_pad")
    if t is None or t.cin_pad != cin_pad:
        t = T.ConvT(conv, cin_pad, 32)
        conv.__dict__["_esn_T_pad"] = t
    n, _, h, w = x.t.shape
    wide = T.V(ops.new_act(n, 32, h, w, x.t.dtype, x.t.device))
    t.forward(tape, x, out=wide)
    return wide.slice(0, classes)


def _conv(tape, conv, x, out=None, need_dx=True, dtype=None):
    t = _convT(conv)
    if isinstance(t, T.GroupedConvT):
        return t.forward(tape, x, out=out)
    return t.forward(tape, x, out=out, need_dx=need_dx, dtype=dtype)


def _cbr(tape, m, x, out=None, need_dx=True, dtype=None):
    return T.BNActT(m.bn, ACT_PRELU, m.act).forward(tape, _conv(tape, m.conv, x, need_dx=need_dx, dtype=dtype), out=out)


def _cb(tape, m, x, out=None):
    return T.BNActT(m.bn, ACT_NONE).forward(tape, _conv(tape, m.conv, x), out=out)


def _br(tape, m, x, out=None):
    return T.BNActT(m.bn, ACT_PRELU, m.act).forward(tape, x, out=out)


def _eesp(tape, m, x, out=None):
    """EESP (Model.py:15-99)."""
    o1 = _cbr(tape, m.proj_1x1, x)
    n_, n, h, w = o1.t.shape
    k = len(m.spp_dw)
    ho, wo = (h - 1) // m.stride + 1, (w - 1) // m.stride + 1
    cat = T.V(ops.new_act(n_, n * k, ho, wo, o1.t.dtype, o1.t.device))
    prev = None
    for i, dw in enumerate(m.spp_dw):
        if i == 0:
            prev = _conv(tape, dw.conv, o1, out=cat.slice(0, n))
        else:
            prev = T.add(tape, _conv(tape, dw.conv, o1), prev, out=cat.slice(i * n, (i + 1) * n))
    e = _cb(tape, m.conv_1x1_exp, _br(tape, m.br_after_cat, cat), out=out if (m.stride == 2 and m.downAvg) else None)
    if m.stride == 2 and m.downAvg:
        return e
    if e.t.shape == x.t.shape:
        e = T.add(tape, e, x)
    return T.BNActT(None, ACT_PRELU, m.module_act).forward(tape, e, out=out)


def _down(tape, m, x, img_levels):
    """DownSampler with input reinforcement (Model.py:102-147)."""
    n_, cin, h, w = x.t.shape
    cout = m.act.num_parameters
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    cat = T.V(ops.new_act(n_, cout, ho, wo, x.t.dtype, x.t.device))
    T.avgpool3x3s2(tape, x, out=cat.slice(0, cin))
    _eesp(tape, m.eesp, x, out=cat.slice(cin, cout))
    img = next(t for t in img_levels if t.shape[2] == ho)
    r = _cbr(tape, m.inp_reinf[0], T.V(img), need_dx=False, dtype=x.t.dtype)
    r = _cb(tape, m.inp_reinf[1], r)
    return T.BNActT(None, ACT_PRELU, m.act).forward(tape, T.add(tape, cat, r))


def _psp(tape, m, x):
    """PSPModule (cnn_utils.py:11-25)."""
    n_, c, h, w = x.t.shape
    k = len(m.stages)
    cat = T.V(ops.new_act(n_, c * (k + 1), h, w, x.t.dtype, x.t.device))
    T.copy_into(tape, x, cat.slice(0, c))
    f = x
    for i, stage in enumerate(m.stages):
        f = T.avgpool3x3s2(tape, f)
        T.bilinear(tape, _conv(tape, stage.conv, f), h, w, True, out=cat.slice((i + 1) * c, (i + 2) * c))
    return _cbr(tape, m.project, cat)


def espnetv2_train_forward(model, input, loss=None):
    ops.require_cuda(input, "EESPNet_Seg")
    if input.dtype != torch.float32 or not input.is_contiguous():
        input = input.float().contiguous()
    n, _, H, W = input.shape
    if (H | W) & 15:
        raise ValueError("EESPNet_Seg: input height and width must be multiples of 16, got %dx%d" % (H, W))
    dt, dev = ops.compute_dtype(input), input.device
    tape = T.Tape(model.__dict__.get("_esn_buckets"), device=input.device)
    net = model.net

    # image pyramid for the input reinforcement (no parameters, no gradient)
    levels, im = [], input
    for _ in range(4):
        y = ops.new_act(n, 3, (im.shape[2] - 1) // 2 + 1, (im.shape[3] - 1) // 2 + 1, torch.float32, dev, c_alloc=4)
        im = ops.avgpool3x3s2(im if (ops.is_nhwc(im) or im.is_contiguous()) else im.contiguous(), y)
        levels.append(im)

    l1 = _cbr(tape, net.level1, T.V(input), need_dx=False, dtype=dt)
    l2 = _down(tape, net.level2_0, l1, levels)
    l3 = _down(tape, net.level3_0, l2, levels)
    for layer in net.level3:
        l3 = _eesp(tape, layer, l3)
    l4 = _down(tape, net.level4_0, l3, levels)
    for layer in net.level4:
        l4 = _eesp(tape, layer, l4)

    def cat_buffer(c, like, c_alloc=None):
        return T.V(ops.new_act(n, c, like.t.shape[2], like.t.shape[3], dt, dev, c_alloc=c_alloc or (c + 7) // 8 * 8, zero=True))

    # padded widths: a multiple of 64 above 64 channels (the input gradient then has a TMA-staged epilogue: 160 output channels
    # ran the element-wise one, 3.1 ms), of 16 below
    pad16 = lambda c: (c + 63) // 64 * 64 if c > 64 else (c + 15) // 16 * 16
    # bf16: the two odd-channel 1x1 convs onto the class scores run over zero-padded widths on the tensor cores
    tc_head = dt == torch.bfloat16 and model.project_l2.conv.bias is None and model.project_l1[1].conv.bias is None \
        and model.project_l2.conv.kernel_size == (1, 1) and model.project_l1[1].conv.kernel_size == (1, 1)

    c3 = l3.t.shape[1]
    cat3 = cat_buffer(2 * c3, l3)
    T.copy_into(tape, l3, cat3.slice(0, c3))
    T.bilinear(tape, _cbr(tape, model.proj_L4_C, l4), l3.t.shape[2], l3.t.shape[3], True, out=cat3.slice(c3, 2 * c3))
    m3 = _psp(tape, model.pspMod[1], _eesp(tape, model.pspMod[0], cat3))
    m3 = T.dropout(tape, m3, model.project_l3[0].p, per_channel=True)
    s3 = _br(tape, model.act_l3, _conv(tape, model.project_l3[1].conv, m3))
    classes = s3.t.shape[1]

    c2 = l2.t.shape[1]
    tc_head = tc_head and classes <= 32 and model.project_l2.conv.out_channels == classes
    cat2 = cat_buffer(c2 + classes, l2, pad16(c2 + classes) if tc_head else None)
    T.copy_into(tape, l2, cat2.slice(0, c2))
    T.bilinear(tape, s3, l2.t.shape[2], l2.t.shape[3], True, out=cat2.slice(c2, c2 + classes))
    if tc_head:
        m2 = _br(tape, model.project_l2, _conv_padded(tape, model.project_l2.conv, cat2, pad16(c2 + classes), classes))
    else:
        m2 = _cbr(tape, model.project_l2, cat2)

    c1 = l1.t.shape[1]
    cat1 = cat_buffer(c1 + classes, l1)
    T.copy_into(tape, l1, cat1.slice(0, c1))
    T.bilinear(tape, m2, l1.t.shape[2], l1.t.shape[3], True, out=cat1.slice(c1, c1 + classes))
    if tc_head and model.project_l1[0].p > 0.0 and (c1 + classes) % 4:
        # (the per-element dropout path, taken for channel counts that are not multiples of 4, can write a padded buffer)
        d1 = T.dropout(tape, cat1, model.project_l1[0].p, per_channel=True, c_alloc=pad16(c1 + classes))
        scores = _conv_padded(tape, model.project_l1[1].conv, d1, pad16(c1 + classes), classes)
    else:
        d1 = T.dropout(tape, cat1, model.project_l1[0].p, per_channel=True)
        scores = T.V(ops.new_act(n, classes, l1.t.shape[2], l1.t.shape[3], dt, dev, c_alloc=32))
        _conv(tape, model.project_l1[1].conv, d1, out=scores)
    # fp32 logits, or the loss sums of the fused close when called from EESPNet_Seg.fused_loss (esn_bilinear_ce)
    return T.bilinear_close(tape, scores, H, W, loss, align_corners=True)
