"""Train-mode forward of DABNet on the training kernels (esn/train.py): same graph as the
reference's DABNet.forward (model/DABNet.py:160-183) with batch-statistics BatchNorm, recording the
backward on a tape.  Concats are channel slices of one buffer (forward and gradient)."""
import torch

from esn import ops
from esn import train as T
from esn._lib import ACT_NONE, ACT_PRELU


def _convT(conv, cin_pad=None, cout_pad=None):
    key = "_esn_T%s_%s" % (cin_pad, cout_pad)
    t = conv.__dict__.get(key)
    if t is None:
        t = T.ConvT(conv, cin_pad, cout_pad)
        conv.__dict__[key] = t
    return t


def _tc_channels(c):
    for v in (16, 32, 64):
        if c <= v:
            return v
    return (c + 63) // 64 * 64


def _bnprelu(tape, m, x, out=None):
    return T.BNActT(m.bn, ACT_PRELU, m.acti).forward(tape, x, out=out)


def _conv(tape, m, x, out=None, need_dx=True, dtype=None):
    """DABNet `Conv` wrapper: conv (bias=False) [+ BNPReLU]."""
    if not m.bn_acti:
        return _convT(m.conv).forward(tape, x, out=out, need_dx=need_dx, dtype=dtype)
    y = _convT(m.conv).forward(tape, x, need_dx=need_dx, dtype=dtype)
    return _bnprelu(tape, m.bn_prelu, y, out=out)


def _dab_module(tape, m, xin, out=None):
    b = _bnprelu(tape, m.bn_relu_1, xin)
    t = _conv(tape, m.conv3x3, b)
    v1 = _conv(tape, m.dconv1x3, _conv(tape, m.dconv3x1, t))
    v2 = _conv(tape, m.ddconv1x3, _conv(tape, m.ddconv3x1, t))
    s = T.add(tape, v1, v2)
    w = _bnprelu(tape, m.bn_relu_2, s)
    return _convT(m.conv1x1.conv).forward(tape, w, out=out, residual=xin)      # output + input


def _down(tape, m, x, out, dt):
    """DownSamplingBlock: conv3x3 s2 [|| maxpool] -> BNPReLU, written into `out` (a V slice)."""
    n, _, h, w = x.t.shape
    pre = T.V(ops.new_act(n, m.nOut, h // 2, w // 2, dt, x.t.device))
    nc = m.conv3x3.conv.out_channels
    # bf16: run the strided conv (odd channel counts 35->29, 131->128) on the tensor cores over the
    # zero-padded width of the concat buffer; extra output channels are overwritten by the pool branch
    cin_p, cout_p = _tc_channels(m.nIn), (nc + 7) // 8 * 8
    padded = (dt == torch.bfloat16 and x.t.stride(3) >= cin_p and (cin_p != m.nIn or cout_p != nc)
              and (cout_p == nc or m.nIn < m.nOut) and (h | w) % 2 == 0)
    ct = _convT(m.conv3x3.conv, cin_p, cout_p) if padded else _convT(m.conv3x3.conv)
    if m.nIn < m.nOut:
        ct.forward(tape, x, out=pre.slice(0, cout_p if padded else nc))
        T.maxpool2x2(tape, x, pre.slice(nc, m.nOut))
    else:
        ct.forward(tape, x, out=pre)
    return _bnprelu(tape, m.bn_prelu, pre, out=out)


def dabnet_train_forward(model, input, loss=None):
    ops.require_cuda(input, "DABNet")
    if input.dtype != torch.float32 or not input.is_contiguous():
        input = input.float().contiguous()
    dt = ops.compute_dtype(input)
    dev = input.device
    n, _, H, W = input.shape
    tape = T.Tape(model.__dict__.get("_esn_buckets"), device=input.device)     # data-parallel gradient buckets, if attached

    d1 = model.down_1(input)          # input-injection pyramid: no parameters, no gradient needed
    d2 = model.down_1(d1)
    d3 = model.down_1(d2)

    def cat_buffer(c, like):
        # every one of the c channels is written by a producer; the <= 7 pad lanes of the 16-byte-aligned pixel stride are read
        # (full vectors) but never used by the BatchNorm kernels, so the buffer needs no fill
        hh, ww = like.shape[2:]
        return T.V(ops.new_act(n, c, hh, ww, dt, dev, c_alloc=(c + 7) // 8 * 8))

    def padded_out(key, c, like, c_alloc):
        # BNPReLU output read by a tensor-core conv over its zero-padded width (35 -> 64, 131 -> 192, 259 -> 320): the BatchNorm
        # kernel only ever writes the c real channels, so the zero tail is written ONCE and the buffer kept on the model
        # (a 134 / 100 / 42 MB fill per step otherwise)
        hh, ww = like.shape[2:]
        cache = model.__dict__.setdefault("_esn_train_bufs", {})
        shape = (n, hh, ww, dt, str(dev))
        held = cache.get(key)
        if held is None or held[0] != shape:          # one buffer per role: a new batch shape replaces the old buffer
            held = (shape, ops.new_act(n, c, hh, ww, dt, dev, c_alloc=c_alloc, zero=True))
            cache[key] = held
        return T.V(held[1])

    x = T.V(input)
    y = _conv(tape, model.init_conv[0], x, need_dx=False, dtype=dt)
    y = _conv(tape, model.init_conv[1], y)
    cat0 = cat_buffer(35, d1)
    _conv(tape, model.init_conv[2], y, out=cat0.slice(0, 32))
    ops.affine_act(d1, None, None, None, ACT_NONE, out=cat0.t[:, 32:35])
    c0 = _bnprelu(tape, model.bn_prelu_1, cat0, out=padded_out("c0", 35, d1, 64))

    cat1 = cat_buffer(131, d2)
    y = _down(tape, model.downsample_1, c0, cat1.slice(64, 128), dt)
    blocks = list(model.DAB_Block_1)
    for i, blk in enumerate(blocks):
        y = _dab_module(tape, blk, y, out=cat1.slice(0, 64) if i == len(blocks) - 1 else None)
    ops.affine_act(d2, None, None, None, ACT_NONE, out=cat1.t[:, 128:131])
    c1 = _bnprelu(tape, model.bn_prelu_2, cat1, out=padded_out("c1", 131, d2, 192))

    cat2 = cat_buffer(259, d3)
    y = _down(tape, model.downsample_2, c1, cat2.slice(128, 256), dt)
    blocks = list(model.DAB_Block_2)
    for i, blk in enumerate(blocks):
        y = _dab_module(tape, blk, y, out=cat2.slice(0, 128) if i == len(blocks) - 1 else None)
    ops.affine_act(d3, None, None, None, ACT_NONE, out=cat2.t[:, 256:259])
    classes = model.classifier[0].conv.out_channels
    hh, ww = d3.shape[2:]
    if dt == torch.bfloat16:
        # 1x1 classifier 259 -> 19 on the tensor cores in all three directions (forward, input gradient, weight gradient):
        # the BNPReLU output lives in a 320-channel buffer with a zero tail, the scores in a 32-channel one (zero weight rows)
        c2 = _bnprelu(tape, model.bn_prelu_3, cat2, out=padded_out("c2", 259, d3, 320))
        wide = T.V(ops.new_act(n, 32, hh, ww, dt, dev))
        _convT(model.classifier[0].conv, 320, 32).forward(tape, c2, out=wide)
        scores = wide.slice(0, classes)
    else:
        c2 = _bnprelu(tape, model.bn_prelu_3, cat2)
        scores = T.V(ops.new_act(n, classes, hh, ww, dt, dev, c_alloc=32))
        _conv(tape, model.classifier[0], c2, out=scores)
    # fp32 logits, or the loss sums of the fused close when called from DABNet.fused_loss (esn_bilinear_ce)
    return T.bilinear_close(tape, scores, H, W, loss)
