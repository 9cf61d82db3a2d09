"""TEST INFRASTRUCTURE ONLY: a CPU model of the C ABI (include/esn.h) at the level of the structs the host passes.

tests/host_emulation.py replaces the Python wrappers; this file goes one level down and keeps ALL of ``esn.ops``
running -- routing between the tcgen05 / sliced / direct conv paths, the shape gates, weight packing for each kernel
family (``w_direct``, ``w_umma``, ``w_umma_scaled``, the phase-fused transposed conv), epilogue parameter blocks,
descriptors of channel slices -- and intercepts only the foreign call itself (``ops._call``).  Each entry point is
modelled from the raw arguments: tensors are rebuilt from EsnTensor descriptors over the caller's memory (host memory
here), weights are decoded from the packed layouts documented in esn.h, arithmetic is fp32 with the result rounded to
the output dtype, as the kernels do.  A bf16 forward of a model under this emulation therefore checks everything the
host contributes to a launch, without a GPU.

Nothing in the package imports this file.  Kernel arithmetic itself is checked on the device (``-m gpu`` tests).
"""
import contextlib
import ctypes as C

import torch
import torch.nn.functional as F

from esn import ops
from esn import _lib as L

_DT = {L.ESN_F32: (torch.float32, 4), L.ESN_BF16: (torch.bfloat16, 2)}


def _buf(ptr, count, dtype, esize):
    if count <= 0:
        return torch.empty(0, dtype=dtype)
    assert ptr, "null pointer passed for a non-empty buffer"
    return torch.frombuffer((C.c_char * (count * esize)).from_address(ptr), dtype=dtype)


ALLOCS = []       # (begin, end) byte ranges of every activation buffer handed out by ops.new_act under the emulation


def _in_bounds(ptr, nbytes):
    """A descriptor that starts inside a tracked activation buffer must end inside it (a widened channel slice that runs
    past the last pixel of its buffer would be an out-of-bounds read on the device)."""
    for lo, hi in ALLOCS:
        if lo <= ptr < hi:
            assert ptr + nbytes <= hi, "descriptor overruns its buffer by %d bytes" % (ptr + nbytes - hi)
            return


def tensor(d):
    """EsnTensor -> logical (N,C,H,W) view over the caller's memory."""
    dtype, es = _DT[d.dtype]
    if d.layout == L.ESN_NHWC:
        assert d.c_stride >= d.c > 0
        span = (d.n * d.h * d.w - 1) * d.c_stride + d.c
        _in_bounds(d.ptr, span * es)
        return _buf(d.ptr, span, dtype, es).as_strided((d.n, d.c, d.h, d.w), (d.h * d.w * d.c_stride, 1, d.w * d.c_stride, d.c_stride))
    return _buf(d.ptr, d.n * d.c * d.h * d.w, dtype, es).view(d.n, d.c, d.h, d.w)


def vec(ptr, n):
    return None if not ptr else _buf(ptr, n, torch.float32, 4)


def _act(v, act, alpha):
    if act == L.ACT_RELU:
        return v.clamp_min(0)
    if act == L.ACT_PRELU:
        return v.clamp_min(0) + alpha.view(1, -1, 1, 1) * v.clamp_max(0)
    return v


def epilogue(acc, ep):
    c = acc.shape[1]
    sc, sh, al = vec(ep.scale, c), vec(ep.shift, c), vec(ep.alpha, c)
    res = tensor(ep.residual).float() if ep.residual.ptr else None
    one = lambda v, dflt: dflt if v is None else v.view(1, -1, 1, 1)
    if ep.flags & 2:                                                    # ESN_EP_RESIDUAL_FIRST
        return _act((acc + res) * one(sc, 1.0) + one(sh, 0.0), ep.act, al)
    v = acc * one(sc, 1.0) + one(sh, 0.0)
    if res is None:
        return _act(v, ep.act, al)
    if ep.flags & 1:                                                    # ESN_EP_ACT_BEFORE_RESIDUAL
        return _act(_act(v, ep.act, al) + res, ep.act, al)
    return _act(v + res, ep.act, al)


def store(y, v):
    assert tuple(y.shape) == tuple(v.shape), (tuple(y.shape), tuple(v.shape))
    y.copy_(v.to(y.dtype))


def _finite(x, what):
    assert torch.isfinite(x).all(), "%s reads non-finite values (uninitialised channel tail?)" % what
    return x


# ------------------------------------------------------------------------------------------------ conv family
def _conv_core(p, x, wt, cout):
    """wt: (Cout, Cin/groups, kh, kw) fp32 (transposed: the adjoint conv's (Cout, Cin, kh, kw))."""
    if p.transposed:
        n, _, h, w = x.shape
        yh = tensor(p.y).shape[2]
        out_pad = yh - ((h - 1) * p.stride - 2 * p.pad_h + p.dil_h * (p.kh - 1) + 1)
        wtt = wt if (p.groups > 1) else wt.permute(1, 0, 2, 3)
        return F.conv_transpose2d(x, wtt, None, p.stride, (p.pad_h, p.pad_w), out_pad, p.groups, (p.dil_h, p.dil_w))
    return F.conv2d(x, wt, None, p.stride, (p.pad_h, p.pad_w), (p.dil_h, p.dil_w), p.groups)


def esn_conv2d_direct(ref):
    p = ref._obj
    x, y = _finite(tensor(p.x).float(), "esn_conv2d_direct"), tensor(p.y)
    cin_g, cout, taps = p.x.c // p.groups, p.y.c, p.kh * p.kw
    wt = _buf(p.w, taps * cin_g * cout, torch.float32, 4).view(p.kh, p.kw, cin_g, cout).permute(3, 2, 0, 1)   # [tap][Cin/g][Cout]
    store(y, epilogue(_conv_core(p, x, wt, cout), p.ep))
    return 0


def esn_conv2d_umma(ref):
    p = ref._obj
    assert p.x.dtype == L.ESN_BF16 and p.y.dtype == L.ESN_BF16 and p.groups == 1
    assert p.x.c_stride % 8 == 0 and p.y.c_stride % 8 == 0 and p.x.ptr % 16 == 0 and p.y.ptr % 16 == 0, "TMA alignment"
    x, y = _finite(tensor(p.x).float(), "esn_conv2d_umma"), tensor(p.y)
    cin, taps = p.x.c, p.kh * p.kw
    assert cin % 16 == 0, cin
    wp = _buf(p.w, taps * p.cout_pad * cin, torch.bfloat16, 2).float().view(p.kh, p.kw, p.cout_pad, cin)       # [tap][Cout_pad][Cin]
    if p.transposed == 2:
        # phase-fused ConvTranspose2d(3, s2, p1, op1): 2x2 taps, 4*Cout outputs ordered (row parity, column parity, c)
        cout = p.y.c
        assert p.cout_pad == 4 * cout and (p.kh, p.kw, p.stride) == (2, 2, 1) and p.y.h == 2 * p.x.h and p.y.w == 2 * p.x.w
        acc = F.conv2d(F.pad(x, (0, 1, 0, 1)), wp.permute(2, 3, 0, 1))                                         # (N, 4*Cout, H, W)
        c4 = 4 * cout
        sc, sh, al = vec(p.ep.scale, c4), vec(p.ep.shift, c4), vec(p.ep.alpha, c4)
        v = acc * (1.0 if sc is None else sc.view(1, -1, 1, 1)) + (0.0 if sh is None else sh.view(1, -1, 1, 1))
        v = _act(v, p.ep.act, al)
        n, _, h, w = v.shape
        v = v.view(n, 2, 2, cout, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, cout, 2 * h, 2 * w)               # pixel shuffle
        store(y, v)
        return 0
    assert p.cout_pad <= 256 and p.y.c <= p.cout_pad
    wt = wp[:, :, :p.y.c].permute(2, 3, 0, 1)
    store(y, epilogue(_conv_core(p, x, wt, p.y.c), p.ep))
    return 0


def esn_conv2d_umma_dual(ref):
    """include/esn.h EsnConvDual: y = epilogue(conv) (stored when store_y), y2 = act2(bf16(y) * scale2 + shift2)."""
    d = ref._obj
    p = d.conv
    assert p.x.dtype == L.ESN_BF16 and p.y.dtype == L.ESN_BF16 and d.y2.dtype == L.ESN_BF16 and p.groups == 1 and not p.transposed
    assert p.x.c_stride % 8 == 0 and p.y.c_stride % 8 == 0 and d.y2.c_stride % 8 == 0 and d.y2.ptr % 16 == 0, "TMA alignment"
    cout = p.y.c
    assert cout % 8 == 0 and (cout <= 64 or cout % 64 == 0) and p.cout_pad <= 256, "staged epilogue only"
    assert (d.y2.n, d.y2.h, d.y2.w, d.y2.c) == (p.y.n, p.y.h, p.y.w, p.y.c)
    x = _finite(tensor(p.x).float(), "esn_conv2d_umma_dual")
    cin, taps = p.x.c, p.kh * p.kw
    assert cin in (16, 32, 64) or cin % 64 == 0, cin
    wp = _buf(p.w, taps * p.cout_pad * cin, torch.bfloat16, 2).float().view(p.kh, p.kw, p.cout_pad, cin)
    v = epilogue(_conv_core(p, x, wp[:, :, :cout].permute(2, 3, 0, 1), cout), p.ep).to(torch.bfloat16)
    if d.store_y:
        store(tensor(p.y), v)
    sc, sh, al = vec(d.scale2, cout), vec(d.shift2, cout), vec(d.alpha2, cout)
    one = lambda t, dflt: dflt if t is None else t.view(1, -1, 1, 1)
    store(tensor(d.y2), _act(v.float() * one(sc, 1.0) + one(sh, 0.0), d.act2, al))
    return 0


def esn_concat_tail(ref, tail_c):
    """include/esn.h: y[..., :c) = act(x*scale + shift), y[..., c:tail_c) = 0 (y = view of a concat buffer at the first
    injected channel; the zero channels lie in the buffer's pixel padding, behind the view's logical channels)."""
    p = ref._obj
    tail_c = int(getattr(tail_c, "value", tail_c))
    assert p.x.dtype == L.ESN_F32 and p.x.c <= 4 and p.x.c_stride == 4 and (tail_c & (tail_c - 1)) == 0 and 8 <= tail_c <= p.y.c_stride
    c = p.x.c
    sc, sh, al = vec(p.ep.scale, c), vec(p.ep.shift, c), vec(p.ep.alpha, c)
    one = lambda t, dflt: dflt if t is None else t.view(1, -1, 1, 1)
    v = _act(tensor(p.x).float() * one(sc, 1.0) + one(sh, 0.0), p.ep.act, al)
    wide = L.EsnTensor.from_buffer_copy(p.y)
    wide.c = tail_c
    yw = tensor(wide)
    yw.zero_()
    store(yw[:, :c], v)
    return 0


def esn_conv_pair_umma(ref):
    p = ref._obj
    x, y = _finite(tensor(p.x).float(), "esn_conv_pair_umma"), tensor(p.y)
    c, d = p.x.c, p.dilation
    cp = (c + 15) // 16 * 16
    assert p.taps == 3 and c in (16, 64)
    w1 = _buf(p.w1, 3 * cp * c, torch.bfloat16, 2).float().view(3, cp, c)[:, :c]                               # 3x1: [tap][Cout][Cin]
    w2 = _buf(p.w2, 3 * cp * c, torch.bfloat16, 2).float().view(3, cp, c)[:, :c]                               # 1x3, scale folded in
    t = F.conv2d(x, w1.permute(1, 2, 0).unsqueeze(3), None, 1, (d, 0), (d, 1))
    t = epilogue(t, p.ep1).to(torch.bfloat16).float()                   # the intermediate row lives in shared memory as bf16
    acc = F.conv2d(t, w2.permute(1, 2, 0).unsqueeze(2), None, 1, (0, d), (1, d))
    assert not p.ep2.scale, "the pair kernel takes the second scale inside w2"
    store(y, epilogue(acc, p.ep2))
    return 0


def esn_stem_conv3x3s2(ref):
    p = ref._obj
    assert p.x.layout == L.ESN_NCHW and p.x.dtype == L.ESN_F32 and p.x.c == 3
    x, y = tensor(p.x), tensor(p.y)
    wt = _buf(p.w, 27 * p.cconv, torch.float32, 4).view(3, 3, 3, p.cconv).permute(3, 2, 0, 1)
    acc = F.conv2d(x, wt, None, 2, 0 if p.with_pool & 256 else 1)
    if p.with_pool & 3 == 1:
        acc = torch.cat([acc, F.max_pool2d(x, 2, 2)], 1)
    elif p.with_pool & 3 == 2:
        acc = torch.cat([acc, F.max_pool2d(x, 3, 2, 1)], 1)
    store(y, epilogue(acc, p.ep))
    return 0


# ------------------------------------------------------------------------------------------------ pointwise / pools / resize
def _pool(fn):
    def run(ref):
        p = ref._obj
        store(tensor(p.y), epilogue(fn(_finite(tensor(p.x).float(), "pool/affine")), p.ep))
        return 0
    return run


def _maxpool2(ref):
    """MaxPool2d(2, 2); when y has ceil(h/2) x ceil(w/2) pixels (odd input), zero padding behind the pooled map (include/esn.h)."""
    p = ref._obj
    v = F.max_pool2d(_finite(tensor(p.x).float(), "pool"), 2, 2)
    if (p.y.h, p.y.w) != tuple(v.shape[2:]):
        assert (p.y.h, p.y.w) == ((p.x.h + 1) // 2, (p.x.w + 1) // 2)
        v = F.pad(v, [0, p.y.w - v.shape[3], 0, p.y.h - v.shape[2]])
    store(tensor(p.y), epilogue(v, p.ep))
    return 0


def esn_convert_layout(xr, yr):
    store(tensor(yr._obj), tensor(xr._obj).float())
    return 0


def esn_adaptive_avgpool(xr, yr):
    y = tensor(yr._obj)
    store(y, F.adaptive_avg_pool2d(tensor(xr._obj).float(), (y.shape[2], y.shape[3])))
    return 0


def esn_bilinear_nhwc(xr, yr, align):
    y = tensor(yr._obj)
    store(y, F.interpolate(tensor(xr._obj).float(), (y.shape[2], y.shape[3]), mode="bilinear", align_corners=bool(align)))
    return 0


def _head_out(p, logits_f32):
    if p.logits.ptr:
        store(tensor(p.logits), logits_f32)
    if p.mask:
        n, _, h, w = logits_f32.shape
        torch.frombuffer((C.c_char * (n * h * w)).from_address(p.mask), dtype=torch.uint8).view(n, h, w).copy_(
            logits_f32.argmax(1).to(torch.uint8))
    return 0


def esn_head_convt2x2(ref):
    p = ref._obj
    x = tensor(p.x).float()
    wt = _buf(p.w, 4 * p.x.c * 32, torch.float32, 4).view(2, 2, p.x.c, 32)[:, :, :, :p.classes].permute(2, 3, 0, 1)
    return _head_out(p, F.conv_transpose2d(x, wt, vec(p.bias, p.classes), 2))


def esn_head_convt3x3s2_mask(ref):
    """ConvTranspose2d(16, classes, 3, 2, 1, 1) + argmax from the packed mma B fragments (include/esn.h): the weights are decoded
    back from fragment order, so a wrong packing on the host shows up here."""
    p = ref._obj
    assert p.x.dtype == L.ESN_BF16 and p.x.c == 16 and p.x.w % 16 == 0 and p.classes <= 24
    x = tensor(p.x).float()
    frag = _buf(p.wfrag, 9 * 3 * 32 * 2 * 2, torch.bfloat16, 2).view(9, 3, 32, 2, 2).float()
    pairs = ((0, 0, 0, 0), (0, 1, 0, 0), (0, 1, 0, 1), (1, 0, 0, 0), (1, 0, 1, 0), (1, 1, 0, 0), (1, 1, 0, 1), (1, 1, 1, 0), (1, 1, 1, 1))
    w = torch.zeros(16, 24, 3, 3)
    for q, (a, b, dy, dx) in enumerate(pairs):
        for nt in range(3):
            for lane in range(32):
                g, t = lane // 4, lane % 4
                for r in range(2):
                    for e in range(2):
                        w[2 * t + 8 * r + e, nt * 8 + g, a + 1 - 2 * dy, b + 1 - 2 * dx] = frag[q, nt, lane, r, e]
    assert (w[:, p.classes:] == 0).all(), "padded classes must carry zero weights"
    logits = F.conv_transpose2d(x, w[:, :p.classes], vec(p.bias, p.classes), 2, 1, 1)
    n, _, h, ww = logits.shape
    torch.frombuffer((C.c_char * (n * h * ww)).from_address(p.mask), dtype=torch.uint8).view(n, h, ww).copy_(
        logits.argmax(1).to(torch.uint8))
    return 0


def esn_head_convt2x2_mask(ref):
    """ConvTranspose2d(16, classes, 2, 2) + argmax from the packed hi / lo mma B fragments (include/esn.h): the weights are
    decoded back from fragment order (channel 4t + 2r + e in register r of lane (g, t)) and summed hi + lo."""
    p = ref._obj
    assert p.x.dtype == L.ESN_BF16 and p.x.c == 16 and p.x.w % 16 == 0 and p.classes <= 24 and p.x.c_stride % 4 == 0
    x = tensor(p.x).float()
    frag = _buf(p.wfrag, 2 * 4 * 3 * 32 * 2 * 2, torch.bfloat16, 2).view(2, 4, 3, 32, 2, 2).float()
    w = torch.zeros(16, 24, 2, 2)
    for hl in range(2):
        for pos in range(4):
            for nt in range(3):
                for lane in range(32):
                    g, t = lane // 4, lane % 4
                    for r in range(2):
                        for e in range(2):
                            w[4 * t + 2 * r + e, nt * 8 + g, pos // 2, pos % 2] += frag[hl, pos, nt, lane, r, e]
    assert (w[:, p.classes:] == 0).all(), "padded classes must carry zero weights"
    logits = F.conv_transpose2d(x, w[:, :p.classes], vec(p.bias, p.classes), 2)
    n, _, h, ww = logits.shape
    torch.frombuffer((C.c_char * (n * h * ww)).from_address(p.mask), dtype=torch.uint8).view(n, h, ww).copy_(
        logits.argmax(1).to(torch.uint8))
    return 0


def esn_bilinear_ce(ref):
    """Bilinear up-sampling + weighted CE sums + d sums[0] / d scores (include/esn.h), by torch autograd in fp32."""
    p = ref._obj
    x = tensor(p.scores).float().contiguous().requires_grad_(True)
    n, c, h, w = x.shape
    H, W = p.out_h, p.out_w
    tgt = _buf(p.target, n * H * W, torch.int64, 8).view(n, H, W).clone()
    tgt[(tgt < 0) | (tgt >= c)] = p.ignore_label
    wt = vec(p.weight, c)
    with torch.enable_grad():
        logits = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=bool(p.align_corners))
        loss = F.cross_entropy(logits, tgt, wt, ignore_index=p.ignore_label, reduction="sum")
        (g,) = torch.autograd.grad(loss, x)
    valid = tgt != p.ignore_label
    sums = _buf(p.sums, 2, torch.float32, 4)
    sums[0] += loss.detach()
    sums[1] += (wt[tgt[valid]].sum() if wt is not None else valid.float().sum())
    _buf(p.dscores.ptr, n * h * w * p.dscores.c_stride, torch.float32, 4).zero_()      # every lane of the pixel stride is written
    tensor(p.dscores).copy_(g)
    return 0


def esn_bottleneck4(ref):
    """ENet's 16-channel RegularBottleneck: y = act(x + act(BN3(W3 . act(BN2(W2 * act(BN1(W1 . x)))))))."""
    p = ref._obj
    assert p.x.dtype == L.ESN_BF16 and p.x.c == 16 and 1 <= p.dilation <= 4
    x, y = _finite(tensor(p.x).float(), "esn_bottleneck4"), tensor(p.y)
    d = p.dilation
    w1 = _buf(p.w1, 64, torch.float32, 4).view(1, 1, 16, 4).permute(3, 2, 0, 1)
    w2 = _buf(p.w2, 144, torch.float32, 4).view(3, 3, 4, 4).permute(3, 2, 0, 1)
    w3 = _buf(p.w3, 64, torch.float32, 4).view(1, 1, 4, 16).permute(3, 2, 0, 1)
    aff = lambda v, s_, b_, c: v * vec(s_, c).view(1, -1, 1, 1) + vec(b_, c).view(1, -1, 1, 1)
    e = _act(aff(F.conv2d(x, w1), p.scale1, p.shift1, 4), p.act, vec(p.alpha1, 4))
    e = _act(aff(F.conv2d(e, w2, padding=d, dilation=d), p.scale2, p.shift2, 4), p.act, vec(p.alpha2, 4))
    e = _act(aff(F.conv2d(e, w3), p.scale3, p.shift3, 16), p.act, vec(p.alpha3, 16))
    store(y, _act(x + e, p.act, vec(p.alpha3, 16)))
    return 0


def esn_head_bilinear(ref):
    p = ref._obj
    x = tensor(p.x).float()[:, :p.classes]
    return _head_out(p, F.interpolate(x, (p.out_h, p.out_w), mode="bilinear", align_corners=bool(p.align_corners)))


def esn_maxpool3x3s2_idx(xr, yr, idx_ptr):
    x, y = tensor(xr._obj).float().contiguous(), tensor(yr._obj)
    v, idx = F.max_pool2d(x, 3, 2, 1, return_indices=True)
    store(y, v)
    n, c, ho, wo = v.shape
    _buf(idx_ptr.value, n * ho * wo * c, torch.int32, 4).view(n, ho, wo, c).copy_(idx.permute(0, 2, 3, 1).to(torch.int32))
    return 0


def esn_max_unpool2x2(ref):
    p = ref._obj
    v, y = tensor(p.v).float().contiguous(), tensor(p.y)
    n, c, h, w = v.shape
    idx = _buf(p.idx, n * h * w * c, torch.int32, 4).view(n, h, w, c).permute(0, 3, 1, 2).contiguous().long()
    up = F.max_unpool2d(v, idx, 2, output_size=(2 * h, 2 * w))
    if p.ext.ptr:
        up = up + tensor(p.ext).float()
    store(y, _act(up, p.act, vec(p.alpha, c)))
    return 0


def esn_dab_dw_pair(ref):
    p = ref._obj
    x, y = tensor(p.x).float(), tensor(p.y)
    c, d = p.x.c, p.dilation
    prm = _buf(p.prm, 27 * c, torch.float32, 4).view(27, c)
    taps, aff, fin = prm[0:12].view(4, 3, c), prm[12:24].view(4, 3, c), prm[24:27]

    def stage(t, i, vertical, dd):
        wt = taps[i].t().reshape(c, 1, 3, 1) if vertical else taps[i].t().reshape(c, 1, 1, 3)
        t = F.conv2d(t, wt, None, 1, (dd, 0) if vertical else (0, dd), (dd, 1) if vertical else (1, dd), c)
        return _act(t * aff[i, 0].view(1, -1, 1, 1) + aff[i, 1].view(1, -1, 1, 1), L.ACT_PRELU, aff[i, 2])
    br = stage(stage(x, 0, True, 1), 1, False, 1) + stage(stage(x, 2, True, d), 3, False, d)
    store(y, _act(br * fin[0].view(1, -1, 1, 1) + fin[1].view(1, -1, 1, 1), L.ACT_PRELU, fin[2]))
    return 0


def esn_gate_bcast(gr, xr, br, yr):
    g, x, y = tensor(gr._obj).float(), tensor(xr._obj).float(), tensor(yr._obj)
    assert xr._obj.dtype == yr._obj.dtype and g.shape[1] == 1
    assert gr._obj.dtype == xr._obj.dtype or (gr._obj.dtype, xr._obj.dtype) == (L.ESN_F32, L.ESN_BF16)     # fp32 gate on bf16 scores
    v = g * x
    if br._obj.ptr:
        assert br._obj.dtype == xr._obj.dtype and (br._obj.h, br._obj.w) == (1, 1)
        v = v + tensor(br._obj).float()
    store(y, v)
    return 0


def esn_weighted_ce(ref):
    """include/esn.h EsnCE: sums[0] += sum w[y]*nll, sums[1] += sum w[y] over pixels with y != ignore; optional
    dlogits = w[y] * (softmax - onehot) * (*gout) / (*gnorm) (unscaled when the scalars are absent)."""
    p = ref._obj
    d = p.logits
    assert d.layout == L.ESN_NCHW
    x = tensor(d).float()
    n, c, h, w = x.shape
    y = _buf(p.target, n * h * w, torch.int64, 8).view(n, h, w)
    wv = vec(p.weight, c) if p.weight else torch.ones(c)
    valid = y != p.ignore_label
    ys = torch.where(valid, y, torch.zeros_like(y))
    logp = torch.log_softmax(x, 1)
    if p.prob_out or p.keep_thresh:      # OHEM: probability of the labelled class (1 where ignored), optional threshold mask
        py = logp.gather(1, ys.unsqueeze(1)).squeeze(1).exp().masked_fill(~valid, 1.0)
        if p.prob_out:
            _buf(p.prob_out, n * h * w, torch.float32, 4).view(n, h, w).copy_(py)
        if p.keep_thresh:
            valid = valid & (py <= _buf(p.keep_thresh, 1, torch.float32, 4)[0])
    wi = wv[ys] * valid.float()
    nll = -logp.gather(1, ys.unsqueeze(1)).squeeze(1)
    sums = _buf(p.sums, 2, torch.float32, 4)
    sums[0] += (wi * nll).sum()
    sums[1] += wi.sum()
    if p.dlogits.ptr:
        g = (logp.exp() - torch.zeros_like(x).scatter_(1, ys.unsqueeze(1), 1.0)) * wi.unsqueeze(1)
        if p.gout:
            g = g * _buf(p.gout, 1, torch.float32, 4)[0]
        if p.gnorm:
            g = g / _buf(p.gnorm, 1, torch.float32, 4)[0]
        store(tensor(p.dlogits), g)
    return 0


def esn_ohem_threshold(prob, n, min_kept, thresh, num_valid, out, workspace):
    """include/esn.h: +inf if min_kept > *num_valid (or min_kept <= 0), else max(thresh, min(n, min_kept)-th smallest prob)."""
    pv = _buf(prob.value, n, torch.float32, 4)
    nv = float(_buf(num_valid.value, 1, torch.float32, 4)[0])
    o = _buf(out.value, 1, torch.float32, 4)
    if min_kept > nv or nv <= 0 or min_kept <= 0:
        o[0] = float("inf")
    else:
        o[0] = max(float(thresh.value if hasattr(thresh, "value") else thresh), float(torch.sort(pv).values[min(n, min_kept) - 1]))
    return 0


def esn_augment_u8(items, n, crop_h, crop_w, mean3, ignore_label, out_img, out_label):
    """include/esn.h EsnAugItem / esn_augment_u8, from the raw structs: oracle.pipeline's cv2 restatement driven by the
    struct's own (rh, rw, scale) -- the resized image is materialised here, the kernel gathers instead."""
    import numpy as np
    from oracle import pipeline as P
    mean = np.array([mean3[0], mean3[1], mean3[2]], dtype=np.float32)
    xo = _buf(out_img.value, n * 3 * crop_h * crop_w, torch.float32, 4).view(n, 3, crop_h, crop_w)
    yo = _buf(out_label.value, n * crop_h * crop_w, torch.int64, 8).view(n, crop_h, crop_w)
    for k in range(n):
        it = items[k]
        img = torch.frombuffer((C.c_char * (it.h * it.w * 3)).from_address(it.img), dtype=torch.uint8).view(it.h, it.w, 3).numpy()
        lab = torch.frombuffer((C.c_char * (it.h * it.w)).from_address(it.label), dtype=torch.uint8).view(it.h, it.w).numpy()
        if it.do_scale:
            xof, xa = P._linear_coeffs(it.rw, it.w, it.scale, True)
            yof, ya = P._linear_coeffs(it.rh, it.h, it.scale, False)
            s = img.astype(np.int32)
            hb = s[:, xof] * xa[:, 0][None, :, None] + s[:, np.minimum(xof + 1, it.w - 1)] * xa[:, 1][None, :, None]
            s0, s1 = hb[np.clip(yof, 0, it.h - 1)], hb[np.clip(yof + 1, 0, it.h - 1)]
            img = ((((ya[:, 0][:, None, None] * (s0 >> 4)) >> 16) + ((ya[:, 1][:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)
            sy = np.minimum(np.floor(np.arange(it.rh) * it.scale).astype(np.int64), it.h - 1)
            sx = np.minimum(np.floor(np.arange(it.rw) * it.scale).astype(np.int64), it.w - 1)
            lab = lab[sy][:, sx]
        x, y = P.train_item(img, lab, None, it.h_off, it.w_off, -1 if it.flip else 1, (crop_h, crop_w), mean, ignore_label)
        xo[k].copy_(torch.from_numpy(x))
        yo[k].copy_(torch.from_numpy(y.astype(np.int64)))
    return 0


def esn_image_u8hwc_to_f32nchw(img, out, n, h, w, mean3, reverse):
    src = torch.frombuffer((C.c_char * (n * h * w * 3)).from_address(img.value), dtype=torch.uint8).view(n, h, w, 3)
    v = src.float() - torch.tensor([mean3[0], mean3[1], mean3[2]], dtype=torch.float32)
    if reverse:
        v = v.flip(3)
    _buf(out.value, n * 3 * h * w, torch.float32, 4).view(n, 3, h, w).copy_(v.permute(0, 3, 1, 2))
    return 0


ENTRY = {
    "esn_conv2d_direct": esn_conv2d_direct, "esn_conv2d_umma": esn_conv2d_umma, "esn_conv2d_umma_dual": esn_conv2d_umma_dual, "esn_concat_tail": esn_concat_tail, "esn_conv_pair_umma": esn_conv_pair_umma,
    "esn_stem_conv3x3s2": esn_stem_conv3x3s2,
    "esn_maxpool2x2_affine_act": _maxpool2,
    "esn_avgpool3x3s2_affine_act": _pool(lambda x: F.avg_pool2d(x, 3, 2, 1)),
    "esn_affine_act": _pool(lambda x: x),
    "esn_convert_layout": esn_convert_layout, "esn_adaptive_avgpool": esn_adaptive_avgpool, "esn_bilinear_nhwc": esn_bilinear_nhwc,
    "esn_head_convt2x2": esn_head_convt2x2, "esn_head_bilinear": esn_head_bilinear,
    "esn_maxpool3x3s2_idx": esn_maxpool3x3s2_idx, "esn_max_unpool2x2": esn_max_unpool2x2, "esn_dab_dw_pair": esn_dab_dw_pair,
    "esn_image_u8hwc_to_f32nchw": esn_image_u8hwc_to_f32nchw, "esn_gate_bcast": esn_gate_bcast, "esn_weighted_ce": esn_weighted_ce,
    "esn_ohem_threshold": esn_ohem_threshold, "esn_augment_u8": esn_augment_u8,
    "esn_head_convt3x3s2_mask": esn_head_convt3x3s2_mask, "esn_head_convt2x2_mask": esn_head_convt2x2_mask, "esn_bilinear_ce": esn_bilinear_ce, "esn_bottleneck4": esn_bottleneck4,
}
CALLS = []        # (entry point, tag) of every emulated launch, for assertions about routing


DECLINE = set()   # entry-point names the model should answer ESN_ERR_UNSUPPORTED for (to exercise the host's next route)


def _call(fn, name, arg_refs, alg_bytes=0, flops=0, tag="", allow_unsupported=False):
    if name in DECLINE and allow_unsupported:      # calls without a fallback route (phase-fused transposed conv) are served
        return False
    CALLS.append((name, tag))
    rc = ENTRY[name](*arg_refs)
    L.check(rc, name)
    return True


def _new_act(n, c, h, w, dtype, device, c_alloc=None, zero=False):
    ca = c if c_alloc is None else c_alloc
    buf = torch.zeros((n, h, w, ca), dtype=dtype, device=device) if zero else torch.full((n, h, w, ca), float("nan"), dtype=dtype, device=device)
    ALLOCS.append((buf.data_ptr(), buf.data_ptr() + buf.numel() * buf.element_size()))
    _KEEP.append(buf)      # keep every buffer alive for the duration of the emulation so that address ranges stay unique
    t = buf.permute(0, 3, 1, 2)
    return t if ca == c else t[:, :c]


_KEEP = []


def _to_nchw(x, dtype=None):
    return x.to(dtype or x.dtype).contiguous()


def _fglo_gate(x, w1, b1, w2, b2, out=None, residual=None):
    # three entry points with device-sized scratch (esn_global_avgpool(+_chunks), esn_fglo_gate, esn_scale_nc): modelled as one
    n, c, h, w = x.shape
    xf = x.float()
    gate = torch.sigmoid(F.linear(F.relu(F.linear(xf.mean(dim=(2, 3)), w1, b1)), w2, b2)).view(n, c, 1, 1)
    if out is None:
        out = ops.new_act(n, c, h, w, x.dtype, x.device)
    v = xf * gate
    out.copy_((v if residual is None else v + residual.float()).to(out.dtype))
    return out


@contextlib.contextmanager
def emulate_abi(bf16=False):
    """Run esn.ops on host tensors with every foreign call answered by the CPU model above.  bf16=True makes
    ops.compute_dtype choose bf16 (as under torch.autocast('cuda', torch.bfloat16) on the device)."""
    swaps = dict(_call=_call, require_cuda=lambda t, what: None, new_act=_new_act, to_nchw=_to_nchw, fglo_gate=_fglo_gate)
    if bf16:
        swaps["compute_dtype"] = lambda x: torch.bfloat16
    saved = {k: getattr(ops, k) for k in swaps}
    saved_profile = ops.PROFILE
    del CALLS[:]
    del ALLOCS[:]
    del _KEEP[:]
    try:
        for k, v in swaps.items():
            setattr(ops, k, v)
        ops.PROFILE = None
        yield CALLS
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
        ops.PROFILE = saved_profile
        del ALLOCS[:]
        del _KEEP[:]
