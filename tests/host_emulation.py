"""TEST INFRASTRUCTURE ONLY: torch-CPU stand-ins for the C-ABI kernel wrappers of ``esn.ops``.

Purpose: check the HOST logic of a model file -- weight packing, BatchNorm folding, channel padding, concat-slice
writes, residual chaining, layer order -- in the build container, which has no GPU.  ``emulate_kernels()`` swaps the
thin wrappers that launch kernels (``ops.conv2d``, ``ops.stem_conv3x3s2``, ...) for functions computing the SAME
documented contract (include/esn.h) with ``torch.nn.functional`` in fp32, keeping every buffer and layout decision
of the host code (NHWC-strided activations, padded pixel strides, ``out=`` slices).  A model forward under the
emulation then must reproduce the reference's golden logits; garbage in an uninitialised channel tail shows up as
NaN / mismatch exactly as it would on the device.

Nothing in the package imports this file; the product has no CPU path (tests/test_host_cpu.py::test_no_cpu_fallback).
The kernels themselves are checked on the GPU (tests/test_ops_gpu.py, tests/test_models_gpu.py).
"""
import contextlib

import torch
import torch.nn.functional as F

from esn import ops
from esn import _lib as L

EP_ACT_BEFORE_RESIDUAL, EP_RESIDUAL_FIRST = 1, 2      # include/esn.h: EsnEpilogue.flags
STEM_PAD0 = 256                                       # include/esn.h: ESN_STEM_PAD0


def _vec(v):
    return None if v is None else v.view(1, -1, 1, 1)


def _act(v, act, alpha):
    if act == L.ACT_RELU:
        return v.clamp_min(0)
    if act == L.ACT_PRELU:
        return v.clamp_min(0) + _vec(alpha) * v.clamp_max(0)
    return v


def _epilogue(acc, scale, shift, alpha, act, residual, flags=0):
    """include/esn.h EsnEpilogue: v = act(acc*scale + shift (+ residual)), with the two flagged variants."""
    c = acc.shape[1]
    sc = 1.0 if scale is None else _vec(scale[:c])
    sh = 0.0 if shift is None else _vec(shift[:c])
    al = None if alpha is None else alpha[:c]
    res = None if residual is None else residual.float()
    if flags & EP_RESIDUAL_FIRST:
        return _act((acc + res) * sc + sh, act, al)
    v = acc * sc + sh
    if res is None:
        return _act(v, act, al)
    if flags & EP_ACT_BEFORE_RESIDUAL:
        return _act(_act(v, act, al) + res, act, al)
    return _act(v + res, act, al)


def _store(out, v):
    assert out.shape == v.shape, (tuple(out.shape), tuple(v.shape))
    out.copy_(v.to(out.dtype))
    return out


def _poison(t):
    """torch.empty on the CPU usually returns zero pages; a device allocation does not.  Fill fresh buffers with NaN so
    reads of never-written channel tails are caught."""
    return t.fill_(float("nan")) if t.is_floating_point() else t


def new_act(n, c, h, w, dtype, device, c_alloc=None, zero=False):
    ca = c if c_alloc is None else c_alloc
    buf = torch.zeros((n, h, w, ca), dtype=dtype, device=device) if zero else _poison(torch.empty((n, h, w, ca), dtype=dtype, device=device))
    t = buf.permute(0, 3, 1, 2)
    return t if ca == c else t[:, :c]


def require_cuda(t, what):
    return None


def as_act(x, dtype=None):
    dtype = dtype or ops.compute_dtype(x)
    if ops.is_nhwc(x) and x.dtype == dtype:
        return x
    n, c, h, w = x.shape
    y = new_act(n, c, h, w, dtype, x.device, c_alloc=(c + 7) // 8 * 8 if c % 8 else None)
    return _store(y, x.float())


def to_nchw(x, dtype=None):
    return x.to(dtype or x.dtype).contiguous()


def conv2d(x, prep, out=None, residual=None, force_direct=False):
    n, c, h, w = x.shape
    assert c == prep.cin, (c, prep.cin)
    ho, wo = prep.out_hw(h, w)
    if out is None:
        out = new_act(n, prep.cout, ho, wo, x.dtype if x.dtype == torch.bfloat16 else torch.float32, x.device)
    wt = prep._w_src
    xf = x.float()
    assert torch.isfinite(xf).all(), "conv input holds non-finite values (uninitialised channel tail?)"
    if prep.transposed:
        wt = wt if prep.depthwise else wt.permute(1, 0, 2, 3)
        acc = F.conv_transpose2d(xf, wt, None, prep.stride, (prep.pad_h, prep.pad_w), prep.out_pad,
                                 prep.groups, (prep.dil_h, prep.dil_w))
    else:
        acc = F.conv2d(xf, wt, None, prep.stride, (prep.pad_h, prep.pad_w), (prep.dil_h, prep.dil_w), prep.groups)
    return _store(out, _epilogue(acc, prep.scale, prep.shift, prep.alpha, prep.act, residual, getattr(prep, "ep_flags", 0)))


def stem_conv3x3s2(x, w_direct, cconv, with_pool, out, scale, shift, alpha, act):
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3
    wt = w_direct.reshape(3, 3, 3, cconv).permute(3, 2, 0, 1)          # [tap][Cin][Cout] -> (Cout, Cin, kh, kw)
    pool = int(with_pool) & 3
    acc = F.conv2d(x, wt, None, 2, 0 if int(with_pool) & STEM_PAD0 else 1)
    if pool == 1:
        acc = torch.cat([acc, F.max_pool2d(x, 2, 2)], 1)
    elif pool == 2:
        acc = torch.cat([acc, F.max_pool2d(x, 3, 2, 1)], 1)
    return _store(out, _epilogue(acc, scale, shift, alpha, act, None))


def maxpool2x2(x, out, scale=None, shift=None, alpha=None, act=L.ACT_NONE):
    v = F.max_pool2d(x.float(), 2, 2)
    if tuple(out.shape[2:]) != tuple(v.shape[2:]):       # odd input: zero padding behind the pooled map (include/esn.h)
        v = F.pad(v, [0, out.shape[3] - v.shape[3], 0, out.shape[2] - v.shape[2]])
    return _store(out, _epilogue(v, scale, shift, alpha, act, None))


def avgpool3x3s2(x, out, scale=None, shift=None, alpha=None, act=L.ACT_NONE):
    return _store(out, _epilogue(F.avg_pool2d(x.float(), 3, 2, 1), scale, shift, alpha, act, None))


def affine_act(x, scale, shift, alpha, act, out=None, residual=None, flags=0):
    if out is None:
        n, c, h, w = x.shape
        out = new_act(n, c, h, w, x.dtype, x.device)
    return _store(out, _epilogue(x.float(), scale, shift, alpha, act, residual, flags))


def concat_tail(x, buf, c0, scale=None, shift=None, alpha=None, act=L.ACT_NONE):
    """include/esn.h esn_concat_tail: the injected channels through their BNPReLU slice, zeros to the end of the padded pixel."""
    n, c, h, w = x.shape
    tail = buf.stride(3) - c0
    assert tail >= 8 and tail & (tail - 1) == 0, tail
    full = buf.as_strided((n, tail, h, w), buf.stride(), buf.storage_offset() + c0)
    full.zero_()
    full[:, :c].copy_(_epilogue(x.float(), scale, shift, alpha, act, None).to(buf.dtype))
    return full[:, :c]


def adaptive_avgpool(x, size, dtype=None):
    n, c, h, w = x.shape
    y = new_act(n, c, size, size, dtype or x.dtype, x.device)
    return _store(y, F.adaptive_avg_pool2d(x.float(), size))


def bilinear(x, out_h, out_w, align_corners, out=None):
    n, c, h, w = x.shape
    if out is None:
        out = new_act(n, c, out_h, out_w, x.dtype, x.device)
    return _store(out, F.interpolate(x.float(), (out_h, out_w), mode="bilinear", align_corners=bool(align_corners)))


def _finish_head(logits_f32, want_logits, want_mask, logits_dtype):
    logits = logits_f32.to(logits_dtype).contiguous() if want_logits else None
    mask = logits_f32.argmax(1).to(torch.uint8) if want_mask else None      # first maximum wins (test.py:79-82)
    return logits, mask


def head_convt2x2(x, w, bias, classes, want_logits=True, want_mask=False, logits_dtype=torch.float32):
    wt = w[:, :, :, :classes].permute(2, 3, 0, 1)                       # [dy][dx][Cin][32] -> (Cin, classes, 2, 2)
    return _finish_head(F.conv_transpose2d(x.float(), wt, None if bias is None else bias[:classes], 2), want_logits, want_mask, logits_dtype)


def head_bilinear(x, classes, out_h, out_w, want_logits=True, want_mask=False, logits_dtype=torch.float32,
                  align_corners=False):
    y = F.interpolate(x[:, :classes].float(), (out_h, out_w), mode="bilinear", align_corners=bool(align_corners))
    return _finish_head(y, want_logits, want_mask, logits_dtype)


def fglo_gate(x, w1, b1, w2, b2, out=None, residual=None):
    """CGNet FGlo (CGNet.py:173-191): y = x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2) (+ residual)."""
    n, c, h, w = x.shape
    xf = x.float()
    gate = torch.sigmoid(F.linear(F.relu(F.linear(xf.mean(dim=(2, 3)), w1, b1)), w2, b2)).view(n, c, 1, 1)
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device)
    v = xf * gate
    return _store(out, v if residual is None else v + residual.float())


def maxpool3x3s2_idx(x):
    """MaxPool2d(3, 2, 1, return_indices=True): (pooled, int32 flat indices h*W+w laid out [N,Ho,Wo,C])."""
    n, c, h, w = x.shape
    v, idx = F.max_pool2d(x.float().contiguous(), 3, 2, 1, return_indices=True)
    y = new_act(n, c, v.shape[2], v.shape[3], x.dtype, x.device)
    return _store(y, v), idx.permute(0, 2, 3, 1).contiguous().to(torch.int32)


def max_unpool2x2(v, idx, ext=None, act=L.ACT_NONE, alpha=None):
    """y = act(MaxUnpool2d(2)(v, idx) + ext); CPU max_unpool2d is last-writer-wins in raster order, the kernel's rule."""
    n, c, h, w = v.shape
    up = F.max_unpool2d(v.float().contiguous(), idx.permute(0, 3, 1, 2).contiguous().long(), 2, output_size=(2 * h, 2 * w))
    y = new_act(n, c, 2 * h, 2 * w, v.dtype, v.device)
    return _store(y, _act(up if ext is None else up + ext.float(), act, alpha))


def gate_bcast(g, x, b=None, out=None):
    """include/esn.h esn_gate_bcast: y = g * x + b, g (N,1,H,W), b (N,C,1,1) or None."""
    n, c, h, w = x.shape
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device, c_alloc=x.stride(3), zero=x.stride(3) != c)
    v = g.float() * x.float()
    return _store(out, v if b is None else v + b.float())


def dab_dw_pair(x, prm, dilation, out=None):
    """include/esn.h EsnDabPair; prm rows: 0-11 taps of (3x1, 1x3, dilated 3x1, dilated 1x3), 12-23 their
    (scale, shift, alpha) triples, 24-26 the closing BN + PReLU (model/DABNet.py DABModule._build_prep)."""
    n, c, h, w = x.shape
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device)
    taps, aff, fin = prm[0:12].view(4, 3, c), prm[12:24].view(4, 3, c), prm[24:27]

    def stage(t, i, vertical, d):
        wt = taps[i].t().reshape(c, 1, 3, 1) if vertical else taps[i].t().reshape(c, 1, 1, 3)
        t = F.conv2d(t, wt, None, 1, (d, 0) if vertical else (0, d), (d, 1) if vertical else (1, d), c)
        return _act(t * _vec(aff[i, 0]) + _vec(aff[i, 1]), L.ACT_PRELU, aff[i, 2])
    xf = x.float()
    br1 = stage(stage(xf, 0, True, 1), 1, False, 1)
    br2 = stage(stage(xf, 2, True, dilation), 3, False, dilation)
    return _store(out, _act((br1 + br2) * _vec(fin[0]) + _vec(fin[1]), L.ACT_PRELU, fin[2]))


_SWAPS = dict(gate_bcast=gate_bcast, fglo_gate=fglo_gate, maxpool3x3s2_idx=maxpool3x3s2_idx, max_unpool2x2=max_unpool2x2,
              dab_dw_pair=dab_dw_pair, new_act=new_act, require_cuda=require_cuda, as_act=as_act, to_nchw=to_nchw, conv2d=conv2d,
              stem_conv3x3s2=stem_conv3x3s2, maxpool2x2=maxpool2x2, avgpool3x3s2=avgpool3x3s2, affine_act=affine_act, concat_tail=concat_tail,
              adaptive_avgpool=adaptive_avgpool, bilinear=bilinear, head_convt2x2=head_convt2x2,
              head_bilinear=head_bilinear)


@contextlib.contextmanager
def emulate_kernels():
    saved = {k: getattr(ops, k) for k in _SWAPS}
    try:
        for k, v in _SWAPS.items():
            setattr(ops, k, v)
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
