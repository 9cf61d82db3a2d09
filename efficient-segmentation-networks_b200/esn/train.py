"""Training path: forward kernels + hand-written backward on a tiny tape.

No torch.autograd inside the model: every activation-sized computation (forward and backward) is
a C-ABI kernel call.  The model's train-mode ``forward`` records backward closures on a `Tape`; one
`torch.autograd.Function` around the whole network hands the parameter gradients back to
PyTorch, so the reference's ``output = model(images); loss = criterion(output, labels);
loss.backward(); optimizer.step()`` (train.py:351-355) works unchanged.

Semantics follow torch: train-mode BatchNorm normalises with the biased batch
variance and updates running_var with the unbiased one; PReLU per channel; MaxPool2d(2,2) routes
the gradient to the first maximum; bilinear align_corners=False.
"""
import contextlib
import ctypes as C

import torch

from . import _lib as L
from . import ops
from .prep import weights_generation


class V:
    """An activation and its gradient slot.  A V may be a channel slice of a wider (concat) V, in
    which case its gradient is the same slice of the parent's gradient buffer."""
    __slots__ = ("t", "_g", "parent", "lo", "hi", "_own")

    def __init__(self, t, parent=None, lo=0, hi=0):
        self.t, self._g, self.parent, self.lo, self.hi = t, None, parent, lo, hi
        self._own = False      # _g is a buffer of this V alone (safe to accumulate into in place)

    @property
    def g(self):
        if self.parent is not None:
            pg = self.parent.g
            return None if pg is None else pg[:, self.lo:self.hi]
        return self._g

    def slice(self, lo, hi):
        return V(self.t[:, lo:hi], self, lo, hi)

    def add_grad(self, fn):
        """fn(existing) -> tensor holding existing + this consumer's contribution.  For slices the
        contribution is accumulated in place into the parent's buffer."""
        if self.parent is not None:
            root = self.parent
            while root.parent is not None:
                root = root.parent
            if root._g is None:     # gradient buffer of the whole concat tensor, zero-filled once
                n, c, h, w = root.t.shape
                root._g = ops.new_act(n, c, h, w, root.t.dtype, root.t.device, c_alloc=root.t.stride(3), zero=True)
                root._own = True
            elif not root._own:
                # the root's gradient was handed over by a whole-tensor consumer and may be shared with another V (a
                # residual add passes the SAME tensor to both operands): slices accumulate in place, so take a private copy
                # first (ESPNet's block: BN(input + cat[...]) -- the in-place sums into cat's slices would otherwise leak
                # into the gradient of `input`)
                n, c, h, w = root.t.shape
                priv = ops.new_act(n, c, h, w, root._g.dtype, root._g.device, c_alloc=root.t.stride(3),
                                   zero=root.t.stride(3) != c)
                root._g = ops.affine_act(root._g, None, None, None, L.ACT_NONE, out=priv)
                root._own = True
            view = self.g
            out = fn(view, view)
            assert out.data_ptr() == view.data_ptr()
        else:
            self._g = fn(self._g, None)
            # ownership is taken ONCE, by the V that receives the buffer from the op that allocated it: when this V's producer
            # later hands the same tensor on (a residual add gives it to both operands), nobody else may accumulate into it
            self._own = bool(getattr(self._g, "_esn_fresh", False))
            if self._own:
                self._g._esn_fresh = False


class _ZeroArena:
    """ONE memset per training step for all the small zero-initialised scratch buffers of the tape (BatchNorm statistic sums,
    weight-gradient accumulators: ~200 per DABNet iteration, each a separate torch.zeros fill kernel in round 1 = 5 % of the
    step).  Sized from what the previous step used; a request that does not fit falls back to its own torch.zeros.  A new
    buffer is allocated per step, so views handed out as parameter gradients stay valid until the optimizer drops them."""

    def __init__(self):
        self.need, self.buf, self.off, self.used, self.dev = {}, None, 0, 0, None

    def begin(self, device):
        if self.dev is not None:
            self.need[self.dev] = self.used
        self.dev, self.off, self.used = device, 0, 0
        size = self.need.get(device, 0)
        self.buf = torch.zeros(size, dtype=torch.uint8, device=device) if size else None

    def take(self, shape, dtype, device):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        n = 1
        for v in shape:
            n *= v
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        aligned = (nbytes + 255) // 256 * 256
        if device == self.dev:
            self.used += aligned
            if self.buf is not None and self.off + aligned <= self.buf.numel():
                view = self.buf[self.off:self.off + nbytes].view(dtype).view(shape)
                self.off += aligned
                return view
        return torch.zeros(shape, dtype=dtype, device=device)


_ARENA = _ZeroArena()


class Tape:
    def __init__(self, buckets=None, device=None):
        if device is not None:
            _ARENA.begin(torch.device(device))
        self.steps = []
        self.param_grads = {}     # parameter -> fp32 gradient tensor
        self.buckets = buckets    # esn.parallel.GradBuckets or None
        self.bn_counters = []     # num_batches_tracked of the BatchNorm layers of this forward: bumped by ONE foreach kernel
        self.side = None          # stream for work off the critical path of the backward (weight gradients)
        self.keep = []            # tensors read on the side stream: kept alive until it has been joined

    def off_critical_path(self, *tensors):
        """Context manager: kernels launched inside run on a side stream that has waited for everything issued so far on the
        current stream.  Weight gradients are leaves of the backward -- nothing downstream in the tape reads them -- while the
        chain BatchNorm backward -> input gradient -> BatchNorm backward ... is a sequence of small latency-bound kernels; run
        concurrently, the weight gradients fill the SMs the chain leaves idle.  The side stream is joined before the gradients
        are handed to anyone (end of backward / a data-parallel bucket being packed)."""
        if not SIDE_WGRAD or not torch.cuda.is_available():
            return contextlib.nullcontext()
        cur = torch.cuda.current_stream()
        if self.side is None:
            self.side = _SIDE_STREAMS.setdefault(cur.device_index, None) or torch.cuda.Stream()
            _SIDE_STREAMS[cur.device_index] = self.side
            if self.buckets is not None:
                self.buckets.wgrad_stream = self.side
        self.side.wait_stream(cur)
        self.keep.extend(tensors)
        return torch.cuda.stream(self.side)

    def count_batch(self, bn):
        if bn.num_batches_tracked is not None:
            self.bn_counters.append(bn.num_batches_tracked)

    def flush_counters(self):
        if self.bn_counters:
            torch._foreach_add_(self.bn_counters, 1)
            self.bn_counters = []

    def push(self, fn):
        self.steps.append(fn)

    def add_param_grad(self, p, g):
        g = g.to(p.dtype) if g.dtype != p.dtype else g
        if p in self.param_grads:
            if self.side is not None:       # the earlier contribution may still be in flight on the side stream
                torch.cuda.current_stream().wait_stream(self.side)
            prev = self.param_grads[p]
            if self.buckets is not None:       # the bucket view is only filled when its bucket is packed
                prev = self.buckets.staged[self.buckets.bucket_of[p]].get(p, prev)
            g = prev + g
        # every parameter of the hot-path nets is used exactly once per step, so its gradient is final
        # here: hand it to the data-parallel buckets (flat copy + all-reduce launch, overlapped)
        self.param_grads[p] = self.buckets.grad_ready(p, g) if self.buckets is not None else g

    def backward(self):
        for fn in reversed(self.steps):
            fn()
        self.steps = []
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self.keep = []
        if self.buckets is not None:
            self.buckets.finish()
        return self.param_grads


def _f64zeros(n, device):
    return _ARENA.take(n, torch.float64, torch.device(device))


def _f32zeros(shape, device):
    return _ARENA.take(shape, torch.float32, torch.device(device))


import os as _os
FUSED_BN = _os.environ.get("ESN_FUSED_BN", "1") != "0"      # one cooperative launch per BatchNorm layer and direction
_FUSED_DIR = _os.environ.get("ESN_FUSED_BN_DIR", "fwd,bwd")  # diagnosis: restrict the fused kernels to one direction
SIDE_WGRAD = _os.environ.get("ESN_SIDE_WGRAD", "1") != "0"   # weight gradients on a side stream, concurrent with the dgrad chain
_SIDE_STREAMS = {}
BN_REPLICAS = 8                                              # ESN_BN_FUSED_REPLICAS (include/esn.h)


def _fresh(t):
    """Mark a gradient buffer as allocated by the op that returns it and referenced by nobody else: V.add_grad then lets
    channel slices accumulate into it in place instead of taking a private copy first."""
    t._esn_fresh = True
    return t


def _v8(t):
    """bf16 NHWC view whose channel vectors are 16-byte aligned (what the 16-byte kernels of esn_bn_fused.cu take)."""
    return (t.dtype == torch.bfloat16 and ops.is_nhwc(t) and t.stride(3) % 8 == 0 and t.data_ptr() % 16 == 0)


# --------------------------------------------------------------------------- convolution
class ConvT:
    """nn.Conv2d (dense or depthwise) for training: forward = raw conv (+bias); backward = input
    gradient through the (tensor-core) conv kernels with flipped / transposed weights, weight
    gradient through esn_conv2d_wgrad."""

    def __init__(self, conv, cin_pad=None, cout_pad=None):
        """cin_pad / cout_pad: run the kernels over zero-padded channel counts (tensor-core friendly);
        the input buffer must then be at least cin_pad wide with zeros in the tail, and the caller's
        output slice cout_pad wide (the tail is overwritten or ignored by the caller)."""
        self.conv = conv
        self.cin_pad, self.cout_pad = cin_pad, cout_pad
        self._key = None

    def preps(self):
        c = self.conv
        w = c.weight
        key = (weights_generation(), w.data_ptr(), w._version, None if c.bias is None else c.bias._version)
        if self._key != key:
            wd = w.detach().float()
            bias = None if c.bias is None else c.bias.detach().float()
            s, pad, dil, g = c.stride[0], tuple(c.padding), tuple(c.dilation), c.groups
            kh, kw = wd.shape[2:]
            if self.cin_pad or self.cout_pad:      # zero-extend (Cout, Cin, kh, kw)
                assert g == 1 and bias is None
                wd = torch.nn.functional.pad(wd, (0, 0, 0, 0, 0, (self.cin_pad or wd.shape[1]) - wd.shape[1],
                                                  0, (self.cout_pad or wd.shape[0]) - wd.shape[0]))
            self.fwd_prep = ops.ConvPrep.from_weight(wd, s, pad, dil, g, bias=bias)
            if s == 1:
                wf = wd.flip(2, 3) if g != 1 else wd.permute(1, 0, 2, 3).flip(2, 3)
                self.dgrad_prep = ops.ConvPrep.from_weight(wf.contiguous(), 1,
                                                           (dil[0] * (kh - 1) - pad[0], dil[1] * (kw - 1) - pad[1]), dil, g)
            else:   # conv_transpose2d(dy, W): W (Cout, Cin, kh, kw) read as (in=Cout, out=Cin)
                if g != 1 and not (g == wd.shape[0] and wd.shape[1] == 1):
                    raise NotImplementedError("strided grouped conv backward")
                self.dgrad_prep = ops.ConvPrep.from_weight(wd, s, pad, dil, g, transposed=True)
            self._key = key
        return self.fwd_prep, self.dgrad_prep

    def forward(self, tape, x, out=None, residual=None, need_dx=True, dtype=None, precomputed=False):
        """x: V (or a raw NCHW input tensor wrapped in V with need_dx=False).  precomputed: `out` already holds the conv
        result (written by a fused kernel, e.g. ENet's conv || max-pool stem); only the backward is recorded."""
        fwd_prep, dgrad_prep = self.preps()
        xt = x.t if not self.cin_pad else ops.widen(x.t, self.cin_pad)
        cin_real, cout_real = self.conv.in_channels, self.conv.out_channels
        if out is None:
            n, _, h, w = xt.shape
            ho, wo = fwd_prep.out_hw(h, w)
            out = ops.new_act(n, fwd_prep.cout, ho, wo, dtype or (xt.dtype if ops.is_nhwc(xt) else torch.float32), xt.device)
        y = out if isinstance(out, V) else V(out)      # `out` may be a channel slice (V) of a concat buffer
        if precomputed:
            pass
        elif (not ops.is_nhwc(xt) and xt.dtype == torch.float32 and xt.is_contiguous() and fwd_prep.cin == 3 and residual is None
                and (fwd_prep.kh, fwd_prep.kw, fwd_prep.stride) == (3, 3, 2) and (fwd_prep.pad_h, fwd_prep.pad_w) in ((0, 0), (1, 1))
                and fwd_prep.cout % 4 == 0 and fwd_prep.cout <= 32 and y.t.shape[1] == fwd_prep.cout and y.t.stride(3) % 4 == 0
                and (fwd_prep.pad_h == 0 or not ((xt.shape[2] | xt.shape[3]) & 1))):
            # network stem on the NCHW image: the dedicated kernel (bias as the epilogue shift)
            ops.stem_conv3x3s2(xt, fwd_prep.w_direct, fwd_prep.cout, 256 if fwd_prep.pad_h == 0 else 0, y.t,
                               fwd_prep.scale, fwd_prep.shift, None, L.ACT_NONE)
        else:
            ops.conv2d(xt, fwd_prep, out=y.t, residual=None if residual is None else residual.t)
        conv = self.conv

        def bwd():
            dy = y.g
            # weight gradient [tap][Cin/g][Cout] -> (Cout, Cin/g, kh, kw)
            kh, kw = fwd_prep.kh, fwd_prep.kw
            cin_g = fwd_prep.cin // fwd_prep.groups
            stem = (not ops.is_nhwc(xt) and fwd_prep.cin == 3 and (kh, kw) == (3, 3) and fwd_prep.cout <= 32
                    and xt.dtype == torch.float32)          # esn_conv2d_wgrad has a kernel for the NCHW fp32 image
            pad8 = not stem and not ops.is_nhwc(xt) and dy.dtype == torch.bfloat16 and fwd_prep.groups == 1 and fwd_prep.cin < 8
            if pad8:
                cin_g = 8
            dwbuf = _f32zeros((kh * kw, cin_g, fwd_prep.cout), dy.device)
            bsums = _f64zeros(fwd_prep.cout, dy.device) if conv.bias is not None else None
            with tape.off_critical_path(dy, xt):
                p = L.EsnConv()
                xw = xt
                if pad8:
                    # network stem (NCHW fp32 image, Cin=3): give the tensor-core wgrad an NHWC bf16 copy
                    # padded to 8 channels; the padded rows of dW are dropped below
                    n_, c_, h_, w_ = xt.shape
                    x8 = ops.new_act(n_, c_, h_, w_, torch.bfloat16, xt.device, c_alloc=8, zero=True)
                    dxs, dx8 = ops.tdesc(xt), ops.tdesc(x8)
                    ops._call(L.lib.esn_convert_layout, "esn_convert_layout", (C.byref(dxs), C.byref(dx8)),
                              ops._nbytes(xt) + ops._nbytes(x8))
                    xw = ops.widen(x8, 8)
                    tape.keep.append(x8)
                flops = 2 * dy.shape[0] * dy.shape[2] * dy.shape[3] * fwd_prep.cout * cin_g * kh * kw
                p.x, p.y = ops.tdesc(xw), ops.tdesc(dy)
                p.w = dwbuf.data_ptr()
                p.kh, p.kw, p.stride = kh, kw, fwd_prep.stride
                p.pad_h, p.pad_w, p.dil_h, p.dil_w = fwd_prep.pad_h, fwd_prep.pad_w, fwd_prep.dil_h, fwd_prep.dil_w
                p.groups, p.transposed, p.cout_pad = fwd_prep.groups, 0, fwd_prep.cout_pad
                ops._call(L.lib.esn_conv2d_wgrad, "esn_conv2d_wgrad", (C.byref(p),), ops._nbytes(xt) + ops._nbytes(dy), flops,
                          "%dx%d c%d-%d s%d g%d" % (kh, kw, fwd_prep.cin, fwd_prep.cout, fwd_prep.stride, fwd_prep.groups))
                if bsums is not None:
                    d = ops.tdesc(dy)
                    ops._call(L.lib.esn_channel_stats, "esn_channel_stats", (C.byref(d), C.c_void_p(bsums.data_ptr()), 0),
                              ops._nbytes(dy))
                    bgrad = bsums.float()
                    tape.keep.append(bsums)
            dw4 = dwbuf.view(kh, kw, cin_g, fwd_prep.cout).permute(3, 2, 0, 1)
            if fwd_prep.groups == 1:
                dw4 = dw4[:cout_real, :cin_real]        # drop the zero-padded channels
            tape.add_param_grad(conv.weight, dw4)
            if conv.bias is not None:
                tape.add_param_grad(conv.bias, bgrad)
            if residual is not None:
                residual.add_grad(lambda ex, dst: dy if ex is None else ops.affine_act(dy, None, None, None, L.ACT_NONE,
                                                                                        out=dst, residual=ex))
            if need_dx:
                if dgrad_prep.transposed:
                    h = xt.shape[2]
                    dgrad_prep.out_pad = h - ((dy.shape[2] - 1) * dgrad_prep.stride - 2 * dgrad_prep.pad_h
                                              + dgrad_prep.dil_h * (dgrad_prep.kh - 1) + 1)
                if not self.cin_pad:
                    x.add_grad(lambda ex, dst: ops.conv2d(dy, dgrad_prep, out=dst, residual=ex))
                else:
                    def run(ex, dst):
                        assert dst is None
                        g = ops.conv2d(dy, dgrad_prep)[:, :cin_real]     # computed over the padded width
                        return g if ex is None else ops.affine_act(g, None, None, None, L.ACT_NONE, out=g, residual=ex)
                    x.add_grad(run)

        tape.push(bwd)
        return y


class ConvTransposeT:
    """nn.ConvTranspose2d (dense, stride 2: ERFNet's UpsamplerBlock 3x3/p1/op1 and its 2x2 output conv) for training.
    Forward = the transposed-conv kernels; input gradient = the strided conv with the same weight (the adjoint);
    weight gradient = esn_conv2d_wgrad with the roles of input and output gradient swapped."""

    def __init__(self, conv):
        self.conv = conv
        self._key = None

    def preps(self):
        c = self.conv
        w = c.weight
        key = (weights_generation(), w.data_ptr(), w._version, None if c.bias is None else c.bias._version)
        if self._key != key:
            self.fwd_prep = ops.ConvPrep(c)
            # (Cin, Cout, kh, kw) read as a conv weight (out = Cin, in = Cout): conv2d(dy, W, stride, padding) = dx
            self.dgrad_prep = ops.ConvPrep.from_weight(w.detach().float(), c.stride[0], tuple(c.padding), (1, 1), 1)
            self._key = key
        return self.fwd_prep, self.dgrad_prep

    def backward_from(self, tape, x, dy):
        """dy: NHWC gradient of the output (channels = conv.out_channels, possibly in a zero-padded wider buffer)."""
        conv = self.conv
        fwd_prep, dgrad_prep = self.preps()
        kh, kw = conv.kernel_size
        cin, cout = conv.in_channels, conv.out_channels
        xt = x.t
        dwbuf = _f32zeros((kh * kw, cout, cin), dy.device)   # [tap][Cout][Cin]
        sums = _f64zeros(cout, dy.device) if conv.bias is not None else None
        with tape.off_critical_path(dy, xt):
            p = L.EsnConv()
            p.x, p.y, p.w = ops.tdesc(dy), ops.tdesc(xt), dwbuf.data_ptr()
            p.kh, p.kw, p.stride = kh, kw, conv.stride[0]
            p.pad_h, p.pad_w, p.dil_h, p.dil_w = conv.padding[0], conv.padding[1], 1, 1
            p.groups, p.transposed, p.cout_pad = 1, 0, (cin + 15) // 16 * 16
            ops._call(L.lib.esn_conv2d_wgrad, "esn_conv2d_wgrad", (C.byref(p),), ops._nbytes(xt) + ops._nbytes(dy),
                      2 * xt.shape[0] * xt.shape[2] * xt.shape[3] * cin * cout * kh * kw, "%dx%dT c%d-%d" % (kh, kw, cin, cout))
            if sums is not None:
                d = ops.tdesc(dy)
                ops._call(L.lib.esn_channel_stats, "esn_channel_stats", (C.byref(d), C.c_void_p(sums.data_ptr()), 0), ops._nbytes(dy))
                bgrad = sums.float()
                tape.keep.append(sums)
        tape.add_param_grad(conv.weight, dwbuf.view(kh, kw, cout, cin).permute(3, 2, 0, 1))
        if conv.bias is not None:
            tape.add_param_grad(conv.bias, bgrad)
        x.add_grad(lambda ex, dst: ops.conv2d(dy, dgrad_prep, out=dst, residual=ex))

    def forward(self, tape, x, out=None):
        fwd_prep, _ = self.preps()
        n, _, h, w = x.t.shape
        if out is None:
            ho, wo = fwd_prep.out_hw(h, w)
            out = V(ops.new_act(n, fwd_prep.cout, ho, wo, x.t.dtype, x.t.device))
        y = out
        ops.conv2d(x.t, fwd_prep, out=y.t)
        tape.push(lambda: self.backward_from(tape, x, y.g))
        return y


def convt2x2_logits(tape, convt, x, w_packed, bias, classes):
    """ERFNet's output conv (ConvTranspose2d(16, classes, 2, stride 2), ERFNet.py:128) -> NCHW fp32 logits through the
    fused head kernel; backward converts d logits to NHWC once and reuses ConvTransposeT's gradients."""
    logits, _ = ops.head_convt2x2(x.t, w_packed, bias, classes, True, False, torch.float32)
    holder = {}

    def bwd():
        dl = holder["dlogits"]
        n, c, h, w = dl.shape
        dy = ops.new_act(n, c, h, w, x.t.dtype, dl.device, c_alloc=(c + 7) // 8 * 8, zero=True)
        a, b = ops.tdesc(dl), ops.tdesc(dy)
        a.layout, a.c_stride = L.ESN_NCHW, 0
        ops._call(L.lib.esn_convert_layout, "esn_convert_layout", (C.byref(a), C.byref(b)), ops._nbytes(dl) + ops._nbytes(dy))
        convt.backward_from(tape, x, dy)
    tape.push(bwd)
    return logits, holder


# --------------------------------------------------------------------------- BatchNorm (+ activation)
class BNActT:
    """Train-mode nn.BatchNorm2d followed by PReLU / ReLU / nothing; or, with bn=None, the
    activation alone (its backward then also yields sum(dz) = the bias gradient of a preceding conv)."""

    def __init__(self, bn, act, prelu=None, alpha_sink=None):
        """prelu: nn.PReLU with one slope per channel, or ONE slope for all channels (nn.PReLU(), ENet: its gradient is the
        sum over channels).  alpha_sink(param, grad): where the slope gradient goes instead of the tape -- a SharedParamGrad
        when one PReLU module serves several call sites of a block (ENet.py:53-89)."""
        self.bn, self.act, self.prelu, self.alpha_sink = bn, act, prelu, alpha_sink

    def forward(self, tape, x, out=None):
        xt = x.t
        n, c, h, w = xt.shape
        dev = xt.device
        bn = self.bn
        alpha = None if self.prelu is None else self.prelu.weight.detach()
        one_slope = alpha is not None and alpha.numel() == 1 and c > 1
        if one_slope:
            alpha = alpha.expand(c).contiguous()
        scale = shift = mean = invstd = None
        fused_out = None
        if bn is not None:
            sums = _f64zeros(BN_REPLICAS * 2 * c + 1, dev)    # replicated sums + the fused kernel's grid barrier word
            scale, shift, mean, invstd = torch.empty(4, c, dtype=torch.float32, device=dev).unbind(0)
            f = L.EsnBnFinalize()
            f.sums, f.count = sums.data_ptr(), n * h * w
            f.gamma, f.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
            f.eps, f.momentum = bn.eps, (bn.momentum if bn.momentum is not None else 0.1)
            f.running_mean, f.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            f.scale, f.shift, f.mean, f.invstd = scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr()
            f.channels = c
            dst = out.t if isinstance(out, V) else out
            if FUSED_BN and "fwd" in _FUSED_DIR and _v8(xt) and (_v8(dst) if dst is not None else c % 8 == 0):
                # one launch: statistics -> grid barrier -> finalize + normalise + activate out of L2 (esn_bn_fused.cu)
                if dst is None:
                    dst = ops.new_act(n, c, h, w, xt.dtype, dev)
                q = L.EsnBnTrainFwd()
                q.x, q.y, q.fin = ops.tdesc(xt), ops.tdesc(dst), f
                q.alpha = alpha.data_ptr() if alpha is not None else None
                q.barrier, q.act = sums.data_ptr() + 8 * BN_REPLICAS * 2 * c, self.act
                ops._call(L.lib.esn_bn_act_train_fwd, "esn_bn_act_train_fwd", (C.byref(q),), ops._nbytes(xt) + ops._nbytes(dst))
                fused_out = dst
            else:
                d = ops.tdesc(xt)
                ops._call(L.lib.esn_channel_stats, "esn_channel_stats", (C.byref(d), C.c_void_p(sums.data_ptr()), 1),
                          ops._nbytes(xt))
                ops._call(L.lib.esn_bn_finalize, "esn_bn_finalize", (C.byref(f),))
            tape.count_batch(bn)
        if fused_out is not None:
            y = out if isinstance(out, V) else V(fused_out)
        elif isinstance(out, V):
            y = out
            ops.affine_act(xt, scale, shift, alpha, self.act, out=y.t)
        else:
            y = V(ops.affine_act(xt, scale, shift, alpha, self.act, out=out))
        act, prelu = self.act, self.prelu

        def bwd():
            dy = y.g
            sums3 = _f64zeros(BN_REPLICAS * 3 * c + 1, dev)   # replicated sums + the fused kernel's grid barrier word
            dgamma = torch.empty(c, dtype=torch.float32, device=dev) if bn is not None else None
            dbeta = torch.empty(c, dtype=torch.float32, device=dev)
            dalpha = torch.empty(c, dtype=torch.float32, device=dev) if prelu is not None else None

            def run(ex, dst):
                fresh = dst is None
                if fresh:       # 16-byte channel vectors also for 35 / 131 / 259 channels (zero tail inside the pixel stride)
                    # (the <= 7 pad lanes are never read: consumers are channel slices of the c real channels)
                    dx = ops.new_act(n, c, h, w, dy.dtype, dev, c_alloc=(c + 7) // 8 * 8)
                else:
                    dx = dst
                p = L.EsnBnBwd()
                p.x, p.dy, p.dx = ops.tdesc(xt), ops.tdesc(dy), ops.tdesc(dx)
                if ex is not None:
                    p.extra = ops.tdesc(ex)
                p.scale = scale.data_ptr() if scale is not None else None
                p.shift = shift.data_ptr() if shift is not None else None
                p.alpha = alpha.data_ptr() if alpha is not None else None
                p.mean = mean.data_ptr() if mean is not None else None
                p.invstd = invstd.data_ptr() if invstd is not None else None
                p.sums = sums3.data_ptr()
                p.dgamma = dgamma.data_ptr() if dgamma is not None else None
                p.dbeta = dbeta.data_ptr()
                p.dalpha = dalpha.data_ptr() if dalpha is not None else None
                p.act, p.train_stats = act, int(bn is not None)
                nb = ops._nbytes(xt) + ops._nbytes(dy)
                if (FUSED_BN and "bwd" in _FUSED_DIR and bn is not None and _v8(xt) and _v8(dy) and _v8(dx) and (ex is None or _v8(ex))
                        and dx.data_ptr() not in (dy.data_ptr(), xt.data_ptr())):
                    ops._call(L.lib.esn_bn_act_bwd_fused, "esn_bn_act_bwd_fused",
                              (C.byref(p), C.c_void_p(sums3.data_ptr() + 8 * BN_REPLICAS * 3 * c)), nb + ops._nbytes(dx))
                    return _fresh(dx) if fresh else dx
                if (bn is None and prelu is None and c % 8 == 0 and _v8(xt) and _v8(dy) and _v8(dx)
                        and (ex is None or _v8(ex))):
                    # a bare ReLU has no sums anyone reads: dx = dy * [x > 0] is one streaming pass
                    ops._call(L.lib.esn_act_bwd, "esn_act_bwd", (C.byref(p),), nb + ops._nbytes(dx))
                    return _fresh(dx) if fresh else dx
                if bn is not None or prelu is not None:
                    ops._call(L.lib.esn_bn_act_bwd_reduce, "esn_bn_act_bwd_reduce", (C.byref(p),), nb)
                ops._call(L.lib.esn_bn_act_bwd_apply, "esn_bn_act_bwd_apply", (C.byref(p),), nb + ops._nbytes(dx))
                return _fresh(dx) if fresh else dx

            x.add_grad(run)
            if bn is not None:
                tape.add_param_grad(bn.weight, dgamma)
                tape.add_param_grad(bn.bias, dbeta)
            if prelu is not None:
                ga = dalpha.sum().reshape(1) if one_slope else dalpha
                (self.alpha_sink or tape.add_param_grad)(prelu.weight, ga)

        tape.push(bwd)
        return y


class SharedParamGrad:
    """One parameter used at several places of a block (ENet's shared activation module): the contributions are summed here
    and handed to the tape ONCE, after the last of them in backward order -- create it BEFORE the block's first use in the
    forward (its flush step is pushed first, so it runs last), so that data-parallel gradient buckets see each parameter
    exactly once per step."""

    def __init__(self, tape, param):
        self.tape, self.param, self.total = tape, param, None
        tape.push(self.flush)

    def __call__(self, param, g):
        assert param is self.param
        self.total = g if self.total is None else self.total + g

    def flush(self):
        if self.total is not None:
            self.tape.add_param_grad(self.param, self.total)
            self.total = None


def maxpool3x3s2_idx(tape, x):
    """MaxPool2d(3, 2, 1, return_indices=True) (ENet.py:126-130): returns (V pooled, int32 indices); the backward is a
    deterministic gather over the input pixels."""
    yt, idx = ops.maxpool3x3s2_idx(x.t)
    y = V(yt)

    def bwd():
        dy = y.g

        def run(ex, dst):
            n, c, h, w = x.t.shape
            dx = dst if dst is not None else (ex if ex is not None else ops.new_act(n, c, h, w, dy.dtype, dy.device))
            a, b = ops.tdesc(dy), ops.tdesc(dx)
            ops._call(L.lib.esn_maxpool3x3s2_idx_bwd, "esn_maxpool3x3s2_idx_bwd",
                      (C.byref(a), C.c_void_p(idx.data_ptr()), C.byref(b), int(ex is not None)), 2 * ops._nbytes(dx))
            return dx
        x.add_grad(run)
    tape.push(bwd)
    return y, idx


def max_unpool2x2(tape, v, idx):
    """MaxUnpool2d(2)(v, idx) (ENet.py:225, 262); backward: every pooled cell reads the gradient at its scatter position."""
    y = V(ops.max_unpool2x2(v.t, idx))

    def bwd():
        dy = y.g

        def run(ex, dst):
            n, c, h, w = v.t.shape
            dv = ops.new_act(n, c, h, w, dy.dtype, dy.device)
            a, b = ops.tdesc(dy), ops.tdesc(dv)
            ops._call(L.lib.esn_max_unpool2x2_bwd, "esn_max_unpool2x2_bwd", (C.byref(a), C.c_void_p(idx.data_ptr()), C.byref(b)),
                      2 * ops._nbytes(dv))
            if ex is None and dst is None:
                return dv
            return ops.affine_act(dv, None, None, None, L.ACT_NONE, out=dst, residual=ex)
        v.add_grad(run)
    tape.push(bwd)
    return y


def add(tape, a, b, out=None):
    """y = a + b (branch merge), optionally written into `out` (a V, e.g. a channel slice of a concat buffer)."""
    if out is not None:
        y = out
        ops.affine_act(a.t, None, None, None, L.ACT_NONE, out=y.t, residual=b.t)
    else:
        y = V(ops.affine_act(a.t, None, None, None, L.ACT_NONE, residual=b.t))

    def bwd():
        dy = y.g
        for v in (a, b):
            v.add_grad(lambda ex, dst: dy if (ex is None and dst is None) else
                       ops.affine_act(dy, None, None, None, L.ACT_NONE, out=dst, residual=ex))
    tape.push(bwd)
    return y


def maxpool2x2(tape, x, out, need_dx=True):
    """MaxPool2d(2,2) of x written into `out` (a channel slice of a concat buffer)."""
    y = out
    ops.maxpool2x2(x.t, out.t)
    if not need_dx:          # the network input: no gradient wanted
        return y

    def bwd():
        dy = y.g

        def run(ex, dst):
            n, c, h, w = x.t.shape
            dx = dst if dst is not None else (ex if ex is not None else
                                              ops.new_act(n, c, h, w, dy.dtype, dy.device, c_alloc=(c + 7) // 8 * 8))
            dxd, dyd, xd = ops.tdesc(dx), ops.tdesc(dy), ops.tdesc(x.t)
            ops._call(L.lib.esn_maxpool2x2_bwd, "esn_maxpool2x2_bwd", (C.byref(xd), C.byref(dyd), C.byref(dxd),
                                                                      int(ex is not None)), 2 * ops._nbytes(dx))
            return dx
        x.add_grad(run)
    tape.push(bwd)
    return y


def bilinear_logits(tape, scores, out_h, out_w, logits_dtype=torch.float32, align_corners=False):
    """F.interpolate(scores, (H, W), bilinear, align_corners) -> NCHW logits; backward gathers."""
    classes = scores.t.shape[1]
    logits, _ = ops.head_bilinear(scores.t, classes, out_h, out_w, True, False, logits_dtype, align_corners=align_corners)
    holder = {}

    def bwd():
        dl = holder["dlogits"]

        def run(ex, dst):
            # the logits are the only consumer of the scores: their gradient is written, not accumulated (when the scores
            # are a channel slice of a wider buffer, `dst` is that slice of the zero-filled gradient buffer)
            assert ex is None or ex is dst
            n, c, h, w = scores.t.shape
            dlow = dst if dst is not None else ops.new_act(n, c, h, w, scores.t.dtype, dl.device, c_alloc=scores.t.stride(3))
            a, b = ops.tdesc(dl), ops.tdesc(dlow)
            a.layout, a.c_stride = L.ESN_NCHW, 0
            if align_corners:
                ops._call(L.lib.esn_bilinear_bwd_nhwc, "esn_bilinear_bwd_nhwc", (C.byref(a), C.byref(b), 1, 0),
                          dl.numel() * dl.element_size())
            else:
                ops._call(L.lib.esn_bilinear_bwd, "esn_bilinear_bwd", (C.byref(a), C.byref(b), C.c_float(1.0)),
                          dl.numel() * dl.element_size())
            return dlow
        scores.add_grad(run)
    tape.push(bwd)
    return logits, holder


def bilinear_ce(tape, scores, out_h, out_w, target, weight, ignore_label, align_corners=False):
    """bilinear_logits + CrossEntropyLoss2d + both backward passes as one launch (ops.bilinear_ce): returns (sums, holder) with
    sums = [sum w*nll, sum w] of this rank, or None when the geometry is not taken.  The tape step scales the stored gradient
    by holder["gscale"] (a device scalar: upstream gradient / normaliser, set by _NetLossFn.backward) while converting it to
    the scores' dtype."""
    res = ops.bilinear_ce(scores.t, target, weight, ignore_label, out_h, out_w, sums=_f32zeros(2, scores.t.device),
                          align_corners=align_corners)
    if res is None:
        return None
    sums, ds = res
    holder = {}

    def bwd():
        n, c, h, w = scores.t.shape
        scale = holder["gscale"].reshape(1).expand(c).contiguous()

        def run(ex, dst):
            assert ex is None or ex is dst          # the loss is the only consumer of the scores
            dlow = dst if dst is not None else ops.new_act(n, c, h, w, scores.t.dtype, ds.device, c_alloc=scores.t.stride(3))
            return ops.affine_act(ds, scale, None, None, L.ACT_NONE, out=dlow)
        scores.add_grad(run)
    tape.push(bwd)
    return sums, holder


def bilinear_close(tape, scores, out_h, out_w, loss=None, align_corners=False):
    """The close of a training forward whose head is F.interpolate(scores, (H, W), bilinear, align_corners): the fp32
    NCHW logits -- or, with loss = (target, class weights, ignore label) from a model's fused_loss, the loss sums of
    bilinear_ce.  Returns what _NetFn / _NetLossFn expect from run_forward: (output, tape, holder)."""
    if loss is not None:
        res = bilinear_ce(tape, scores, out_h, out_w, *loss, align_corners=align_corners)
        if res is None:
            raise RuntimeError("fused_loss: esn_bilinear_ce does not take %dx%d scores for a %dx%d target"
                               % (scores.t.shape[2], scores.t.shape[3], out_h, out_w))
        return res[0], tape, res[1]
    logits, holder = bilinear_logits(tape, scores, out_h, out_w, torch.float32, align_corners=align_corners)
    return logits, tape, holder


def fused_bilinear_loss(model, train_forward, input, target, criterion, classes):
    """Body of a model's fused_loss(input, target, criterion): criterion(model(input), target) (train.py:351-352) with the
    bilinear head, CrossEntropyLoss2d and both their backward passes as ONE launch (esn_bilinear_ce).  train_forward(model,
    input, loss=...) is the model's tape forward ending in bilinear_close.  Eval mode, no-grad, any other criterion, more than
    32 classes or a target that is not (N, H, W) take the two-module form."""
    from utils.losses.loss import fused_head_spec
    spec = fused_head_spec(criterion, input.device, target, classes) if (model.training and torch.is_grad_enabled()) else None
    if spec is None or classes > 32 or target.dim() != 3 or tuple(target.shape[1:]) != tuple(input.shape[2:]):
        return criterion(model(input), target)
    tgt, w, ignore, reduction, distributed = spec
    return run_network_loss(model, lambda x: train_forward(model, x, loss=(tgt, w, ignore)), input, reduction, distributed)


def _accumulating(fn_into):
    """add_grad adapter for kernels with an `accumulate` flag: fn_into(dx, accumulate) fills / adds into dx."""
    def run(ex, dst, alloc):
        if dst is not None:           # slice of a concat gradient buffer (zero-filled at creation): add in place
            fn_into(dst, 1)
            return dst
        if ex is not None:
            fn_into(ex, 1)
            return ex
        dx = alloc()
        fn_into(dx, 0)
        return dx
    return run


def bilinear(tape, x, out_h, out_w, align_corners, out=None):
    """F.interpolate(x, (out_h, out_w), bilinear, align_corners) on NHWC activations (optionally into a concat slice)."""
    y = out if isinstance(out, V) else V(ops.bilinear(x.t, out_h, out_w, align_corners, out=out))
    if isinstance(out, V):
        ops.bilinear(x.t, out_h, out_w, align_corners, out=out.t)
    n, c, h, w = x.t.shape

    def bwd():
        dy = y.g

        def into(dx, acc):
            a, b = ops.tdesc(dy), ops.tdesc(dx)
            ops._call(L.lib.esn_bilinear_bwd_nhwc, "esn_bilinear_bwd_nhwc", (C.byref(a), C.byref(b), int(bool(align_corners)), acc),
                      ops._nbytes(dy) + ops._nbytes(dx))
        run = _accumulating(into)
        x.add_grad(lambda ex, dst: run(ex, dst, lambda: ops.new_act(n, c, h, w, dy.dtype, dy.device)))
    tape.push(bwd)
    return y


def adaptive_avgpool(tape, x, size):
    """F.adaptive_avg_pool2d(x, size)."""
    y = V(ops.adaptive_avgpool(x.t, size))
    n, c, h, w = x.t.shape

    def bwd():
        dy = y.g

        def into(dx, acc):
            a, b = ops.tdesc(dy), ops.tdesc(dx)
            ops._call(L.lib.esn_adaptive_avgpool_bwd, "esn_adaptive_avgpool_bwd", (C.byref(a), C.byref(b), acc),
                      ops._nbytes(dy) + ops._nbytes(dx))
        run = _accumulating(into)
        x.add_grad(lambda ex, dst: run(ex, dst, lambda: ops.new_act(n, c, h, w, dy.dtype, dy.device)))
    tape.push(bwd)
    return y


def avgpool3x3s2(tape, x, out=None):
    """AvgPool2d(3, stride 2, pad 1) (count_include_pad), optionally into a concat slice."""
    n, c, h, w = x.t.shape
    if out is None:
        out = V(ops.new_act(n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, x.t.dtype, x.t.device))
    y = out
    ops.avgpool3x3s2(x.t, y.t)

    def bwd():
        dy = y.g

        def into(dx, acc):
            a, b = ops.tdesc(dy), ops.tdesc(dx)
            ops._call(L.lib.esn_avgpool3x3s2_bwd, "esn_avgpool3x3s2_bwd", (C.byref(a), C.byref(b), acc),
                      ops._nbytes(dy) + ops._nbytes(dx))
        run = _accumulating(into)
        x.add_grad(lambda ex, dst: run(ex, dst, lambda: ops.new_act(n, c, h, w, dy.dtype, dy.device)))
    tape.push(bwd)
    return y


class _SubConv:
    """One group of a grouped nn.Conv2d, seen as a dense conv over channel slices (weight is a view)."""

    def __init__(self, conv, g):
        og, cg = conv.out_channels // conv.groups, conv.in_channels // conv.groups
        self.weight = conv.weight.detach()[g * og:(g + 1) * og]
        self.bias = None
        self.stride, self.padding, self.dilation, self.groups = conv.stride, conv.padding, conv.dilation, 1
        self.in_channels, self.out_channels = cg, og


class _GradCollector:
    """Stands in for the tape while the groups of a grouped conv run their backward."""

    def __init__(self, tape):
        self.tape, self.grads = tape, {}

    def push(self, fn):
        self.tape.push(fn)

    def add_param_grad(self, p, g):
        self.grads[id(p)] = g

    def off_critical_path(self, *tensors):
        return self.tape.off_critical_path(*tensors)

    @property
    def keep(self):
        return self.tape.keep


class GroupedConvT:
    """nn.Conv2d with 1 < groups < channels (ESPNetv2's g=4 1x1 convs, cnn_utils.py:27-110): one dense ConvT per
    group over channel slices of the input / output; the weight gradient is assembled once all groups are done."""

    def __init__(self, conv):
        assert conv.bias is None
        self.conv = conv
        self.subs = [_SubConv(conv, g) for g in range(conv.groups)]
        self.convts = [ConvT(sc) for sc in self.subs]
        self._ptr = conv.weight.data_ptr()

    def forward(self, tape, x, out=None):
        conv = self.conv
        if conv.weight.data_ptr() != self._ptr:          # parameter storage replaced (.to(), load): rebuild the views
            self.__init__(conv)
        n, _, h, w = x.t.shape
        if out is None:
            ho, wo = self.convts[0].preps()[0].out_hw(h, w)
            out = V(ops.new_act(n, conv.out_channels, ho, wo, x.t.dtype, x.t.device))
        coll = _GradCollector(tape)
        og, cg = conv.out_channels // conv.groups, conv.in_channels // conv.groups

        def assemble():      # pushed first => runs after every group's backward
            with tape.off_critical_path():      # same stream as the groups' weight-gradient kernels, after them
                full = torch.cat([coll.grads[id(sc.weight)] for sc in self.subs], 0)
            tape.add_param_grad(conv.weight, full)
        tape.push(assemble)
        for g, ct in enumerate(self.convts):
            ct.forward(coll, x.slice(g * cg, (g + 1) * cg), out=out.slice(g * og, (g + 1) * og))
        return out


def copy_into(tape, x, out):
    """out (a channel slice of a concat buffer) = x; the slice's gradient flows back to x."""
    ops.affine_act(x.t, None, None, None, L.ACT_NONE, out=out.t)

    def bwd():
        dy = out.g
        x.add_grad(lambda ex, dst: ops.affine_act(dy, None, None, None, L.ACT_NONE, out=dst, residual=ex)
                   if (ex is not None or dst is not None) else ops.affine_act(dy, None, None, None, L.ACT_NONE))
    tape.push(bwd)
    return out


_DROPOUT_CALLS = [0]


def _scale_nc(src, mask, residual=None, out=None):
    """out = src * mask[n][c] (+ residual): the per-plane scaling kernel of CGNet's FGlo gate (esn_fglo.cu)."""
    n, c, h, w = src.shape
    if out is None:
        out = ops.new_act(n, c, h, w, src.dtype, src.device)
    a, b = ops.tdesc(src), ops.tdesc(out)
    r = ops.tdesc(residual) if residual is not None else ops._NULL
    ops._call(L.lib.esn_scale_nc, "esn_scale_nc", (C.byref(a), C.c_void_p(mask.data_ptr()), C.byref(r), C.byref(b)),
              ops._nbytes(src) + ops._nbytes(out) + (ops._nbytes(residual) if residual is not None else 0))
    return out


def dropout(tape, x, p, per_channel=False, training=True, residual=None, c_alloc=None):
    """nn.Dropout (element-wise) / nn.Dropout2d (per (n, c) plane); with `residual` (a V) the result is dropout(x) + residual in
    the same pass (ERFNet's `output + input`, ERFNet.py:62-65).  Element-wise: the keep mask is regenerated from the seed in
    backward.  Per plane: the N x C scale factors (0 or 1 / (1 - p)) are drawn once into a small table and forward / backward
    are one per-plane scaling each -- the per-element kernel hashed every element (62 us per ERFNet block on a 17 MB tensor).
    The seed is drawn from torch's CPU generator, so torch.manual_seed makes runs repeatable; the iteration counter lives on
    the device, so CUDA-graph replays draw new masks.  c_alloc: pixel stride of the output buffer, zero behind the channels (for a
    consumer that reads a zero-padded channel count on the tensor cores); per-element path only."""
    if not training or p <= 0.0:
        return x if residual is None else add(tape, x, residual)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) + _DROPOUT_CALLS[0]
    _DROPOUT_CALLS[0] += 1
    n, c, h, w = x.t.shape
    def apply(src, pad=None):
        dst = ops.new_act(n, c, h, w, src.dtype, src.device, c_alloc=pad, zero=bool(pad) and pad != c)
        a, b = ops.tdesc(src), ops.tdesc(dst)
        ops._call(L.lib.esn_dropout_step, "esn_dropout", (C.byref(a), C.byref(b), C.c_uint64(seed),
                                                          C.c_void_p(ops.step_counter(src.device).data_ptr()), C.c_float(p),
                                                          int(per_channel)), ops._nbytes(src) + ops._nbytes(dst))
        return dst

    def nc_ok(t):       # what esn_scale_nc takes: 4-channel vectors
        return t is None or (ops.is_nhwc(t) and t.shape[1] % 4 == 0 and t.stride(3) % 4 == 0
                             and t.data_ptr() % (4 * t.element_size()) == 0)

    if per_channel and nc_ok(x.t) and (residual is None or nc_ok(residual.t)):
        mask = torch.empty(n * c, dtype=torch.float32, device=x.t.device)
        ops._call(L.lib.esn_dropout_mask_nc, "esn_dropout_mask_nc",
                  (C.c_void_p(mask.data_ptr()), C.c_int64(n * c), C.c_uint64(seed),
                   C.c_void_p(ops.step_counter(x.t.device).data_ptr()), C.c_float(p)))
        y = V(_scale_nc(x.t, mask, residual=None if residual is None else residual.t))

        def bwd_nc():
            dy = y.g
            if residual is not None:
                residual.add_grad(lambda ex, dst: dy if (ex is None and dst is None) else
                                  ops.affine_act(dy, None, None, None, L.ACT_NONE, out=dst, residual=ex))

            def run(ex, dst):
                if nc_ok(dy) and nc_ok(ex) and nc_ok(dst):
                    return _scale_nc(dy, mask, residual=ex, out=dst)
                g = apply(dy)          # odd layouts: the per-element kernel draws the same mask (same keys, hash and counter)
                return g if (ex is None and dst is None) else ops.affine_act(g, None, None, None, L.ACT_NONE, out=dst, residual=ex)
            x.add_grad(run)
        tape.push(bwd_nc)
        return y

    y = V(apply(x.t, c_alloc))

    def bwd():
        g = apply(y.g)
        x.add_grad(lambda ex, dst: g if (ex is None and dst is None) else
                   ops.affine_act(g, None, None, None, L.ACT_NONE, out=dst, residual=ex))
    tape.push(bwd)
    return y if residual is None else add(tape, y, residual)


# --------------------------------------------------------------------------- autograd glue
class _NetFn(torch.autograd.Function):
    """One autograd node for the whole network: forward runs the kernels and records the tape,
    backward replays it and returns the parameter gradients."""

    @staticmethod
    def forward(ctx, run_forward, x, *params):
        logits, tape, holder = run_forward(x)
        tape.flush_counters()
        ctx.tape, ctx.holder, ctx.params = tape, holder, params
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ctx.holder["dlogits"] = dlogits.contiguous()
        grads = ctx.tape.backward()
        out = tuple(grads.get(p) for p in ctx.params)
        # the tape must not keep the gradients alive: AccumulateGrad adopts a returned tensor only when nobody else holds it
        # (otherwise it clones: one device copy per parameter, 213 of them in a DABNet step)
        ctx.tape.param_grads = {}
        del grads
        return (None, None) + out


class _NetLossFn(torch.autograd.Function):
    """The whole network AND its loss as one autograd node (the fused close of ops.bilinear_ce): forward returns the
    weighted-mean (or summed) cross-entropy, backward hands d loss / d scores to the tape.  Same global-batch normalisation
    as _CEFn when `distributed`."""

    @staticmethod
    def forward(ctx, run_forward, x, reduction, distributed, *params):
        sums, tape, holder = run_forward(x)
        tape.flush_counters()
        if (distributed and torch.distributed.is_available() and torch.distributed.is_initialized()
                and torch.distributed.get_world_size() > 1):
            torch.distributed.all_reduce(sums)
        ctx.tape, ctx.holder, ctx.params, ctx.reduction = tape, holder, params, reduction
        ctx.save_for_backward(sums)
        return sums[0] / sums[1] if reduction == "mean" else sums[0].clone()

    @staticmethod
    def backward(ctx, gout):
        (sums,) = ctx.saved_tensors
        g = gout.detach().float().reshape(())
        ctx.holder["gscale"] = g / sums[1] if ctx.reduction == "mean" else g
        grads = ctx.tape.backward()
        out = tuple(grads.get(p) for p in ctx.params)
        ctx.tape.param_grads = {}
        del grads
        return (None, None, None, None) + out


def run_network_loss(model, run_forward, x, reduction="mean", distributed=False):
    """run_forward(x) -> (sums, tape, holder) with the loss sums of T.bilinear_ce; returns the loss (a scalar with grad_fn)."""
    params = [p for p in model.parameters() if p.requires_grad]
    return _NetLossFn.apply(run_forward, x, reduction, distributed, *params)


def run_network(model, run_forward, x):
    params = [p for p in model.parameters() if p.requires_grad]
    return _NetFn.apply(run_forward, x, *params)


class _CEFn(torch.autograd.Function):
    """CrossEntropyLoss2d (utils/losses/loss.py:15-32): weighted mean over non-ignored pixels of the
    GLOBAL batch -- under torch.distributed the two sums are all-reduced before the division so the
    loss and its gradient equal the reference's gathered-batch value."""

    @staticmethod
    def forward(ctx, logits, target, weight, ignore_label, distributed=False, reduction="mean", keep_thresh=None):
        lg = logits.detach().contiguous()
        sums, _ = ops.weighted_ce(lg, target, weight, ignore_label, want_grad=False, keep_thresh=keep_thresh)
        ctx.keep_thresh = keep_thresh
        if (distributed and torch.distributed.is_available() and torch.distributed.is_initialized()
                and torch.distributed.get_world_size() > 1):
            torch.distributed.all_reduce(sums)
        ctx.save_for_backward(lg, target, sums)
        ctx.weight, ctx.ignore, ctx.reduction = weight, ignore_label, reduction
        return sums[0] / sums[1] if reduction == "mean" else sums[0].clone()

    @staticmethod
    def backward(ctx, gout):
        lg, target, sums = ctx.saved_tensors
        gout = gout.detach().float().reshape(1).contiguous()
        scratch = torch.zeros(2, dtype=torch.float32, device=lg.device)
        gnorm = sums[1:2] if ctx.reduction == "mean" else torch.ones(1, dtype=torch.float32, device=lg.device)
        _, g = ops.weighted_ce(lg, target, ctx.weight, ctx.ignore, want_grad=True, sums=scratch, gnorm=gnorm, gout=gout,
                               keep_thresh=ctx.keep_thresh)
        return g, None, None, None, None, None, None


def cross_entropy(logits, target, weight=None, ignore_label=255, distributed=False, reduction="mean", keep_thresh=None):
    """distributed=True all-reduces (sum w*nll, sum w) over the default process group: only correct together with
    SUM-reduced gradients (esn.parallel); see utils/losses/loss.py.  keep_thresh: device scalar, pixels whose labelled-class
    probability exceeds it are ignored (OHEM)."""
    return _CEFn.apply(logits, target, weight, ignore_label, distributed, reduction, keep_thresh)


def fglo(tape, fc, x, out=None, residual=None):
    """CGNet's global-context gate FGlo (CGNet.py:173-191): y = x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2) (+ residual).
    The activation-sized work is three kernels forward (partial sums, gate multiply) and two backward (sum_hw dy*x, dy*g + the
    pooled-mean gradient); the two nn.Linear layers act on (N, C) vectors -- a few kFLOP -- and go through torch autograd on a
    local graph, which also yields their weight / bias gradients.  fc = the module's nn.Sequential(Linear, ReLU, Linear, Sigmoid)."""
    xt = x.t
    n, c, h, w = xt.shape
    dev = xt.device
    dx_ = ops.tdesc(xt)
    chunks = L.lib.esn_global_avgpool_chunks(C.byref(dx_))
    sums = torch.empty((chunks, n, c), dtype=torch.float32, device=dev)
    ops._call(L.lib.esn_global_avgpool, "esn_global_avgpool", (C.byref(dx_), C.c_void_p(sums.data_ptr())), ops._nbytes(xt))
    lin1, lin2 = fc[0], fc[2]
    params = [lin1.weight, lin1.bias, lin2.weight, lin2.bias]
    with torch.enable_grad(), torch.autocast("cuda", enabled=False):      # the gate is fp32 in every precision mode, as in inference
        pooled = (sums.sum(0) / float(h * w)).requires_grad_(True)
        gate = torch.sigmoid(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(pooled, lin1.weight.float(), lin1.bias.float())),
                                                        lin2.weight.float(), lin2.bias.float()))
    g = gate.detach().contiguous()
    y = out if isinstance(out, V) else V(out if out is not None else ops.new_act(n, c, h, w, xt.dtype, dev))
    dy_ = ops.tdesc(y.t)
    dr_ = ops.tdesc(residual.t) if residual is not None else ops._NULL
    ops._call(L.lib.esn_scale_add_nc, "esn_scale_add_nc", (C.byref(dx_), C.c_void_p(g.data_ptr()), None, C.byref(dr_), C.byref(dy_)),
              ops._nbytes(xt) + ops._nbytes(y.t) + (ops._nbytes(residual.t) if residual is not None else 0))

    def bwd():
        dy = y.g
        dg = _f32zeros((n, c), dev)
        a, b = ops.tdesc(dy), ops.tdesc(xt)
        ops._call(L.lib.esn_dot_nc, "esn_dot_nc", (C.byref(a), C.byref(b), C.c_void_p(dg.data_ptr())), ops._nbytes(dy) + ops._nbytes(xt))
        grads = torch.autograd.grad(gate, [pooled] + params, dg)
        dpool = (grads[0] / float(h * w)).contiguous()
        for p_, g_ in zip(params, grads[1:]):
            tape.add_param_grad(p_, g_)
        if residual is not None:
            residual.add_grad(lambda ex, dst: dy if (ex is None and dst is None) else
                              ops.affine_act(dy, None, None, None, L.ACT_NONE, out=dst, residual=ex))

        def run(ex, dst):
            dx = dst if dst is not None else ops.new_act(n, c, h, w, dy.dtype, dev)
            e = ops.tdesc(ex) if ex is not None else ops._NULL
            o = ops.tdesc(dx)
            ops._call(L.lib.esn_scale_add_nc, "esn_scale_add_nc",
                      (C.byref(a), C.c_void_p(g.data_ptr()), C.c_void_p(dpool.data_ptr()), C.byref(e), C.byref(o)),
                      2 * ops._nbytes(dy) + (ops._nbytes(ex) if ex is not None else 0))
            return dx
        x.add_grad(run)
    tape.push(bwd)
    return y
