"""Host-side wrappers: torch tensors in, C-ABI calls out (no torch compute on the path).

Internal activations are torch tensors of logical shape (N,C,H,W) whose memory
is NHWC ("channels_last"), possibly a channel slice of a wider concat buffer.
PyTorch is used for device memory and streams only.
"""
import ctypes as C
import os

import torch

from . import _lib as L

_DT = {torch.float32: L.ESN_F32, torch.bfloat16: L.ESN_BF16, torch.uint8: L.ESN_U8, torch.int64: L.ESN_I64}
_NULL = L.EsnTensor()
UMMA_ENABLED = os.environ.get("ESN_DISABLE_UMMA", "0") != "1"
PAIR_DISABLED = os.environ.get("ESN_DISABLE_PAIR", "0") == "1"
CONVT_FUSED = os.environ.get("ESN_DISABLE_FUSED_CONVT", "0") != "1"


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_STEP = {}


def step_counter(device=None):
    """Device-resident training-iteration counter (int64 scalar per device) mixed into the dropout seeds; a graphed
    training step (esn/graph.py) advances it inside the graph so that replays draw fresh masks."""
    dev = torch.device(device if device is not None else "cuda")
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    t = _STEP.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("esn.ops.step_counter must exist before CUDA-graph capture (run one eager iteration first)")
        t = torch.zeros((), dtype=torch.int64, device=torch.device("cuda", key))
        _STEP[key] = t
    return t


def advance_step_counter(device=None):
    step_counter(device).add_(1)


# Optional per-launch profile (bench.py's roofline leg): when PROFILE is a list every C-ABI call is
# bracketed by CUDA events on the launching stream and (kernel, algorithmic bytes, flops, events)
# is appended.  None (the default) adds no work to the hot path.
PROFILE = None


def _nbytes(t):
    return 0 if t is None else t.shape[0] * t.shape[1] * t.shape[2] * t.shape[3] * t.element_size()


_WARNED_UNSUPPORTED = set()


def _call(fn, name, arg_refs, alg_bytes=0, flops=0, tag="", allow_unsupported=False):
    """Launch one C-ABI entry point on the current stream.  Returns True.  With allow_unsupported, an
    ESN_ERR_UNSUPPORTED answer (the entry point's own shape gate, checked before anything is launched) returns False
    instead of raising, so the caller can take the next CUDA route; it is reported once per (entry point, shape tag)."""
    if PROFILE is None:
        rc = fn(*arg_refs, stream())
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*arg_refs, stream())
        e1.record()
    if rc == L.ERR_UNSUPPORTED and allow_unsupported:
        if (name, tag) not in _WARNED_UNSUPPORTED:
            _WARNED_UNSUPPORTED.add((name, tag))
            import warnings
            warnings.warn("%s declined %s (host-side gate out of date?); taking the direct CUDA kernel" % (name, tag))
        return False
    L.check(rc, name)
    if PROFILE is not None:
        PROFILE.append({"kernel": name, "tag": tag, "bytes": int(alg_bytes), "flops": int(flops), "ev": (e0, e1)})
    return True


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s: tensors must live on a CUDA device; this framework has no CPU path" % what)
    # the C side launches on the CURRENT device and on torch's current stream of it (esn_current_device, stream()): a tensor
    # on another device would be read through a stale peer mapping or fault -- one process per GPU, or torch.cuda.device(...)
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise RuntimeError("%s: tensor lives on cuda:%d but the current device is cuda:%d; wrap the call in "
                           "torch.cuda.device(tensor.device) (one process per GPU is the supported form)"
                           % (what, t.device.index, torch.cuda.current_device()))


def widen(t, c):
    """The same NHWC buffer seen with `c` logical channels (c <= pixel stride): used to run
    vectorised / tensor-core kernels over zero-padded channel tails of concat buffers."""
    assert is_nhwc(t) and c <= t.stride(3)
    n, _, h, w = t.shape
    return t.as_strided((n, c, h, w), t.stride(), t.storage_offset())


def is_nhwc(t):
    if t.dim() != 4:
        return False
    n, c, h, w = t.shape
    s = t.stride()
    return s[1] == 1 and s[3] >= c and s[2] == w * s[3] and s[0] == h * s[2] and (c > 1 or s[3] == 1)


def tdesc(t):
    """EsnTensor view of a 4-D torch tensor (NHWC-strided or NCHW-contiguous)."""
    n, c, h, w = t.shape
    d = L.EsnTensor()
    d.ptr = t.data_ptr()
    d.dtype = _DT[t.dtype]
    d.n, d.c, d.h, d.w = n, c, h, w
    if is_nhwc(t):
        d.layout, d.c_stride = L.ESN_NHWC, t.stride(3)
    elif t.is_contiguous():
        d.layout, d.c_stride = L.ESN_NCHW, 0
    else:
        raise RuntimeError("tensor is neither NHWC-strided nor NCHW-contiguous: shape %s stride %s" % (tuple(t.shape), t.stride()))
    return d


def new_act(n, c, h, w, dtype, device, c_alloc=None, zero=False):
    """Fresh NHWC activation with logical shape (N,C,H,W); c_alloc pads the pixel stride."""
    ca = c if c_alloc is None else c_alloc
    buf = (torch.zeros if zero else torch.empty)((n, h, w, ca), dtype=dtype, device=device)
    t = buf.permute(0, 3, 1, 2)
    return t if ca == c else t[:, :c]


def compute_dtype(x):
    """fp32 unless bf16 autocast is active, ESN_COMPUTE=bf16, or the activations already are bf16."""
    if x.dtype == torch.bfloat16:
        return torch.bfloat16
    if torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16:
        return torch.bfloat16
    if os.environ.get("ESN_COMPUTE", "").lower() == "bf16":
        return torch.bfloat16
    return torch.float32


def as_act(x, dtype=None):
    """Accept the caller's NCHW tensor or an internal NHWC one; return NHWC in `dtype`."""
    require_cuda(x, "forward")
    dtype = dtype or compute_dtype(x)
    if is_nhwc(x) and x.dtype == dtype:
        return x
    if not x.is_contiguous() and not is_nhwc(x):
        x = x.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    n, c, h, w = x.shape
    # channel counts that are not a multiple of 8 get a padded pixel stride; a single channel stays dense (for C = 1 the
    # NHWC and NCHW layouts coincide and tdesc accepts only the dense form)
    y = new_act(n, c, h, w, dtype, x.device, c_alloc=(c + 7) // 8 * 8 if (c % 8 and c > 1) else None)
    dx, dy = tdesc(x), tdesc(y)
    if dx.layout == L.ESN_NHWC:  # NHWC but other dtype: elementwise copy through the affine kernel
        return affine_act(x, None, None, None, L.ACT_NONE, out=y)
    _call(L.lib.esn_convert_layout, "esn_convert_layout", (C.byref(dx), C.byref(dy)), _nbytes(x) + _nbytes(y))
    return y


def to_nchw(x, dtype=None):
    """NHWC internal activation -> NCHW-contiguous tensor (what the reference returns)."""
    n, c, h, w = x.shape
    y = torch.empty((n, c, h, w), dtype=dtype or x.dtype, device=x.device)
    dx, dy = tdesc(x), tdesc(y)
    if dx.layout == L.ESN_NCHW:
        return x if dtype in (None, x.dtype) else x.to(dtype)
    L.check(L.lib.esn_convert_layout(C.byref(dx), C.byref(dy), stream()), "esn_convert_layout")
    return y


# --------------------------------------------------------------------------- prepared (packed) parameters
def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def bn_affine(bn, device):
    """Eval-mode BatchNorm2d as y = x*scale + shift (reads bn.eps: SURVEY H11)."""
    w, b = _f32(bn.weight, device), _f32(bn.bias, device)
    m, v = _f32(bn.running_mean, device), _f32(bn.running_var, device)
    scale = w / torch.sqrt(v + bn.eps)
    return scale, b - m * scale


class ConvPrep:
    """Packed weights + folded epilogue of one conv (+BN slice) (+activation)."""

    def __init__(self, conv, scale=None, shift=None, act=L.ACT_NONE, alpha=None, device=None, cin_pad=None,
                 cout_pad=None):
        """cin_pad / cout_pad: zero-extend the weight so the kernel sees a channel-padded NHWC view
        (padded input channels must hold finite values; padded outputs evaluate to act(0))."""
        device = device or conv.weight.device
        transposed = isinstance(conv, torch.nn.ConvTranspose2d)
        self._init(_f32(conv.weight, device), None if conv.bias is None else _f32(conv.bias, device), conv.stride[0],
                   conv.padding, conv.dilation, conv.groups, transposed,
                   conv.output_padding[0] if transposed else 0, scale, shift, act, alpha, device, cin_pad, cout_pad)

    @classmethod
    def from_weight(cls, w, stride=1, padding=(0, 0), dilation=(1, 1), groups=1, transposed=False, out_pad=0,
                    bias=None, scale=None, shift=None, act=L.ACT_NONE, alpha=None, cin_pad=None, cout_pad=None):
        """Build from a raw weight tensor: (Cout, Cin/groups, kh, kw), or (Cin, Cout, kh, kw) if transposed."""
        self = cls.__new__(cls)
        self._init(w.detach().float(), bias, stride, tuple(padding), tuple(dilation), groups, transposed, out_pad,
                   scale, shift, act, alpha, w.device, cin_pad, cout_pad)
        return self

    def _init(self, w, bias, stride, padding, dilation, groups, transposed, out_pad, scale, shift, act, alpha, device,
              cin_pad, cout_pad):
        if transposed and not (groups > 1 and groups == w.shape[0] and w.shape[1] == 1):
            w = w.permute(1, 0, 2, 3).contiguous()          # (Cin, Cout, kh, kw) -> (Cout, Cin, kh, kw); depthwise: (C, 1, kh, kw) as is
        if cin_pad is not None and cin_pad > w.shape[1]:
            assert groups == 1
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin_pad - w.shape[1]))
        if cout_pad is not None and cout_pad > w.shape[0]:
            assert groups == 1
            extra = cout_pad - w.shape[0]
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, extra))
            zeros = torch.zeros(extra, device=device)
            scale = torch.cat([torch.ones(w.shape[0] - extra, device=device) if scale is None else scale, zeros + 1])
            shift = torch.cat([torch.zeros(w.shape[0] - extra, device=device) if shift is None else shift, zeros])
            if alpha is not None:
                alpha = torch.cat([_f32(alpha, device), zeros])
            if bias is not None:
                bias = torch.cat([bias, zeros])
        w = w.contiguous()
        self.cout, cin_g, self.kh, self.kw = w.shape
        self.groups = groups
        self.cin = cin_g * groups
        self.depthwise = self.groups > 1 and self.groups == self.cin == self.cout
        self.grouped = self.groups > 1 and not self.depthwise
        if self.grouped and (transposed or self.cout % groups):
            raise NotImplementedError("grouped transposed convs are not supported")
        self.transposed = int(transposed)
        self.stride = stride
        self.pad_h, self.pad_w = padding
        self.dil_h, self.dil_w = dilation
        # a dilation along an axis with a single tap has no effect (EDANet.py:53-54 passes dilation=d to 3x1 / 1x3 convs)
        if self.kh == 1:
            self.dil_h = 1
        if self.kw == 1:
            self.dil_w = 1
        self.out_pad = out_pad
        taps = self.kh * self.kw
        # direct layout: [tap][Cin/groups][Cout]
        self.w_direct = w.permute(2, 3, 1, 0).reshape(taps, cin_g, self.cout).contiguous()
        self._w_umma = None
        self._w_src = w
        self.cout_pad = (self.cout + 15) // 16 * 16
        sc = torch.ones(self.cout, device=device) if scale is None else scale.clone()
        sh = torch.zeros(self.cout, device=device) if shift is None else shift.clone()
        if bias is not None:
            sh = sh + bias * sc
        self.scale, self.shift = sc.contiguous(), sh.contiguous()
        self.act = act
        self.alpha = None if alpha is None else _f32(alpha, device)

    def cin_split(self):
        """Sub-convs over 64-channel-multiple input slices: scale on every part, shift + activation on the last."""
        if getattr(self, "_parts", None) is None:
            import copy
            taps = self.kh * self.kw
            # weights of one part stay resident in shared memory next to the A ring and the output staging
            staging = 128 * self.cout * 2 if (self.cout % 8 == 0 and (self.cout <= 64 or self.cout % 64 == 0)) else 0
            budget = 226 * 1024 - 6144 - 2 * 128 * 64 * 2 - staging
            per = max(64, (budget // (taps * self.cout_pad * 2)) // 64 * 64)
            parts, lo = [], 0
            while lo < self.cin:
                hi = min(self.cin, lo + per)
                q = copy.copy(self)
                q._parts, q._w_umma = None, None
                q._w_src = self._w_src[:, lo:hi].contiguous()
                q.cin, q.cin_lo = hi - lo, lo // 64
                q.w_direct = q._w_src.permute(2, 3, 1, 0).reshape(taps, hi - lo, self.cout).contiguous()
                last = hi == self.cin
                if not last:
                    q.shift = torch.zeros_like(self.shift)
                    q.act, q.alpha = L.ACT_NONE, None
                parts.append(q)
                lo = hi
            self._parts = parts
        return self._parts

    def group_split(self):
        """A grouped conv as `groups` dense sub-convs over channel slices of the input and output."""
        if getattr(self, "_gparts", None) is None:
            import copy
            taps = self.kh * self.kw
            cg, og = self.cin // self.groups, self.cout // self.groups
            parts = []
            for g in range(self.groups):
                q = copy.copy(self)
                q._gparts, q._oparts, q._parts, q._w_umma = None, None, None, None
                q.groups, q.grouped, q.depthwise = 1, False, False
                q._w_src = self._w_src[g * og:(g + 1) * og].contiguous()
                q.cin, q.cout, q.cout_pad = cg, og, (og + 15) // 16 * 16
                q.cin_lo_g, q.cout_lo = g * cg, g * og
                q.w_direct = q._w_src.permute(2, 3, 1, 0).reshape(taps, cg, og).contiguous()
                q.scale, q.shift = self.scale[g * og:(g + 1) * og].clone(), self.shift[g * og:(g + 1) * og].clone()
                q.alpha = None if self.alpha is None else self.alpha[g * og:(g + 1) * og].clone()
                parts.append(q)
            self._gparts = parts
        return self._gparts

    def cout_split(self):
        """Sub-convs over <=256-output-channel slices (each with its slice of the epilogue parameters)."""
        if getattr(self, "_oparts", None) is None:
            import copy
            taps = self.kh * self.kw
            parts, lo = [], 0
            while lo < self.cout:
                hi = min(self.cout, lo + 256)
                q = copy.copy(self)
                q._oparts, q._parts, q._w_umma = None, None, None
                q._w_src = self._w_src[lo:hi].contiguous()
                q.cout, q.cout_lo, q.cout_pad = hi - lo, lo, (hi - lo + 15) // 16 * 16
                q.w_direct = q._w_src.permute(2, 3, 1, 0).reshape(taps, self.cin // self.groups, hi - lo).contiguous()
                q.scale, q.shift = self.scale[lo:hi].contiguous(), self.shift[lo:hi].contiguous()
                q.alpha = None if self.alpha is None else self.alpha[lo:hi].contiguous()
                parts.append(q)
                lo = hi
            self._oparts = parts
        return self._oparts

    @property
    def w_umma(self):
        if self._w_umma is None:     # bf16 [tap][Cout_pad][Cin]
            w = self._w_src
            taps = self.kh * self.kw
            p = torch.zeros((taps, self.cout_pad, self.cin), dtype=torch.bfloat16, device=w.device)
            p[:, :self.cout] = w.permute(2, 3, 0, 1).reshape(taps, self.cout, self.cin).to(torch.bfloat16)
            self._w_umma = p.contiguous()
        return self._w_umma

    @property
    def w_umma_scaled(self):
        """bf16 [tap][Cout_pad][Cin] with the epilogue scale folded into the output channels (the fused pair
        kernel adds the residual inside the accumulator, so the scale cannot be applied afterwards)."""
        if getattr(self, "_w_umma_s", None) is None:
            w = self._w_src * self.scale.view(-1, 1, 1, 1)
            taps = self.kh * self.kw
            p = torch.zeros((taps, self.cout_pad, self.cin), dtype=torch.bfloat16, device=w.device)
            p[:, :self.cout] = w.permute(2, 3, 0, 1).reshape(taps, self.cout, self.cin).to(torch.bfloat16)
            self._w_umma_s = p.contiguous()
        return self._w_umma_s

    def fused_convt(self):
        """ConvTranspose2d(3, stride 2, pad 1, output_padding 1) as ONE stride-1 conv with 2x2 taps and 4*Cout outputs
        ordered (row parity a, column parity b, c): out[2i+a, 2j+b, c] = sum_{dy,dx} x[i+dy, j+dx] . w[a+1-2dy, b+1-2dx]
        (taps outside 0..2 are zero blocks).  The kernel stores the two row parities through pixel-shuffle tensor maps."""
        if getattr(self, "_fused", None) is None:
            w = self._w_src                      # (Cout, Cin, 3, 3) (already transposed from the module's (Cin, Cout, 3, 3))
            co, ci = w.shape[0], w.shape[1]
            f = torch.zeros((2, 2, 2, 2, co, ci), dtype=torch.float32, device=w.device)      # [dy][dx][a][b][c][ci]
            for dy in range(2):
                for dx in range(2):
                    for a in range(2):
                        for b in range(2):
                            r, q = a + 1 - 2 * dy, b + 1 - 2 * dx
                            if 0 <= r <= 2 and 0 <= q <= 2:
                                f[dy, dx, a, b] = w[:, :, r, q]
            wf = f.reshape(4, 4 * co, ci).to(torch.bfloat16).contiguous()                      # [tap][4*Cout][Cin]
            rep = lambda v: None if v is None else v.repeat(4).contiguous()
            self._fused = (wf, rep(self.scale), rep(self.shift), rep(self.alpha))
        return self._fused

    def out_hw(self, h, w):
        if self.transposed:
            return ((h - 1) * self.stride - 2 * self.pad_h + self.dil_h * (self.kh - 1) + self.out_pad + 1,
                    (w - 1) * self.stride - 2 * self.pad_w + self.dil_w * (self.kw - 1) + self.out_pad + 1)
        return ((h + 2 * self.pad_h - self.dil_h * (self.kh - 1) - 1) // self.stride + 1,
                (w + 2 * self.pad_w - self.dil_w * (self.kw - 1) - 1) // self.stride + 1)


def _epilogue(ep, scale, shift, alpha, act, residual, flags=0):
    ep.flags = flags
    ep.scale = scale.data_ptr() if scale is not None else None
    ep.shift = shift.data_ptr() if shift is not None else None
    ep.alpha = alpha.data_ptr() if alpha is not None else None
    ep.act = act
    ep.residual = tdesc(residual) if residual is not None else _NULL


def conv2d(x, prep, out=None, residual=None, force_direct=False):
    """y = act(conv(x)*scale + shift (+ residual)); routes to tcgen05 when the shape allows."""
    n, c, h, w = x.shape
    assert c == prep.cin, (c, prep.cin)
    ho, wo = prep.out_hw(h, w)
    if out is None:
        odt = x.dtype if x.dtype == torch.bfloat16 else torch.float32
        out = new_act(n, prep.cout, ho, wo, odt, x.device)
    if prep.grouped:
        for q in prep.group_split():
            conv2d(x[:, q.cin_lo_g:q.cin_lo_g + q.cin], q, out=out[:, q.cout_lo:q.cout_lo + q.cout],
                   residual=None if residual is None else residual[:, q.cout_lo:q.cout_lo + q.cout],
                   force_direct=force_direct)
        return out
    p = L.EsnConv()
    p.x, p.y = tdesc(x), tdesc(out)
    p.kh, p.kw, p.stride = prep.kh, prep.kw, prep.stride
    p.pad_h, p.pad_w, p.dil_h, p.dil_w = prep.pad_h, prep.pad_w, prep.dil_h, prep.dil_w
    p.groups, p.transposed, p.cout_pad = prep.groups, prep.transposed, prep.cout_pad
    _epilogue(p.ep, prep.scale, prep.shift, prep.alpha, prep.act, residual, getattr(prep, "ep_flags", 0))
    alg = _nbytes(x) + _nbytes(out) + _nbytes(residual)
    flops = 2 * n * ho * wo * prep.cout * (prep.cin // prep.groups) * prep.kh * prep.kw
    if prep.transposed:
        flops = 2 * n * h * w * prep.cout * prep.cin * prep.kh * prep.kw
    tag = "%dx%d c%d-%d s%d d%d%s" % (prep.kh, prep.kw, prep.cin, prep.cout, prep.stride,
                                       max(prep.dil_h, prep.dil_w), "T" if prep.transposed else "")
    tc_ok = (UMMA_ENABLED and not force_direct and x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
             and prep.groups == 1 and p.x.layout == L.ESN_NHWC)
    if (tc_ok and prep.transposed and CONVT_FUSED and residual is None and (prep.kh, prep.kw, prep.stride) == (3, 3, 2)
            and (prep.pad_h, prep.pad_w) == (1, 1) and (ho, wo) == (2 * h, 2 * w) and prep.cout in (8, 16, 32, 64)
            and (prep.cin in (16, 32, 64) or prep.cin % 64 == 0) and out.stride(3) == prep.cout
            and 4 * prep.cin * 4 * prep.cout * 2 <= 120 * 1024 and x.stride(3) % 8 == 0 and x.data_ptr() % 16 == 0
            and out.data_ptr() % 16 == 0):
        # one launch instead of four output-parity phases: x is read once
        wf, sc4, sh4, al4 = prep.fused_convt()
        p.w = wf.data_ptr()
        p.kh, p.kw, p.stride, p.pad_h, p.pad_w, p.transposed, p.cout_pad = 2, 2, 1, 0, 0, 2, 4 * prep.cout
        _epilogue(p.ep, sc4, sh4, al4, prep.act, None, getattr(prep, "ep_flags", 0))
        _call(L.lib.esn_conv2d_umma, "esn_conv2d_umma", (C.byref(p),), alg, flops, tag + "f")
        return out
    if tc_ok and umma_supported(prep, p):
        p.w = prep.w_umma.data_ptr()
        if _call(L.lib.esn_conv2d_umma, "esn_conv2d_umma", (C.byref(p),), alg, flops, tag, allow_unsupported=True):
            return out
        tc_ok = False          # declined by the entry point's own gate: straight to the direct kernel, no slicing
    if tc_ok and prep.cout_pad > 256:
        # more output channels than one UMMA N tile: run 256-channel slices of the weight
        for q in prep.cout_split():
            conv2d(x, q, out=out[:, q.cout_lo:q.cout_lo + q.cout],
                   residual=None if residual is None else residual[:, q.cout_lo:q.cout_lo + q.cout])
        return out
    if tc_ok and prep.cin % 64 == 0 and prep.cin > 64:
        # all taps of the full-Cin weight do not fit in shared memory: run the conv as a sum over
        # 64-channel input slices, carrying the scaled partial sum through the residual operand
        parts = prep.cin_split()
        if parts is not None and all(umma_supported(q, p) for q in parts):
            acc = residual
            for i, q in enumerate(parts):
                last = i == len(parts) - 1
                xi = x[:, 64 * q.cin_lo:64 * q.cin_lo + q.cin]
                # partial sums live in private buffers whose pixel stride is a multiple of 8 channels (TMA alignment)
                yi = out if last else new_act(n, prep.cout, ho, wo, out.dtype, x.device,
                                              c_alloc=(prep.cout + 7) // 8 * 8 if prep.cout % 8 else None)
                conv2d(xi, q, out=yi, residual=acc)
                acc = yi
            return out
    p.w = prep.w_direct.data_ptr()
    _call(L.lib.esn_conv2d_direct, "esn_conv2d_direct", (C.byref(p),), alg, flops, tag)
    return out


DUAL_ENABLED = os.environ.get("ESN_DUAL", "1") != "0"


def conv2d_then_affine(x, prep, scale2, shift2, alpha2, act2, out2, out=None, residual=None, store_y=True):
    """y = conv2d(x, prep, residual) and y2 = act2(y * scale2 + shift2) in one tcgen05 launch (esn_conv2d_umma_dual): the
    second stage reads the value as it is stored, so the result is bit-identical to conv2d followed by affine_act.  With
    store_y=False only y2 is written.  Shapes the dual entry point does not take run as those two launches.
    Returns (y or None, y2)."""
    n, c, h, w = x.shape
    assert c == prep.cin, (c, prep.cin)
    ho, wo = prep.out_hw(h, w)
    if out is None and store_y:
        out = new_act(n, prep.cout, ho, wo, x.dtype if x.dtype == torch.bfloat16 else torch.float32, x.device)
    yref = out if store_y else out2
    ok = (DUAL_ENABLED and UMMA_ENABLED and x.dtype == torch.bfloat16 and yref.dtype == torch.bfloat16
          and out2.dtype == torch.bfloat16 and not prep.grouped and prep.groups == 1 and not prep.transposed
          and prep.cout % 8 == 0 and (prep.cout <= 64 or prep.cout % 64 == 0) and prep.cout_pad <= 256 and is_nhwc(x)
          and is_nhwc(out2) and out2.stride(3) % 8 == 0 and out2.data_ptr() % 16 == 0)
    if ok:
        d = L.EsnConvDual()
        p = d.conv
        p.x, p.y = tdesc(x), tdesc(yref)
        p.kh, p.kw, p.stride = prep.kh, prep.kw, prep.stride
        p.pad_h, p.pad_w, p.dil_h, p.dil_w = prep.pad_h, prep.pad_w, prep.dil_h, prep.dil_w
        p.groups, p.transposed, p.cout_pad = prep.groups, prep.transposed, prep.cout_pad
        _epilogue(p.ep, prep.scale, prep.shift, prep.alpha, prep.act, residual, getattr(prep, "ep_flags", 0))
        ok = umma_supported(prep, p)
    if ok:
        p.w = prep.w_umma.data_ptr()
        d.y2 = tdesc(out2)
        d.scale2 = scale2.data_ptr() if scale2 is not None else None
        d.shift2 = shift2.data_ptr() if shift2 is not None else None
        d.alpha2 = alpha2.data_ptr() if alpha2 is not None else None
        d.act2, d.store_y = act2, 1 if store_y else 0
        alg = _nbytes(x) + _nbytes(out2) + _nbytes(residual) + (_nbytes(out) if store_y else 0)
        flops = 2 * n * ho * wo * prep.cout * prep.cin * prep.kh * prep.kw
        tag = "%dx%d c%d-%d s%d d%d %s" % (prep.kh, prep.kw, prep.cin, prep.cout, prep.stride, max(prep.dil_h, prep.dil_w),
                                           "dual" if store_y else "chain")
        if _call(L.lib.esn_conv2d_umma_dual, "esn_conv2d_umma_dual", (C.byref(d),), alg, flops, tag, allow_unsupported=True):
            return (out if store_y else None), out2
    if store_y:
        conv2d(x, prep, out=out, residual=residual)
        affine_act(out, scale2, shift2, alpha2, act2, out=out2)
        return out, out2
    conv2d(x, prep, out=out2, residual=residual)
    affine_act(out2, scale2, shift2, alpha2, act2, out=out2)
    return None, out2


def pair_supported(x, p1, p2, out, residual):
    """Mirror of esn_conv_pair_umma's gate (csrc/esn_umma_pair.cu)."""
    if not UMMA_ENABLED or PAIR_DISABLED or x.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        return False
    n, c, h, w = x.shape
    if c not in (16, 64) or p1.cout != c or p2.cout != c or p1.cin != c or p2.cin != c:
        return False
    if (p1.kh, p1.kw, p2.kh, p2.kw) != (3, 1, 1, 3) or p1.stride != 1 or p2.stride != 1 or p1.groups != 1 or p2.groups != 1:
        return False
    d = p1.dil_h
    if d != p2.dil_w or d < 1 or d > 8 or p1.pad_h != d or p2.pad_w != d or p1.pad_w != 0 or p2.pad_h != 0:
        return False
    if p1.act not in (L.ACT_NONE, L.ACT_RELU) or getattr(p1, "ep_flags", 0) or getattr(p2, "ep_flags", 0):
        return False
    bw = 128 * (64 // c)
    if w % bw or w // bw > 8:
        return False
    rb = 2 * c
    inter = ((w + 16) * rb + 1023) // 1024 * 1024
    wreg = (2 * 3 * c * rb + 1023) // 1024 * 1024
    if 1024 + wreg + inter + 2 * bw * rb + 1280 + 1024 + 4 * bw * rb > 226 * 1024:
        return False
    for t in (x, out) + ((residual,) if residual is not None else ()):
        if not is_nhwc(t) or t.stride(3) % 8 or t.data_ptr() % 16 or t.dtype != torch.bfloat16:
            return False
    return True


def conv_pair(x, p1, p2, out=None, residual=None):
    """Fused k x 1 -> act -> 1 x k -> affine (+residual) -> act on the tcgen05 pair kernel (caller checked pair_supported)."""
    n, c, h, w = x.shape
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device)
    p = L.EsnConvPair()
    p.x, p.y = tdesc(x), tdesc(out)
    p.w1, p.w2 = p1.w_umma.data_ptr(), p2.w_umma_scaled.data_ptr()
    p.taps, p.dilation = 3, p1.dil_h
    _epilogue(p.ep1, p1.scale, p1.shift, p1.alpha, p1.act, None)
    _epilogue(p.ep2, None, p2.shift, p2.alpha, p2.act, residual)      # scale2 lives in w2
    alg = _nbytes(x) + _nbytes(out) + _nbytes(residual)
    flops = 2 * 2 * n * h * w * c * c * 3
    _call(L.lib.esn_conv_pair_umma, "esn_conv_pair_umma", (C.byref(p),), alg, flops, "3x1+1x3 c%d d%d" % (c, p1.dil_h))
    return out


def umma_supported(prep, p):
    """Mirror of esn_conv2d_umma's shape gate (csrc/esn_umma.cu) so routing costs no failed call."""
    cin = prep.cin
    if cin % 16 or prep.cout_pad > 256:
        return False
    if prep.kh * prep.kw > 9 or prep.stride not in (1, 2):
        return False
    if prep.transposed and (prep.stride != 2 or prep.dil_h != 1 or prep.dil_w != 1 or p.y.h != 2 * p.x.h):
        return False
    if not prep.transposed and prep.stride == 2 and ((p.x.h | p.x.w) & 1):
        return False
    if p.x.c_stride % 8 or p.y.c_stride % 8 or p.x.ptr % 16 or p.y.ptr % 16:
        return False
    if p.ep.residual.ptr and (p.ep.residual.c_stride % 8 or p.ep.residual.ptr % 16 or p.ep.residual.dtype != L.ESN_BF16):
        return False
    taps = prep.kh * prep.kw if not prep.transposed else 4
    kb = 64 if cin % 64 == 0 else (32 if cin % 32 == 0 else 16)            # K block of the kernel (esn_umma.cu)
    mt = 64 // kb
    while mt > 1 and mt * prep.cout_pad > 256:
        mt //= 2
    staged = prep.cout % 8 == 0 and (prep.cout <= 64 or prep.cout % 64 == 0)
    staging = mt * 128 * prep.cout * 2 if staged else 0     # the planner can go down to one staging buffer
    if taps * cin * prep.cout_pad * 2 + staging + 2 * mt * 128 * kb * 2 + 6144 > 226 * 1024:
        return False
    return True


def stem_conv3x3s2(x, w_direct, cconv, with_pool, out, scale, shift, alpha, act):
    """3x3/s2 stem conv (+ 2x2 max-pool concat) on the NCHW fp32 image -> NHWC `out`."""
    p = L.EsnStem()
    p.x, p.y = tdesc(x), tdesc(out)
    p.w, p.cconv, p.with_pool = w_direct.data_ptr(), cconv, int(with_pool)
    _epilogue(p.ep, scale, shift, alpha, act, None)
    flops = 2 * out.shape[0] * out.shape[2] * out.shape[3] * cconv * 27
    _call(L.lib.esn_stem_conv3x3s2, "esn_stem_conv3x3s2", (C.byref(p),), _nbytes(x) + _nbytes(out), flops)
    return out


def _pool_call(fn, name, x, out, scale, shift, alpha, act, residual=None, flags=0):
    p = L.EsnPool()
    p.x, p.y = tdesc(x), tdesc(out)
    _epilogue(p.ep, scale, shift, alpha, act, residual, flags)
    _call(fn, name, (C.byref(p),), _nbytes(x) + _nbytes(out) + _nbytes(residual))
    return out


def maxpool2x2(x, out, scale=None, shift=None, alpha=None, act=L.ACT_NONE):
    return _pool_call(L.lib.esn_maxpool2x2_affine_act, "esn_maxpool2x2_affine_act", x, out, scale, shift, alpha, act)


def avgpool3x3s2(x, out, scale=None, shift=None, alpha=None, act=L.ACT_NONE):
    return _pool_call(L.lib.esn_avgpool3x3s2_affine_act, "esn_avgpool3x3s2_affine_act", x, out, scale, shift, alpha, act)


def affine_act(x, scale, shift, alpha, act, out=None, residual=None, flags=0):
    if out is None:
        n, c, h, w = x.shape
        out = new_act(n, c, h, w, x.dtype, x.device)
    return _pool_call(L.lib.esn_affine_act, "esn_affine_act", x, out, scale, shift, alpha, act, residual, flags)


def concat_tail(x, buf, c0, scale=None, shift=None, alpha=None, act=L.ACT_NONE):
    """buf[:, c0:c0+c) = act(x*scale + shift) and zeros from there to the end of buf's padded pixel (esn_concat_tail):
    x = the c <= 4 injected channels (fp32 NHWC, pixel stride 4), buf = a concat buffer whose pixel stride covers its
    channel padding.  Whole-vector writes: the buffer needs no zero fill."""
    n, c, h, w = x.shape
    stride = buf.stride(3)
    y = buf.as_strided((n, c, h, w), buf.stride(), buf.storage_offset() + c0)
    p = L.EsnPool()
    p.x, p.y = tdesc(x), tdesc(y)
    _epilogue(p.ep, scale, shift, alpha, act, None, 0)
    tail = stride - c0
    _call(L.lib.esn_concat_tail, "esn_concat_tail", (C.byref(p), tail), _nbytes(x) + n * h * w * tail * buf.element_size())
    return y


def fglo_gate(x, w1, b1, w2, b2, out=None, residual=None):
    """CGNet FGlo: y = x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2) (+ residual)."""
    n, c, h, w = x.shape
    dx = tdesc(x)
    chunks = L.lib.esn_global_avgpool_chunks(C.byref(dx))
    sums = torch.empty((chunks, n, c), dtype=torch.float32, device=x.device)
    gate = torch.empty((n, c), dtype=torch.float32, device=x.device)
    _call(L.lib.esn_global_avgpool, "esn_global_avgpool", (C.byref(dx), C.c_void_p(sums.data_ptr())), _nbytes(x))
    p = L.EsnFGlo()
    p.sums, p.w1, p.b1, p.w2, p.b2, p.gate = (sums.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                              gate.data_ptr())
    p.n, p.channels, p.hidden, p.hw, p.chunks = n, c, w1.shape[0], h * w, chunks
    _call(L.lib.esn_fglo_gate, "esn_fglo_gate", (C.byref(p),))
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device)
    dy = tdesc(out)
    dr = tdesc(residual) if residual is not None else _NULL
    _call(L.lib.esn_scale_nc, "esn_scale_nc", (C.byref(dx), C.c_void_p(gate.data_ptr()), C.byref(dr), C.byref(dy)),
          _nbytes(x) + _nbytes(out) + _nbytes(residual))
    return out


def maxpool3x3s2_idx(x):
    """MaxPool2d(3, 2, 1, return_indices=True) on NHWC -> (pooled, int32 indices [N,Ho,Wo,C])."""
    n, c, h, w = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = new_act(n, c, ho, wo, x.dtype, x.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.int32, device=x.device)
    dx, dy = tdesc(x), tdesc(y)
    _call(L.lib.esn_maxpool3x3s2_idx, "esn_maxpool3x3s2_idx", (C.byref(dx), C.byref(dy), C.c_void_p(idx.data_ptr())),
          _nbytes(x) + _nbytes(y) + idx.numel() * 4)
    return y, idx


def max_unpool2x2(v, idx, ext=None, act=L.ACT_NONE, alpha=None):
    """y = act(MaxUnpool2d(2)(v, idx) + ext): deterministic gather (last writer in raster order wins)."""
    n, c, h, w = v.shape
    y = new_act(n, c, 2 * h, 2 * w, v.dtype, v.device)
    p = L.EsnUnpool()
    p.v, p.y, p.idx = tdesc(v), tdesc(y), idx.data_ptr()
    if ext is not None:
        p.ext = tdesc(ext)
    p.alpha = alpha.data_ptr() if alpha is not None else None
    p.act = act
    _call(L.lib.esn_max_unpool2x2, "esn_max_unpool2x2", (C.byref(p),), _nbytes(v) + idx.numel() * 4 + _nbytes(ext) + _nbytes(y))
    return y


def dab_dw_pair(x, prm, dilation, out=None):
    if out is None:
        n, c, h, w = x.shape
        out = new_act(n, c, h, w, x.dtype, x.device)
    p = L.EsnDabPair()
    p.x, p.y, p.prm, p.dilation = tdesc(x), tdesc(out), prm.data_ptr(), dilation
    _call(L.lib.esn_dab_dw_pair, "esn_dab_dw_pair", (C.byref(p),), _nbytes(x) + _nbytes(out), 0, "d%d" % dilation)
    return out


def adaptive_avgpool(x, size, dtype=None):
    n, c, h, w = x.shape
    y = new_act(n, c, size, size, dtype or x.dtype, x.device)
    dx, dy = tdesc(x), tdesc(y)
    _call(L.lib.esn_adaptive_avgpool, "esn_adaptive_avgpool", (C.byref(dx), C.byref(dy)), _nbytes(x) + _nbytes(y))
    return y


def bilinear(x, out_h, out_w, align_corners, out=None):
    n, c, h, w = x.shape
    if out is None:
        out = new_act(n, c, out_h, out_w, x.dtype, x.device)
    dx, dy = tdesc(x), tdesc(out)
    _call(L.lib.esn_bilinear_nhwc, "esn_bilinear_nhwc", (C.byref(dx), C.byref(dy), int(bool(align_corners))),
          _nbytes(x) + _nbytes(out))
    return out


def _head(fn, name, x, w, bias, classes, out_h, out_w, want_logits, want_mask, logits_dtype, align_corners=False):
    n = x.shape[0]
    p = L.EsnHead()
    p.align_corners = int(bool(align_corners))
    p.x = tdesc(x)
    p.w = w.data_ptr() if w is not None else None
    p.bias = bias.data_ptr() if bias is not None else None
    logits = mask = None
    if want_logits:
        logits = torch.empty((n, classes, out_h, out_w), dtype=logits_dtype, device=x.device)
        p.logits = tdesc(logits)
        p.logits.layout, p.logits.c_stride = L.ESN_NCHW, 0
    if want_mask:
        mask = torch.empty((n, out_h, out_w), dtype=torch.uint8, device=x.device)
        p.mask = mask.data_ptr()
    p.classes, p.out_h, p.out_w = classes, out_h, out_w
    alg = _nbytes(x) + (logits.numel() * logits.element_size() if logits is not None else 0) + \
        (mask.numel() if mask is not None else 0)
    _call(fn, name, (C.byref(p),), alg)
    return logits, mask


def head_convt2x2(x, w, bias, classes, want_logits=True, want_mask=False, logits_dtype=torch.float32):
    return _head(L.lib.esn_head_convt2x2, "esn_head_convt2x2", x, w, bias, classes, 2 * x.shape[2], 2 * x.shape[3],
                 want_logits, want_mask, logits_dtype)


def head_bilinear(x, classes, out_h, out_w, want_logits=True, want_mask=False, logits_dtype=torch.float32,
                  align_corners=False):
    return _head(L.lib.esn_head_bilinear, "esn_head_bilinear", x, None, None, classes, out_h, out_w,
                 want_logits, want_mask, logits_dtype, align_corners)


def pack_convt3x3s2_frags(weight, classes):
    """ConvTranspose2d(16, classes, 3, 2, 1, 1).weight (16, classes, 3, 3) -> the bf16 B fragments of esn_head_convt3x3s2_mask
    (include/esn.h): int32 [9 pairs][3 class tiles][32 lanes][2]."""
    w = weight.detach().float()
    assert w.shape[0] == 16 and w.shape[2:] == (3, 3) and classes <= 24
    pairs = ((0, 0, 0, 0), (0, 1, 0, 0), (0, 1, 0, 1), (1, 0, 0, 0), (1, 0, 1, 0), (1, 1, 0, 0), (1, 1, 0, 1), (1, 1, 1, 0), (1, 1, 1, 1))
    B = torch.zeros(9, 16, 24, dtype=torch.float32, device=w.device)
    for p, (a, b, dy, dx) in enumerate(pairs):
        B[p, :, :classes] = w[:, :classes, a + 1 - 2 * dy, b + 1 - 2 * dx]
    lane = torch.arange(32, device=w.device)
    g, t = lane // 4, lane % 4
    frag = torch.empty(9, 3, 32, 2, 2, dtype=torch.bfloat16, device=w.device)       # [..., register, (low, high)]
    for nt in range(3):
        for r in range(2):
            k0 = 2 * t + 8 * r
            frag[:, nt, :, r, 0] = B[:, k0, nt * 8 + g].to(torch.bfloat16)
            frag[:, nt, :, r, 1] = B[:, k0 + 1, nt * 8 + g].to(torch.bfloat16)
    return frag.contiguous().view(torch.int32).reshape(9, 3, 32, 2).contiguous()


def head_convt3x3s2_mask(x, wfrag, bias, classes):
    """uint8 argmax mask (N, 2h, 2w) of ConvTranspose2d(16, classes, 3, 2, 1, 1)(x) in one launch (tensor cores, the scores
    stay in registers); None when the entry point does not take the shape (the caller then runs conv + head)."""
    n, c, h, w = x.shape
    if not (x.dtype == torch.bfloat16 and is_nhwc(x) and c == 16 and w % 16 == 0 and classes <= 24):
        return None
    mask = torch.empty((n, 2 * h, 2 * w), dtype=torch.uint8, device=x.device)
    p = L.EsnHeadT3()
    p.x, p.wfrag, p.mask, p.classes = tdesc(x), wfrag.data_ptr(), mask.data_ptr(), classes
    p.bias = bias.data_ptr() if bias is not None else None
    _call(L.lib.esn_head_convt3x3s2_mask, "esn_head_convt3x3s2_mask", (C.byref(p),), _nbytes(x) + mask.numel(),
          2 * n * h * w * 9 * 16 * classes)
    return mask


def pack_convt2x2_frags(weight, classes):
    """ConvTranspose2d(16, classes, 2, 2).weight (16, classes, 2, 2) -> the bf16 B fragments of esn_head_convt2x2_mask
    (include/esn.h): int32 [2 hi/lo][4 positions][3 class tiles][32 lanes][2]; w = hi + lo keeps the fp32 weights to ~2^-17."""
    w = weight.detach().float()
    assert w.shape[0] == 16 and w.shape[2:] == (2, 2) and classes <= 24
    B = torch.zeros(4, 16, 24, dtype=torch.float32, device=w.device)
    for a in range(2):
        for b in range(2):
            B[a * 2 + b, :, :classes] = w[:, :classes, a, b]
    hi = B.to(torch.bfloat16)
    lo = (B - hi.float()).to(torch.bfloat16)
    lane = torch.arange(32, device=w.device)
    g, t = lane // 4, lane % 4
    frag = torch.empty(2, 4, 3, 32, 2, 2, dtype=torch.bfloat16, device=w.device)    # [..., register, (low, high)]
    for hl, part in enumerate((hi, lo)):
        for nt in range(3):
            for r in range(2):
                frag[hl, :, nt, :, r, 0] = part[:, 4 * t + 2 * r, nt * 8 + g]
                frag[hl, :, nt, :, r, 1] = part[:, 4 * t + 2 * r + 1, nt * 8 + g]
    return frag.contiguous().view(torch.int32).reshape(2, 4, 3, 32, 2).contiguous()


def head_convt2x2_mask(x, wfrag, bias, classes):
    """uint8 argmax mask (N, 2h, 2w) of ConvTranspose2d(16, classes, 2, 2)(x) in one tensor-core launch; None when the entry
    point does not take the shape (the caller then runs esn_head_convt2x2)."""
    n, c, h, w = x.shape
    if not (x.dtype == torch.bfloat16 and is_nhwc(x) and c == 16 and w % 16 == 0 and classes <= 24 and x.stride(3) % 4 == 0):
        return None
    mask = torch.empty((n, 2 * h, 2 * w), dtype=torch.uint8, device=x.device)
    p = L.EsnHeadT3()
    p.x, p.wfrag, p.mask, p.classes = tdesc(x), wfrag.data_ptr(), mask.data_ptr(), classes
    p.bias = bias.data_ptr() if bias is not None else None
    _call(L.lib.esn_head_convt2x2_mask, "esn_head_convt2x2_mask", (C.byref(p),), _nbytes(x) + mask.numel(),
          2 * n * h * w * 4 * 16 * classes)
    return mask


def weighted_ce(logits, target, weight=None, ignore_label=255, want_grad=False, sums=None, gnorm=None, gout=None,
                prob_out=None, keep_thresh=None):
    """Returns (sums[2] = [sum w*nll, sum w], dlogits or None); dlogits are scaled by gout/gnorm
    (device scalars) when given, unnormalised otherwise.  OHEM (loss.py:163-216): prob_out (N,H,W) fp32 receives the softmax
    probability of the labelled class (1 where ignored); keep_thresh (device scalar) makes pixels above it count as ignored."""
    require_cuda(logits, "weighted_ce")
    logits = logits.contiguous()
    target = target.contiguous()
    if sums is None:
        sums = torch.zeros(2, dtype=torch.float32, device=logits.device)
    p = L.EsnCE()
    p.gnorm = gnorm.data_ptr() if gnorm is not None else None
    p.gout = gout.data_ptr() if gout is not None else None
    p.logits = tdesc(logits)
    p.logits.layout, p.logits.c_stride = L.ESN_NCHW, 0
    p.target = target.data_ptr()
    p.weight = weight.data_ptr() if weight is not None else None
    p.sums = sums.data_ptr()
    g = None
    if want_grad:
        g = torch.empty_like(logits)
        p.dlogits = tdesc(g)
        p.dlogits.layout, p.dlogits.c_stride = L.ESN_NCHW, 0
    p.ignore_label = ignore_label
    p.prob_out = prob_out.data_ptr() if prob_out is not None else None
    p.keep_thresh = keep_thresh.data_ptr() if keep_thresh is not None else None
    _call(L.lib.esn_weighted_ce, "esn_weighted_ce", (C.byref(p),), logits.numel() * logits.element_size() * (2 if want_grad else 1))
    return sums, g


def bilinear_ce(scores, target, weight, ignore_label, out_h, out_w, sums=None, align_corners=False):
    """F.interpolate(scores, (out_h, out_w), bilinear, align_corners) -> weighted cross-entropy sums + the gradient of the
    scores, in one launch (esn_bilinear_ce; no full-resolution tensor is written).  Returns (sums[2] = [sum w*nll, sum w],
    dscores fp32 NHWC = d sums[0] / d scores), or None for what the entry point does not take (more than 32 classes, a
    target that is not int64 (N, out_h, out_w))."""
    require_cuda(scores, "bilinear_ce")
    n, c, h, w = scores.shape
    if not (is_nhwc(scores) and c <= 32 and out_h >= 1 and out_w >= 1 and tuple(target.shape) == (n, out_h, out_w)
            and target.dtype == torch.int64):
        return None
    target = target.contiguous()
    if sums is None:
        sums = torch.zeros(2, dtype=torch.float32, device=scores.device)
    ds = new_act(n, c, h, w, torch.float32, scores.device, c_alloc=(c + 3) // 4 * 4)
    p = L.EsnBilinearCE()
    p.scores, p.dscores = tdesc(scores), tdesc(ds)
    p.target = target.data_ptr()
    p.weight = weight.data_ptr() if weight is not None else None
    p.sums = sums.data_ptr()
    p.out_h, p.out_w, p.ignore_label, p.align_corners = out_h, out_w, ignore_label, int(bool(align_corners))
    _call(L.lib.esn_bilinear_ce, "esn_bilinear_ce", (C.byref(p),), _nbytes(scores) + target.numel() * 8 + _nbytes(ds))
    return sums, ds


def ohem_threshold(prob, min_kept, thresh, num_valid):
    """Device scalar max(thresh, min_kept-th smallest of prob), or +inf when min_kept exceeds *num_valid (loss.py:199-203);
    radix select on the device, no host synchronisation."""
    require_cuda(prob, "ohem_threshold")
    prob = prob.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=prob.device)
    ws = torch.zeros((int(L.lib.esn_ohem_workspace_bytes()) + 7) // 8, dtype=torch.int64, device=prob.device)
    _call(L.lib.esn_ohem_threshold, "esn_ohem_threshold",
          (C.c_void_p(prob.data_ptr()), prob.numel(), int(min_kept), float(thresh), C.c_void_p(num_valid.data_ptr()),
           C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr())), 3 * prob.numel() * 4)
    return out


def gate_bcast(g, x, b=None, out=None):
    """y = g * x + b with g (N,1,H,W), x (N,C,H,W), b (N,C,1,1) or None (LEDNet's APN close, LEDNet.py:279-281)."""
    n, c, h, w = x.shape
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device, c_alloc=x.stride(3), zero=x.stride(3) != c)
    dg, dx, dy = tdesc(g), tdesc(x), tdesc(out)
    db = tdesc(b) if b is not None else _NULL
    _call(L.lib.esn_gate_bcast, "esn_gate_bcast", (C.byref(dg), C.byref(dx), C.byref(db), C.byref(dy)),
          _nbytes(g) + _nbytes(x) + _nbytes(out))
    return out


def image_u8_to_f32(img, mean, reverse_channels=True, out=None):
    """Device half of the reference's dataset classes (dataset/cityscapes.py:74-78,164-170,208-214): uint8 HWC batch
    (N,H,W,3) in cv2's BGR order -> fp32 NCHW (N,3,H,W) = (img - mean)[..., ::-1] transposed; `mean` holds three values in
    the input's channel order (the pickle's fp32 BGR mean).  The result is what the models' stem kernels read."""
    require_cuda(img, "image_u8_to_f32")
    if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] != 3:
        raise TypeError("image_u8_to_f32: expected a uint8 (N,H,W,3) tensor, got %s %s" % (img.dtype, tuple(img.shape)))
    img = img.contiguous()
    n, h, w, _ = img.shape
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=img.device)
    elif out.shape != (n, 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != img.device:
        raise ValueError("image_u8_to_f32: out must be a contiguous fp32 (N,3,H,W) tensor on the input's device")
    if img.numel() == 0:          # empty batch: nothing to launch (an empty tensor has no device pointer to pass)
        return out
    m = (C.c_float * 3)(*[float(v) for v in mean])
    _call(L.lib.esn_image_u8hwc_to_f32nchw, "esn_image_u8hwc_to_f32nchw",
          (C.c_void_p(img.data_ptr()), C.c_void_p(out.data_ptr()), n, h, w, m, int(bool(reverse_channels))),
          img.numel() + out.numel() * 4)
    return out


def launch_count():
    return int(L.lib.esn_launch_count())


def launch_count_reset():
    L.lib.esn_launch_count_reset()
