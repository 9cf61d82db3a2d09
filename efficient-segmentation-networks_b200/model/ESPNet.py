"""ESPNet (ESPNet-C encoder + light-weight decoder) on B200 kernels -- drop-in for the reference's model/ESPNet.py.

Same class names, constructor signatures and attribute names (identical ``state_dict`` keys) as
/root/reference/model/ESPNet.py:13-385.

Layout.  The ESP block concatenates five dilated-conv branches of n1, n, n, n, n channels (16+4*12 at level 2,
28+4*25 at level 3): widths that no vector load, TMA box or UMMA K block likes.  Inside the network every such
tensor is therefore kept in a *physical* NHWC layout whose five slices are padded to S = 16 / 32 channels
(64 -> 80, 128 -> 160 channels) with exact-zero tails; a `_Map` records where each logical channel lives, and the
weights / BatchNorm / PReLU vectors of every consumer are scattered through it once, at prep time (zero weight
columns over the tails).  The branch convs (K = 16 / 32) then run on the tcgen05 kernel, each writing one
aligned slice of the concat buffer and taking the previous slice as its residual operand (the hierarchical
sums); "BN(input + combine) -> PReLU" is one pointwise pass (ESN_EP_RESIDUAL_FIRST).  Blocks called on their
own take and return ordinary logical NCHW/NHWC tensors.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (kept for parity with the reference module's namespace)

from esn import ops
from esn._lib import ACT_NONE, ACT_PRELU
from esn.prep import PrepMixin

__all__ = ["ESPNet_Encoder", "ESPNet"]

_RES_FIRST = 2   # ESN_EP_RESIDUAL_FIRST


def _no_train(mod):
    if mod.training:
        raise NotImplementedError(
            "%s: training-mode kernels are not built yet for this model; call .eval(). "
            "There is no eager-PyTorch fallback." % type(mod).__name__)


def _tc(c):
    """Channel count the tensor-core conv accepts as Cin: 16/32/64 or a multiple of 64."""
    for v in (16, 32, 64):
        if c <= v:
            return v
    return (c + 63) // 64 * 64


class _Map:
    """Where each logical channel of a tensor lives in its physical NHWC buffer of `width` channels."""

    def __init__(self, pos, width):
        self.pos, self.width = list(pos), int(width)
        self.key = (tuple(self.pos), self.width)
        assert len(set(self.pos)) == len(self.pos) and max(self.pos) < self.width

    @staticmethod
    def identity(c, width=None):
        return _Map(range(c), width or c)

    @staticmethod
    def cat(parts, width=None):
        """parts: [(map, offset)] -- maps placed at the given physical channel offsets."""
        pos, end = [], 0
        for m, off in parts:
            pos += [off + p for p in m.pos]
            end = max(end, off + m.width)
        return _Map(pos, width or end)

    @staticmethod
    def branches(n, n1):
        s = 8 if n1 <= 8 else (16 if n1 <= 16 else (n1 + 31) // 32 * 32)
        pos = list(range(n1))
        for j in range(1, 5):
            pos += list(range(j * s, j * s + n))
        return _Map(pos, 5 * s), s

    def vec(self, v, fill, device):
        out = torch.full((self.width,), float(fill), dtype=torch.float32, device=device)
        out[torch.tensor(self.pos, device=device)] = v.detach().float().to(device)
        return out.contiguous()

    def cin(self, w):
        """(Cout, C_logical, kh, kw) -> (Cout, width, kh, kw), zero columns over the padding."""
        out = torch.zeros((w.shape[0], self.width) + tuple(w.shape[2:]), dtype=torch.float32, device=w.device)
        out[:, torch.tensor(self.pos, device=w.device)] = w.detach().float()
        return out

    @property
    def logical(self):
        return len(self.pos)

    @property
    def is_identity(self):
        return self.pos == list(range(len(self.pos)))

    def runs(self):
        """Contiguous (logical_start, physical_start, length) runs."""
        out, i = [], 0
        while i < len(self.pos):
            j = i
            while j + 1 < len(self.pos) and self.pos[j + 1] == self.pos[j] + 1:
                j += 1
            out.append((i, self.pos[i], j - i + 1))
            i = j + 1
        return out


def _to_physical(x, m):
    if m.is_identity and x.stride(3) >= m.width:
        return ops.widen(x, m.width) if m.width != x.shape[1] else x
    n, _, h, w = x.shape
    y = ops.new_act(n, m.width, h, w, x.dtype, x.device, zero=True)
    for lo, po, ln in m.runs():
        ops.affine_act(x[:, lo:lo + ln], None, None, None, ACT_NONE, out=y[:, po:po + ln])
    return y


def _to_logical(y, m):
    if m.is_identity:
        return y[:, :m.logical] if y.shape[1] != m.logical else y
    n, _, h, w = y.shape
    x = ops.new_act(n, m.logical, h, w, y.dtype, y.device)
    for lo, po, ln in m.runs():
        ops.affine_act(y[:, po:po + ln], None, None, None, ACT_NONE, out=x[:, lo:lo + ln])
    return x


def _bn_vecs(bn, act, m, device):
    s, b = ops.bn_affine(bn, device)
    a = None if act is None else m.vec(act.weight, 0.0, device)
    return m.vec(s, 1.0, device), m.vec(b, 0.0, device), a


def _pieces(vecs, spans):
    """Slices (offset, count) of a (scale, shift, alpha) triple as separate, aligned tensors."""
    return [tuple(t[o:o + c].clone() for t in vecs) for o, c in spans]


def _conv(w, m_in, cout_pad=None, stride=1, pad=0, dil=1, transposed=False, **epi):
    """ConvPrep of a weight (Cout, Cin_logical, k, k) [(Cin, Cout, k, k) if transposed] reading a physical input."""
    w = w.detach().float()
    if transposed:
        w = w.permute(1, 0, 2, 3)
    w = m_in.cin(w)
    if transposed:
        w = w.permute(1, 0, 2, 3).contiguous()
    return ops.ConvPrep.from_weight(w, stride=stride, padding=(pad, pad), dilation=(dil, dil), transposed=transposed,
                                    cout_pad=cout_pad, **epi)


class _MapPrep(PrepMixin):
    """Per-module cache of preps keyed by the physical layout of the input."""

    def _build_prep(self, device):
        return {}

    def prep_for(self, m_in, device):
        cache = self.prep(device)
        if m_in.key not in cache:
            with torch.no_grad():
                cache[m_in.key] = self._build_for(m_in, device)
        return cache[m_in.key]


class CBR(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False)
        self.bn = nn.BatchNorm2d(nOut, eps=1e-03)
        self.act = nn.PReLU(nOut)

    def _build_prep(self, device):
        s, b = ops.bn_affine(self.bn, device)
        return ops.ConvPrep(self.conv, s, b, ACT_PRELU, self.act.weight, device=device)

    def forward(self, input, out=None):
        _no_train(self)
        c = self.conv
        if (input.shape[1] == 3 and input.dtype == torch.float32 and input.is_contiguous() and not ops.is_nhwc(input)
                and c.kernel_size == (3, 3) and c.stride == (2, 2) and c.out_channels % 4 == 0 and c.out_channels <= 32
                and not ((input.shape[2] | input.shape[3]) & 1)):
            ops.require_cuda(input, "CBR")
            prep = self.prep(input.device)
            n, _, h, w = input.shape
            if out is None:
                out = ops.new_act(n, prep.cout, h // 2, w // 2, ops.compute_dtype(input), input.device)
            return ops.stem_conv3x3s2(input, prep.w_direct, prep.cout, 0, out, prep.scale, prep.shift, prep.alpha, ACT_PRELU)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device), out=out)


class BR(PrepMixin, nn.Module):
    def __init__(self, nOut):
        super().__init__()
        self.bn = nn.BatchNorm2d(nOut, eps=1e-03)
        self.act = nn.PReLU(nOut)

    def _build_prep(self, device):
        return _bn_vecs(self.bn, self.act, _Map.identity(self.bn.num_features), device)

    def forward(self, input, out=None):
        _no_train(self)
        x = ops.as_act(input)
        s, b, a = self.prep(x.device)
        return ops.affine_act(x, s, b, a, ACT_PRELU, out=out)


class CB(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False)
        self.bn = nn.BatchNorm2d(nOut, eps=1e-03)

    def _build_prep(self, device):
        s, b = ops.bn_affine(self.bn, device)
        return ops.ConvPrep(self.conv, s, b, ACT_NONE, device=device)

    def forward(self, input):
        _no_train(self)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device))


class C(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1):
        super().__init__()
        padding = int((kSize - 1) / 2)
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False)

    def _build_prep(self, device):
        return ops.ConvPrep(self.conv, device=device)

    def forward(self, input):
        _no_train(self)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device))


class CDilated(PrepMixin, nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1, d=1):
        super().__init__()
        padding = int((kSize - 1) / 2) * d
        self.conv = nn.Conv2d(nIn, nOut, (kSize, kSize), stride=stride, padding=(padding, padding), bias=False, dilation=d)

    def _build_prep(self, device):
        return ops.ConvPrep(self.conv, device=device)

    def forward(self, input):
        _no_train(self)
        x = ops.as_act(input)
        return ops.conv2d(x, self.prep(x.device))


class _FiveBranch(_MapPrep, nn.Module):
    """Reduce -> five dilated 3x3 branches with hierarchical sums -> concat, shared by DownSamplerB and the ESP block."""

    def _make(self, nIn, nOut, ksize, stride):
        n = int(nOut / 5)
        n1 = nOut - 4 * n
        self.c1 = C(nIn, n, ksize, stride)
        self.d1 = CDilated(n, n1, 3, 1, 1)
        self.d2 = CDilated(n, n, 3, 1, 2)
        self.d4 = CDilated(n, n, 3, 1, 4)
        self.d8 = CDilated(n, n, 3, 1, 8)
        self.d16 = CDilated(n, n, 3, 1, 16)

    def out_map(self):
        return _Map.branches(self.d2.conv.out_channels, self.d1.conv.out_channels)[0]

    def _branch_preps(self, m_in, device):
        n, n1 = self.d2.conv.out_channels, self.d1.conv.out_channels
        m_out, s = _Map.branches(n, n1)
        n_p = _tc(n) if n > 8 else n                 # the reduced tensor, padded for the tensor-core branches
        c1 = self.c1.conv
        reduce = _conv(c1.weight.to(device), m_in, cout_pad=n_p if n_p != n else None, stride=c1.stride[0], pad=c1.padding[0])
        m_red = _Map.identity(n, n_p)
        br = []
        for name, d in (("d1", 1), ("d2", 2), ("d4", 4), ("d8", 8), ("d16", 16)):
            w = getattr(self, name).conv.weight.to(device)
            br.append(_conv(w, m_red, cout_pad=s if s != w.shape[0] else None, pad=d, dil=d))
        return reduce, br, m_out, s, n, n_p

    def _branches(self, x, reduce, br, s, n, n_p):
        nb, _, h, w = x.shape
        ho, wo = reduce.out_hw(h, w)
        tail = n_p != reduce.cout                     # direct path writes only the logical channels
        o1 = ops.new_act(nb, reduce.cout, ho, wo, x.dtype, x.device, c_alloc=n_p, zero=tail)
        ops.conv2d(x, reduce, out=o1)
        o1 = ops.widen(o1, n_p)
        cat = ops.new_act(nb, 5 * s, ho, wo, x.dtype, x.device)
        for j, p in enumerate(br):
            ops.conv2d(o1, p, out=cat[:, j * s:(j + 1) * s], residual=cat[:, (j - 1) * s:j * s] if j >= 2 else None)
        return cat


class DownSamplerB(_FiveBranch):
    def __init__(self, nIn, nOut):
        super().__init__()
        self._make(nIn, nOut, 3, 2)
        self.bn = nn.BatchNorm2d(nOut, eps=1e-3)
        self.act = nn.PReLU(nOut)

    def _build_for(self, m_in, device):
        reduce, br, m_out, s, n, n_p = self._branch_preps(m_in, device)
        return reduce, br, s, n, n_p, _bn_vecs(self.bn, self.act, m_out, device)

    def run(self, x, m_in):
        reduce, br, s, n, n_p, (sc, sh, al) = self.prep_for(m_in, x.device)
        cat = self._branches(x, reduce, br, s, n, n_p)
        return ops.affine_act(cat, sc, sh, al, ACT_PRELU, out=cat)

    def forward(self, input):
        _no_train(self)
        x = ops.as_act(input)
        return _to_logical(self.run(x, _Map.identity(x.shape[1])), self.out_map())


class DilatedParllelResidualBlockB(_FiveBranch):
    def __init__(self, nIn, nOut, add=True):
        super().__init__()
        self._make(nIn, nOut, 1, 1)
        self.bn = BR(nOut)
        self.add = add

    def _build_for(self, m_in, device):
        reduce, br, m_out, s, n, n_p = self._branch_preps(m_in, device)
        if self.add and m_in.key != m_out.key:
            raise ValueError("ESP block: the residual input must be in the block's own concat layout")
        return reduce, br, s, n, n_p, _bn_vecs(self.bn.bn, self.bn.act, m_out, device)

    def run(self, x, m_in):
        reduce, br, s, n, n_p, (sc, sh, al) = self.prep_for(m_in, x.device)
        cat = self._branches(x, reduce, br, s, n, n_p)
        return ops.affine_act(cat, sc, sh, al, ACT_PRELU, out=cat, residual=x if self.add else None,
                              flags=_RES_FIRST if self.add else 0)

    def forward(self, input):
        _no_train(self)
        x = ops.as_act(input)
        m_out = self.out_map()
        m_in = m_out if self.add else _Map.identity(x.shape[1])
        return _to_logical(self.run(_to_physical(x, m_in), m_in), m_out)


class InputProjectionA(nn.Module):
    def __init__(self, samplingTimes):
        super().__init__()
        self.pool = nn.ModuleList()
        for i in range(0, samplingTimes):
            self.pool.append(nn.AvgPool2d(3, stride=2, padding=1))

    def forward(self, input, start=0):
        """`start`: number of leading pools already applied to `input` (the encoder shares the pyramid)."""
        ops.require_cuda(input, "InputProjectionA")
        x = input
        for _ in list(self.pool)[start:]:
            n, c, h, w = x.shape
            y = ops.new_act(n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, torch.float32, x.device, c_alloc=4)
            x = ops.avgpool3x3s2(x if (ops.is_nhwc(x) or x.is_contiguous()) else x.contiguous(), y)
        return x


class ESPNet_Encoder(PrepMixin, nn.Module):
    def __init__(self, classes=19, p=5, q=3):
        super().__init__()
        self.level1 = CBR(3, 16, 3, 2)
        self.sample1 = InputProjectionA(1)
        self.sample2 = InputProjectionA(2)
        self.b1 = BR(16 + 3)
        self.level2_0 = DownSamplerB(16 + 3, 64)
        self.level2 = nn.ModuleList()
        for i in range(0, p):
            self.level2.append(DilatedParllelResidualBlockB(64, 64))
        self.b2 = BR(128 + 3)
        self.level3_0 = DownSamplerB(128 + 3, 128)
        self.level3 = nn.ModuleList()
        for i in range(0, q):
            self.level3.append(DilatedParllelResidualBlockB(128, 128))
        self.b3 = BR(256)
        self.classifier = C(256, classes, 1, 1)

    # physical layouts of the three concat stages
    def maps(self):
        c0 = self.b1.bn.num_features
        m0 = _Map.identity(c0, _tc(c0))                                   # [level1 | image] + zero tail
        m2 = self.level2_0.out_map()
        m1 = _Map.cat([(m2, 0), (m2, m2.width), (_Map.identity(3), 2 * m2.width)], _tc(2 * m2.width + 3))
        m3 = self.level3_0.out_map()
        mc = _Map.cat([(m3, 0), (m3, m3.width)], _tc(2 * m3.width))
        return m0, m1, mc

    def _build_prep(self, device):
        m0, m1, mc = self.maps()
        cls = _conv(self.classifier.conv.weight.to(device), mc)
        c1 = self.level1.conv.out_channels
        w2, w3 = self.level2_0.out_map().width, self.level3_0.out_map().width
        return dict(b1=_pieces(_bn_vecs(self.b1.bn, self.b1.act, m0, device), [(0, c1), (c1, 3)]),
                    b2=_pieces(_bn_vecs(self.b2.bn, self.b2.act, m1, device), [(0, w2), (w2, w2), (2 * w2, 3)]),
                    b3=_pieces(_bn_vecs(self.b3.bn, self.b3.act, mc, device), [(0, w3), (w3, w3)]), cls=cls,
                    clsp=_conv(self.classifier.conv.weight.to(device), mc, cout_pad=32))

    def stages(self, input, cat0=None):
        """The three concat stages in physical layout: (cat0, cat1, cat2) + their maps.
        `cat0`: optional pre-allocated zeroed buffer view for the level-1 concat."""
        ops.require_cuda(input, "ESPNet_Encoder")
        _no_train(self)
        if input.dtype != torch.float32 or not input.is_contiguous():
            input = input.float().contiguous()
        n, _, h, w = input.shape
        if (h | w) & 7:
            raise ValueError("ESPNet: input height and width must be multiples of 8, got %dx%d" % (h, w))
        dt, dev = ops.compute_dtype(input), input.device
        P = self.prep(dev)
        m0, m1, mc = self.maps()
        c1 = self.level1.conv.out_channels
        if cat0 is None:
            cat0 = ops.new_act(n, m0.width, h // 2, w // 2, dt, dev, zero=True)
        self.level1(input, out=cat0[:, :c1])
        ops.affine_act(cat0[:, :c1], *P["b1"][0], ACT_PRELU, out=cat0[:, :c1])
        inp1 = self.sample1(input)
        ops.affine_act(inp1, *P["b1"][1], ACT_PRELU, out=cat0[:, c1:c1 + 3])
        inp2 = self.sample2(inp1, start=1)
        o1_0 = self.level2_0.run(cat0, m0)
        m2 = self.level2_0.out_map()
        o1 = o1_0
        for layer in self.level2:
            o1 = layer.run(o1, m2)
        w2 = m2.width
        cat1 = ops.new_act(n, m1.width, h // 4, w // 4, dt, dev, zero=True)
        for (src, off, cw), vecs in zip(((o1, 0, w2), (o1_0, w2, w2), (inp2, 2 * w2, 3)), P["b2"]):
            ops.affine_act(src, *vecs, ACT_PRELU, out=cat1[:, off:off + cw])
        o2_0 = self.level3_0.run(cat1, m1)
        m3 = self.level3_0.out_map()
        o2 = o2_0
        for layer in self.level3:
            o2 = layer.run(o2, m3)
        w3 = m3.width
        cat2 = ops.new_act(n, mc.width, h // 8, w // 8, dt, dev, zero=mc.width != 2 * w3)
        for (src, off), vecs in zip(((o2_0, 0), (o2, w3)), P["b3"]):
            ops.affine_act(src, *vecs, ACT_PRELU, out=cat2[:, off:off + w3])
        return cat0, cat1, cat2

    def _scores(self, input):
        _, _, cat2 = self.stages(input)
        P = self.prep(cat2.device)
        n, _, h, w = cat2.shape
        classes = P["cls"].cout
        sc = ops.new_act(n, classes, h, w, cat2.dtype, cat2.device, c_alloc=32)
        if cat2.dtype == torch.bfloat16:
            ops.conv2d(cat2, P["clsp"], out=ops.widen(sc, 32))
        else:
            ops.conv2d(cat2, P["cls"], out=sc)
        return sc

    def forward(self, input):
        sc = self._scores(input)
        ldt = torch.bfloat16 if sc.dtype == torch.bfloat16 else torch.float32
        return ops.head_bilinear(sc, sc.shape[1], input.shape[2], input.shape[3], True, False, ldt)[0]


class ESPNet(PrepMixin, nn.Module):
    def __init__(self, classes=19, p=2, q=3, encoderFile=None):
        super().__init__()
        self.encoder = ESPNet_Encoder(classes, p, q)
        if encoderFile is not None:
            self.encoder.load_state_dict(torch.load(encoderFile, map_location="cpu"))
            print('Encoder loaded!')
        self.en_modules = []
        for i, m in enumerate(self.encoder.children()):
            self.en_modules.append(m)
        self.level3_C = C(128 + 3, classes, 1, 1)
        self.br = nn.BatchNorm2d(classes, eps=1e-03)
        self.conv = CBR(19 + classes, classes, 3, 1)
        self.up_l3 = nn.Sequential(nn.ConvTranspose2d(classes, classes, 2, stride=2, padding=0, output_padding=0, bias=False))
        self.combine_l2_l3 = nn.Sequential(BR(2 * classes), DilatedParllelResidualBlockB(2 * classes, classes, add=False))
        self.up_l2 = nn.Sequential(nn.ConvTranspose2d(classes, classes, 2, stride=2, padding=0, output_padding=0, bias=False),
                                   BR(classes))
        self.classifier = nn.ConvTranspose2d(classes, classes, 2, stride=2, padding=0, output_padding=0, bias=False)

    def _build_prep(self, device):
        enc = self.encoder
        m0, m1, mc = enc.maps()
        classes = self.classifier.out_channels
        c0 = m0.logical
        s_br, b_br = ops.bn_affine(self.br, device)
        comb = self.combine_l2_l3[0]
        s_c, b_c = ops.bn_affine(comb.bn, device)
        a_c = comb.act.weight.detach().float().to(device)
        up2 = self.up_l2[1]
        s_u, b_u = ops.bn_affine(up2.bn, device)
        cw = enc.classifier.conv.weight.to(device)
        l3w = self.level3_C.conv.weight.to(device)
        # [comb (classes) | zero tail to 32 | level-1 concat (c0) | zero tail]: the 3x3 "conv" reads both halves
        m_e = _Map.cat([(_Map.identity(classes), 0), (_Map.identity(c0), 32)], 32 + m0.width)
        m_esp = self.combine_l2_l3[1].out_map()
        s_v, b_v = ops.bn_affine(self.conv.bn, device)
        P = dict(
            m_e=m_e, m_esp=m_esp, classes=classes,
            cls=_conv(cw, mc, scale=s_br, shift=b_br), clsp=_conv(cw, mc, cout_pad=32, scale=s_br, shift=b_br),
            up_l3=_conv(self.up_l3[0].weight.to(device), _Map.identity(classes), stride=2, transposed=True,
                        scale=s_c[classes:].clone(), shift=b_c[classes:].clone(), act=ACT_PRELU, alpha=a_c[classes:].clone()),
            l3c=_conv(l3w, m1, scale=s_c[:classes].clone(), shift=b_c[:classes].clone(), act=ACT_PRELU, alpha=a_c[:classes].clone()),
            l3cp=_conv(l3w, m1, cout_pad=32, scale=s_c[:classes].clone(), shift=b_c[:classes].clone(), act=ACT_PRELU,
                       alpha=a_c[:classes].clone()),
            up_l2=_conv(self.up_l2[0].weight.to(device), m_esp, stride=2, transposed=True, scale=s_u, shift=b_u, act=ACT_PRELU,
                        alpha=up2.act.weight.detach().float().to(device)),
            conv=_conv(self.conv.conv.weight.to(device), m_e, pad=1, scale=s_v, shift=b_v, act=ACT_PRELU,
                       alpha=self.conv.act.weight.detach().float().to(device)),
            convp=_conv(self.conv.conv.weight.to(device), m_e, cout_pad=32, pad=1, scale=s_v, shift=b_v, act=ACT_PRELU,
                        alpha=self.conv.act.weight.detach().float().to(device)),
        )
        w = self.classifier.weight.detach().to(device=device, dtype=torch.float32)      # (Cin, classes, 2, 2)
        if classes == 19:
            packed = torch.zeros((2, 2, 20, 32), dtype=torch.float32, device=device)
            packed[:, :, :w.shape[0], :classes] = w.permute(2, 3, 0, 1)
            P["head_w"], P["head_b"] = packed.contiguous(), torch.zeros(32, dtype=torch.float32, device=device)
        else:
            P["cls_t"] = ops.ConvPrep(self.classifier, device=device)
        return P

    def features(self, input):
        ops.require_cuda(input, "ESPNet")
        _no_train(self)
        enc = self.encoder
        n, _, h, w = input.shape
        dt, dev = ops.compute_dtype(input), input.device
        P = self.prep(dev)
        classes = P["classes"]
        if classes > 32:
            raise NotImplementedError("ESPNet: the decoder buffers are laid out for at most 32 classes, got %d" % classes)
        tc = dt == torch.bfloat16
        m_e = P["m_e"]
        cat_e = ops.new_act(n, m_e.width, h // 2, w // 2, dt, dev, zero=True)
        cat0, cat1, cat2 = enc.stages(input, cat0=cat_e[:, 32:])
        # RUM 1: classifier (+br BN) -> up_l3 (+ second half of the combine BR) ; level3_C (+ first half)
        s = ops.new_act(n, classes, h // 8, w // 8, dt, dev, c_alloc=32)
        if tc:
            ops.conv2d(cat2, P["clsp"], out=ops.widen(s, 32))
        else:
            ops.conv2d(cat2, P["cls"], out=s)
        cat_d = ops.new_act(n, 2 * classes, h // 4, w // 4, dt, dev, c_alloc=64)
        if tc:      # writes channels 0..31 (zeros past `classes`); up_l3 then overwrites classes..2*classes-1
            ops.conv2d(cat1, P["l3cp"], out=cat_d[:, :32])
        else:
            ops.conv2d(cat1, P["l3c"], out=cat_d[:, :classes])
        ops.conv2d(s, P["up_l3"], out=cat_d[:, classes:])
        comb = self.combine_l2_l3[1].run(cat_d, _Map.identity(2 * classes))
        ops.conv2d(comb, P["up_l2"], out=cat_e[:, :classes])
        y = ops.new_act(n, classes, h // 2, w // 2, dt, dev, c_alloc=32, zero=not tc)
        if tc:
            ops.conv2d(cat_e, P["convp"], out=ops.widen(y, 32))
        else:
            ops.conv2d(cat_e, P["conv"], out=y)
        return (ops.widen(y, 20) if classes == 19 else y), P

    def _head(self, y, P, want_logits, want_mask):
        """The closing ConvTranspose2d(classes, classes, 2, stride 2) (ESPNet.py:353) -> NCHW logits and / or the argmax mask:
        the fused 2x2 transposed-conv head for the 19-class layout, otherwise (any other class count, ESPNet.py:350 takes it
        as a constructor argument) the transposed conv through esn_conv2d_direct followed by the same-size head kernel."""
        ldt = torch.bfloat16 if y.dtype == torch.bfloat16 else torch.float32
        classes = P["classes"]
        if classes == 19:
            return ops.head_convt2x2(y, P["head_w"], P["head_b"], classes, want_logits, want_mask, ldt)
        n, _, h, w = y.shape
        s = ops.new_act(n, classes, 2 * h, 2 * w, y.dtype, y.device, c_alloc=(classes + 7) // 8 * 8)
        ops.conv2d(y, P["cls_t"], out=s)
        return ops.head_bilinear(s, classes, 2 * h, 2 * w, want_logits, want_mask, ldt)

    def forward(self, input):
        if self.training:
            # batch-statistics BatchNorm and the recorded backward (esn/train.py); one autograd node for the net
            from esn import train as T
            from model._espnet_train import espnet_train_forward
            return T.run_network(self, lambda inp: espnet_train_forward(self, inp), input)
        y, P = self.features(input)
        return self._head(y, P, True, False)[0]

    @torch.no_grad()
    def predict_mask(self, input, with_logits=False):
        y, P = self.features(input)
        logits, mask = self._head(y, P, with_logits, True)
        return (logits, mask) if with_logits else mask
