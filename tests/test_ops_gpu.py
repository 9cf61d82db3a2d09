"""Kernel-level parity: each C-ABI op against the same op in plain torch fp32 (GPU tests)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from esn import ops as o
    return o


def _rand_conv(cin, cout, k, stride=1, pad=0, dil=1, groups=1, bias=True, transposed=False, out_pad=0, seed=0):
    torch.manual_seed(seed)
    if transposed:
        m = nn.ConvTranspose2d(cin, cout, k, stride=stride, padding=pad, output_padding=out_pad, bias=bias)
    else:
        m = nn.Conv2d(cin, cout, k, stride=stride, padding=pad, dilation=dil, groups=groups, bias=bias)
    return m.cuda().float()


def _nhwc(x, dtype, ops, c_alloc=None):
    n, c, h, w = x.shape
    y = ops.new_act(n, c, h, w, dtype, x.device, c_alloc=c_alloc)
    y.copy_(x)
    return y


CONV_CASES = [
    # cin, cout, k, stride, pad, dil, groups, transposed, out_pad, H, W
    (16, 16, (3, 1), 1, (1, 0), (1, 1), 1, False, 0, 20, 36),
    (64, 64, (1, 3), 1, (0, 2), (1, 2), 1, False, 0, 17, 40),
    (128, 128, (3, 1), 1, (16, 0), (16, 1), 1, False, 0, 24, 32),
    (16, 48, 3, 2, 1, 1, 1, False, 0, 32, 48),
    (64, 64, 3, 2, 1, 1, 1, False, 0, 16, 32),
    (35, 29, 3, 2, 1, 1, 1, False, 0, 16, 32),
    (128, 64, 3, 2, 1, 1, 1, True, 1, 9, 14),
    (64, 16, 3, 2, 1, 1, 1, True, 1, 12, 20),
    (32, 32, (3, 1), 1, (4, 0), (4, 1), 32, False, 0, 20, 24),   # depthwise dilated
    (64, 32, 3, 1, 1, 1, 1, False, 0, 16, 24),
    (32, 64, 1, 1, 0, 1, 1, False, 0, 16, 24),
    (259, 19, 1, 1, 0, 1, 1, False, 0, 8, 16),
    (32, 24, 1, 1, 0, 1, 4, False, 0, 16, 24),      # grouped 1x1 (ESPNetv2 reduce: 8 -> 6 per group)
    (96, 96, 1, 1, 0, 1, 4, False, 0, 12, 20),      # grouped 1x1 expand
    (24, 24, 3, 2, 6, 6, 24, False, 0, 20, 28),     # depthwise, stride 2, dilation 6
    (40, 19, 2, 2, 0, 1, 1, True, 0, 12, 20),       # ESPNet 2x2/s2 transposed convs
    (19, 19, 2, 2, 0, 1, 1, True, 0, 9, 14),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_direct_matches_torch(ops, case, dtype):
    from esn._lib import ACT_PRELU
    cin, cout, k, stride, pad, dil, groups, tr, op, H, W = case
    m = _rand_conv(cin, cout, k, stride, pad, dil, groups, True, tr, op)
    torch.manual_seed(1)
    x = torch.randn(2, cin, H, W, device="cuda")
    scale = torch.rand(cout, device="cuda") + 0.5
    shift = torch.randn(cout, device="cuda") * 0.1
    alpha = torch.rand(cout, device="cuda") * 0.4
    xa = _nhwc(x, dtype, ops)
    with torch.no_grad():
        ref = m(xa.float())
        res = torch.randn_like(ref)
        resa = _nhwc(res, dtype, ops)
        ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + resa.float()
        ref = torch.where(ref >= 0, ref, ref * alpha.view(1, -1, 1, 1))
    prep = ops.ConvPrep(m, scale, shift, ACT_PRELU, alpha)
    y = ops.conv2d(xa, prep, residual=resa, force_direct=True)
    assert y.shape == ref.shape
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    err = (y.float() - ref).abs().max() / ref.abs().max()
    assert err < tol, err


@pytest.mark.parametrize("cin,cout", [(256, 64), (512, 512), (256, 256)])
def test_grouped_1x1_on_tensor_cores(ops, cin, cout):
    """Grouped (g=4) 1x1 conv as per-group tcgen05 convs over channel slices, bf16, residual + PReLU epilogue."""
    from esn._lib import ACT_PRELU
    m = _rand_conv(cin, cout, 1, groups=4, bias=False)
    torch.manual_seed(2)
    x = torch.randn(2, cin, 24, 40, device="cuda")
    scale = torch.rand(cout, device="cuda") + 0.5
    shift = torch.randn(cout, device="cuda") * 0.1
    alpha = torch.rand(cout, device="cuda") * 0.4
    xa = _nhwc(x, torch.bfloat16, ops)
    with torch.no_grad():
        mb = nn.Conv2d(cin, cout, 1, groups=4, bias=False).cuda()
        mb.weight.copy_(m.weight.to(torch.bfloat16).float())
        ref = mb(xa.float())
        resa = _nhwc(torch.randn_like(ref), torch.bfloat16, ops)
        ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + resa.float()
        ref = torch.where(ref >= 0, ref, ref * alpha.view(1, -1, 1, 1))
    prep = ops.ConvPrep(m, scale, shift, ACT_PRELU, alpha)
    ops.launch_count_reset()
    y = ops.conv2d(xa, prep, residual=resa)
    assert ops.launch_count() == 4
    err = (y.float() - ref).abs().max() / ref.abs().max()
    assert err < 1e-2, err


def test_conv_direct_nchw_input_and_slice_output(ops):
    from esn._lib import ACT_RELU
    m = _rand_conv(3, 13, 3, 2, 1)
    x = torch.randn(2, 3, 32, 48, device="cuda") * 50
    prep = ops.ConvPrep(m, act=ACT_RELU)
    y = ops.new_act(2, 16, 16, 24, torch.float32, x.device)
    y.zero_()
    ops.conv2d(x, prep, out=y[:, :13])
    scale = torch.rand(3, device="cuda") + 0.5
    shift = torch.randn(3, device="cuda")
    ops.maxpool2x2(x, y[:, 13:], scale, shift, None, ACT_RELU)
    with torch.no_grad():
        ref = torch.cat([F.relu(m(x)), F.relu(F.max_pool2d(x, 2, 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))], 1)
    assert (y - ref).abs().max() / ref.abs().max() < 1e-5


def test_pools_affine_convert(ops):
    from esn._lib import ACT_NONE, ACT_PRELU
    x = torch.randn(2, 35, 16, 24, device="cuda")
    xa = _nhwc(x, torch.float32, ops, c_alloc=40)
    s, b, a = torch.rand(35, device="cuda") + 0.5, torch.randn(35, device="cuda"), torch.rand(35, device="cuda") * 0.3
    out = ops.affine_act(xa, s, b, a, ACT_PRELU)
    ref = x * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    ref = torch.where(ref >= 0, ref, ref * a.view(1, -1, 1, 1))
    assert torch.allclose(out, ref, atol=1e-6, rtol=1e-5)
    # in place
    ops.affine_act(xa, s, b, a, ACT_PRELU, out=xa)
    assert torch.allclose(xa, ref, atol=1e-6, rtol=1e-5)
    # avg pool 3x3 s2 p1 (count_include_pad) on NCHW input, odd sizes
    img = torch.randn(2, 3, 33, 47, device="cuda") * 30
    y = ops.new_act(2, 3, 17, 24, torch.float32, img.device, c_alloc=4)
    ops.avgpool3x3s2(img, y)
    assert torch.allclose(y, F.avg_pool2d(img, 3, 2, 1), atol=1e-4, rtol=1e-5)
    y2 = ops.new_act(2, 3, 9, 12, torch.float32, img.device, c_alloc=4)
    ops.avgpool3x3s2(y, y2)
    assert torch.allclose(y2, F.avg_pool2d(F.avg_pool2d(img, 3, 2, 1), 3, 2, 1), atol=1e-4, rtol=1e-5)
    # max pool on NHWC bf16
    xb = _nhwc(x, torch.bfloat16, ops, c_alloc=40)
    yb = ops.new_act(2, 35, 8, 12, torch.bfloat16, x.device, c_alloc=40)
    ops.maxpool2x2(xb, yb)
    assert torch.equal(yb.float(), F.max_pool2d(xb.float(), 2, 2))
    # layout round trip
    for dt in (torch.float32, torch.bfloat16):
        t = ops.as_act(x, dt)
        assert ops.is_nhwc(t) and t.dtype == dt
        back = ops.to_nchw(t, torch.float32)
        assert back.is_contiguous()
        assert torch.allclose(back, x.to(dt).float())


@pytest.mark.parametrize("cconv,pool,pad,act,H,W", [
    (13, 1, 1, "relu", 24, 44),      # ERFNet DownsamplerBlock(3,16): conv || MaxPool2d(2,2)
    (32, 0, 1, "prelu", 20, 70),     # DABNet Conv(3,32,3,2)+BNPReLU, ragged 16-pixel tiles
    (13, 2, 1, "prelu", 18, 36),     # ENet InitialBlock: conv || MaxPool2d(3,2,1) (-inf padding)
    (32, 0, 0, "relu", 33, 67),      # Fast-SCNN stem: padding 0, odd sizes
    (16, 0, 1, "none", 8, 16),
    (29, 1, 1, "prelu", 22, 300),    # 29 conv + 3 pooled channels = 32: the pool branch of the mma.sync kernel, several segments
    (29, 2, 1, "relu", 21, 139),     # ... with the 3x3 pool (-inf padding), odd sizes
    (24, 0, 1, "relu", 16, 260),     # 24 channels in a 32-wide tile
])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_stem_conv_matches_torch(ops, cconv, pool, pad, act, H, W, dt):
    """esn_stem_conv3x3s2 on the NCHW fp32 image (fp32 output: the exact FMA kernel; bf16 output: the mma.sync kernel with
    bf16 weights and the image split into bf16 hi + lo parts) against torch conv2d (+ pool concat) + affine + activation in
    fp32, borders included.  bf16: tight against the same arithmetic with bf16-rounded weights, and within the tensor-core
    tolerance of the other bf16 convs (1.5e-2 of the maximum) against the fp32 weights."""
    from esn._lib import ACT_NONE, ACT_RELU, ACT_PRELU
    torch.manual_seed(9)
    x = (torch.randint(0, 256, (2, 3, H, W), device="cuda").float() - 80.0).contiguous()
    w = torch.randn(cconv, 3, 3, 3, device="cuda") * 0.2
    ctot = cconv + (3 if pool else 0)
    scale = torch.rand(ctot, device="cuda") * 0.02 + 0.005
    shift = torch.randn(ctot, device="cuda") * 0.3
    alpha = torch.rand(ctot, device="cuda") * 0.4
    Ho, Wo = (H + 2 * pad - 3) // 2 + 1, (W + 2 * pad - 3) // 2 + 1
    out = ops.new_act(2, ctot, Ho, Wo, dt, x.device)
    wd = w.permute(2, 3, 1, 0).reshape(9, 3, cconv).contiguous()
    code = {"none": ACT_NONE, "relu": ACT_RELU, "prelu": ACT_PRELU}[act]
    ops.stem_conv3x3s2(x, wd, cconv, pool | (0 if pad else 256), out, scale, shift, alpha if act == "prelu" else None, code)
    def reference(wt):
        ref = F.conv2d(x, wt, None, 2, pad)
        if pool == 1:
            ref = torch.cat([ref, F.max_pool2d(x, 2, 2)], 1)
        elif pool == 2:
            ref = torch.cat([ref, F.max_pool2d(x, 3, 2, 1)], 1)
        ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        if act == "relu":
            ref = torch.relu(ref)
        elif act == "prelu":
            ref = torch.where(ref >= 0, ref, ref * alpha.view(1, -1, 1, 1))
        return ref
    got = out.float()
    ref = reference(w)
    assert got.shape == ref.shape
    if dt == torch.float32:
        err = (got - ref).abs()
        assert bool((err <= 1e-5 * ref.abs() + 1e-5 * ref.abs().max()).all()), err.max().item()
    else:
        # > 16 output channels: the mma.sync kernel (bf16 weights); otherwise the CUDA-core kernel with fp32 weights
        ref_b = reference(w.bfloat16().float()) if ctot > 16 else ref
        err = (got - ref_b).abs()
        assert bool((err <= 8e-3 * ref_b.abs() + 2e-3 * ref_b.abs().max()).all()), err.max().item()
        assert ((got - ref).abs().max() / ref.abs().max()).item() < 1.5e-2


@pytest.mark.parametrize("c,start,ctot", [(35, 29, 64), (3, 13, 16), (19, 2, 32), (40, 24, 64), (35, 0, 40)])
def test_maxpool_into_unaligned_channel_slice(ops, c, start, ctot):
    """MaxPool2d(2,2) + BN + PReLU written into a channel slice that does not start on a 16-byte boundary
    (DABNet.py:104-108: 35 pooled channels at channel 29 of the 64-channel block output): aligned-vector kernel with a
    channel shift; the neighbouring channels of the buffer must stay untouched."""
    from esn._lib import ACT_PRELU
    torch.manual_seed(5)
    x = torch.randn(2, c, 12, 20, device="cuda")
    xb = _nhwc(x, torch.bfloat16, ops, c_alloc=(c + 7) // 8 * 8)
    out = ops.new_act(2, ctot, 6, 10, torch.bfloat16, x.device)
    out.fill_(7.0)
    scale = torch.rand(c, device="cuda") + 0.5
    shift = torch.randn(c, device="cuda")
    alpha = torch.rand(c, device="cuda") * 0.4
    ops.maxpool2x2(xb, out[:, start:start + c], scale, shift, alpha, ACT_PRELU)
    ref = F.max_pool2d(xb.float(), 2, 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    ref = torch.where(ref >= 0, ref, ref * alpha.view(1, -1, 1, 1))
    got = out.float()
    assert (got[:, start:start + c] - ref).abs().max() <= 1e-2 * ref.abs().max()
    assert torch.equal(got[:, :start], torch.full_like(got[:, :start], 7.0))
    assert torch.equal(got[:, start + c:], torch.full_like(got[:, start + c:], 7.0))


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_dab_pair_matches_torch(ops, dt):
    """esn_dab_dw_pair (row kernel: stage-1 rows in shared memory) against torch: both DABNet module shapes, dilation 1,
    dilation wider than the image, odd widths, the narrowest channel chunk (4) and a 3-chunk split (C/2 = 48)."""
    torch.manual_seed(3)
    from model.DABNet import DABModule
    from oracle import fixture
    for c, d, hh, ww in ((64, 2, 20, 28), (128, 16, 20, 28), (64, 1, 5, 7), (8, 4, 9, 33), (128, 8, 3, 130), (96, 4, 17, 70),
                         (64, 40, 6, 12)):
        m = DABModule(c, d=d)
        m.load_state_dict(fixture.randomize_state_dict(m.state_dict(), 5))
        m = m.cuda().eval()
        x = torch.randn(2, c // 2, hh, ww, device="cuda").to(dt).float()
        with torch.no_grad():
            def cbr(conv, t, pad, dil):
                y = F.conv2d(t, conv.conv.weight, None, 1, pad, dil, c // 2)
                y = F.batch_norm(y, conv.bn_prelu.bn.running_mean, conv.bn_prelu.bn.running_var,
                                 conv.bn_prelu.bn.weight, conv.bn_prelu.bn.bias, False, 0.0, 1e-3)
                return F.prelu(y, conv.bn_prelu.acti.weight)
            b1 = cbr(m.dconv1x3, cbr(m.dconv3x1, x, (1, 0), 1), (0, 1), 1)
            b2 = cbr(m.ddconv1x3, cbr(m.ddconv3x1, x, (d, 0), (d, 1)), (0, d), (1, d))
            s = b1 + b2
            s = F.batch_norm(s, m.bn_relu_2.bn.running_mean, m.bn_relu_2.bn.running_var, m.bn_relu_2.bn.weight,
                             m.bn_relu_2.bn.bias, False, 0.0, 1e-3)
            ref = F.prelu(s, m.bn_relu_2.acti.weight)
        y = ops.dab_dw_pair(ops.as_act(x, dt), m.prep(x.device), d)
        assert y.dtype == dt
        tol = 1e-5 if dt == torch.float32 else 1e-2      # bf16: output rounding only (fp32 arithmetic inside)
        assert (y.float() - ref).abs().max() / ref.abs().max() < tol, (c, d, hh, ww)


def test_heads_and_ce(ops):
    from oracle import fixture, loss as oloss
    torch.manual_seed(0)
    # convT 2x2 head
    m = nn.ConvTranspose2d(16, 19, 2, stride=2).cuda()
    x = torch.randn(2, 16, 12, 20, device="cuda")
    w = torch.zeros(2, 2, 16, 32, device="cuda")
    w[:, :, :, :19] = m.weight.detach().permute(2, 3, 0, 1)
    logits, mask = ops.head_convt2x2(ops.as_act(x, torch.float32), w.contiguous(), m.bias.detach().contiguous(), 19, True, True)
    with torch.no_grad():
        ref = m(x)
    assert torch.allclose(logits, ref, atol=1e-5, rtol=1e-5)
    assert (mask.long() == ref.argmax(1)).float().mean() > 0.999
    # bilinear head (align_corners=False), x8
    s = torch.randn(2, 19, 8, 16, device="cuda")
    sa = ops.new_act(2, 19, 8, 16, torch.float32, s.device, c_alloc=32)
    sa.copy_(s)
    logits, mask = ops.head_bilinear(sa, 19, 64, 128, True, True)
    ref = F.interpolate(s, (64, 128), mode="bilinear", align_corners=False)
    assert torch.allclose(logits, ref, atol=1e-5, rtol=1e-5)
    assert (mask.long() == ref.argmax(1)).float().mean() > 0.999
    # weighted CE fwd + grad vs the CPU oracle
    lg = torch.randn(2, 19, 16, 32) * 3
    lab = fixture.make_labels(2, 16, 32, 19, seed=5)
    wt = torch.tensor(fixture.CLASS_WEIGHTS)
    sums, g = ops.weighted_ce(lg.cuda(), lab.cuda(), wt.cuda(), 255, want_grad=True)
    l, swl, sw = oloss.weighted_ce(lg.double(), lab, wt.double())
    assert abs(sums[0].item() / sums[1].item() - l.item()) < 1e-5 * abs(l.item())
    gref = oloss.weighted_ce_grad(lg.double(), lab, wt.double())
    assert ((g.cpu().double() / sums[1].item()) - gref).abs().max() < 1e-6


def test_enet_pool_unpool_matches_cpu_reference(ops):
    """MaxPool2d(3,2,1,return_indices) and the gather-form MaxUnpool2d against torch on the CPU
    (NCHW, sequential raster order: last writer wins -- SURVEY.md H5), with duplicate indices present."""
    from esn._lib import ACT_RELU
    torch.manual_seed(4)
    x = torch.randn(2, 16, 12, 20).round()        # rounding creates ties -> exercises first-max / duplicate paths
    ref_y, ref_i = F.max_pool2d(x, 3, 2, 1, return_indices=True)
    xa = _nhwc(x.cuda(), torch.float32, ops)
    y, idx = ops.maxpool3x3s2_idx(xa)
    assert torch.equal(y.cpu(), ref_y)
    assert torch.equal(idx.permute(0, 3, 1, 2).cpu().long(), ref_i)
    v = torch.randn(2, 16, 6, 10)
    ext = torch.randn(2, 16, 12, 20)
    ref = F.relu(F.max_unpool2d(v.contiguous(), ref_i, 2) + ext)
    out = ops.max_unpool2x2(_nhwc(v.cuda(), torch.float32, ops), idx, ext=_nhwc(ext.cuda(), torch.float32, ops), act=ACT_RELU)
    assert torch.equal(out.cpu(), ref)
    dup = (ref_i.flatten(2).sort(dim=2).values.diff(dim=2) == 0).float().mean().item()
    assert dup > 0.05, "test input must contain duplicate indices"
    # bf16: the 16-byte (8-channel) pool / unpool paths, values exactly representable
    yb, idxb = ops.maxpool3x3s2_idx(_nhwc(x.cuda(), torch.bfloat16, ops))
    assert torch.equal(yb.float().cpu(), ref_y) and torch.equal(idxb.permute(0, 3, 1, 2).cpu().long(), ref_i)
    vb, eb = v.round(), ext.round()
    refb = F.relu(F.max_unpool2d(vb.contiguous(), ref_i, 2) + eb)
    outb = ops.max_unpool2x2(_nhwc(vb.cuda(), torch.bfloat16, ops), idx, ext=_nhwc(eb.cuda(), torch.bfloat16, ops), act=ACT_RELU)
    assert outb.dtype == torch.bfloat16 and torch.equal(outb.float().cpu(), refb)


DW_STRIP_CASES = [
    # C, k, dil, H, W, act, residual mode (None / "post" / "pre_act")
    (64, (3, 3), (1, 1), 33, 50, "prelu", None),        # CGNet F_loc
    (64, (3, 3), (4, 4), 33, 50, "prelu", None),        # CGNet F_sur
    (128, (3, 3), (2, 2), 40, 36, "relu", "post"),
    (384, (3, 3), (1, 1), 16, 24, "relu", None),        # Fast-SCNN bottleneck expansion
    (32, (3, 1), (16, 1), 40, 24, "prelu", None),       # DABNet branch, dilation 16
    (32, (1, 3), (1, 16), 24, 40, "prelu", "pre_act"),
    (64, (3, 1), (1, 1), 70, 9, "none", "post"),        # chains longer than one segment
    (8, (3, 3), (3, 3), 5, 7, "none", None),            # dilation ~ image size
    (16, (3, 3), (1, 1), 130, 140, "relu", None),
]


@pytest.mark.parametrize("case", DW_STRIP_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_depthwise_strip_matches_torch(ops, case, dtype):
    """The register-strip depthwise kernel (csrc/esn_dw_strip.cu) behind esn_conv2d_direct: every tap shape, dilated column
    chains, segment halos, ragged widths, both residual orders -- against torch's depthwise conv2d in fp32, and against the
    gather kernel it replaces (same C-ABI call with ESN_DISABLE_DW_STRIP=1)."""
    import os
    from esn._lib import ACT_NONE, ACT_PRELU, ACT_RELU, EP_ACT_BEFORE_RESIDUAL
    C, k, dil, H, W, act, rmode = case
    pad = ((k[0] // 2) * dil[0], (k[1] // 2) * dil[1])
    m = _rand_conv(C, C, k, 1, pad, dil, C, False)
    torch.manual_seed(2)
    x = torch.randn(3, C, H, W, device="cuda")
    scale = torch.rand(C, device="cuda") + 0.5
    shift = torch.randn(C, device="cuda") * 0.1
    alpha = torch.rand(C, device="cuda") * 0.4
    xa = _nhwc(x, dtype, ops)
    f = {"none": lambda t: t, "relu": torch.relu, "prelu": lambda t: torch.where(t >= 0, t, t * alpha.view(1, -1, 1, 1))}[act]
    with torch.no_grad():
        ref = m(xa.float()) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        resa = None
        if rmode:
            resa = _nhwc(torch.randn_like(ref), dtype, ops)
            ref = (f(ref) if rmode == "pre_act" else ref) + resa.float()
        ref = f(ref)
    prep = ops.ConvPrep(m, scale, shift, {"none": ACT_NONE, "relu": ACT_RELU, "prelu": ACT_PRELU}[act], alpha if act == "prelu" else None)
    if rmode == "pre_act":
        prep.ep_flags = EP_ACT_BEFORE_RESIDUAL
    y = ops.conv2d(xa, prep, residual=resa, force_direct=True)
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    err = (y.float() - ref).abs().max() / ref.abs().max()
    assert err < tol, err
    os.environ["ESN_DISABLE_DW_STRIP"] = "1"
    try:
        y_old = ops.conv2d(xa, prep, residual=resa, force_direct=True)
    finally:
        del os.environ["ESN_DISABLE_DW_STRIP"]
    assert (y.float() - y_old.float()).abs().max() / ref.abs().max() < tol


@pytest.mark.parametrize("classes,bias,shape", [(19, True, (2, 24, 48)), (5, False, (1, 7, 16)), (24, True, (3, 9, 32)),
                                                (19, True, (2, 64, 4096))])
def test_convt2x2_argmax_head_on_tensor_cores(classes, bias, shape):
    """ERFNet's / ESNet's head in one launch (esn_head_convt2x2_mask: transposed conv 2x2 / s2 + argmax on mma.sync, fp32
    weights as hi + lo bf16 fragments) against torch's conv_transpose2d with the FP32 weights on the same bf16 activations,
    and against the CUDA-core head kernel it replaces."""
    import torch.nn.functional as F
    from esn import ops
    torch.manual_seed(12)
    n, h, w = shape
    x = ops.new_act(n, 16, h, w, torch.bfloat16, "cuda").normal_()
    wt = (torch.randn(16, classes, 2, 2, device="cuda") * 0.3)
    b = torch.randn(classes, device="cuda") if bias else None
    frags = ops.pack_convt2x2_frags(wt, classes)
    mask = ops.head_convt2x2_mask(x, frags, b, classes)
    assert mask is not None and mask.shape == (n, 2 * h, 2 * w) and mask.dtype == torch.uint8
    ref = F.conv_transpose2d(x.double(), wt.double(), None if b is None else b.double(), stride=2)
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref.abs().amax(dim=1).clamp_min(1.0)
    want = ref.argmax(1)
    assert clear.float().mean().item() > 0.99
    assert torch.equal(mask.long()[clear], want[clear])
    assert (mask.long() == want).float().mean().item() > 0.9995
    packed = torch.zeros((2, 2, 16, 32), dtype=torch.float32, device="cuda")
    packed[:, :, :, :classes] = wt.permute(2, 3, 0, 1)
    bb = b if b is not None else torch.zeros(classes, device="cuda")
    _, old = ops.head_convt2x2(x, packed.contiguous(), bb, classes, False, True, torch.bfloat16)
    assert torch.equal(mask[clear], old[clear])


@pytest.mark.parametrize("classes,bias,shape", [(19, False, (2, 24, 48)), (5, True, (1, 7, 16)), (24, True, (3, 9, 32))])
def test_convt3x3s2_argmax_head_on_tensor_cores(classes, bias, shape):
    """ENet's head in one launch (esn_head_convt3x3s2_mask: transposed conv 3x3 / s2 + argmax, scores in registers) against
    torch's conv_transpose2d on the same bf16 operands: identical masks wherever the top-2 margin exceeds fp32 rounding."""
    import torch.nn.functional as F
    from esn import ops
    torch.manual_seed(11)
    n, h, w = shape
    x = ops.new_act(n, 16, h, w, torch.bfloat16, "cuda").normal_()
    wt = (torch.randn(16, classes, 3, 3, device="cuda") * 0.3)
    b = torch.randn(classes, device="cuda") if bias else None
    frags = ops.pack_convt3x3s2_frags(wt, classes)
    mask = ops.head_convt3x3s2_mask(x, frags, b, classes)
    assert mask is not None and mask.shape == (n, 2 * h, 2 * w) and mask.dtype == torch.uint8
    ref = F.conv_transpose2d(x.float(), wt.bfloat16().float(), b, stride=2, padding=1, output_padding=1)
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref.abs().amax(dim=1).clamp_min(1.0)
    want = ref.argmax(1)
    assert clear.float().mean().item() > 0.99
    assert torch.equal(mask.long()[clear], want[clear])
    assert (mask.long() == want).float().mean().item() > 0.999


@pytest.mark.parametrize("relu,dilation,shape", [(True, 1, (2, 19, 70)), (False, 1, (1, 8, 32)), (False, 2, (2, 13, 45))])
def test_enet_bottleneck4_in_one_launch(relu, dilation, shape):
    """ENet's RegularBottleneck(16) (four internal channels) as one launch against the reference arithmetic in fp32 on the same
    bf16 input, and against the three-launch path of the same module."""
    import model.ENet as E
    torch.manual_seed(12)
    n, h, w = shape
    blk = E.RegularBottleneck(16, padding=dilation, dilation=dilation, dropout_prob=0.1, relu=relu).cuda().eval()
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
            if isinstance(m, nn.PReLU):
                m.weight.fill_(0.2)
    x = torch.randn(n, 16, h, w, device="cuda")
    xb = x.bfloat16().float()
    with torch.no_grad():
        act = blk.out_prelu
        e = act(blk.ext_conv1[1](F.conv2d(xb, blk.ext_conv1[0].weight)))
        e = act(blk.ext_conv2[1](F.conv2d(e, blk.ext_conv2[0].weight, padding=dilation, dilation=dilation)))
        e = act(blk.ext_conv3[1](F.conv2d(e, blk.ext_conv3[0].weight)))
        ref = act(xb + e)
        outs = {}
        for fused in (True, False):
            E.FUSED_BNECK4 = fused
            try:
                from esn import ops
                xa = ops.as_act(x, torch.bfloat16)
                n0 = ops.L.lib.esn_launch_count()
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y = blk(xa)
                outs[fused] = (y.float(), ops.L.lib.esn_launch_count() - n0)
            finally:
                E.FUSED_BNECK4 = True
    assert outs[True][1] == 1 and outs[False][1] == 3
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(outs[True][0], ref) < 6e-3, rel(outs[True][0], ref)          # bf16 output rounding only
    assert rel(outs[False][0], ref) < 2e-2
    assert rel(outs[True][0], outs[False][0]) < 2e-2


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_concat_tail_writes_injected_channels_and_zero_padding(ops, dt):
    """esn_concat_tail: the three input-injection channels through the concat's BNPReLU slice, zeros up to the end of the
    padded pixel, nothing before the first injected channel (DABNet.py:166,171,176 on a 64 / 192-channel padded buffer)."""
    from esn._lib import ACT_PRELU
    torch.manual_seed(2)
    for c0, alloc, hh, ww in ((32, 64, 9, 13), (128, 192, 5, 16), (256, 320, 3, 7)):
        x = ops.new_act(2, 3, hh, ww, torch.float32, "cuda", c_alloc=4)
        x.copy_(torch.randn(2, 3, hh, ww, device="cuda"))
        s, b, a = torch.rand(3, device="cuda") + 0.5, torch.randn(3, device="cuda"), torch.rand(3, device="cuda") * 0.3
        buf = ops.new_act(2, c0 + 3, hh, ww, dt, "cuda", c_alloc=alloc)
        whole = ops.widen(buf, alloc)
        whole.fill_(7.0)
        ops.concat_tail(x, buf, c0, s, b, a, ACT_PRELU)
        torch.cuda.synchronize()
        v = x.float() * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        ref = torch.where(v >= 0, v, v * a.view(1, -1, 1, 1)).to(dt)
        assert torch.equal(whole[:, c0:c0 + 3], ref)
        assert float(whole[:, c0 + 3:].abs().max()) == 0.0
        assert bool((whole[:, :c0] == 7.0).all())
