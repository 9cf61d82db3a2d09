"""esn_bilinear_ce -- bilinear up-sampling of the class scores + weighted cross-entropy + the gradient of the scores in one
launch (the fused close of a training iteration: DABNet.py:181 + utils/losses/loss.py:15-32, train.py:351-353) -- against
torch's interpolate / cross_entropy / autograd in fp64, and DABNet.fused_loss against criterion(model(x), y)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import spec_state_dict
from oracle import fixture

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("n,c,h,w,s,weighted", [(2, 19, 8, 16, 8, True), (1, 19, 5, 18, 8, True), (2, 11, 7, 33, 4, False),
                                               (1, 19, 4, 4, 16, True), (1, 3, 9, 21, 2, True), (1, 32, 4, 17, 8, True),
                                               (2, 19, 64, 128, 8, True), (1, 19, 5, 7, (33, 50), True), (1, 19, 9, 9, (9, 9), True),
                                               (1, 19, 1, 1, (6, 5), True), (1, 19, 12, 20, (7, 31), False)])
def test_bilinear_ce_matches_torch_autograd(dt, align, n, c, h, w, s, weighted):
    from esn import ops
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + c * 10 + h)
    x = ops.new_act(n, c, h, w, dt, "cuda", c_alloc=32)
    x.copy_(torch.randn(n, c, h, w, device="cuda", generator=g) * 3)
    H, W = (s * h, s * w) if isinstance(s, int) else s        # integer scales, or any output size (incl. down-sampling)
    tgt = torch.randint(0, c, (n, H, W), device="cuda", generator=g)
    tgt[torch.rand(n, H, W, device="cuda", generator=g) < 0.15] = 255          # ignored pixels
    tgt[0, :3, :] = 255                                                          # a whole border strip ignored
    wt = (torch.rand(c, device="cuda", generator=g) + 0.5) if weighted else None
    res = ops.bilinear_ce(x, tgt, wt, 255, H, W, align_corners=align)
    assert res is not None
    sums, ds = res
    xr = x.double().detach().contiguous().requires_grad_(True)
    logits = F.interpolate(xr, size=(H, W), mode="bilinear", align_corners=align)
    loss = F.cross_entropy(logits, tgt, None if wt is None else wt.double(), ignore_index=255, reduction="sum")
    (gr,) = torch.autograd.grad(loss, xr)
    valid = tgt != 255
    wsum = (wt[tgt[valid]].double().sum() if wt is not None else valid.double().sum()).item()
    assert abs(sums[0].item() - loss.item()) <= 2e-5 * abs(loss.item()), (sums[0].item(), loss.item())
    assert abs(sums[1].item() - wsum) <= 1e-5 * wsum
    assert ds.dtype == torch.float32 and ds.shape == (n, c, h, w)
    assert _rel(ds, gr) < 2e-5, _rel(ds, gr)
    pad = ds.as_strided((n, ds.stride(3), h, w), ds.stride(), ds.storage_offset())[:, c:]
    assert pad.numel() == 0 or float(pad.abs().max()) == 0.0                     # the padded lanes are written as zeros


def test_bilinear_ce_declines_what_it_does_not_take():
    from esn import ops
    x = ops.new_act(1, 19, 8, 16, torch.float32, "cuda").normal_()
    assert ops.bilinear_ce(x, torch.zeros(1, 64, 128, dtype=torch.int32, device="cuda"), None, 255, 64, 128) is None
    assert ops.bilinear_ce(x, torch.zeros(1, 64, 96, dtype=torch.int64, device="cuda"), None, 255, 64, 128) is None


@pytest.mark.parametrize("net", ["DABNet", "CGNet", "FastSCNN"])
@pytest.mark.parametrize("dt", [None, torch.bfloat16])
def test_fused_loss_equals_the_two_module_form(spec, dt, net):
    """loss and parameter gradients of model.fused_loss (one esn_bilinear_ce launch) against criterion(model(x), y)."""
    import contextlib
    from builders.model_builder import build_model
    from utils.losses.loss import CrossEntropyLoss2d
    from esn import ops
    x = fixture.make_input(2, 64, 128).cuda()
    lab = fixture.make_labels(2, 64, 128, 19).cuda()
    crit = CrossEntropyLoss2d(weight=torch.tensor(fixture.CLASS_WEIGHTS), ignore_label=255).cuda()
    ctx = (lambda: torch.autocast("cuda", dtype=dt)) if dt is not None else contextlib.nullcontext

    def build():
        m = build_model(net, 19)
        m.load_state_dict(spec_state_dict(spec, net))
        m = m.cuda().train()
        for mod in m.modules():                       # dropout off: every forward would draw another mask
            if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
                mod.p = 0.0
        return m

    def grads(m):                                     # parameters that take no part in the forward have no gradient
        return [torch.zeros_like(p) if p.grad is None else p.grad.detach().clone() for p in m.parameters()]

    def run(fused):
        m = build()
        ops.PROFILE = []
        try:
            with ctx():
                loss = m.fused_loss(x, lab, crit) if fused else crit(m(x), lab)
            loss.backward()
            torch.cuda.synchronize()
            names = [r["kernel"] for r in ops.PROFILE]
        finally:
            ops.PROFILE = None
        return loss.item(), grads(m), names

    l0, g0, n0 = run(False)
    l1, g1, n1 = run(True)
    assert "esn_bilinear_ce" in n1 and "esn_weighted_ce" not in n1 and "esn_head_bilinear" not in n1 and "esn_bilinear_bwd" not in n1
    assert "esn_weighted_ce" in n0 and "esn_bilinear_ce" not in n0
    assert len(n1) == len(n0) - 2                     # head, CE forward, CE backward, bilinear backward -> fused + scale
    assert abs(l1 - l0) <= (1e-5 if dt is None else 2e-3) * abs(l0), (l0, l1)
    live = [i for i, b in enumerate(g0) if float(b.abs().max()) > 0]
    cat = lambda gs: torch.cat([gs[i].flatten().double() for i in live])
    if dt is None:
        # fp32: only the summation order differs.  Fast-SCNN's pyramid has a 1x1 level whose train-mode BatchNorm sees two values
        # at batch 2: its input gradient is analytically zero and numerically rounding noise times 1 / sigma, in
        # either form -- so there the parameters are compared as a whole and by their 90 % quantile
        rels = sorted(_rel(g1[i], g0[i]) for i in live)
        assert _rel(cat(g1), cat(g0)) < 1e-3, _rel(cat(g1), cat(g0))
        assert (rels[int(0.9 * len(rels))] if net == "FastSCNN" else rels[-1]) < 2e-3, rels[-5:]
        return
    # bf16: both forms round the same d scores to bf16 from values that differ in the last fp32 bits (atomics order of the
    # weight sum), so single roundings flip.  Some parameter gradients are sums of cancelling terms (train-mode BatchNorm) and
    # move by 10-20 % under such flips -- in the two-module form as much as in the fused one.  So: all gradients together
    # agree closely, and parameter by parameter the fused form is as close to the FP32 gradients as the two-module form is.
    assert _rel(cat(g1), cat(g0)) < 2e-2, _rel(cat(g1), cat(g0))
    m32 = build()
    crit(m32(x), lab).backward()                      # the two-module form in fp32
    g32 = grads(m32)
    for i in live:
        e1, e0 = _rel(g1[i], g32[i]), _rel(g0[i], g32[i])
        assert e1 < max(5e-2, 2.0 * e0), (i, e1, e0)


def test_dabnet_fused_loss_falls_back(spec):
    """Eval mode, another criterion or an input that is not a multiple of 8 take criterion(model(x), y)."""
    from builders.model_builder import build_model
    from utils.losses.loss import CrossEntropyLoss2d, FocalLoss2d
    x = fixture.make_input(1, 64, 128).cuda()
    lab = fixture.make_labels(1, 64, 128, 19).cuda()
    m = build_model("DABNet", 19)
    m.load_state_dict(spec_state_dict(spec, "DABNet"))
    m = m.cuda().train()
    crit = CrossEntropyLoss2d(ignore_label=255).cuda()
    a = m.fused_loss(x, lab, FocalLoss2d(ignore_index=255).cuda())
    assert a.requires_grad and a.item() == a.item()
    m.eval()
    with torch.no_grad():
        b = m.fused_loss(x, lab, crit)
        c = crit(m(x), lab)
    assert abs(b.item() - c.item()) <= 1e-6 * abs(c.item())
