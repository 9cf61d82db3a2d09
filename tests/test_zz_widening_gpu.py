"""Model-level parity (GPU) for the nets of SURVEY 8f-1 / 8f-2 (ESNet, EDANet, LEDNet, ContextNet): same checks and tolerances as
tests/test_models_gpu.py -- fp32 logits vs the unmodified reference's golden (1e-3), argmax >= 99.9 %, bf16 vs the
CPU oracle (5e-2 or torch's own bf16-autocast error on the same graph).

First device run: round-1 driver run on B200 (23 of 25 passed under a non-strict xfail blanket, removed in round 2).
"""
import pytest
import torch

import test_models_gpu as T
from conftest import spec_state_dict
from oracle import fixture, nets

pytestmark = pytest.mark.gpu

NETS = ["ESNet", "ContextNet", "EDANet", "LEDNet"]


@pytest.mark.parametrize("name", NETS)
def test_fp32_matches_reference_golden(name, spec, golden):
    T.test_fp32_matches_reference_golden(name, spec, golden)


@pytest.mark.parametrize("name", ["ESNet", "ContextNet", "EDANet"])
def test_bf16_matches_oracle(name, spec):
    T.test_bf16_matches_oracle(name, spec)


def _lednet_features(m, x):
    y = m.initial_block(x)
    for layer in m.layers:
        y = layer(y)
    return y


def _lednet_oracle_features(sd, x):
    p = nets.SD(sd, "", x.dtype)
    y = nets.erf_downsampler(p.sub("initial_block"), x)
    for i, d in enumerate(nets.LED_LAYERS):
        q = p.sub("layers.%d" % i)
        y = nets.erf_downsampler(q, y) if d is None else nets.led_ssnbt(q, y, d)
    return y


@pytest.mark.parametrize("size", [(2, 128, 256), (1, 512, 1024)])
def test_lednet_bf16_matches_oracle_stage_by_stage(size, spec):
    """LEDNet in bf16.  With random-init weights its attention pyramid (a 128 -> 1 channel reduction with heavy cancellation
    that gates every class score, LEDNet.py:189-283) is ill-conditioned: on the ORACLE, an 8e-3 perturbation of the 128-channel
    features comes out of the module as 4e-2 ... 9e-2, and torch's own bf16 autocast of the reference graph is 3e-2 (128x256)
    to 1.1e-1 (512x1024) off the fp32 logits (profiles/r02_diag_lednet.log).  So north_star's 5e-2 is checked where it is
    meaningful -- (1) the features entering the module, accumulated over all 15 blocks, (2) the module itself on identical
    inputs -- and the end-to-end logits are bounded by what the fp32 oracle module itself makes of our features (the
    conditioning term, no arithmetic of ours in it) plus the module's own error, and by torch-autocast's error."""
    n, h, w = size
    m = T._model("LEDNet", spec)
    sd = spec_state_dict(spec, "LEDNet")
    x = fixture.make_input(n, h, w)
    with torch.no_grad():
        ref_feat = _lednet_oracle_features(sd, x)
        ref_scores = nets.led_apn(nets.SD(sd, "apn."), ref_feat)
        ref = nets.forward("LEDNet", sd, x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            feat = _lednet_features(m, x.cuda())
            scores = m.apn(feat)
            y = m(x.cuda())
            mask = m.predict_mask(x.cuda())
        feat_c = feat.float().cpu()
        cond_scores = nets.led_apn(nets.SD(sd, "apn."), feat_c)        # exact module on OUR features
        sd_gpu = {k: v.cuda() for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y_ac = nets.forward("LEDNet", sd_gpu, x.cuda())
    r_feat = T._rel(feat_c, ref_feat)
    r_mod = T._rel(scores.float().cpu()[:, :19], cond_scores)
    r_cond = T._rel(cond_scores, ref_scores)
    r_all = T._rel(y.float().cpu(), ref)
    r_ac = T._rel(y_ac.float().cpu(), ref)
    print("LEDNet bf16 %s: features %.3e  module on identical input %.3e  conditioning (oracle module on our features) %.3e  "
          "logits %.3e  torch-autocast logits %.3e" % (size, r_feat, r_mod, r_cond, r_all, r_ac))
    assert y.dtype == torch.bfloat16
    assert r_feat < T.BF16_LOGIT_TOL, r_feat
    assert r_mod < T.BF16_LOGIT_TOL, r_mod
    assert r_all < max(T.BF16_LOGIT_TOL, 1.5 * r_ac, r_cond + T.BF16_LOGIT_TOL), (r_all, r_ac, r_cond)
    ref_mask = torch.from_numpy(nets.argmax_mask(ref))
    safe = T._margin_mask(ref, max(T.BF16_LOGIT_TOL, r_cond))
    aware = (mask.cpu() == ref_mask)[safe].float().mean().item()
    ac_aware = (torch.from_numpy(nets.argmax_mask(y_ac.float())) == ref_mask)[safe].float().mean().item()
    print("LEDNet bf16 argmax: raw %.4f  margin-aware %.4f (torch-autocast %.4f)" %
          ((mask.cpu() == ref_mask).float().mean().item(), aware, ac_aware))
    assert aware >= min(T.ARGMAX_MIN, ac_aware - 5e-3), (aware, ac_aware)


def test_esnet_blocks_are_drop_in(spec):
    from model.ESNet import FCU, PFCU
    sd = spec_state_dict(spec, "ESNet")
    torch.manual_seed(0)
    for cls, args, idx, ref_fn in ((FCU, (16, 3, 0.03, 1), 0, lambda p, x: nets.es_fcu(p, x, 3, 1)),
                                   (FCU, (64, 5, 0.03, 1), 4, lambda p, x: nets.es_fcu(p, x, 5, 1)),
                                   (PFCU, (128,), 7, nets.es_pfcu)):
        pre = "layers.%d." % idx
        blk = cls(*args)
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        blk = blk.cuda().eval()
        x = torch.randn(2, args[0], 24, 40)
        ref = ref_fn(nets.SD(sd, pre), x)
        assert T._rel(blk(x.cuda()).float().cpu(), ref) < 1e-4, cls.__name__
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yb = blk(x.cuda())
        assert yb.dtype == torch.bfloat16 and T._rel(yb.float().cpu(), ref) < T.BF16_LOGIT_TOL, cls.__name__


def test_edanet_blocks_are_drop_in(spec):
    """EDAModule / EDANetBlock called on their own (NCHW in, cat([new, input]) out)."""
    from model.EDANet import EDAModule, EDANetBlock
    sd = spec_state_dict(spec, "EDANet")
    torch.manual_seed(0)
    pre = "layers.2.residual_dense_layers.3."
    blk = EDAModule(180, 2)
    blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    x = torch.randn(2, 180, 24, 40)
    ref = nets.eda_module(nets.SD(sd, pre), x, 2)
    y = blk.cuda().eval()(x.cuda())
    assert y.shape == ref.shape and T._rel(y.float().cpu(), ref) < 1e-4
    pre = "layers.2."
    blk = EDANetBlock(60, 5, [1, 1, 1, 2, 2], 40)
    blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    x = torch.randn(2, 60, 24, 40)
    ref = x
    for i, d in enumerate(nets.EDA_BLOCKS[2]):
        ref = nets.eda_module(nets.SD(sd, pre + "residual_dense_layers.%d." % i), ref, d)
    blk = blk.cuda().eval()
    y = blk(x.cuda())
    assert y.shape == ref.shape and T._rel(y.float().cpu(), ref) < 1e-4
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yb = blk(x.cuda())
    assert yb.dtype == torch.bfloat16 and T._rel(yb.float().cpu(), ref) < T.BF16_LOGIT_TOL


def test_lednet_blocks_are_drop_in(spec):
    """SS_nbt_module_paper (split / two factorized branches / merge + input / shuffle) and APNModule on their own."""
    from model.LEDNet import APNModule, SS_nbt_module_paper
    sd = spec_state_dict(spec, "LEDNet")
    torch.manual_seed(0)
    for idx, chann, d in ((1, 32, 1), (5, 64, 1), (9, 128, 5)):
        pre = "layers.%d." % idx
        blk = SS_nbt_module_paper(chann, 0.03, d)
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        blk = blk.cuda().eval()
        x = torch.randn(2, chann, 24, 40)
        ref = nets.led_ssnbt(nets.SD(sd, pre), x, d)
        assert T._rel(blk(x.cuda()).float().cpu(), ref) < 1e-4, (chann, d)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yb = blk(x.cuda())
        assert yb.dtype == torch.bfloat16 and T._rel(yb.float().cpu(), ref) < T.BF16_LOGIT_TOL, (chann, d)
    apn = APNModule(128, 19)
    apn.load_state_dict({k[len("apn."):]: v for k, v in sd.items() if k.startswith("apn.")})
    x = torch.randn(2, 128, 24, 40).relu()
    ref = nets.led_apn(nets.SD(sd, "apn."), x)
    y = apn.cuda().eval()(x.cuda())
    assert y.shape == ref.shape and T._rel(y.float().cpu(), ref) < 1e-4


def test_gate_bcast_matches_torch():
    from esn import ops
    torch.manual_seed(1)
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
        g = ops.as_act(torch.randn(2, 1, 9, 13).cuda(), dt)
        x = ops.as_act(torch.randn(2, 19, 9, 13).cuda(), dt)
        b = ops.as_act(torch.randn(2, 19, 1, 1).cuda(), dt)
        want = g.float() * x.float() + b.float()
        assert T._rel(ops.gate_bcast(g, x, b).float(), want) < tol
        assert T._rel(ops.gate_bcast(g, x).float(), g.float() * x.float()) < tol


def test_contextnet_quarter_scale_image():
    import torch.nn.functional as F
    from model.ContextNet import quarter_scale_image
    x = fixture.make_input(2, 64, 136).cuda()
    y = quarter_scale_image(x)
    ref = F.interpolate(x, scale_factor=0.25, mode="bilinear", align_corners=True)
    assert y.shape == ref.shape and y.is_contiguous()
    assert torch.allclose(y, ref, atol=1e-3)          # inputs span [-83, 183]


@pytest.mark.parametrize("name", NETS)
def test_full_size_properties(name, spec):
    """512x1024: batch-permutation equivariance, fused argmax == argmax of the logits, bf16 close to fp32."""
    m = T._model(name, spec)
    x = fixture.make_input(2, 512, 1024).cuda()
    with torch.no_grad():
        y = m(x)
        assert torch.equal(y, m(x.flip(0)).flip(0)), "batch-permutation equivariance"
        logits, mask = m.predict_mask(x, with_logits=True)
        assert torch.equal(mask.long(), logits.argmax(1)), "fused argmax"
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yb = m(x)
        tol = T.BF16_LOGIT_TOL
        if name == "LEDNet":
            # ill-conditioned attention pyramid (see test_lednet_bf16_matches_oracle_stage_by_stage): the bound is what the
            # fp32 module makes of the bf16 features, plus the usual tolerance for the module's own bf16 arithmetic
            f32 = _lednet_features(m, x)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                fb = _lednet_features(m, x)
            assert T._rel(fb.float(), f32) < T.BF16_LOGIT_TOL
            tol += T._rel(m.apn(fb.float())[:, :19], m.apn(f32)[:, :19])
    r = T._rel(yb.float(), y)
    print("%s 2x512x1024: bf16 vs own fp32 rel-L2 %.3e (tolerance %.3e)" % (name, r, tol))
    assert r < tol, (r, tol)


def test_focal_loss_matches_reference_golden(golden):
    """FocalLoss2d drop-in (loss.py:96-127) on the fused weighted-CE kernel: value and d-logits vs the
    reference class in fp64."""
    from utils.losses.loss import FocalLoss2d
    g = golden("loss")
    w = torch.tensor(fixture.CLASS_WEIGHTS)
    crit = FocalLoss2d(alpha=0.5, gamma=2, weight=w, ignore_index=255).cuda()
    logits = torch.from_numpy(g["logits"]).float().cuda().requires_grad_(True)
    loss = crit(logits, torch.from_numpy(g["labels"]).cuda())
    loss.backward()
    assert abs(loss.item() - g["focal_loss"][0]) < 1e-4 * abs(g["focal_loss"][0])
    assert T._rel(logits.grad.cpu(), torch.from_numpy(g["focal_grad"])) < 1e-4


@pytest.mark.parametrize("size", [(1, 51, 77), (2, 36, 50), (1, 33, 64)])
def test_esnet_odd_input_sizes(size, spec, golden):
    """ESNet.py:22-29 on the device: inputs that are odd at one or more levels.  fp32 logits against the unmodified
    reference (tests/golden/oddsize.npz), bf16 logits and the fused mask against the CPU oracle."""
    n, h, w = size
    m = T._model("ESNet", spec)
    ref = torch.from_numpy(golden("oddsize")["ESNet_%dx%dx%d_logits" % size])
    x = fixture.make_input(n, h, w).cuda()
    with torch.no_grad():
        y = m(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16, mask = m.predict_mask(x, with_logits=True)
    assert y.shape == ref.shape and mask.shape == (n,) + tuple(ref.shape[2:])
    assert T._rel(y.cpu(), ref) < 1e-3
    assert (y.cpu().argmax(1) == ref.argmax(1)).float().mean().item() >= 0.999
    assert T._rel(y16.float().cpu(), ref) < 5e-2
    assert torch.equal(mask.cpu().long(), y16.float().cpu().argmax(1)) or \
        (mask.cpu().long() == y16.float().cpu().argmax(1)).float().mean().item() > 0.995     # fp32 accumulators vs bf16-rounded logits


@pytest.mark.parametrize("size", [(1, 51, 77), (2, 36, 50), (1, 33, 64)])
def test_lednet_odd_input_sizes(size, spec, golden):
    """LEDNet on inputs that are odd at one or more levels (LEDNet.py:76-96, 245-264): fp32 logits against the unmodified
    reference; the bf16 path runs and stays finite (its accuracy is bounded stage by stage above)."""
    n, h, w = size
    m = T._model("LEDNet", spec)
    ref = torch.from_numpy(golden("oddsize")["LEDNet_%dx%dx%d_logits" % size])
    x = fixture.make_input(n, h, w).cuda()
    with torch.no_grad():
        y = m(x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            mask = m.predict_mask(x)
    assert y.shape == ref.shape and T._rel(y.cpu(), ref) < 1e-3
    assert mask.shape == (n, h, w) and mask.dtype == torch.uint8 and int(mask.max()) < 19
