"""Host-side composition of the model files, checked WITHOUT a GPU (tests/host_emulation.py): with the kernel
wrappers replaced by torch-CPU functions of the same contract, each model's forward -- packing, BN folding, channel
padding, concat slices, residual chaining, layer order -- must reproduce the unmodified reference's golden logits.

The seven nets of SURVEY 8a are GPU-verified (tests/test_models_gpu.py); here they validate the emulation itself and
guard later host-side refactors.  ESNet and ContextNet reuse their kernels.
"""
import pytest
import torch

from conftest import spec_state_dict
from host_emulation import emulate_kernels
from oracle import fixture, nets


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("name", ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet", "ESNet", "ContextNet", "EDANet", "LEDNet"])
def test_model_composition_matches_reference_golden(name, spec, golden):
    from builders.model_builder import build_model
    m = build_model(name, 19)
    m.load_state_dict(spec_state_dict(spec, name))
    m.eval()
    g = golden(name)
    with emulate_kernels(), torch.no_grad():
        for (n, h, w) in ((1, 64, 128), (2, 128, 256)):
            x = fixture.make_input(n, h, w)
            y = m(x)
            logits, mask = m.predict_mask(x, with_logits=True)
            assert y.shape == (n, 19, h, w) and y.dtype == torch.float32 and y.is_contiguous()
            assert torch.equal(y, logits)
            tag = "eval_%dx%dx%d" % (n, h, w)
            if h == 64:
                assert _rel(y, torch.from_numpy(g[tag + "_logits"])) < 1e-5
            else:
                assert _rel(y[:, :, ::4, ::4], torch.from_numpy(g[tag + "_logits_s4"])) < 1e-5
            assert (mask.numpy() == g[tag + "_argmax"]).mean() > 0.9999


def test_emulation_is_scoped():
    """Outside the context manager the product path is back: CPU tensors are refused."""
    from builders.model_builder import build_model
    m = build_model("ESNet", 19).eval()
    with emulate_kernels():
        pass
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 32, 64))


def test_esnet_blocks_match_oracle(spec):
    """FCU (k = 3 and 5) and PFCU called on their own, NCHW in."""
    from model.ESNet import FCU, PFCU
    sd = spec_state_dict(spec, "ESNet")
    torch.manual_seed(0)
    with emulate_kernels(), torch.no_grad():
        for cls, args, idx, ref_fn in ((FCU, (16, 3, 0.03, 1), 0, lambda p, x: nets.es_fcu(p, x, 3, 1)),
                                       (FCU, (64, 5, 0.03, 1), 4, lambda p, x: nets.es_fcu(p, x, 5, 1)),
                                       (PFCU, (128,), 7, nets.es_pfcu)):
            pre = "layers.%d." % idx
            blk = cls(*args).eval()
            blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
            x = torch.randn(2, args[0], 24, 40)
            assert _rel(blk(x), ref_fn(nets.SD(sd, pre), x)) < 1e-5, cls.__name__


def test_contextnet_quarter_scale_image():
    from model.ContextNet import quarter_scale_image
    import torch.nn.functional as F
    x = fixture.make_input(2, 64, 136)
    with emulate_kernels():
        y = quarter_scale_image(x)
    ref = F.interpolate(x, scale_factor=0.25, mode="bilinear", align_corners=True)
    assert y.shape == ref.shape and y.is_contiguous() and y.dtype == torch.float32
    assert torch.allclose(y, ref, atol=1e-4)


def test_edanet_blocks_match_oracle(spec):
    """EDAModule / EDANetBlock on their own return cat([new, input]) like the reference modules."""
    from model.EDANet import EDAModule, EDANetBlock
    sd = spec_state_dict(spec, "EDANet")
    torch.manual_seed(0)
    with emulate_kernels(), torch.no_grad():
        pre = "layers.2.residual_dense_layers.3."
        blk = EDAModule(180, 2).eval()
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        x = torch.randn(2, 180, 24, 40)
        assert _rel(blk(x), nets.eda_module(nets.SD(sd, pre), x, 2)) < 1e-5
        pre = "layers.2."
        blk = EDANetBlock(60, 5, [1, 1, 1, 2, 2], 40).eval()
        blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        x = torch.randn(2, 60, 24, 40)
        ref = x
        for i, d in enumerate(nets.EDA_BLOCKS[2]):
            ref = nets.eda_module(nets.SD(sd, pre + "residual_dense_layers.%d." % i), ref, d)
        y = blk(x)
        assert y.shape == ref.shape and _rel(y, ref) < 1e-5


def test_lednet_blocks_match_oracle(spec):
    from model.LEDNet import APNModule, SS_nbt_module_paper
    sd = spec_state_dict(spec, "LEDNet")
    torch.manual_seed(0)
    with emulate_kernels(), torch.no_grad():
        for idx, chann, d in ((1, 32, 1), (5, 64, 1), (9, 128, 5)):
            pre = "layers.%d." % idx
            blk = SS_nbt_module_paper(chann, 0.03, d).eval()
            blk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
            x = torch.randn(2, chann, 24, 40)
            assert _rel(blk(x), nets.led_ssnbt(nets.SD(sd, pre), x, d)) < 1e-5, (chann, d)
        apn = APNModule(128, 19).eval()
        apn.load_state_dict({k[len("apn."):]: v for k, v in sd.items() if k.startswith("apn.")})
        for hw in ((24, 40), (9, 13)):            # odd sizes: the pyramid levels are ceil(h/2), ceil(h/4), ceil(h/8)
            x = torch.randn(2, 128, *hw).relu()
            y = apn(x)
            ref = nets.led_apn(nets.SD(sd, "apn."), x)
            assert y.shape == ref.shape and _rel(y, ref) < 1e-5, hw
