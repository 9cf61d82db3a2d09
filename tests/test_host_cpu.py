"""CPU-side checks: the C-ABI library loads and exports every symbol include/esn.h declares,
the drop-in modules expose the reference's state_dict keys, and the product path fails loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, spec_state_dict


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "esn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(esn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from esn import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "libesn_sm100.so does not export %s" % name
    assert set(declared) == set(_lib.SYMBOLS), set(declared) ^ set(_lib.SYMBOLS)
    assert _lib.lib.esn_version() == 100
    assert _lib.lib.esn_strerror(-3).decode().startswith("configuration not supported")


def test_struct_layouts_match_header_sizes():
    from esn import _lib as L
    assert ctypes.sizeof(L.EsnTensor) == 40
    assert ctypes.sizeof(L.EsnEpilogue) == 24 + 8 + 40
    assert ctypes.sizeof(L.EsnConv) == 2 * 40 + 8 + 10 * 4 + 72


@pytest.mark.parametrize("name", ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet", "ESNet", "ContextNet", "EDANet", "LEDNet"])
def test_state_dict_keys_match_reference(name, spec):
    from builders.model_builder import build_model
    m = build_model(name, 19)
    sd = m.state_dict()
    ref = spec[name]["keys"]
    assert list(sd.keys()) == [k for k, _, _ in ref]
    for k, shape, dt in ref:
        assert list(sd[k].shape) == shape, k
        assert str(sd[k].dtype) == dt, k
    assert sum(p.numel() for p in m.parameters()) == spec[name]["n_params"]
    m.load_state_dict(spec_state_dict(spec, name))      # reference checkpoints load unchanged


def test_builder_contract():
    from builders.model_builder import build_model
    with pytest.raises(NotImplementedError):
        build_model("UNet", 19)
    with pytest.raises(NotImplementedError):
        build_model("NoSuchNet", 19)


def test_no_cpu_fallback():
    from builders.model_builder import build_model
    m = build_model("ERFNet", 19).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 32, 64))
    m.train()
    if torch.cuda.is_available():
        return
    with pytest.raises((RuntimeError, NotImplementedError)):
        m(torch.zeros(1, 3, 32, 64))


def test_fused_transposed_conv_packing_equals_conv_transpose():
    """Host logic of the phase-fused transposed conv (ops.ConvPrep.fused_convt, used for ERFNet's UpsamplerBlock:
    ERFNet.py:109-112): ConvTranspose2d(3, s2, p1, op1) == ONE stride-1 conv with 2x2 taps and 4*Cout outputs ordered
    (row parity, column parity, c) + pixel shuffle.  Pure torch on the CPU -- no kernel call."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from esn import ops
    from esn._lib import ACT_RELU
    torch.manual_seed(0)
    for cin, cout in ((8, 4), (16, 16), (64, 16)):
        m = nn.ConvTranspose2d(cin, cout, 3, stride=2, padding=1, output_padding=1, bias=True)
        prep = ops.ConvPrep(m, None, None, ACT_RELU, device="cpu")
        wf, sc4, sh4, _ = prep.fused_convt()
        assert tuple(wf.shape) == (4, 4 * cout, cin) and sh4.shape[0] == 4 * cout
        x = torch.randn(2, cin, 5, 7).to(torch.bfloat16).float()
        w = wf.float().view(2, 2, 4 * cout, cin).permute(2, 3, 0, 1)                  # [4*Cout][Cin][dy][dx]
        y4 = F.conv2d(F.pad(x, (0, 1, 0, 1)), w) * sc4.view(1, -1, 1, 1) + sh4.view(1, -1, 1, 1)
        y = y4.view(2, 2, 2, cout, 5, 7).permute(0, 3, 4, 1, 5, 2).reshape(2, cout, 10, 14)   # out[2i+a, 2j+b]
        ref = F.conv_transpose2d(x, m.weight.detach().to(torch.bfloat16).float(), m.bias.detach(), 2, 1, 1)
        assert (y - ref).abs().max() <= 2e-6 * ref.abs().max()


def test_call_wrapper_contract(monkeypatch):
    """esn.ops._call (the one place every kernel launch goes through): 0 -> True; an error code raises; with
    allow_unsupported an ESN_ERR_UNSUPPORTED answer returns False (the caller takes its next CUDA route) and warns once."""
    from esn import ops, _lib
    monkeypatch.setattr(ops, "stream", lambda: None)
    monkeypatch.setattr(ops, "PROFILE", None)
    seen = []
    assert ops._call(lambda a, st: seen.append((a, st)) or 0, "esn_fake", (7,)) is True and seen == [(7, None)]
    with pytest.raises(_lib.EsnError, match="configuration not supported"):
        ops._call(lambda st: -3, "esn_fake", ())
    with pytest.raises(_lib.EsnError):
        ops._call(lambda st: -4, "esn_fake", (), allow_unsupported=True)      # only UNSUPPORTED is negotiable
    with pytest.warns(UserWarning, match="declined"):
        assert ops._call(lambda st: -3, "esn_fake", (), tag="shape-a", allow_unsupported=True) is False
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert ops._call(lambda st: -3, "esn_fake", (), tag="shape-a", allow_unsupported=True) is False     # warned once already


def test_cross_entropy_module_is_a_drop_in():
    """Same state_dict key as the reference's CrossEntropyLoss2d (its nn.CrossEntropyLoss child is named nll_loss,
    utils/losses/loss.py:23); the global-batch normalisation is opt-in / tied to esn.parallel, never implied by an
    initialised process group alone."""
    from utils.losses.loss import CrossEntropyLoss2d, FocalLoss2d
    from esn import parallel
    w = torch.arange(1, 20, dtype=torch.float32)
    crit = CrossEntropyLoss2d(weight=w, ignore_label=255)
    assert list(crit.state_dict().keys()) == ["nll_loss.weight"]
    assert torch.equal(crit.weight, w) and crit.distributed is None and not parallel.is_active()
    assert list(CrossEntropyLoss2d().state_dict().keys()) == []
    assert CrossEntropyLoss2d(reduction="sum").reduction == "sum"
    with pytest.raises(NotImplementedError):
        CrossEntropyLoss2d(reduction="none")
    with pytest.raises(RuntimeError, match="no CPU path"):
        crit(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    assert list(FocalLoss2d(weight=w).state_dict().keys()) == ["ce_fn.weight"]       # loss.py:104
