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


@pytest.mark.parametrize("name", ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet"])
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
