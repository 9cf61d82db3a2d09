import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-segmentation-networks_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")
# torch's GPU fp32 convs default to TF32 (10-bit mantissa); the torch reference ops used by the
# kernel-level tests must be true fp32
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def spec():
    return json.load(open(os.path.join(GOLD, "state_dict_spec.json")))


def spec_state_dict(spec, name, seed=1234):
    """Rebuild the seeded fixture weights from the committed key/shape list."""
    from oracle import fixture
    proto = {k: torch.empty(shape, dtype=getattr(torch, dt.split(".")[1])) for k, shape, dt in spec[name]["keys"]}
    return fixture.randomize_state_dict(proto, seed)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLD, name + ".npz"))
        return cache[name]
    return load
