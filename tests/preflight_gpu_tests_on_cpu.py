"""TEST INFRASTRUCTURE ONLY (a script, run in its own process by tests/test_abi_emulation_cpu.py; never imported).

Pre-flight for GPU test CODE that has not met a GPU yet: executes the test functions of tests/test_zz_*_gpu.py on the
CPU with `.cuda()` as the identity, `torch.autocast("cuda", bf16)` mapped to a flag that ops.compute_dtype reads, and every
C-ABI call answered by the CPU model of tests/abi_emulation.py.  It catches mistakes in the tests themselves (shapes, keys,
tolerances that the host code cannot meet, API misuse) before they cost GPU minutes; it says nothing about the kernels.
Monkeypatches torch globally -- hence its own process.  Skips the full-size and device-allocation tests.
"""
import sys, json, os, inspect, warnings
warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, 'efficient-segmentation-networks_b200')]
import numpy as np, torch
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
import contextlib
AC = {"on": False}
@contextlib.contextmanager
def fake_autocast(device_type, dtype=None, **k):
    prev = AC["on"]; AC["on"] = (device_type == "cuda"); 
    try:
        if device_type == "cuda":
            with REAL_AUTOCAST("cpu", dtype=dtype): yield
        else:
            with REAL_AUTOCAST(device_type, dtype=dtype, **k): yield
    finally: AC["on"] = prev
REAL_AUTOCAST = torch.autocast
torch.autocast = fake_autocast
import abi_emulation as A
import conftest
spec = json.load(open(os.path.join(HERE, 'golden', 'state_dict_spec.json')))
cache = {}
def golden(name):
    if name not in cache: cache[name] = np.load(os.path.join(HERE, 'golden', '%s.npz' % name))
    return cache[name]
import test_zz_widening_gpu as W
import test_zz_pipeline_gpu as P
import pytest
def params(fn):
    for m in getattr(fn, "pytestmark", []):
        if m.name == "parametrize": return m.args[0], m.args[1]
    return None, [None]
ok = True
for mod in (W, P):
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith("test_") or fn.__module__ != mod.__name__: continue
        if "full_size" in name or "empty_batch" in name: continue          # too slow on the CPU
        argname, values = params(fn)
        for v in values:
            kw = {}
            sig = inspect.signature(fn).parameters
            if "spec" in sig: kw["spec"] = spec
            if "golden" in sig: kw["golden"] = golden
            if argname: kw[argname] = v
            try:
                with A.emulate_abi():
                    # bf16 when the test enters autocast: the real compute_dtype consults torch's autocast state
                    from esn import ops
                    ops.compute_dtype = (lambda x, _real=A.ops.__dict__.get("compute_dtype"): torch.bfloat16 if (x.dtype == torch.bfloat16 or AC["on"]) else torch.float32)
                    fn(**kw)
                print("PASS", name, v if v is not None else "")
            except Exception as e:
                ok = False
                print("FAIL", name, v, type(e).__name__, str(e)[:300])
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
