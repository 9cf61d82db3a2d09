"""esn_gate_bcast (LEDNet's APN close): the kernel's DEVICE SOURCE (csrc/esn_gate_kernel.cuh) compiled with g++
behind tests/cuda_cpu_shim.h and compared with numpy -- strides, the per-image bias index, the grid-stride tail and that
nothing outside the logical channels is written.  The device run itself is tests/test_zz_widening_gpu.py (pending)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gate_on_cpu(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("gate") / "gate_kernel_host")
    subprocess.run([gxx, "-O1", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "tests"),
                    "-I" + os.path.join(ROOT, "efficient-segmentation-networks_b200", "csrc"),
                    os.path.join(ROOT, "tests", "gate_kernel_host.cpp"), "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("case", [
    # n, h, w, C, g_cs, x_cs, b_cs, y_cs, has_b, grid
    (2, 5, 7, 19, 1, 32, 32, 32, 1, 3),        # LEDNet's shapes in small: 19 classes in 32-channel buffers, 1-channel gate
    (1, 3, 3, 19, 1, 19, 19, 19, 0, 1),        # no bias, dense strides, one CTA walking the whole tensor
    (3, 4, 4, 5, 1, 8, 5, 6, 1, 7),            # every stride different
])
def test_gate_device_source_on_cpu(gate_on_cpu, case):
    n, h, w, C, g_cs, x_cs, b_cs, y_cs, has_b, grid = case
    rng = np.random.RandomState(sum(case))
    npix = n * h * w
    g = rng.randn(npix, g_cs).astype(np.float32)
    x = rng.randn(npix, x_cs).astype(np.float32)
    b = rng.randn(n, b_cs).astype(np.float32)
    r = subprocess.run([gate_on_cpu] + [str(v) for v in case], input=g.tobytes() + x.tobytes() + (b.tobytes() if has_b else b""),
                       capture_output=True, timeout=300, check=True)
    y = np.frombuffer(r.stdout, dtype=np.float32).reshape(npix, y_cs)
    want = g[:, :1] * x[:, :C]
    if has_b:
        want = want + np.repeat(b[:, :C], h * w, axis=0)
    assert np.array_equal(y[:, :C], want.astype(np.float32))
    assert (y[:, C:] == -12345.0).all()                # channels beyond C are never written
