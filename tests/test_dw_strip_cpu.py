"""Register-strip depthwise conv (csrc/esn_dw_strip_kernel.cuh): the kernel's DEVICE SOURCE, float instantiation, compiled
with g++ (tests/dw_strip_host.cpp) and compared with torch's depthwise conv2d + the same epilogue -- work decomposition
(dilated column chains, segments, column groups), zero padding on all four sides, ragged widths, channel strides, halo rows
between segments, residual / activation order, and that nothing outside the logical channels is written.  The bf16
instantiation differs in the vector width (8 channels) and the conversions only; its device run is tests/test_ops_gpu.py."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_exe(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("dws") / "dw_strip_host")
    subprocess.run([gxx, "-O1", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "tests"),
                    "-I" + os.path.join(ROOT, "efficient-segmentation-networks_b200", "csrc"),
                    os.path.join(ROOT, "tests", "dw_strip_host.cpp"), "-o", exe], check=True)
    return exe


CASES = [
    # n, h, w, C, x_cs, y_cs, kh, kw, dil_h, dil_w, act, has_res, pre_act, tw, seg_max, min_threads
    (1, 9, 11, 8, 8, 8, 3, 3, 1, 1, 0, 0, 0, 2, 4, 0),            # several segments per chain, ragged width
    (2, 13, 10, 8, 12, 16, 3, 3, 2, 2, 2, 1, 0, 2, 4, 0),         # dilation 2, padded strides, PReLU + residual
    (1, 7, 9, 4, 4, 4, 3, 3, 4, 4, 1, 1, 1, 4, 32, 0),            # dilation 4 ~ image size, ReLU before and after residual
    (1, 5, 6, 8, 8, 8, 3, 3, 16, 16, 0, 0, 0, 2, 32, 0),          # dilation beyond the image: only the centre tap lands
    (2, 17, 8, 8, 8, 8, 3, 1, 3, 1, 2, 0, 0, 2, 4, 0),            # 3x1, vertical dilation 3
    (1, 6, 21, 12, 12, 12, 1, 3, 1, 5, 1, 1, 0, 4, 8, 0),         # 1x3, horizontal dilation 5
    (1, 40, 7, 4, 8, 4, 3, 3, 1, 1, 0, 0, 0, 4, 32, 10 ** 9),     # planner shortens the chains when the grid is small
    (3, 4, 4, 4, 4, 4, 3, 1, 1, 1, 0, 1, 0, 4, 32, 0),
]


@pytest.mark.parametrize("case", CASES)
def test_dw_strip_device_source_on_cpu(host_exe, case):
    n, h, w, C, x_cs, y_cs, kh, kw, dh, dw, act, has_res, pre_act, tw, seg_max, min_threads = case
    g = torch.Generator().manual_seed(sum(case[:14]))
    x = torch.randn(n, h, w, x_cs, generator=g)
    wt = torch.randn(kh * kw, C, generator=g)
    sc = torch.rand(C, generator=g) + 0.5
    sh = torch.randn(C, generator=g)
    al = torch.rand(C, generator=g) * 0.4 + 0.05
    res = torch.randn(n, h, w, y_cs, generator=g)
    blob = b"".join(t.numpy().astype(np.float32).tobytes() for t in (x, wt, sc, sh, al)) + (res.numpy().tobytes() if has_res else b"")
    r = subprocess.run([host_exe] + [str(v) for v in case], input=blob, capture_output=True, timeout=300, check=True)
    y = torch.from_numpy(np.frombuffer(r.stdout, dtype=np.float32).copy()).view(n, h, w, y_cs)
    weight = wt.view(kh, kw, C).permute(2, 0, 1).unsqueeze(1).contiguous()
    dh_e, dw_e = (dh if kh == 3 else 1), (dw if kw == 3 else 1)
    ref = F.conv2d(x[..., :C].permute(0, 3, 1, 2), weight, None, 1, ((kh // 2) * dh_e, (kw // 2) * dw_e), (dh_e, dw_e), C)
    ref = ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    f = {0: lambda t: t, 1: torch.relu, 2: lambda t: torch.where(t >= 0, t, t * al.view(1, -1, 1, 1))}[act]
    if has_res:
        if pre_act:
            ref = f(ref)
        ref = ref + res[..., :C].permute(0, 3, 1, 2)
    ref = f(ref).permute(0, 2, 3, 1)
    assert torch.allclose(y[..., :C], ref, rtol=1e-5, atol=1e-5), (y[..., :C] - ref).abs().max()
    assert (y[..., C:] == -12345.0).all()              # channels beyond C are never written
