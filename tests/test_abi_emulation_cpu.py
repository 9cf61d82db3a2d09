"""Host side of the C ABI, checked without a GPU (tests/abi_emulation.py): all of esn.ops runs -- routing, shape gates,
weight packing per kernel family, epilogue blocks, channel-slice descriptors -- and only the foreign call is answered by
a CPU model of the entry point working from the raw structs.

* fp32: every conv goes through esn_conv2d_direct's packed weights; logits must match the reference golden to 1e-5.
* bf16: the tcgen05 routes (plain, 64-channel input slices, 256-channel output slices, the phase-fused transposed
  conv, the factorized pair) are taken exactly as on the device; logits must match the fp32 oracle within the bf16
  tolerance of tests/test_models_gpu.py (5e-2, or 1.5x torch's own bf16-autocast error on the same graph).
"""
import pytest
import torch

from abi_emulation import emulate_abi
from conftest import spec_state_dict
from oracle import fixture, nets

NETS = ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet", "ESNet", "ContextNet", "EDANet", "LEDNet"]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _model(name, spec):
    from builders.model_builder import build_model
    m = build_model(name, 19)
    m.load_state_dict(spec_state_dict(spec, name))
    return m.eval()


@pytest.mark.parametrize("name", NETS)
def test_fp32_through_the_abi_matches_reference_golden(name, spec, golden):
    m, g = _model(name, spec), golden(name)
    x = fixture.make_input(1, 64, 128)
    with emulate_abi() as calls, torch.no_grad():
        logits, mask = m.predict_mask(x, with_logits=True)
    assert _rel(logits, torch.from_numpy(g["eval_1x64x128_logits"])) < 1e-5
    assert (mask.numpy() == g["eval_1x64x128_argmax"]).mean() > 0.9999
    assert not any(n in ("esn_conv2d_umma", "esn_conv_pair_umma") for n, _ in calls)      # fp32 never takes a bf16 kernel


@pytest.mark.parametrize("name", NETS)
def test_bf16_routes_through_the_abi(name, spec):
    m = _model(name, spec)
    sd = spec_state_dict(spec, name)
    x = fixture.make_input(2, 128, 256)
    with torch.no_grad():
        ref = nets.forward(name, sd, x)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            rel_ac = _rel(nets.forward(name, sd, x).float(), ref)
        with emulate_abi(bf16=True) as calls:
            y = m(x)
            names = [n for n, _ in calls]
    assert y.dtype == torch.bfloat16 and y.shape == ref.shape
    rel = _rel(y.float(), ref)
    print("%s bf16 through the ABI model: rel-L2 %.3e (torch bf16 autocast of the oracle graph: %.3e); launches %d, tcgen05 %d"
          % (name, rel, rel_ac, len(names), sum(n in ("esn_conv2d_umma", "esn_conv_pair_umma") for n in names)))
    assert rel < max(5e-2, 1.5 * rel_ac), (rel, rel_ac)
    assert "esn_conv2d_umma" in names                        # the tensor-core route is the one exercised


def test_erfnet_routing_is_the_benchmarked_one(spec):
    """At a pair-kernel shape (W multiple of 128) ERFNet's bf16 step makes the entry-point calls that
    profiles/r01_bench_erfnet.json lists under "kernels": 18 pair launches (C = 16 / 64 blocks), 36 single tcgen05 convs
    (8 C=128 blocks x 4, two downsamplers, two transposed convs), the stem, two max-pool branches, the head -- and no
    direct conv."""
    m = _model("ERFNet", spec)
    x = fixture.make_input(1, 128, 1024)
    with emulate_abi(bf16=True) as calls, torch.no_grad():
        m.predict_mask(x)
        names = [n for n, _ in calls]
    assert names.count("esn_conv_pair_umma") == 18
    assert names.count("esn_conv2d_umma") == 36
    assert names.count("esn_conv2d_direct") == 0
    assert names.count("esn_stem_conv3x3s2") == 1 and names.count("esn_head_convt2x2_mask") == 1
    assert names.count("esn_maxpool2x2_affine_act") == 2 and len(names) == 58


@pytest.mark.parametrize("name", ["ERFNet", "ESNet"])
def test_tensor_core_mask_head_equals_the_logits_head(name, spec):
    """predict_mask (esn_head_convt2x2_mask: hi / lo bf16 weight fragments in mma order) against the argmax of the logits of
    esn_head_convt2x2 on the same features: the host packing of the fragments, decoded by the ABI model, must reproduce the
    weights (masks equal wherever the top-2 margin exceeds the 2^-17 of the hi + lo split)."""
    m = _model(name, spec)
    x = fixture.make_input(1, 64, 128)
    with emulate_abi(bf16=True) as calls, torch.no_grad():
        mask = m.predict_mask(x)
        assert [n for n, _ in calls].count("esn_head_convt2x2_mask") == 1
        logits, mask2 = m.predict_mask(x, with_logits=True)
    assert mask.dtype == torch.uint8 and mask.shape == mask2.shape
    top2 = logits.float().topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-2 * logits.float().abs().amax(dim=1).clamp_min(1.0)      # logits are bf16-rounded
    assert clear.float().mean().item() > 0.9
    assert torch.equal(mask[clear], mask2[clear])


def test_input_pipeline_through_the_abi(golden):
    from esn import ops
    g = golden("pipeline")
    with emulate_abi():
        for i in range(3):
            y = ops.image_u8_to_f32(torch.from_numpy(g["image%d" % i][None].copy()), g["mean"])
            assert (y[0].numpy() == g["input%d" % i]).all()


def test_declined_tcgen05_call_takes_the_direct_kernel(spec, golden):
    """If esn_conv2d_umma answers ESN_ERR_UNSUPPORTED for a shape the host-side gate let through, ops.conv2d goes on to
    the direct CUDA kernel instead of failing the forward (the sliced / phase-fused routes keep raising)."""
    import abi_emulation as A
    m = _model("ESNet", spec)
    sd = spec_state_dict(spec, "ESNet")
    x = fixture.make_input(1, 64, 128)
    A.DECLINE.add("esn_conv2d_umma")
    try:
        with torch.no_grad(), emulate_abi(bf16=True) as calls:
            y = m(x)
            names = [n for n, _ in calls]
    finally:
        A.DECLINE.clear()
    assert names.count("esn_conv2d_umma") == 1 and names.count("esn_conv2d_direct") > 30     # 1 = the phase-fused 64 -> 16 transposed conv
    with torch.no_grad():
        ref = nets.forward("ESNet", sd, x)
    assert _rel(y.float(), ref) < 5e-2


def test_real_library_planner_accepts_every_call_at_benchmark_shapes():
    """tests/planner_dry_run.py in a subprocess (ESN_DRY_RUN=1 must not leak into this process): every C-ABI call of a
    bf16 forward at 1024x2048 goes to the REAL libesn_sm100.so; argument validation of every entry point and the full
    launch planner of esn_conv2d_umma (tile shapes, reuse mode, shared-memory / TMEM budget, nominal B200 limits) must
    accept it.  -4 = validation passed and the launch failed for lack of a GPU; 0 = planner dry run accepted."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    args = []
    for name in NETS:
        args += [name, "1024", "2048"]
    for name in ("ESNet", "ContextNet", "EDANet", "LEDNet"):
        args += [name, "64", "128", name, "512", "1024"]
    # the dry-run switch exists only in the test build of the library (csrc/Makefile: -DESN_TESTING)
    tlib = os.path.join(os.path.dirname(here), "efficient-segmentation-networks_b200", "esn", "libesn_sm100_testing.so")
    assert os.path.exists(tlib), "run `make -C efficient-segmentation-networks_b200/csrc` (builds the testing library too)"
    env = dict(os.environ, ESN_DRY_RUN="1", ESN_LIB_PATH=tlib)
    r = subprocess.run([sys.executable, os.path.join(here, "planner_dry_run.py")] + args, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(res) == len(NETS) + 8
    for key, v in res.items():
        assert v["refused"] == [], (key, v["refused"][:5])
        assert all(k.endswith(":0") or k.endswith(":-4") for k in v["calls"]), (key, v["calls"])
        assert v["calls"].get("esn_conv2d_umma:0", 0) > 10, (key, v["calls"])
    # the verified default workload launches what the B200 bench recorded (profiles/r01_bench_erfnet.json "kernels")
    erf = res["ERFNet@1024x2048"]["calls"]
    assert erf["esn_conv2d_umma:0"] == 36 and erf["esn_conv_pair_umma:-4"] == 18
    assert "ESN_DRY_RUN" not in os.environ


def test_a_view_that_overruns_its_buffer_is_refused():
    """A channel slice widened past the end of its pixel would read out of bounds at the last pixel of the buffer.
    torch refuses to build such a view when the buffer is the whole storage (ops.widen -> as_strided); the C-ABI model
    checks descriptors against the tracked activation buffers for the remaining case (a view into a larger storage)."""
    import abi_emulation as A
    from esn import ops
    from esn._lib import ACT_NONE
    with emulate_abi():
        buf = ops.new_act(1, 8, 2, 2, torch.float32, "cpu", zero=True)
        tail = buf[:, 4:8]
        ops.affine_act(tail, None, None, None, ACT_NONE)                       # in bounds
        with pytest.raises(RuntimeError, match="out of bounds"):
            ops.widen(tail, 8)                                                 # 4 channels past the last pixel
        d = ops.tdesc(tail)
        d.c = 8                                                                # the same overrun, forged at descriptor level
        with pytest.raises(AssertionError, match="overruns"):
            A.tensor(d)


def test_losses_through_the_abi_match_reference(golden):
    """utils.losses.loss.CrossEntropyLoss2d / FocalLoss2d (drop-ins for loss.py:15-32 / :96-127): forward value and the
    gradient autograd returns -- the scalar focal factor reaches the CE kernel's backward as its upstream gradient --
    against the reference classes' fp64 golden, with esn_weighted_ce answered by the C-ABI model."""
    from utils.losses.loss import CrossEntropyLoss2d, FocalLoss2d
    g = golden("loss")
    lab = torch.from_numpy(g["labels"])
    w = torch.tensor(fixture.CLASS_WEIGHTS)
    for crit, lkey, gkey in ((CrossEntropyLoss2d(weight=w, ignore_label=255), "loss", "grad"),
                             (FocalLoss2d(alpha=0.5, gamma=2, weight=w, ignore_index=255), "focal_loss", "focal_grad")):
        logits = torch.from_numpy(g["logits"]).float().requires_grad_(True)
        with emulate_abi():
            loss = crit(logits, lab)
            loss.backward()
        assert abs(loss.item() - g[lkey][0]) < 1e-5 * abs(g[lkey][0]), lkey
        ref = torch.from_numpy(g[gkey])
        assert _rel(logits.grad, ref) < 1e-5, gkey
    assert list(FocalLoss2d(weight=w).state_dict().keys()) == ["ce_fn.weight"]       # the reference's only key


@pytest.mark.parametrize("case", ["kth_above_thresh", "thresh_wins", "nothing_filtered", "unweighted"])
def test_ohem_loss_through_the_abi_matches_reference(golden, case):
    """utils.losses.loss.ProbOhemCrossEntropy2d (drop-in for loss.py:163-216): value and gradient against the reference
    class's fp64 golden (tools/make_golden_ohem.py), the three entry points answered by the C-ABI model."""
    from utils.losses.loss import ProbOhemCrossEntropy2d
    g, o = golden("loss"), golden("ohem")
    thresh, min_kept, use_weight = o[case + "_cfg"]
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=float(thresh), min_kept=int(min_kept), use_weight=bool(use_weight))
    assert list(crit.state_dict().keys()) == (["criterion.weight"] if use_weight else [])
    logits = torch.from_numpy(g["logits"]).float().requires_grad_(True)
    with emulate_abi():
        loss = crit(logits, torch.from_numpy(g["labels"]))
        loss.backward()
    assert abs(loss.item() - o[case + "_loss"][0]) < 1e-5 * abs(o[case + "_loss"][0])
    assert _rel(logits.grad, torch.from_numpy(o[case + "_grad"])) < 1e-5


def test_pending_gpu_test_code_runs_on_the_abi_model():
    """tests/preflight_gpu_tests_on_cpu.py: the functions of tests/test_zz_*_gpu.py (not yet run on a B200) executed on
    the CPU against the C-ABI model, in a subprocess because the script patches torch globally."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "preflight_gpu_tests_on_cpu.py")], capture_output=True, text=True,
                       timeout=1200)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith(("PASS", "FAIL"))]
    assert r.returncode == 0 and lines and not any(ln.startswith("FAIL") for ln in lines), "\n".join(lines[-30:]) + r.stderr[-1500:]
    assert sum(ln.startswith("PASS") for ln in lines) >= 20


@pytest.mark.parametrize("classes", [11, 32])
def test_espnet_with_another_class_count(classes):
    """build_model('ESPNet', C) for C != 19 (ESPNet.py:350 takes the class count as an argument; only the 19-class layout has
    the fused 2x2 transposed-conv head, the others close with esn_conv2d_direct + the same-size head kernel): fp32 and bf16
    against the oracle on the model's own seeded weights, through the C-ABI model."""
    from builders.model_builder import build_model
    m = build_model("ESPNet", classes)
    sd = fixture.randomize_state_dict(m.state_dict(), 77)
    m.load_state_dict(sd)
    m = m.eval()
    x = fixture.make_input(1, 64, 128)
    with torch.no_grad():
        ref = nets.forward("ESPNet", sd, x)
        with emulate_abi() as calls:
            logits, mask = m.predict_mask(x, with_logits=True)
        assert logits.shape == (1, classes, 64, 128) and _rel(logits, ref) < 1e-5
        assert (mask.numpy() == nets.argmax_mask(ref)).mean() > 0.9999
        assert "esn_head_convt2x2" not in [n for n, _ in calls]
        with emulate_abi(bf16=True):
            yb = m(x)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            rel_ac = _rel(nets.forward("ESPNet", sd, x).float(), ref)
    assert yb.dtype == torch.bfloat16 and _rel(yb.float(), ref) < max(5e-2, 1.5 * rel_ac)


def test_dabnet_dual_epilogue_routing_equals_two_launches(spec, monkeypatch):
    """DABNet's bf16 step: 7 of the 9 conv1x1 + residual launches also emit the next module's bn_relu_1 (dual), the last
    module of each block and init_conv[2] write only the concat's BNPReLU slice (chain), no BNPReLU pass runs over a whole
    concat buffer -- and the logits are bit-identical to the route with one launch per op (ESN_DUAL=0)."""
    from esn import ops
    m = _model("DABNet", spec)
    x = fixture.make_input(2, 128, 256)
    with emulate_abi(bf16=True) as calls, torch.no_grad():
        y1 = m(x)
        names = [n for n, _ in calls]
    assert names.count("esn_conv2d_umma_dual") == 10          # 7 dual + 2 chained block outputs + init_conv[2]
    n_affine = names.count("esn_affine_act")
    monkeypatch.setattr(ops, "DUAL_ENABLED", False)
    with emulate_abi(bf16=True) as calls, torch.no_grad():
        y0 = m(x)
        names0 = [n for n, _ in calls]
    assert "esn_conv2d_umma_dual" not in names0
    assert names0.count("esn_affine_act") == n_affine + 10
    assert torch.equal(y0, y1)


def test_bilinear_ce_wrapper_through_the_abi():
    """ops.bilinear_ce (the fused close of a training iteration, esn_bilinear_ce): descriptors, target / weight pointers and the
    padded fp32 gradient buffer as the host passes them, answered by the ABI model (torch autograd)."""
    import torch.nn.functional as F
    from esn import ops
    torch.manual_seed(3)
    n, c, h, w, s = 2, 19, 4, 6, 8
    tgt = torch.randint(0, c, (n, s * h, s * w))
    tgt[:, :5] = 255
    wt = torch.rand(c) + 0.5
    with emulate_abi(bf16=False) as calls:
        x = ops.new_act(n, c, h, w, torch.float32, "cpu", c_alloc=32)
        x.copy_(torch.randn(n, c, h, w))
        sums, ds = ops.bilinear_ce(x, tgt, wt, 255, s * h, s * w)
        assert ops.bilinear_ce(x, tgt.int(), wt, 255, s * h, s * w) is None          # int32 target: declined on the host
        assert [nm for nm, _ in calls] == ["esn_bilinear_ce"]
        # align_corners=True at a non-integer scale (Fast-SCNN's / ESPNetv2's close)
        tgt2 = torch.randint(0, c, (n, 29, 41))
        sums2, ds2 = ops.bilinear_ce(x, tgt2, None, 255, 29, 41, align_corners=True)
    xr2 = x.detach().clone().contiguous().requires_grad_(True)
    loss2 = F.cross_entropy(F.interpolate(xr2, size=(29, 41), mode="bilinear", align_corners=True), tgt2, reduction="sum")
    loss2.backward()
    assert abs(sums2[0].item() - loss2.item()) < 1e-4 * abs(loss2.item()) and _rel(ds2, xr2.grad) < 1e-5
    with emulate_abi(bf16=False):
        pass
    xr = x.detach().clone().contiguous().requires_grad_(True)
    loss = F.cross_entropy(F.interpolate(xr, scale_factor=s, mode="bilinear", align_corners=False), tgt, wt, ignore_index=255,
                           reduction="sum")
    loss.backward()
    assert abs(sums[0].item() - loss.item()) < 1e-4 * abs(loss.item())
    assert abs(sums[1].item() - wt[tgt[tgt != 255]].sum().item()) < 1e-3
    assert ds.shape == (n, c, h, w) and ds.stride(3) == 20 and _rel(ds, xr.grad) < 1e-5


@pytest.mark.parametrize("size", [(1, 51, 77), (2, 36, 50), (1, 33, 64)])
def test_lednet_odd_input_sizes_through_the_abi(size, spec, golden):
    """LEDNet.py:76-96 + 245-264 on odd sizes, fp32 against the reference's golden (bf16: the attention pyramid is
    ill-conditioned with random weights, tests/test_zz_widening_gpu.py)."""
    m = _model("LEDNet", spec)
    ref = torch.from_numpy(golden("oddsize")["LEDNet_%dx%dx%d_logits" % size])
    with emulate_abi() as calls, torch.no_grad():
        y = m(fixture.make_input(*size))
    assert y.shape == ref.shape and _rel(y.float(), ref) < 1e-5


@pytest.mark.parametrize("size", [(1, 51, 77), (2, 36, 50), (1, 33, 64)])
@pytest.mark.parametrize("bf16", [False, True])
def test_esnet_odd_input_sizes_through_the_abi(size, bf16, spec, golden):
    """ESNet.py:22-29: odd heights / widths.  The stride-2 conv of the block takes the direct kernel, the pool kernel writes
    the zero-padded map (esn_maxpool2x2_affine_act with a ceil-sized output); fp32 must reproduce the reference's golden."""
    n, h, w = size
    m = _model("ESNet", spec)
    ref = torch.from_numpy(golden("oddsize")["ESNet_%dx%dx%d_logits" % size])
    with emulate_abi(bf16=bf16) as calls, torch.no_grad():
        y = m(fixture.make_input(n, h, w))
        mask = m.predict_mask(fixture.make_input(n, h, w))
    assert y.shape == ref.shape and mask.shape == (n,) + tuple(ref.shape[2:])
    assert _rel(y.float(), ref) < (1e-5 if not bf16 else 5e-2)
    assert any(nm == "esn_maxpool2x2_affine_act" for nm, _ in calls)
