"""Pin the CPU oracle (oracle/) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tools/make_golden.py in the build container)."""
import json

import numpy as np
import pytest
import torch

from conftest import spec_state_dict
from oracle import fixture, loss as oloss, nets


@pytest.mark.parametrize("name", ["ERFNet", "DABNet", "ENet", "CGNet", "FastSCNN", "ESPNet_v2", "ESPNet", "ESNet", "ContextNet", "EDANet", "LEDNet"])
def test_eval_forward_matches_reference(name, spec, golden):
    sd = spec_state_dict(spec, name)
    g = golden(name)
    with torch.no_grad():
        y = nets.forward(name, sd, fixture.make_input(1, 64, 128))
    ref = torch.from_numpy(g["eval_1x64x128_logits"])
    assert y.shape == ref.shape
    rel = (y - ref).norm() / ref.norm()
    assert rel < 1e-5, rel  # fp32 re-association noise (BN as scale+shift)
    assert (nets.argmax_mask(y) == g["eval_1x64x128_argmax"]).mean() > 0.9999
    with torch.no_grad():
        y = nets.forward(name, sd, fixture.make_input(2, 128, 256))
    ref = torch.from_numpy(g["eval_2x128x256_logits_s4"])
    rel = (y[:, :, ::4, ::4] - ref).norm() / ref.norm()
    assert rel < 1e-5, rel  # fp32 re-association noise (BN as scale+shift)
    s = g["eval_2x128x256_sum"]
    assert abs(y.double().abs().sum().item() - s[1]) / s[1] < 1e-6
    assert (nets.argmax_mask(y) == g["eval_2x128x256_argmax"]).mean() > 0.9999


@pytest.mark.parametrize("name", ["ESNet", "LEDNet"])
@pytest.mark.parametrize("size", [(1, 51, 77), (2, 36, 50), (1, 33, 64)])
def test_odd_sizes_match_reference(name, size, spec, golden):
    """The F.pad path of ESNet's / LEDNet's DownsamplerBlock (model/ESNet.py:22-29, LEDNet.py:76-96): inputs that are odd at one or
    more levels.  ESNet's logits have 8 * ceil(ceil(ceil(H / 2) / 2) / 2) rows, LEDNet interpolates back to the input size
    (tools/make_golden_oddsize.py, unmodified reference)."""
    n, h, w = size
    ref = torch.from_numpy(golden("oddsize")["%s_%dx%dx%d_logits" % ((name,) + size)])
    with torch.no_grad():
        y = nets.forward(name, spec_state_dict(spec, name), fixture.make_input(n, h, w))
    up = lambda v: 8 * -(-(-(-(-(-v // 2)) // 2)) // 2)
    assert y.shape == ref.shape == ((n, 19, up(h), up(w)) if name == "ESNet" else (n, 19, h, w))
    assert (y - ref).norm() / ref.norm() < 1e-5


@pytest.mark.parametrize("name", ["ERFNet", "DABNet"])
def test_train_forward_backward_matches_reference_fp64(name, spec, golden):
    sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v)
          for k, v in spec_state_dict(spec, name).items()}
    g = golden(name)
    x = fixture.make_input(2, 64, 128).double()
    lab = fixture.make_labels(2, 64, 128, 19)
    y = nets.forward(name, sd, x, train=True)
    l, _, _ = oloss.weighted_ce(y, lab, torch.tensor(fixture.CLASS_WEIGHTS, dtype=torch.float64))
    assert abs(l.item() - g["train_2x64x128_loss"][0]) < 1e-9 * max(1.0, abs(l.item()))
    ref = torch.from_numpy(g["train_2x64x128_logits_s4"]).double()
    assert ((y.detach()[:, :, ::4, ::4] - ref).norm() / ref.norm()) < 1e-6
    l.backward()
    stats = json.loads(bytes(g["train_2x64x128_gradstats"]).decode())
    checked = 0
    for k, (gnorm, gsum, wnorm) in stats.items():
        gr = sd[k].grad
        assert gr is not None, k
        if gnorm < 1e-10 * max(wnorm, 1e-30):   # mathematically-zero grads (bias feeding train-mode BN), SURVEY H8
            continue
        assert abs(gr.norm().item() - gnorm) / gnorm < 1e-6, k
        checked += 1
    assert checked > 20
    for key in g.files:
        if key.startswith("train_2x64x128_grad::"):
            k = key.split("::")[1]
            ref = torch.from_numpy(g[key]).double()
            assert ((sd[k].grad - ref).norm() / ref.norm()) < 1e-5, k


def test_state_dict_spec_counts(spec):
    # parameter counts published by the reference (usage.txt:91-109)
    assert spec["ERFNet"]["n_params"] == 2066642
    assert spec["DABNet"]["n_params"] == 756643
    assert spec["ENet"]["n_params"] == 360422
    assert spec["ContextNet"]["n_params"] == 876563          # usage.txt:95
    assert spec["EDANet"]["n_params"] == 689485              # usage.txt:97
    assert spec["LEDNet"]["n_params"] == 917387              # usage.txt:105
    # usage.txt:100 lists 1,660,607 for ESNet: its hook-based counter sees each PFCU's shared bn2 three times
    assert spec["ESNet"]["n_params"] == 1660607 - 3 * 2 * 256


def test_weighted_ce_matches_reference(golden):
    g = golden("loss")
    logits = torch.from_numpy(g["logits"])
    lab = torch.from_numpy(g["labels"])
    w = torch.tensor(fixture.CLASS_WEIGHTS, dtype=torch.float64)
    l, swl, sw = oloss.weighted_ce(logits, lab, w)
    assert abs(l.item() - g["loss"][0]) < 1e-12
    gr = oloss.weighted_ce_grad(logits, lab, w)
    assert np.abs(gr.numpy() - g["grad"]).max() < 1e-14


def test_focal_loss_matches_reference(golden):
    g = golden("loss")
    logits = torch.from_numpy(g["logits"])
    lab = torch.from_numpy(g["labels"])
    w = torch.tensor(fixture.CLASS_WEIGHTS, dtype=torch.float64)
    l, gr = oloss.focal(logits, lab, w, 255, alpha=0.5, gamma=2)
    assert abs(l.item() - g["focal_loss"][0]) < 1e-12
    assert np.abs(gr.numpy() - g["focal_grad"]).max() < 1e-14


@pytest.mark.parametrize("case", ["kth_above_thresh", "thresh_wins", "nothing_filtered", "unweighted"])
def test_ohem_matches_reference(golden, case):
    """oracle/loss.py:ohem against the unmodified reference ProbOhemCrossEntropy2d (tools/make_golden_ohem.py), fp64."""
    g, o = golden("loss"), golden("ohem")
    thresh, min_kept, use_weight = o[case + "_cfg"]
    w = torch.tensor(oloss.OHEM_CLASS_WEIGHTS, dtype=torch.float32).double() if use_weight else None
    l, gr, thr = oloss.ohem(torch.from_numpy(g["logits"]), torch.from_numpy(g["labels"]), w, 255, float(thresh), int(min_kept))
    assert abs(l.item() - o[case + "_loss"][0]) < 1e-12
    assert np.abs(gr.numpy() - o[case + "_grad"]).max() < 1e-14
    assert (thr is None) == (case == "nothing_filtered")
    if case == "kth_above_thresh":
        assert thr > thresh
    if case == "thresh_wins":
        assert thr == thresh
