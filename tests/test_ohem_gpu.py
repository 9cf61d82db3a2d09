"""ProbOhemCrossEntropy2d on the device (utils/losses/loss.py:163-216): the radix-select threshold kernel against
torch.sort, the drop-in class against the reference class's fp64 golden (tools/make_golden_ohem.py), and at the size of
BASELINE.json configs[2] (8 x 19 x 512 x 1024) against the same selection done by torch on the device."""
import numpy as np
import pytest
import torch

from oracle import fixture, loss as oloss

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("n,k", [(1, 1), (7, 3), (1000, 1), (1000, 1000), (4097, 2048), (1 << 20, 12345), (3_000_001, 250_000)])
def test_ohem_threshold_is_the_exact_kth_value(n, k):
    from esn import ops
    g = torch.Generator(device="cuda").manual_seed(n + k)
    p = torch.rand(n, device="cuda", generator=g) ** 4          # skewed towards 0, like hard-pixel probabilities
    p[::5] = 1.0                                                 # "ignored" pixels
    if n > 100:
        p[3:60] = p[3]                                           # ties across the k-th value's neighbourhood
    nv = torch.tensor([float(n)], device="cuda")
    kth = torch.sort(p).values[min(n, k) - 1].item()
    for thresh in (0.0, 0.5):
        out = ops.ohem_threshold(p, k, thresh, nv)
        assert out.item() == max(float(np.float32(thresh)), kth), (out.item(), kth)
    # min_kept larger than the number of valid pixels, or not positive: nothing is filtered
    assert ops.ohem_threshold(p, n + 1, 0.5, nv).item() == float("inf")
    assert ops.ohem_threshold(p, 0, 0.5, nv).item() == float("inf")


@pytest.mark.parametrize("case", ["kth_above_thresh", "thresh_wins", "nothing_filtered", "unweighted"])
def test_ohem_matches_reference_golden(golden, case):
    from utils.losses.loss import ProbOhemCrossEntropy2d
    g, o = golden("loss"), golden("ohem")
    thresh, min_kept, use_weight = o[case + "_cfg"]
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=float(thresh), min_kept=int(min_kept), use_weight=bool(use_weight)).cuda()
    logits = torch.from_numpy(g["logits"]).float().cuda().requires_grad_(True)
    loss = crit(logits, torch.from_numpy(g["labels"]).cuda())
    loss.backward()
    assert abs(loss.item() - o[case + "_loss"][0]) < 1e-4 * abs(o[case + "_loss"][0])
    assert _rel(logits.grad.cpu(), torch.from_numpy(o[case + "_grad"])) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ohem_full_size_properties(dtype):
    """8 x 19 x 512 x 1024 (configs[2]; train.py:147-149 uses thresh 0.7, min_kept = batch*h*w/16): the device threshold is
    the k-th smallest labelled-class probability that torch computes from the same logits, the loss is the weighted CE over
    exactly the pixels at or below it, and gradients vanish on every filtered pixel."""
    from esn import ops
    from utils.losses.loss import ProbOhemCrossEntropy2d
    n, c, h, w = 8, 19, 512, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    logits = (torch.randn(n, c, h, w, device="cuda", generator=g) * 2).to(dtype).requires_grad_(True)
    lab = fixture.make_labels(n, h, w, c).cuda()
    min_kept = n * h * w // 16
    for thresh in (0.7, 1e-4):
        crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=thresh, min_kept=min_kept, use_weight=True).cuda()
        logits.grad = None
        loss = crit(logits, lab)
        loss.backward()
        prob = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
        sums, _ = ops.weighted_ce(logits.detach(), lab, None, 255, prob_out=prob)
        valid = lab != 255
        assert sums[1].item() == valid.sum().item()
        kth = torch.kthvalue(prob.flatten(), min_kept).values.item()
        thr = ops.ohem_threshold(prob, min_kept, thresh, sums[1:2]).item()
        assert thr == max(float(np.float32(thresh)), kth)          # the threshold is an fp32 device scalar
        keep = valid & (prob <= thr)
        assert keep.sum().item() >= min_kept
        tgt = torch.where(keep, lab, torch.full_like(lab, 255))
        ref = torch.nn.functional.cross_entropy(logits.detach().float(), tgt, crit.criterion.weight, ignore_index=255)
        assert abs(loss.item() - ref.item()) < 2e-4 * abs(ref.item()), (loss.item(), ref.item())
        gsum = logits.grad.float().abs().sum(1)
        assert (gsum[~keep] == 0).all() and (gsum[keep] > 0).float().mean().item() > 0.999
