"""Host side of esn.optim.Adam without a GPU: the address table / CTA map handed to esn_adam_step (include/esn.h), checked
by running a numpy model of the kernel's contract over host memory and comparing with torch.optim.Adam on the CPU."""
import ctypes as C

import numpy as np
import torch

from esn import _lib as L
from esn.optim import build_tables

SHAPES = [(1,), (3,), (2047,), (2048,), (2049,), (5000,), (8, 3, 3, 3), (4099,)]


def _f32(ptr, n):
    return np.frombuffer((C.c_char * (4 * n)).from_address(ptr), dtype=np.float32)


def _model_adam_step(tab, blk, chunk, lr, step, b1, b2, eps, wd):
    """include/esn.h esn_adam_step, CTA by CTA."""
    t = step + 1.0
    step_size = np.float32(lr) / np.float32(1.0 - b1 ** t)
    bc2_sqrt = np.float32(np.sqrt(1.0 - b2 ** t))
    for ti, ch in blk:
        p_, g_, m_, v_, n = (int(x) for x in tab[ti])
        lo, hi = ch * chunk, min(n, (ch + 1) * chunk)
        p, g, m, v = (_f32(a, n)[lo:hi] for a in (p_, g_, m_, v_))
        gg = g + np.float32(wd) * p
        m[:] = np.float32(b1) * m + np.float32(1 - b1) * gg
        v[:] = np.float32(b2) * v + np.float32(1 - b2) * gg * gg
        p[:] = p - step_size * m / (np.sqrt(v) / bc2_sqrt + np.float32(eps))
    return t


def test_tables_cover_every_element_once_and_reproduce_torch_adam():
    chunk = int(L.lib.esn_adam_chunk())
    assert chunk == 2048
    g = torch.Generator().manual_seed(0)
    ours = [torch.randn(s, generator=g) for s in SHAPES]
    theirs = [torch.nn.Parameter(p.clone()) for p in ours]
    offs, total = [], 0
    for p in ours:
        offs.append(total)
        total += (p.numel() + 3) // 4 * 4
    m, v = torch.zeros(total), torch.zeros(total)
    opt = torch.optim.Adam(theirs, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    step = 0.0
    for it in range(3):
        grads = [torch.randn(p.shape, generator=g) for p in ours]
        tab, blk = build_tables([p.data_ptr() for p in ours], [x.data_ptr() for x in grads], [p.numel() for p in ours], offs,
                                m.data_ptr(), v.data_ptr(), chunk)
        assert tab.dtype == np.int64 and tab.shape == (len(ours), 5) and blk.dtype == np.int32 and blk.shape[1] == 2
        assert C.sizeof(C.c_void_p) * 4 + 8 == 40 == tab.strides[0]          # sizeof(EsnAdamTensor)
        # coverage: every element of every tensor belongs to exactly one CTA
        seen = [np.zeros(p.numel(), dtype=np.int32) for p in ours]
        for ti, ch in blk:
            seen[ti][ch * chunk:(ch + 1) * chunk] += 1
        assert all((s == 1).all() for s in seen)
        assert len(blk) == sum((p.numel() + chunk - 1) // chunk for p in ours)
        step = _model_adam_step(tab, blk, chunk, 1e-3, step, 0.9, 0.999, 1e-8, 1e-4)
        for q, x in zip(theirs, grads):
            q.grad = x.clone()
        opt.step()
        for p, q in zip(ours, theirs):
            assert (p - q.detach()).abs().max().item() <= 2e-6 * max(1.0, q.detach().abs().max().item())
    assert step == 3.0
    # the moments of tensor i sit at its 16-byte aligned slot of the flat buffers
    for i, p in enumerate(ours):
        assert offs[i] % 4 == 0
        assert torch.allclose(m[offs[i]:offs[i] + p.numel()].view_as(p), opt.state[theirs[i]]["exp_avg"], atol=1e-6)


def test_adam_constructor_mirrors_torch_and_has_no_cpu_path():
    """Same argument checks as torch.optim.Adam (train.py:212-215 builds it with lr, betas, eps, weight_decay); amsgrad /
    maximize are not on the path; stepping CPU parameters raises (the product has no CPU path)."""
    import pytest
    from esn.optim import Adam
    p = [torch.nn.Parameter(torch.randn(4, 3))]
    opt = Adam(p, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    g = opt.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"]) == (5e-4, (0.9, 0.999), 1e-8, 1e-4)
    assert isinstance(opt, torch.optim.Optimizer) and opt.state_dict()["state"] == {}
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda i: 0.5 ** i)      # the reference's per-iteration schedulers attach
    assert sched.get_last_lr() == [5e-4]
    for bad in (dict(lr=-1.0), dict(eps=-1.0), dict(betas=(1.0, 0.999)), dict(betas=(0.9, -0.1)), dict(weight_decay=-1.0)):
        with pytest.raises(ValueError):
            Adam(p, **bad)
    for unsupported in (dict(amsgrad=True), dict(maximize=True)):
        with pytest.raises(NotImplementedError):
            Adam(p, **unsupported)
    p[0].grad = torch.randn(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()


def test_fused_head_spec_only_for_the_plain_weighted_ce():
    """model.fused_loss takes the fused close only for CrossEntropyLoss2d itself; focal / OHEM criteria keep the two-module
    form (utils/losses/loss.py)."""
    from utils.losses.loss import CrossEntropyLoss2d, FocalLoss2d, ProbOhemCrossEntropy2d, fused_head_spec
    tgt = torch.randint(0, 19, (1, 8, 8)).to(torch.uint8)
    w = torch.rand(19)
    spec = fused_head_spec(CrossEntropyLoss2d(weight=w, ignore_label=255, reduction="sum", distributed=False), torch.device("cpu"), tgt, 19)
    t, wt, ignore, reduction, distributed = spec
    assert t.dtype == torch.int64 and torch.equal(wt, w) and (ignore, reduction, distributed) == (255, "sum", False)
    assert fused_head_spec(FocalLoss2d(), torch.device("cpu"), tgt, 19) is None
    assert fused_head_spec(ProbOhemCrossEntropy2d(), torch.device("cpu"), tgt, 19) is None

    class Sub(CrossEntropyLoss2d):      # a subclass may change forward: no fused form is assumed for it
        pass
    assert fused_head_spec(Sub(), torch.device("cpu"), tgt, 19) is None
