"""esn.optim.Adam (one esn_adam_step launch per param group; replaces the torch.optim.Adam of train.py:212-215, stepped at
train.py:355) against torch.optim.Adam on the same parameters and gradients: eager steps, ragged / unaligned tensors, a host
learning-rate schedule, CUDA-graph capture, and state_dict exchange in both directions."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1,), (3,), (19,), (2047,), (2048,), (2049,), (5000,), (32, 3, 3, 3), (128, 192, 3, 3), (64,), (4099,), (7, 5)]


def _params(seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in SHAPES]
    # one parameter whose storage is not 16-byte aligned (a view one element into a larger buffer)
    base = torch.randn(3001, device="cuda", generator=g)
    ps.append(torch.nn.Parameter(base[1:]))
    return ps


def _set_grads(ps, it, unaligned=False):
    g = torch.Generator(device="cuda").manual_seed(1000 + it)
    for p in ps:
        if unaligned:
            buf = torch.randn(p.numel() + 1, device="cuda", generator=g)
            p.grad = buf[1:].view_as(p)
        else:
            p.grad = torch.randn(p.shape, device="cuda", generator=g) * (0.1 + it)


def _close(a, b, tol=2e-6):
    for x, y in zip(a, b):
        err = (x.detach() - y.detach()).abs().max().item()
        ref = y.detach().abs().max().item()
        assert err <= tol * max(ref, 1.0), (tuple(x.shape), err, ref)


@pytest.mark.parametrize("wd", [0.0, 1e-4])
@pytest.mark.parametrize("unaligned", [False, True])
def test_adam_matches_torch(wd, unaligned):
    from esn.optim import Adam
    ours, theirs = _params(), _params()
    a = Adam(ours, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    b = torch.optim.Adam(theirs, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    for it in range(6):
        _set_grads(ours, it, unaligned)
        _set_grads(theirs, it, unaligned)
        a.step()
        b.step()
        _close(ours, theirs)
    assert a.state[ours[0]]["step"].item() == 6.0
    _close([a.state[p]["exp_avg"] for p in ours], [b.state[p]["exp_avg"] for p in theirs])
    _close([a.state[p]["exp_avg_sq"] for p in ours], [b.state[p]["exp_avg_sq"] for p in theirs])


def test_adam_one_launch_per_group_and_host_schedule():
    from esn import ops
    from esn.optim import Adam
    ours, theirs = _params(), _params()
    a = Adam([{"params": ours[:5]}, {"params": ours[5:], "weight_decay": 0.0}], lr=1e-3, weight_decay=1e-4)
    b = torch.optim.Adam([{"params": theirs[:5]}, {"params": theirs[5:], "weight_decay": 0.0}], lr=1e-3, weight_decay=1e-4)
    sa = torch.optim.lr_scheduler.LambdaLR(a, lambda i: 0.9 ** i)
    sb = torch.optim.lr_scheduler.LambdaLR(b, lambda i: 0.9 ** i)
    for it in range(4):
        _set_grads(ours, it)
        _set_grads(theirs, it)
        ops.launch_count_reset()
        a.step()
        assert ops.launch_count() == 2           # one launch per param group
        b.step()
        sa.step()
        sb.step()
    _close(ours, theirs)


def test_adam_in_a_cuda_graph_follows_the_host_schedule():
    """The gradients of the captured step are allocated inside the capture (as esn.graph.GraphedTrainStep's are): the address
    table is rebuilt there and its upload becomes a node of the graph."""
    from esn.optim import Adam
    ours, theirs = _params(), _params()
    a = Adam(ours, lr=1e-3, weight_decay=1e-4)
    b = torch.optim.Adam(theirs, lr=1e-3, weight_decay=1e-4)
    static_g = [torch.zeros_like(p) for p in ours]

    def iteration():
        for p, g in zip(ours, static_g):
            p.grad = g * 1.0
        a.step()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        iteration()                               # warm-up (zero gradients: only the weight decay moves anything)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for p, g in zip(theirs, static_g):
        p.grad = g.clone()
    b.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        iteration()
    for it in range(4):
        lr = 1e-3 * 0.8 ** it
        a.param_groups[0]["lr"] = lr
        b.param_groups[0]["lr"] = lr
        gen = torch.Generator(device="cuda").manual_seed(77 + it)
        for p, q, g in zip(ours, theirs, static_g):
            g.copy_(torch.randn(p.shape, device="cuda", generator=gen))
            q.grad = g.clone()
        a.sync_lr()
        graph.replay()
        b.step()
    torch.cuda.synchronize()
    _close(ours, theirs)
    assert a.state[ours[0]]["step"].item() == 5.0        # warm-up + four replays (the capture pass itself does not execute)


def test_state_dict_moves_both_ways():
    from esn.optim import Adam
    ours, theirs = _params(), _params()
    a = Adam(ours, lr=1e-3, weight_decay=1e-4)
    b = torch.optim.Adam(theirs, lr=1e-3, weight_decay=1e-4)
    for it in range(3):
        _set_grads(ours, it)
        _set_grads(theirs, it)
        a.step()
        b.step()
    # ours -> torch, torch -> ours, then three more steps each: all four optimizers agree
    ours2, theirs2 = [torch.nn.Parameter(p.detach().clone()) for p in ours], [torch.nn.Parameter(p.detach().clone()) for p in theirs]
    a2 = Adam(ours2, lr=1e-3, weight_decay=1e-4)
    b2 = torch.optim.Adam(theirs2, lr=1e-3, weight_decay=1e-4)
    a2.load_state_dict(copy.deepcopy(b.state_dict()))
    b2.load_state_dict(copy.deepcopy(a.state_dict()))
    for it in range(3, 6):
        for ps in (ours, theirs, ours2, theirs2):
            _set_grads(ps, it)
        for o in (a, b, a2, b2):
            o.step()
    _close(ours, theirs)
    _close(ours2, theirs)
    _close(theirs2, theirs)


def test_parameters_without_a_gradient_are_left_alone():
    """ERFNet's encoder.output_conv takes no part in the forward (ERFNet.py:88-89) and never gets a gradient: torch.optim.Adam
    skips such parameters, so does the one-launch step."""
    from esn.optim import Adam
    ours, theirs = _params(), _params()
    a = Adam(ours, lr=1e-3, weight_decay=1e-4)
    b = torch.optim.Adam(theirs, lr=1e-3, weight_decay=1e-4)
    frozen = (2, 7)
    before = [ours[i].detach().clone() for i in frozen]
    for it in range(3):
        _set_grads(ours, it)
        _set_grads(theirs, it)
        for i in frozen:
            ours[i].grad = None
            theirs[i].grad = None
        a.step()
        b.step()
    _close(ours, theirs)
    for i, p0 in zip(frozen, before):
        assert torch.equal(ours[i].detach(), p0)
