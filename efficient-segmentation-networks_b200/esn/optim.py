"""Adam as ONE kernel launch per param group -- drop-in for the ``torch.optim.Adam`` that the reference's training script
builds (train.py:212-215: ``torch.optim.Adam(params, lr, (0.9, 0.999), eps=1e-08, weight_decay=...)``) and steps once per
iteration (train.py:355).

Same constructor, same update rule (L2 weight decay folded into the gradient, bias-corrected first / second moments, no
amsgrad), same ``state_dict`` layout (per parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so checkpoints move both
ways.  What differs is where the work happens: the first and second moments of a group live in two flat fp32 buffers (the
per-parameter state entries are views), the addresses of (parameter, gradient, exp_avg, exp_avg_sq) sit in a device-side
table, and ``esn_adam_step`` (include/esn.h) updates every tensor of the group in one launch whose last CTA advances the
device-resident step counter.  A DABNet step is 1 launch of ~770 CTAs instead of 8 multi-tensor launches (0.38 ms of a
7.7 ms iteration, on the critical path behind the gradient all-reduce).

Everything the kernel reads is on the device, so the step can be captured into a CUDA graph (esn.graph.GraphedTrainStep)
as it is; ``param_group["lr"]`` may be a float (uploaded when it changes) or a device tensor that a schedule writes with
``fill_`` between replays.  There is no CPU path: parameters must be CUDA fp32 tensors.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import ops

__all__ = ["Adam"]


def build_tables(p_ptrs, g_ptrs, numels, offs, m0, v0, chunk):
    """Host side of esn_adam_step (include/esn.h): the EsnAdamTensor table (int64 [T][5]: p, g, m, v, n) and the CTA map
    (int32 [B][2]: tensor index, chunk index) -- one CTA per ``chunk`` elements of a tensor, every element exactly once.
    m0 / v0: base addresses of the flat moment buffers, offs: element offset of each tensor's slot in them."""
    tab = np.empty((len(p_ptrs), 5), dtype=np.int64)
    blocks = []
    for i, (p, g, n, o) in enumerate(zip(p_ptrs, g_ptrs, numels, offs)):
        tab[i] = (p, g, m0 + 4 * o, v0 + 4 * o, n)
        blocks.extend((i, c) for c in range((n + chunk - 1) // chunk))
    return tab, np.asarray(blocks, dtype=np.int32).reshape(-1, 2)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, foreach=None,
                 maximize=False, capturable=True, differentiable=False, fused=None):
        if amsgrad or maximize or differentiable:
            raise NotImplementedError("esn.optim.Adam: amsgrad / maximize / differentiable are not on the hot path "
                                      "(train.py:212-215 uses none of them)")
        if not isinstance(lr, torch.Tensor) and lr < 0.0:
            raise ValueError("Invalid learning rate: %r" % (lr,))
        if eps < 0.0:
            raise ValueError("Invalid epsilon value: %r" % (eps,))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: %r" % (betas,))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: %r" % (weight_decay,))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._chunk = int(L.lib.esn_adam_chunk())
        self._g = {}           # id(group) -> per-group device state (flat moments, step, lr, table)
        self._pinned = []      # host tables whose upload was captured into a CUDA graph (read again by every replay)

    # ---- per-group device state ---------------------------------------------------------------------------------
    def _group_state(self, group):
        gs = self._g.get(id(group))
        plist = [p for p in group["params"] if p.requires_grad]
        if gs is not None and gs["plist_ids"] == [id(p) for p in plist]:
            return gs
        if not plist:
            return None
        dev = plist[0].device
        for p in plist:
            ops.require_cuda(p, "esn.optim.Adam")
            if p.dtype != torch.float32 or p.device != dev or not p.is_contiguous():
                raise NotImplementedError("esn.optim.Adam: parameters of a group must be contiguous fp32 tensors on one "
                                          "CUDA device (got %s %s)" % (p.dtype, p.device))
        # 16-byte aligned slots so that whole chunks take the vector path
        offs, total = [], 0
        for p in plist:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        old = gs
        gs = dict(plist_ids=[id(p) for p in plist], plist=plist, offs=offs,
                  m=torch.zeros(total, dtype=torch.float32, device=dev), v=torch.zeros(total, dtype=torch.float32, device=dev),
                  step=torch.zeros((), dtype=torch.float32, device=dev), done=torch.zeros(1, dtype=torch.int32, device=dev),
                  lr_dev=torch.zeros((), dtype=torch.float32, device=dev), lr_host=None, key=None, table=None, blocks=None,
                  n_blocks=0)
        for p, o in zip(plist, offs):
            st = self.state[p]
            n = p.numel()
            mv, vv = gs["m"][o:o + n].view_as(p), gs["v"][o:o + n].view_as(p)
            if "exp_avg" in st:                 # state loaded from a checkpoint (or a re-grouping): keep its values
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                if old is None and "step" in st:
                    gs["step"].copy_(torch.as_tensor(st["step"], dtype=torch.float32))
            st["exp_avg"], st["exp_avg_sq"], st["step"] = mv, vv, gs["step"]
        if old is not None:
            gs["step"].copy_(old["step"])
        self._g[id(group)] = gs
        return gs

    def _tables(self, gs, plist, offs, grads):
        key = tuple(g.data_ptr() for g in grads) + tuple(p.data_ptr() for p in plist)
        if key == gs["key"]:
            return
        tab, blk = build_tables([p.data_ptr() for p in plist], [g.data_ptr() for g in grads], [p.numel() for p in plist], offs,
                                gs["m"].data_ptr(), gs["v"].data_ptr(), self._chunk)
        host = torch.empty(tab.size * 8 + blk.size * 4, dtype=torch.uint8).pin_memory()
        host[:tab.size * 8].copy_(torch.from_numpy(tab.reshape(-1).view(np.uint8)))
        host[tab.size * 8:].copy_(torch.from_numpy(blk.reshape(-1).view(np.uint8)))
        dev = host.to(gs["m"].device, non_blocking=True)
        if torch.cuda.is_current_stream_capturing():
            self._pinned.append(host)       # the captured copy node reads this buffer on every replay
        gs["table"], gs["blocks"], gs["n_blocks"], gs["key"] = dev[:tab.size * 8], dev[tab.size * 8:], int(blk.shape[0]), key

    def sync_lr(self):
        """Upload float learning rates that changed since the last upload (a scheduler writes ``param_group["lr"]`` on the
        host, train.py:356).  step() calls this; esn.graph.GraphedTrainStep calls it before every replay, so the reference's
        per-iteration schedules keep working when the step itself is a replayed CUDA graph."""
        for group in self.param_groups:
            gs = self._g.get(id(group))
            lr = group["lr"]
            if gs is None or isinstance(lr, torch.Tensor):
                continue
            lr = float(lr)
            if lr != gs["lr_host"]:
                gs["lr_dev"].fill_(lr)
                gs["lr_host"] = lr

    # ---- the step -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            self._group_state(group)
        self.sync_lr()
        for group in self.param_groups:
            gs = self._g.get(id(group))
            if gs is None:
                continue
            # parameters without a gradient take no part in this step, as in torch.optim.Adam (ERFNet's encoder.output_conv,
            # ERFNet.py:88-89); the step counter is the group's, so a parameter that only sometimes receives a gradient sees
            # the group's bias corrections rather than its own count (no such parameter on the hot-path nets)
            plist, offs, grads = [], [], []
            for p, o in zip(gs["plist"], gs["offs"]):
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse or g.dtype != torch.float32:
                    raise NotImplementedError("esn.optim.Adam: dense fp32 gradients only")
                plist.append(p)
                offs.append(o)
                grads.append(g if g.is_contiguous() else g.contiguous())
            if not plist:
                continue
            self._tables(gs, plist, offs, grads)
            lr = group["lr"]
            if isinstance(lr, torch.Tensor):
                if not (lr.is_cuda and lr.dtype == torch.float32 and lr.numel() == 1):
                    raise NotImplementedError("esn.optim.Adam: a tensor lr must be one fp32 value on the device")
                lr_ptr = lr.data_ptr()
            else:
                lr_ptr = gs["lr_dev"].data_ptr()
            b1, b2 = group["betas"]
            with torch.cuda.device(gs["m"].device):
                ops._call(L.lib.esn_adam_step, "esn_adam_step",
                          (C.c_void_p(gs["table"].data_ptr()), C.c_void_p(gs["blocks"].data_ptr()), gs["n_blocks"],
                           C.c_void_p(lr_ptr), C.c_void_p(gs["step"].data_ptr()), C.c_void_p(gs["done"].data_ptr()),
                           float(b1), float(b2), float(group["eps"]), float(group["weight_decay"])),
                          28 * sum(p.numel() for p in plist))        # p, g, m, v read; p, m, v written
            gs["keep"] = grads               # gradients made contiguous for this launch stay alive until the next one
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout.  The group's device step counter is ONE tensor here; every parameter entry gets its own
        copy (torch's foreach implementation increments each entry of the list it is handed)."""
        sd = super().state_dict()
        sd["state"] = {k: {**st, "step": st["step"].clone()} if isinstance(st.get("step"), torch.Tensor) else dict(st)
                       for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._g.clear()                     # loaded moments are copied into fresh flat buffers at the next step
        for group in self.param_groups:
            self._group_state(group)
