"""One training iteration (`train.py:343-357`: forward, CrossEntropyLoss2d, backward, optimizer step) captured ONCE into
a CUDA graph and replayed per batch.

The eager iteration issues 400-700 kernel launches through ctypes + the autograd tape (~35 us of host time each), which
makes a step host-bound as soon as the kernels take less than ~20 ms and makes N processes on one box compete for host
cores.  A replay costs one launch: the gradient all-reduces (NCCL, side stream) and the fused Adam update are part of
the graph, so a data-parallel step has no host work besides copying the batch into the static input buffers.

    step = GraphedTrainStep(model, criterion, optimizer, images, labels)
    for images, labels in loader:
        loss = step(images, labels)        # device scalar; .item() only when you log

The optimizer must keep its state on the device: `esn.optim.Adam` (one launch per step; float learning rates written by a
host-side schedule are uploaded before each replay) or `torch.optim.Adam(..., capturable=True)`.  Dropout masks are a function
of (seed, element, step): the step counter lives on the device and is advanced inside the graph, so replays draw new
masks (esn.train.dropout).
"""
import torch

from . import ops
from .prep import bump_weights_generation


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, images, labels, autocast_dtype=torch.bfloat16, warmup=3, fuse_loss=True):
        if not images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (there is no CPU path)")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.autocast_dtype = autocast_dtype
        self.fuse_loss = fuse_loss
        self.images = images.clone()
        self.labels = labels.clone()
        self.graph = None
        self.loss = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # builds packed-weight caches, optimizer state, NCCL communicators
                self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = self._iteration()
        self.graph = g
        self.loss = loss
        self.launches = None

    def _loss(self):
        # a model may offer its head and the criterion as one fused launch (DABNet.fused_loss -> esn_bilinear_ce); it falls
        # back to criterion(model(images), labels) itself for criteria / shapes it has no fused form for
        fused = getattr(self.model, "fused_loss", None) if self.fuse_loss else None
        if fused is not None:
            return fused(self.images, self.labels, self.criterion)
        return self.criterion(self.model(self.images), self.labels)

    def _iteration(self):
        self.optimizer.zero_grad(set_to_none=True)
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                loss = self._loss()
        else:
            loss = self._loss()
        loss.backward()
        self.optimizer.step()
        ops.advance_step_counter()
        return loss.detach()

    def __call__(self, images=None, labels=None):
        if images is not None and images is not self.images:
            self.images.copy_(images, non_blocking=True)
        if labels is not None and labels is not self.labels:
            self.labels.copy_(labels, non_blocking=True)
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()        # esn.optim.Adam: float learning rates written by a host-side schedule
        self.graph.replay()
        # the replay changed weights and BN buffers behind autograd's back (no `_version` bump): invalidate every packed-
        # weight / folded-BN cache so that a following model.eval() or eager iteration rebuilds from the new values
        bump_weights_generation()
        return self.loss
