"""Batch-sharded data parallelism: one process per GPU, NCCL all-reduce of flat gradient buckets on a side stream --
launched behind the last backward kernel (default, see ESN_DP_DEFER below) or from inside the backward tape as soon as the
last gradient of a bucket exists, so that communication overlaps the remaining backward kernels.

Replaces the reference's single-process nn.DataParallel (train.py:166-168): no per-step parameter
broadcast (replicas stay identical by construction), no logits gather (the loss is local; only the
two scalars sum(w*nll), sum(w) are all-reduced inside CrossEntropyLoss2d), BatchNorm statistics stay
per-GPU as in the reference.  Gradients are SUMMED: every rank's loss is already divided by the
global sum of class weights.
"""
import os

import torch
import torch.distributed as dist

# ESN_DP_DEFER (default 1): all-reduce every bucket after the last backward kernel instead of from inside the backward
# (ESN_DP_DEFER=0).  The backward's BatchNorm layers are co-resident (cooperative) grids that need every SM; an NCCL kernel
# spinning on a few SMs makes them wait for it, so overlapping 3 MB of all-reduce with the backward costs more than it hides
# once the optimizer step behind it is a single launch: DABNet training on 8 B200 7.72 ms per step overlapped, 7.60 deferred
# (2 GPUs: 7.67 / 7.58; profiles/r02_allreduce_schedule.json).  With torch's 8-launch fused Adam behind the
# reductions the overlap was the faster one (8.55 against 8.73 ms on 2 GPUs), which is why it used to be the default.
DEFER_ALLREDUCE = os.environ.get("ESN_DP_DEFER", "1") == "1"
# ESN_DP_BUCKET_BYTES: size of a flat gradient bucket when data_parallel() is not given one (default 1 MB)
DEFAULT_BUCKET_BYTES = int(os.environ.get("ESN_DP_BUCKET_BYTES", str(1 << 20)))


class GradBuckets:
    """Flat fp32 buckets over the model's parameters in REVERSE registration order (the order the
    backward tape produces gradients in), ~bucket_bytes each."""

    def __init__(self, model, bucket_bytes=1 << 20, process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in model.parameters() if p.requires_grad][::-1]
        self.buckets, cur, size = [], [], 0
        for p in params:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {p: i for i, b in enumerate(self.buckets) for p in b}
        dev = params[0].device
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=dev) for b in self.buckets]
        self.views = []
        for b, f in zip(self.buckets, self.flat):
            off, vs = 0, {}
            for p in b:
                vs[p] = f[off:off + p.numel()].view(p.shape)
                off += p.numel()
            self.views.append(vs)
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.wgrad_stream = None     # set by esn.train.Tape when it runs weight gradients off the critical path
        self.reset()

    def reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.seen = set()
        self.works = []
        self.staged = [dict() for _ in self.buckets]     # parameter -> gradient waiting to be packed
        self.deferred = []

    def grad_ready(self, p, g):
        """Called by the tape when parameter p's gradient is final.  The copy into the flat bucket is deferred until
        the bucket is complete and then done with ONE multi-tensor copy for all dense gradients (a few hundred tiny
        copy kernels per step otherwise), followed by the bucket's all-reduce.  Returns the bucket view (the gradient;
        valid once the tape's backward has finished)."""
        i = self.bucket_of[p]
        view = self.views[i][p]
        self.staged[i][p] = g          # a second contribution replaces the first (the tape has already summed them)
        if p not in self.seen:
            self.seen.add(p)
            self.pending[i] -= 1
            if self.pending[i] == 0:
                if DEFER_ALLREDUCE:
                    self.deferred.append(i)
                else:
                    self._pack(i)
                    self._launch(i)
        elif self.pending[i] == 0:        # bucket already reduced: late extra contribution (not on the hot-path nets)
            raise RuntimeError("gradient for %r arrived after its bucket was all-reduced" % (tuple(p.shape),))
        return view

    def _pack(self, i):
        if self.wgrad_stream is not None:       # the tape computes weight gradients on a side stream: join it before reading them
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)
        dense_v, dense_g = [], []
        for p, g in self.staged[i].items():
            v = self.views[i][p]
            if g is v:
                continue
            if g.dtype == v.dtype and g.is_contiguous() and g.shape == v.shape:
                dense_v.append(v)
                dense_g.append(g)
            else:
                v.copy_(g)
        if dense_v:
            torch._foreach_copy_(dense_v, dense_g)
        self.staged[i] = dict()

    def _launch(self, i):
        if self.world == 1:
            return
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self.works.append(dist.all_reduce(self.flat[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.works.append(dist.all_reduce(self.flat[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Parameters that produced no gradient this step count as zeros; wait for all reductions."""
        for i in self.deferred:
            self._pack(i)
            self._launch(i)
        self.deferred = []
        for i, b in enumerate(self.buckets):
            if self.pending[i] > 0:
                for p in b:
                    if p not in self.seen:
                        self.views[i][p].zero_()
                self._pack(i)
                self.pending[i] = 0
                self._launch(i)
        for w in self.works:
            w.wait()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.reset()


_ACTIVE = False


def is_active():
    """True once data_parallel() attached SUM-reduced gradient buckets over more than one rank in this process;
    CrossEntropyLoss2d(distributed=None) then normalises by the global sum of class weights."""
    return _ACTIVE


def data_parallel(model, bucket_bytes=None, process_group=None):
    """Attach gradient buckets to a model built by build_model(); its train-mode forward/backward then
    all-reduces gradients across the process group.  Parameters are broadcast once from rank 0."""
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0, group=process_group)
    model.__dict__["_esn_buckets"] = GradBuckets(model, bucket_bytes or DEFAULT_BUCKET_BYTES, process_group)
    global _ACTIVE
    _ACTIVE = _ACTIVE or model.__dict__["_esn_buckets"].world > 1
    return model
