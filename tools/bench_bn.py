#!/usr/bin/env python
"""Micro-benchmark of the train-mode BatchNorm layer kernels (one-launch cooperative kernels of esn_bn_fused.cu vs the
multi-launch path) on DABNet's tensor shapes, each timed as 20 back-to-back layers replayed from ONE CUDA graph (no host
launch cost), CUDA events.  ESN_BN_FWD_VARIANT / ESN_BN_BWD_VARIANT pick the kernel variant (channels per thread, loads in flight, CTAs per SM).

    python tools/bench_bn.py [out.json]
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")]
from esn import _lib as L, ops  # noqa: E402

SHAPES = [(8, 64, 128, 64), (8, 128, 256, 32), (8, 64, 128, 128), (8, 128, 256, 64), (8, 256, 512, 32)]   # n, h, w, c
LAYERS = 20
dev = "cuda"


def build(shape, mode):
    n, h, w, c = shape
    xs = [ops.new_act(n, c, h, w, torch.bfloat16, dev).normal_() for _ in range(LAYERS)]
    dys = [ops.new_act(n, c, h, w, torch.bfloat16, dev).normal_() for _ in range(LAYERS)]
    ys = [ops.new_act(n, c, h, w, torch.bfloat16, dev) for _ in range(LAYERS)]
    scratch = torch.zeros(LAYERS, 2, 8 * 3 * c + 1, dtype=torch.float64, device=dev)      # ESN_BN_FUSED_REPLICAS = 8
    par = torch.ones(LAYERS, 10, c, dtype=torch.float32, device=dev)
    keep = []

    def run():
        scratch.zero_()
        for i in range(LAYERS):
            gamma, beta, alpha, rm, rv, scale, shift, mean, invstd, dpar = (par[i, k] for k in range(10))
            f = L.EsnBnFinalize()
            f.sums, f.count = scratch[i, 0].data_ptr(), n * h * w
            f.gamma, f.beta, f.eps, f.momentum = gamma.data_ptr(), beta.data_ptr(), 1e-3, 0.1
            f.running_mean, f.running_var = rm.data_ptr(), rv.data_ptr()
            f.scale, f.shift, f.mean, f.invstd = scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr()
            f.channels = c
            if mode == "fwd_fused":
                q = L.EsnBnTrainFwd()
                q.x, q.y, q.fin = ops.tdesc(xs[i]), ops.tdesc(ys[i]), f
                q.alpha, q.barrier, q.act = alpha.data_ptr(), scratch[i, 0].data_ptr() + 8 * 8 * 2 * c, L.ACT_PRELU
                L.check(L.lib.esn_bn_act_train_fwd(C.byref(q), ops.stream()), "fwd")
            elif mode == "fwd_multi":
                d = ops.tdesc(xs[i])
                L.check(L.lib.esn_channel_stats(C.byref(d), C.c_void_p(scratch[i, 0].data_ptr()), 1, ops.stream()), "stats")
                L.check(L.lib.esn_bn_finalize(C.byref(f), ops.stream()), "fin")
                ops.affine_act(xs[i], scale, shift, alpha, L.ACT_PRELU, out=ys[i])
            else:
                p = L.EsnBnBwd()
                p.x, p.dy, p.dx = ops.tdesc(xs[i]), ops.tdesc(dys[i]), ops.tdesc(ys[i])
                p.scale, p.shift, p.alpha, p.mean, p.invstd = (t.data_ptr() for t in (scale, shift, alpha, mean, invstd))
                p.sums = scratch[i, 1].data_ptr()
                p.dgamma = p.dbeta = p.dalpha = dpar.data_ptr()
                p.act, p.train_stats = L.ACT_PRELU, 1
                if mode == "bwd_fused":
                    L.check(L.lib.esn_bn_act_bwd_fused(C.byref(p), C.c_void_p(scratch[i, 1].data_ptr() + 8 * 8 * 3 * c), ops.stream()), "bwd")
                else:
                    L.check(L.lib.esn_bn_act_bwd_reduce(C.byref(p), ops.stream()), "red")
                    L.check(L.lib.esn_bn_act_bwd_apply(C.byref(p), ops.stream()), "app")
    keep.append((xs, dys, ys, scratch, par))
    return run, keep


def time_graph(run):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            run()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 / LAYERS * 1e3      # us per layer


res = {}
dbg = os.environ.get("ESN_BN_FWD_VARIANT", "0") + "/" + os.environ.get("ESN_BN_BWD_VARIANT", "0")
for shape in SHAPES:
    for mode in (("fwd_fused", "bwd_fused") if os.environ.get("BN_ONLY_FUSED") else ("fwd_fused", "fwd_multi", "bwd_fused", "bwd_multi")):
        run, keep = build(shape, mode)
        us = time_graph(run)
        n, h, w, c = shape
        mb = n * h * w * c * 2 / 1e6
        res["%dx%dx%dx%d %s" % (n, h, w, c, mode)] = round(us, 2)
        print("dbg=%s %-18s %-10s %7.2f us/layer  (tensor %.1f MB)" % (dbg, "x".join(map(str, shape)), mode, us, mb), flush=True)
        del run, keep
        torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump({"dbg": dbg, "us_per_layer": res}, open(sys.argv[1], "w"), indent=1)
