#!/usr/bin/env python
"""bench.py -- images/s of the segmentation hot path on B200 (driver contract in the task brief).

Default workload = BASELINE.json configs[2]: DABNet (19 classes), bf16 training, batch 8 per GPU, 512x1024 (the metric's
resolution and the only config that names 1/2/4/8 GPUs), weighted cross-entropy, Adam, data parallel: a "step" is one
training iteration (forward, loss, backward, bucketed NCCL gradient all-reduce, fused Adam) of one batch per GPU, so the
driver's --gpus N runs exercise the path that communicates ("weak" scaling: 8 images per GPU).  The same JSON line carries,
under `legs`, BASELINE.json configs[1] -- ERFNet bf16 inference, batch 16 per GPU, 3x1024x2048, fused argmax head (images
sharded over ranks, no collective) -- measured the same way; any workload can be made the primary with --workload.

    python bench.py                                   # N=1
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference                  # the reference's CPU path (oracle port) on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "efficient-segmentation-networks_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (model, batch per GPU, H, W, mode)
    "erfnet_infer_bf16_b16_1024x2048": ("ERFNet", 16, 1024, 2048, "infer"),
    "dabnet_infer_bf16_b16_1024x2048": ("DABNet", 16, 1024, 2048, "infer"),
    "erfnet_infer_bf16_b16_512x1024": ("ERFNet", 16, 512, 1024, "infer"),
    "dabnet_infer_bf16_b16_512x1024": ("DABNet", 16, 512, 1024, "infer"),
    # BASELINE.json configs[3]: ENet + CGNet inference, batch 32, 1024x2048
    "enet_infer_bf16_b32_1024x2048": ("ENet", 32, 1024, 2048, "infer"),
    "cgnet_infer_bf16_b32_1024x2048": ("CGNet", 32, 1024, 2048, "infer"),
    "fastscnn_infer_bf16_b16_1024x2048": ("FastSCNN", 16, 1024, 2048, "infer"),
    "espnet_infer_bf16_b16_1024x2048": ("ESPNet", 16, 1024, 2048, "infer"),
    "espnetv2_infer_bf16_b16_1024x2048": ("ESPNet_v2", 16, 1024, 2048, "infer"),
    # the nets that reuse the ERFNet / Fast-SCNN kernels
    "esnet_infer_bf16_b16_1024x2048": ("ESNet", 16, 1024, 2048, "infer"),
    "contextnet_infer_bf16_b16_1024x2048": ("ContextNet", 16, 1024, 2048, "infer"),
    "edanet_infer_bf16_b16_1024x2048": ("EDANet", 16, 1024, 2048, "infer"),
    "lednet_infer_bf16_b16_1024x2048": ("LEDNet", 16, 1024, 2048, "infer"),
    # BASELINE.json configs[2]: DABNet bf16 training, batch 8/GPU, 512x1024, weighted CE, Adam, data parallel
    "dabnet_train_bf16_b8_512x1024": ("DABNet", 8, 512, 1024, "train"),
    "erfnet_train_bf16_b8_512x1024": ("ERFNet", 8, 512, 1024, "train"),
    # BASELINE.json configs[4] (first half): Fast-SCNN bf16 training, batch 16/GPU, 1024x2048
    "fastscnn_train_bf16_b16_1024x2048": ("FastSCNN", 16, 1024, 2048, "train"),
    "espnetv2_train_bf16_b16_1024x2048": ("ESPNet_v2", 16, 1024, 2048, "train"),
}
# block-fused algorithmic elements per input pixel (forward), bf16 storage; the
# logits term (19 elements/pixel) is replaced by the 1-byte argmax mask because the head is fused.
ALG_ELEMS_PER_PX = {"ERFNet": 162.0, "DABNet": 189.5, "ENet": 178.0, "CGNet": 234.3, "FastSCNN": 54.6,
                    "ESPNet": 141.2, "ESPNet_v2": 186.9,
                    "ESNet": 142.0, "ContextNet": 67.2, "EDANet": 231.7, "LEDNet": 146.6}       # tools/probe_alg_elems.py (reproduces 162.0 / 54.6 for ERFNet / FastSCNN)
GMAC_512x1024 = {"ERFNet": 26.60, "DABNet": 10.22, "ENet": 4.12, "CGNet": 6.76, "FastSCNN": 1.68, "ESPNet": 3.36,
                 "ESPNet_v2": 5.65, "ESNet": 24.05, "ContextNet": 1.68, "EDANet": 8.83, "LEDNet": 11.20}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def fixture_state_dict(name):
    from oracle import fixture
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))
    proto = {k: torch.empty(shape, dtype=getattr(torch, dt.split(".")[1])) for k, shape, dt in spec[name]["keys"]}
    return fixture.randomize_state_dict(proto, 1234)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_leg(model, h, w, steps, warmup, budget_s=25.0, train=False):
    """The reference's CPU path (oracle port of model/*.py, fp32, all host threads) on a bounded
    sample of the workload: one image of the workload's resolution per step (inference: forward +
    numpy argmax; training: train-mode forward + weighted CE + backward through torch autograd)."""
    from oracle import fixture, nets, loss as oloss
    try:        # the GPU arm pins the process next to its GPU (numa_local_affinity); the CPU reference gets every core back
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
        threads = len(os.sched_getaffinity(0))
    except Exception:      # noqa: BLE001
        threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = fixture_state_dict(model)
    x = fixture.make_input(1, h, w)
    times = []
    if train:
        sd = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
        lab = fixture.make_labels(1, h, w, 19)
        wt = torch.tensor(fixture.CLASS_WEIGHTS)
    with torch.set_grad_enabled(train):
        t_start = time.perf_counter()
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            if train:
                l, _, _ = oloss.weighted_ce(nets.forward(model, sd, x, train=True), lab, wt)
                l.backward()
                for v in sd.values():
                    v.grad = None
            else:
                y = nets.forward(model, sd, x)
                nets.argmax_mask(y)
            t1 = time.perf_counter()
            if i >= warmup:
                times.append(t1 - t0)
            if i >= warmup and (t1 - t_start) > budget_s and len(times) >= 2:
                break
    mean = sum(times) / len(times)
    return {"value": 1.0 / mean, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "%d x (1 image 3x%dx%d fp32, %s), oracle/nets.py on %d host threads"
                      % (len(times), h, w, "train-mode forward + weighted CE + backward" if train else "forward + numpy argmax",
                         threads)}, mean, len(times)


def gpu_eager_reference_leg(model, batch, h, w, train, warmup=10, steps=20, variants=None, seed=1234):
    """The number to beat: the reference's op-by-op PyTorch graph run EAGERLY ON THE SAME GPU with
    cuDNN (`cudnn.benchmark=True`), protocol of tools/fps_test/eval_forward_time.py:9-34 (warm-up, synchronise, timed loop,
    synchronise).  /root/reference does not exist on the GPU box, so the graph is the oracle's functional restatement of the
    same modules (oracle/nets.py) with BatchNorm / PReLU going through the reference's own fused ATen calls.  Variants:
    fp32 with TF32 off / on (torch's conv default is TF32 on: the reference's literal protocol), bf16 autocast, bf16
    autocast + channels_last, pure bf16 weights + channels_last.  None of this repo's kernels run here."""
    from oracle import fixture, nets
    nets.REFERENCE_ATEN_CALLS = True
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    sd0 = fixture_state_dict(model)
    x0 = fixture.make_input(batch, h, w, seed=seed)
    lab = fixture.make_labels(batch, h, w, 19, seed=seed).cuda() if train else None
    wt = torch.tensor(fixture.CLASS_WEIGHTS).cuda()
    table = {"fp32": (False, None, False, False), "fp32_tf32": (True, None, False, False),
             "bf16_autocast": (True, torch.bfloat16, False, False),
             "bf16_autocast_channels_last": (True, torch.bfloat16, True, False),
             "bf16_weights_channels_last": (True, None, True, True)}
    out = {}
    for name in variants or ["fp32_tf32", "bf16_autocast", "bf16_autocast_channels_last", "bf16_weights_channels_last"]:
        tf32, ac, cl, pure = table[name]
        if pure and train:
            continue                      # pure-bf16 master weights is not a training configuration anyone runs
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            dt = torch.bfloat16 if pure else torch.float32
            sd = {}
            for k, v in sd0.items():
                v = v.cuda()
                if v.is_floating_point():
                    v = v.to(dt)
                    if cl and v.dim() == 4:
                        v = v.contiguous(memory_format=torch.channels_last)
                    if train:
                        v.requires_grad_(True)
                sd[k] = v
            x = x0.cuda().to(dt)
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
            opt = None
            if train:
                params = [v for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
                opt = torch.optim.Adam(params, lr=5e-4, weight_decay=1e-4, fused=True)

            def one():
                if train:
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=ac, enabled=ac is not None):
                        logits = nets.forward(model, sd, x, train=True)
                    l = torch.nn.functional.cross_entropy(logits.float(), lab, wt, ignore_index=255)     # loss.py:23-32
                    l.backward()
                    opt.step()
                    return l
                with torch.no_grad(), torch.autocast("cuda", dtype=ac, enabled=ac is not None):
                    return nets.forward(model, sd, x).argmax(1).to(torch.uint8)      # argmax on the device (kinder than test.py:79-82)
            for _ in range(warmup):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"images_per_s": round(batch / (ms / 1e3), 2), "ms_per_step": round(ms, 3)}
            del sd, x, opt
        except Exception as exc:      # noqa: BLE001
            out[name] = {"error": repr(exc)[:200]}
        torch.cuda.empty_cache()
    nets.REFERENCE_ATEN_CALLS = False
    torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    ok = {k: v for k, v in out.items() if "images_per_s" in v}
    best = max(ok, key=lambda k: ok[k]["images_per_s"]) if ok else None
    return {"variants": out, "fastest": best, "value": ok[best]["images_per_s"] if best else None, "unit": "images/s",
            "batch": batch, "warmup": warmup, "steps": steps,
            "what": "oracle/nets.py graph (= the reference's modules op by op) eager on this GPU, cudnn.benchmark=True, "
                    "protocol of tools/fps_test/eval_forward_time.py:9-34; %s" %
                    ("train-mode forward + weighted CE + backward + fused Adam" if train else "forward + device argmax")}


# SURVEY 8(a) block units: class names whose OUTERMOST instances are the units of SURVEY 8(d)'s algorithmic-byte count
# (tensors entering + leaving the unit, each once).  Launches outside any unit (heads, input-injection pyramid, stand-alone
# BN/PReLU passes on concat tensors) are their own units.
UNIT_CLASSES = {
    "ERFNet": ("DownsamplerBlock", "non_bottleneck_1d", "UpsamplerBlock"),
    "DABNet": ("DABModule", "DownSamplingBlock"),
    "ENet": ("InitialBlock", "RegularBottleneck", "DownsamplingBottleneck", "UpsamplingBottleneck"),
    "CGNet": ("ContextGuidedBlock", "ContextGuidedBlock_Down"),
    "FastSCNN": ("_ConvBNReLU", "_DSConv", "LinearBottleneck", "PyramidPooling", "FeatureFusionModule", "Classifer"),
    "ESPNet": ("DownSamplerB", "DilatedParllelResidualBlockB"),
    "ESPNet_v2": ("EESP", "DownSampler", "PSPModule"),
    "ESNet": ("DownsamplerBlock", "FCU", "PFCU", "UpsamplerBlock"),
    "ContextNet": ("Custom_Conv", "DepthSepConv", "LinearBottleneck", "FeatureFusionModule", "Classifer"),
    "EDANet": ("DownsamplerBlock", "EDAModule"),
    "LEDNet": ("DownsamplerBlock", "SS_nbt_module_paper", "APNModule"),
}


def numa_local_affinity(local_rank):
    """Pin this process to the CPUs next to its GPU BEFORE any pinned host buffer is allocated, so that the pages the H2D
    copies read are first-touched on the GPU's NUMA node (round 1: every rank allocated on node 0 and the 8-GPU e2e figure
    was bound by one socket's memory).  Best effort: returns a short description for the JSON line."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = open(base + "/numa_node").read().strip()
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus), "pci": bdf}
        return {"numa_node": node, "cpus": 0, "pci": bdf, "note": "no usable CPU in local_cpulist; affinity unchanged"}
    except Exception as exc:      # noqa: BLE001
        return {"error": repr(exc)[:120]}


def unit_roofline(model, model_name, prof_step, hbm_peak, tc_peak):
    """per block unit, algorithmic bytes = the tensors entering + leaving the unit (each once), FLOPs = 2 x the
    MACs of its convs; t_HBM = bytes / measured HBM peak, t_TC = FLOPs / measured bf16 peak; the binding roof is the larger,
    frac = t_binding / measured time.  Unit boundaries come from forward hooks on the outermost block modules; times are CUDA
    events around every C-ABI launch inside them (one instrumented eager step)."""
    from esn import ops
    names = UNIT_CLASSES.get(model_name, ())
    state = {"depth": 0, "cur": None}
    handles = []

    def numel_bytes(t):
        if torch.is_tensor(t):
            return t.numel() * t.element_size() if t.is_floating_point() else 0
        if isinstance(t, (tuple, list)):
            return sum(numel_bytes(v) for v in t)
        return 0

    def first_tensor(t):
        # a block that also emits the next block's BNPReLU returns (output, BNPReLU(output)) or (None, BNPReLU(output)):
        # the unit's logical output is one tensor of the block's shape (SURVEY 8d counts it once)
        if isinstance(t, (tuple, list)):
            for v in t:
                r = first_tensor(v)
                if r is not None:
                    return r
            return None
        return t if torch.is_tensor(t) else None

    def shape_of(t):
        t = first_tensor(t)
        return tuple(t.shape) if t is not None else ()

    units = []

    def pre(mod, inp):
        if state["depth"] == 0:
            state["cur"] = {"kind": type(mod).__name__, "first": len(ops.PROFILE), "in": numel_bytes(inp), "shape": shape_of(inp)}
        state["depth"] += 1

    def post(mod, inp, out):
        state["depth"] -= 1
        if state["depth"] == 0 and state["cur"] is not None:
            u = state["cur"]
            u["last"] = len(ops.PROFILE)
            u["out"] = numel_bytes(first_tensor(out))
            u["oshape"] = shape_of(out)
            units.append(u)
            state["cur"] = None
    for mod in model.modules():
        if type(mod).__name__ in names:
            handles.append(mod.register_forward_pre_hook(pre))
            handles.append(mod.register_forward_hook(post))
    try:
        for _ in range(2):
            del units[:]
            ops.PROFILE = []
            prof_step()
            torch.cuda.synchronize()
            prof, ops.PROFILE = ops.PROFILE, None
    finally:
        ops.PROFILE = None
        for hd in handles:
            hd.remove()
    owner = [None] * len(prof)
    for ui, u in enumerate(units):
        for i in range(u["first"], u["last"]):
            owner[i] = ui
    agg = {}

    def add(key, ms, nbytes, flops, launches, count):
        a = agg.setdefault(key, {"units": 0, "launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
        a["units"] += count
        a["launches"] += launches
        a["ms"] += ms
        a["bytes"] += nbytes
        a["flops"] += flops
    for ui, u in enumerate(units):
        rs = prof[u["first"]:u["last"]]
        ms = sum(r["ev"][0].elapsed_time(r["ev"][1]) for r in rs)
        s = u["shape"]
        key = "%s c%d->%d @%dx%d" % (u["kind"], s[1], u["oshape"][1], s[2], s[3]) if len(s) == 4 else u["kind"]
        add(key, ms, u["in"] + u["out"], sum(r["flops"] for r in rs), len(rs), 1)
    for i, r in enumerate(prof):
        if owner[i] is None:
            add("(top level) " + r["kernel"], r["ev"][0].elapsed_time(r["ev"][1]), r["bytes"], r["flops"], 1, 1)
    tot = sum(a["ms"] for a in agg.values())
    table = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        t_hbm = a["bytes"] / (hbm_peak * 1e9) * 1e3
        t_tc = a["flops"] / (tc_peak * 1e12) * 1e3
        bound = "tensor" if t_tc > t_hbm else "hbm"
        table[k] = {"units": a["units"], "launches": a["launches"], "ms": round(a["ms"], 4), "share": round(a["ms"] / tot, 4),
                    "alg_bytes": a["bytes"], "flops": a["flops"], "bound": bound,
                    "GBps": round(a["bytes"] / a["ms"] / 1e6, 1), "TFLOPs": round(a["flops"] / a["ms"] / 1e9, 1),
                    "hbm_frac": round(t_hbm / a["ms"], 4), "tensor_frac": round(t_tc / a["ms"], 4),
                    "frac": round(max(t_hbm, t_tc) / a["ms"], 4)}
    return prof, table, tot


def kernel_tables(prof):
    agg = {}
    for r in prof:
        ms = r["ev"][0].elapsed_time(r["ev"][1])
        a = agg.setdefault(r["kernel"], {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
        a["launches"] += 1
        a["ms"] += ms
        a["bytes"] += r["bytes"]
        a["flops"] += r["flops"]
    tot_ms = sum(a["ms"] for a in agg.values())
    kernels = {k: {"launches": a["launches"], "ms": round(a["ms"], 4), "share": round(a["ms"] / tot_ms, 4),
                   "alg_GBps": round(a["bytes"] / a["ms"] / 1e6, 1), "TFLOPs": round(a["flops"] / a["ms"] / 1e9, 2)}
               for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    by_tag = {}
    for r in prof:
        ms = r["ev"][0].elapsed_time(r["ev"][1])
        a = by_tag.setdefault(r["kernel"].replace("esn_", "") + " " + r["tag"], [0, 0.0, 0, 0])
        a[0] += 1
        a[1] += ms
        a[2] += r["bytes"]
        a[3] += r["flops"]
    layers = {k: {"n": v[0], "ms": round(v[1], 3), "alg_GBps": round(v[2] / v[1] / 1e6, 1), "TFLOPs": round(v[3] / v[1] / 1e9, 1)}
              for k, v in sorted(by_tag.items(), key=lambda kv: -kv[1][1])[:24]}
    return agg, kernels, layers, tot_ms


# C-ABI entry point -> substrings of the CUDA kernels it launches (for the in-graph kernel table of training workloads)
ENTRY_KERNELS = {
    "esn_bn_act_bwd_fused": ("bn_act_bwd_fused_kernel",), "esn_bn_act_train_fwd": ("bn_act_train_fwd_kernel",),
    "esn_conv2d_wgrad": ("wgrad_",), "esn_conv2d_umma": ("conv_umma_kernel",), "esn_conv_pair_umma": ("conv_pair_kernel",),
    "esn_conv2d_direct": ("conv_direct_kernel", "dw_strip_kernel", "dwconv_kernel"), "esn_weighted_ce": ("weighted_ce_kernel",),
    "esn_bilinear_bwd": ("bilinear_bwd_rows_kernel", "bilinear_bwd_kernel"), "esn_bilinear_bwd_nhwc": ("bilinear_bwd2_kernel",),
    "esn_bn_act_bwd_apply": ("bn_act_bwd_apply",), "esn_bn_act_bwd_reduce": ("bn_act_bwd_reduce",),
    "esn_channel_stats": ("channel_stats",), "esn_act_bwd": ("act_bwd_kernel",), "esn_scale_nc": ("scale_nc_kernel",),
    "esn_dropout": ("dropout_kernel",), "esn_head_bilinear": ("bilinear_head",), "esn_maxpool2x2_bwd": ("maxpool2x2_bwd",),
    "esn_affine_act": ("pw_vec_kernel", "pw_kernel"), "esn_stem_conv3x3s2": ("stem_",),
    "esn_bilinear_ce": ("bilinear_ce_kernel",), "esn_adam_step": ("adam_table_kernel",),
}


def graph_kernel_table(graph, agg, step_ms, reps=3):
    """Device durations of the kernels INSIDE the replayed CUDA graph (CUPTI activity records through torch.profiler, taken
    after the timed region; no kernel replay, caches warm).  CUDA events cannot bracket kernels inside a graph, and the
    events-around-eager-launches table over-states short kernels by the launch latency, so this is the table the training
    roofline is read from.  Returns (per-entry table, per-kernel table, summary)."""
    from collections import defaultdict
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            graph.replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    per = defaultdict(lambda: [0, 0.0])
    for e in ev:
        name = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("at::native::", "").split("(")[0][:96]
        per[name][0] += 1
        per[name][1] += e.time_range.elapsed_us()
    kernels = {k: {"launches": round(v[0] / reps, 1), "ms": round(v[1] / reps / 1e3, 4)}
               for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:40]}
    entries = {}
    for entry, subs in ENTRY_KERNELS.items():
        if entry not in agg:
            continue
        ms = sum(v[1] for k, v in per.items() if any(s_ in k for s_ in subs)) / reps / 1e3
        if ms <= 0:
            continue
        entries[entry] = {"launches": agg[entry]["launches"], "ms": round(ms, 4), "share_of_step": round(ms / step_ms, 4),
                          "alg_GBps": round(agg[entry]["bytes"] / ms / 1e6, 1), "TFLOPs": round(agg[entry]["flops"] / ms / 1e9, 2)}
    entries = dict(sorted(entries.items(), key=lambda kv: -kv[1]["ms"]))
    total = sum(v[1] for v in per.values()) / reps / 1e3
    return entries, kernels, {"sum_of_kernel_ms": round(total, 3), "kernels_per_step": round(len(ev) / reps, 1),
                              "note": "kernels on the side streams (weight gradients, NCCL) overlap the main chain, so the sum can exceed the step"}


def ncu_traffic(workload, kernel):
    """Per-launch DRAM bytes (read + write) of `kernel` from the committed ncu launch list of the same command, or None."""
    fam_names = {"esn_conv2d_umma": "conv_umma_kernel", "esn_conv_pair_umma": "conv_pair_kernel", "esn_conv2d_direct": "conv_direct_kernel",
                 "esn_dab_dw_pair": "dab_dw_pair", "esn_affine_act": "pw_", "esn_conv2d_wgrad": "wgrad", "esn_nb1d_umma": "nb1d_kernel",
                 "esn_dw_conv": "dw_", "esn_dwconv": "dw", "esn_bn_act_bwd_fused": "bn_act_bwd_fused_kernel",
                 "esn_bn_act_train_fwd": "bn_act_train_fwd_kernel"}
    for rnd in ("r02", "r01"):
        for stem in (workload, workload.split("_")[0] + ("_train" if "_train_" in workload else "")):
            tpath = os.path.join(ROOT, "profiles", "%s_traffic_%s.json" % (rnd, stem))
            if not os.path.exists(tpath) or kernel not in fam_names:
                continue
            tj = json.load(open(tpath))
            if tj.get("workload") != workload:
                continue
            fams = [v for k, v in tj["families"].items() if fam_names[kernel] in k]
            n_l = sum(v["launches"] for v in fams)
            if n_l:
                return {"dram_bytes_per_launch": int(sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in fams) / n_l),
                        "kernel_launches": n_l, "source": tj["source"]}
    return None


def measure(args, wl_name, rank, world, local_rank, dist, primary=True):
    """One workload -> the JSON line's fields (rank 0) or None (other ranks)."""
    from builders.model_builder import build_model
    from esn import ops
    from oracle import fixture
    model_name, batch, H, W, mode = WORKLOADS[wl_name]
    train = mode == "train"
    u8_in = args.e2e_input == "u8"
    config = {"workload": wl_name, "net": model_name, "classes": 19, "batch_per_gpu": batch,
              "input": "3x%dx%d fp32 NCHW" % (H, W),
              "mode": "training step: forward + weighted CE + backward + Adam" if train else "inference",
              "head": ("head + loss: the model's fused close where it has one (DABNet: esn_bilinear_ce = bilinear up-sampling + "
                       "weighted CE + both gradients in one launch), else head kernel -> fp32 logits -> weighted-CE kernel"
                       if not args.no_fused_loss else "bilinear -> fp32 logits -> fused weighted-CE kernel")
              if train else "argmax fused (uint8 mask)",
              "optimizer": ("torch.optim.Adam(fused=True)" if args.torch_adam else "esn.optim.Adam (one esn_adam_step launch)")
              if train else None,
              "sharding": ("data parallel: flat fp32 gradient buckets all-reduced (NCCL) from inside the backward tape, "
                           "2-scalar loss all-reduce, per-GPU BatchNorm") if train else "images across ranks, no collective",
              "l2": "no flush: per-step working set (input %.0f MB + activations) >> 126 MB L2" % (batch * 3 * H * W * 4 / 1e6)}
    m = build_model(model_name, 19)
    m.load_state_dict(fixture_state_dict(model_name))
    m = m.cuda()
    x_host = fixture.make_input(batch, H, W, seed=1234 + rank).pin_memory()
    x = x_host.cuda(non_blocking=True)
    y_host = y = None
    if train:
        from esn import parallel
        from utils.losses.loss import CrossEntropyLoss2d
        m.train()
        if world > 1:
            parallel.data_parallel(m)
        crit = CrossEntropyLoss2d(weight=torch.tensor(fixture.CLASS_WEIGHTS), ignore_label=255).cuda()
        if args.torch_adam:
            opt = torch.optim.Adam(m.parameters(), lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, fused=True,
                                   capturable=not args.no_graph)
        else:
            from esn.optim import Adam      # train.py:212-215's torch.optim.Adam as one esn_adam_step launch
            opt = Adam(m.parameters(), lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
        y_host = fixture.make_labels(batch, H, W, 19, seed=1234 + rank).pin_memory()
        y = y_host.cuda(non_blocking=True)
    else:
        m.eval()
    torch.cuda.synchronize()

    def step(inp, lab=None):
        if train:
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = crit(m(inp), y if lab is None else lab)
            loss.backward()
            opt.step()
            return loss.detach()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return m.predict_mask(inp)

    # ---- warm-up (also builds the packed-weight caches), then CUDA-graph capture
    for _ in range(max(args.warmup, 3)):
        mask = step(x)
    torch.cuda.synchronize()
    ops.launch_count_reset()
    step(x)
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    graph = gstep = None
    if train and not args.no_graph:
        # the whole iteration (forward, loss, backward, gradient all-reduce, Adam) as one CUDA graph (esn/graph.py)
        from esn.graph import GraphedTrainStep
        gstep = GraphedTrainStep(m, crit, opt, x, y, warmup=1, fuse_loss=not args.no_fused_loss)
        graph = gstep.graph
        mask = gstep.loss
    elif not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step(x)
            with torch.cuda.graph(graph, stream=s):
                mask = step(x)
        torch.cuda.current_stream().wait_stream(s)
    if graph is not None:
        torch.cuda.synchronize()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
            return mask
        return step(x)

    # ---- timed region: K steps, CUDA events on the launching stream, barrier + sync both sides
    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
    except Exception:
        gpu_id = str(local_rank)
    sampler = ClockSampler(gpu_id)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([elapsed_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = t.item()
    ms_per_step = elapsed_ms / args.steps
    value = world * batch / (ms_per_step / 1e3)

    # ---- e2e through the public API with HOST buffers: every step copies that step's input from pinned host memory and
    # reads the result back.  Inference (default --e2e-input u8): decoded uint8 HWC BGR images, as cv2.imread hands them to
    # the reference's dataset class, -> H2D (copy stream) -> esn_image_u8hwc_to_f32nchw (the dataset class's arithmetic
    # tail, dataset/cityscapes.py:74-78, on the device) -> model.predict_mask (the captured graph) -> uint8 masks -> D2H
    # (third stream).  --e2e-input f32 ships the reference's pre-processed fp32 NCHW batch instead (4x the bytes).
    mean_bgr = [72.3924, 82.90902, 73.158325]      # dataset/inform/cityscapes_inform.pkl['mean'] (BGR, fp32)
    copy_s, d2h_s, main_s = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    if u8_in:
        g8 = torch.Generator().manual_seed(1234 + rank)
        in_host = torch.randint(0, 256, (batch, H, W, 3), dtype=torch.uint8, generator=g8).pin_memory()
        in_dev = [torch.empty((batch, H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    else:
        in_host = x_host
        in_dev = [torch.empty_like(x) for _ in range(2)]
    # training labels: the uint8 label map the dataset class decodes (dataset/cityscapes.py:70) when the images travel as uint8,
    # cast to int64 on the device (train.py:349 does `labels.long()` on the host and ships 8 bytes per pixel); int64 otherwise
    lab_host = (y_host.to(torch.uint8).pin_memory() if u8_in else y_host) if train else None
    lab_dev = [torch.empty_like(y, dtype=lab_host.dtype) for _ in range(2)] if train else [None, None]
    h2d = in_host.numel() * in_host.element_size() + (lab_host.numel() * lab_host.element_size() if train else 0)
    d2h = 4 if train else batch * H * W
    out_host = [(torch.empty((), dtype=torch.float32) if train else torch.empty((batch, H, W), dtype=torch.uint8)).pin_memory()
                for _ in range(2)]
    out_stage = [torch.empty_like(mask) for _ in range(2)]
    exact = None
    if u8_in:       # self-check of the device pre-processing: bit-equal to the same fp32 arithmetic done by torch
        in_dev[0].copy_(in_host)
        sub = in_dev[0][:2]
        got = ops.image_u8_to_f32(sub, mean_bgr, True)
        want = (sub.float() - torch.tensor(mean_bgr, device="cuda")).flip(3).permute(0, 3, 1, 2)
        exact = bool(torch.equal(got, want))
        del got, want

    def e2e_loop(k):
        ready, freed, read = [None, None], [None, None], [None, None]
        for i in range(k + 1):
            b = i & 1
            if i < k:
                with torch.cuda.stream(copy_s):
                    if freed[b] is not None:
                        copy_s.wait_event(freed[b])      # input buffer b was consumed by its step
                    in_dev[b].copy_(in_host, non_blocking=True)
                    if train:
                        lab_dev[b].copy_(lab_host, non_blocking=True)
                    ready[b] = torch.cuda.Event()
                    ready[b].record(copy_s)
            if i >= 1:
                pb = (i - 1) & 1
                main_s.wait_event(ready[pb])
                if gstep is not None:
                    if u8_in:       # pre-processing straight into the graph's static image buffer
                        ops.image_u8_to_f32(in_dev[pb], mean_bgr, True, out=gstep.images)
                        res = gstep(None, lab_dev[pb])
                    else:
                        res = gstep(in_dev[pb], lab_dev[pb])
                elif graph is not None:
                    if u8_in:
                        ops.image_u8_to_f32(in_dev[pb], mean_bgr, True, out=x)      # into the graph's static input
                    else:
                        x.copy_(in_dev[pb], non_blocking=True)
                    graph.replay()
                    res = mask
                else:
                    xi = ops.image_u8_to_f32(in_dev[pb], mean_bgr, True, out=x) if u8_in else in_dev[pb]
                    res = step(xi, lab_dev[pb])
                freed[pb] = torch.cuda.Event()
                freed[pb].record(main_s)
                if read[pb] is not None:
                    main_s.wait_event(read[pb])          # the previous D2H out of this staging buffer has finished
                out_stage[pb].copy_(res, non_blocking=True)     # D2D: the graph's static output is free for the next replay
                done = torch.cuda.Event()
                done.record(main_s)
                with torch.cuda.stream(d2h_s):
                    d2h_s.wait_event(done)
                    out_host[pb].copy_(out_stage[pb], non_blocking=True)
                    read[pb] = torch.cuda.Event()
                    read[pb].record(d2h_s)
        main_s.wait_stream(copy_s)
        main_s.wait_stream(d2h_s)

    e2e_loop(2)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    k_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    e2e_loop(k_e2e)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / k_e2e
    if dist:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_value = world * batch / (e2e_ms / 1e3)

    # ---- roofline: one instrumented eager step, CUDA events around every launch
    hbm_peak, tc_peak, peak_src = peaks()
    roofline = model_roof = kernels = layers = units = None
    if train:
        # (every rank runs the step -- it contains collectives -- only rank 0 keeps the profile)
        for _ in range(2):
            ops.PROFILE = []
            step(x)
            torch.cuda.synchronize()
            prof, ops.PROFILE = ops.PROFILE, None
    else:
        prof, units, _ = unit_roofline(m, model_name, lambda: step(x), hbm_peak, tc_peak)
    if rank == 0:
        agg, kernels, layers, tot_ms = kernel_tables(prof)
        top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        per_launch = {"kernel": top[0], "achieved_GBps": round(top[1]["bytes"] / top[1]["ms"] / 1e6, 1),
                      "frac_of_hbm_peak": round(top[1]["bytes"] / top[1]["ms"] / 1e6 / hbm_peak, 4),
                      "TFLOPs": round(top[1]["flops"] / top[1]["ms"] / 1e9, 1),
                      "frac_of_bf16_peak": round(top[1]["flops"] / top[1]["ms"] / 1e9 / tc_peak, 4),
                      "launches_per_step": top[1]["launches"], "share_of_step": round(top[1]["ms"] / tot_ms, 4),
                      "definition": "sum over the kernel's launches of (|x|+|y|+|residual| bytes, conv FLOPs) / sum of their "
                                    "CUDA-event durations (what each launch itself must move; round-1 definition)"}
        traffic = ncu_traffic(wl_name, top[0])
        if units:
            uk, u = next(iter(units.items()))
            tensor = u["bound"] == "tensor"
            roofline = {"unit": "TFLOP/s" if tensor else "GB/s", "bound": u["bound"], "dominant_unit": uk,
                        "achieved": u["TFLOPs"] if tensor else u["GBps"], "peak": tc_peak if tensor else hbm_peak,
                        "frac": u["frac"], "hbm_frac": u["hbm_frac"], "tensor_frac": u["tensor_frac"],
                        "share_of_step": u["share"], "units_per_step": u["units"], "launches_per_unit": round(u["launches"] / u["units"], 2),
                        "definition": "SURVEY 8(d): algorithmic bytes of a block unit = tensors entering + leaving it, each "
                                      "once (2*C*h*w elements for a non_bottleneck_1d / DABModule), FLOPs = 2 x its conv MACs; "
                                      "the binding roof is the larger of bytes/HBM-peak and FLOPs/bf16-peak; frac = that time "
                                      "/ the CUDA-event time of the unit's launches"}
        else:
            achieved = top[1]["bytes"] / (top[1]["ms"] / 1e3) / 1e9
            roofline = {"unit": "GB/s", "bound": "hbm", "dominant_unit": top[0], "achieved": round(achieved, 1), "peak": hbm_peak,
                        "frac": round(achieved / hbm_peak, 4), "share_of_step": round(top[1]["ms"] / tot_ms, 4),
                        "definition": "training step: per-kernel algorithmic bytes (|x|+|dy| for a weight gradient, in+out for "
                                      "the others) / CUDA-event time, for the kernel with the largest share"}
        graph_tables = None
        if train and graph is not None and world == 1:      # (the graph holds collectives when world > 1: every rank would have to replay)
            try:
                g_entries, g_kernels, g_sum = graph_kernel_table(graph, agg, ms_per_step)
                graph_tables = {"entries": g_entries, "kernels": g_kernels, "summary": g_sum}
                # the dominant KERNEL: the CUDA kernel with the largest device time in the replayed graph, reported through the
                # entry point that launches it when that entry point launches nothing else (esn_conv2d_wgrad dispatches to six
                # kernels, two thirds of them on the side stream: it is listed under graph_kernels.entries, not here)
                gk, gv = next(iter(g_entries.items()))
                for kname in g_kernels:
                    owners = [e for e, subs in ENTRY_KERNELS.items() if e in g_entries and any(s_ in kname for s_ in subs)]
                    if len(owners) == 1 and len(ENTRY_KERNELS[owners[0]]) == 1 and all(
                            (ENTRY_KERNELS[owners[0]][0] not in k2) or k2.split("<")[0] == kname.split("<")[0] for k2 in g_kernels):
                        gk, gv = owners[0], g_entries[owners[0]]
                        break
                achieved = agg[gk]["bytes"] / (gv["ms"] / 1e3) / 1e9
                roofline = {"unit": "GB/s", "bound": "hbm", "dominant_unit": gk, "achieved": round(achieved, 1), "peak": hbm_peak,
                            "frac": round(achieved / hbm_peak, 4), "share_of_step": gv["share_of_step"],
                            "launches_per_step": gv["launches"],
                            "definition": "training step = one replayed CUDA graph: algorithmic bytes of the entry point's launches "
                                          "(in + out per launch; |x|+|dy| for a weight gradient; x + dy + dx for a BatchNorm "
                                          "backward) / their device time inside the replayed graph (CUPTI activity records, "
                                          "graph_kernels), for the CUDA kernel with the largest in-graph time; the "
                                          "events-around-eager-launches figure of the eager step's top kernel is under per_launch"}
                traffic = ncu_traffic(wl_name, gk)
            except Exception as exc:      # noqa: BLE001 -- a diagnostic table must not lose the measured line
                graph_tables = {"error": repr(exc)[:300]}
        roofline.update({"traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_detail": traffic,
                         "peak_source": peak_src, "per_launch": per_launch})
        # whole-network figure against SURVEY 8(d)'s block-fused algorithmic bytes
        px = batch * H * W
        alg_bytes = px * ((ALG_ELEMS_PER_PX[model_name] - 19.0) * 2 + 1 + 3 * 4 - 3 * 2)
        flops = 2 * GMAC_512x1024[model_name] * 1e9 * (H * W) / (512 * 1024) * batch
        if train:   # ~3.5x the forward block-fused traffic (saved activations + gradients), 3x the FLOPs
            alg_bytes = px * 3.5 * ALG_ELEMS_PER_PX[model_name] * 2
            flops *= 3
        t_hbm, t_tc = alg_bytes / (hbm_peak * 1e9) * 1e3, flops / (tc_peak * 1e12) * 1e3
        model_roof = {"alg_bytes_per_step": int(alg_bytes), "hbm_frac": round(t_hbm / ms_per_step, 4),
                      "tensor_frac": round(t_tc / ms_per_step, 4), "bound": "tensor" if t_tc > t_hbm else "hbm",
                      "frac": round(max(t_hbm, t_tc) / ms_per_step, 4),
                      "note": "block-fused algorithmic bytes (SURVEY 8d) and conv FLOPs of the whole step / step time"}
    if rank != 0:
        del graph, gstep
        return None
    line = {"metric": "images/s", "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(config, cuda_graph=graph is not None),
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "steps": k_e2e, "input": "u8" if u8_in else "f32",
                    "preprocessing_bit_exact_vs_torch": exact,
                    "note": (("pinned uint8 HWC BGR images + uint8 label maps (what cv2.imread gives the reference's dataset class) -> H2D "
                              "-> esn_image_u8hwc_to_f32nchw + label cast on the device" if u8_in else
                              "pinned fp32 NCHW images + int64 labels -> H2D") +
                             " -> one training iteration (forward, weighted CE, backward, all-reduce, Adam; CUDA graph: %s) -> D2H "
                             "loss scalar; copies double-buffered on side streams" % (gstep is not None))
                    if train else
                    ("pinned uint8 HWC BGR images (what cv2.imread gives the reference's dataset class) -> H2D -> "
                     "esn_image_u8hwc_to_f32nchw -> model.predict_mask -> D2H uint8 masks; copies double-buffered on side streams"
                     if u8_in else
                     "pinned fp32 NCHW images -> H2D -> model.predict_mask -> D2H uint8 masks; copies double-buffered on side streams")},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "model_roofline": model_roof, "units": dict(list(units.items())[:12]) if units else None,
            "kernels": kernels, "layers": layers}
    if train and graph is not None and world == 1:
        line["graph_kernels"] = graph_tables
    if primary and not args.no_cpu_baseline and world == 1:      # contract: rank 0 at N=1 only
        base, _, _ = cpu_reference_leg(model_name, H, W, 4, 1, budget_s=20.0, train=train)
        line["cpu_baseline"] = base
    del graph, gstep, m
    torch.cuda.empty_cache()
    if not args.no_gpu_eager and world == 1:
        try:
            g = gpu_eager_reference_leg(model_name, batch, H, W, train)
            g["speedup_vs_fastest"] = round(value / g["value"], 3) if g["value"] else None
            line["gpu_eager_baseline"] = g
        except Exception as exc:      # noqa: BLE001 -- report, do not lose the measured line
            line["gpu_eager_baseline"] = {"error": repr(exc)[:300]}
    return line


DEFAULT_WORKLOAD = "dabnet_train_bf16_b8_512x1024"
SECONDARY = {"dabnet_train_bf16_b8_512x1024": ["erfnet_infer_bf16_b16_1024x2048"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the gpu_eager_baseline leg (N=1)")
    ap.add_argument("--no-legs", action="store_true", help="skip the secondary workloads reported under `legs`")
    ap.add_argument("--e2e-input", default="u8", choices=["f32", "u8"],
                    help="e2e leg: what crosses PCIe -- decoded uint8 HWC BGR images with mean subtraction / BGR->RGB / "
                         "CHW done on the device (esn_image_u8hwc_to_f32nchw, SURVEY 8f-4; default), or the reference's "
                         "pre-processed fp32 NCHW batch")
    ap.add_argument("--no-fused-loss", action="store_true",
                    help="training workloads: criterion(model(x), y) as two modules instead of the model's fused head + loss (A/B)")
    ap.add_argument("--torch-adam", action="store_true",
                    help="training workloads: step torch.optim.Adam(fused=True) instead of esn.optim.Adam (A/B)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to stdout) are
    # redirected to stderr for the duration of the run; the JSON goes to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        _real_stdout.write(json.dumps(obj) + "\n")
        _real_stdout.flush()
    model_name, batch, H, W, mode = WORKLOADS[args.workload]
    train = mode == "train"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        wu = max(args.warmup, 3)
        base, mean, n = cpu_reference_leg(model_name, H, W, args.steps, wu, budget_s=90.0, train=train)
        emit({"impl": "reference", "metric": "images/s", "value": base["value"], "unit": "images/s",
              "n_gpus": args.gpus, "steps": n, "warmup": wu, "ms_per_step": mean * 1e3,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": args.workload, "net": model_name, "classes": 19, "batch_per_gpu": 1,
                         "input": "3x%dx%d fp32 NCHW" % (H, W),
                         "mode": ("training step (CPU, reference arithmetic): train-mode forward + weighted CE + backward"
                                  if train else "inference (CPU, reference arithmetic)"),
                         "sample": "each step = ONE image of the workload's resolution (bounded sample of the %d-image batch)" % batch},
              "cpu_baseline": base,
              "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if args.impl == "reference-gpu":
        # the reference graph eager on this GPU (every variant incl. true fp32); one process, rank 0 only
        if rank != 0:
            return
        g = gpu_eager_reference_leg(model_name, batch, H, W, train, warmup=max(args.warmup, 5), steps=args.steps,
                                    variants=["fp32", "fp32_tf32", "bf16_autocast", "bf16_autocast_channels_last",
                                              "bf16_weights_channels_last"])
        emit({"impl": "reference-gpu", "metric": "images/s", "value": g["value"], "unit": "images/s", "n_gpus": 1,
              "steps": args.steps, "warmup": max(args.warmup, 5),
              "ms_per_step": g["variants"][g["fastest"]]["ms_per_step"] if g["fastest"] else None,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": g["fastest"], "data": "synthetic",
              "config": {"workload": args.workload, "net": model_name, "batch_per_gpu": batch, "input": "3x%dx%d" % (H, W)},
              "gpu_eager_baseline": g, "gpu_launches": 0})
        return
    affinity = numa_local_affinity(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = measure(args, args.workload, rank, world, local_rank, dist, primary=True)
    legs = {}
    if not args.no_legs:
        for name in SECONDARY.get(args.workload, []):
            try:
                leg = measure(args, name, rank, world, local_rank, dist, primary=False)
            except Exception as exc:      # noqa: BLE001 -- a failing secondary leg must not cost the primary line
                leg = {"error": repr(exc)[:300]}
            if rank == 0:
                legs[name] = leg
    if rank == 0:
        line["host_affinity"] = affinity
        if legs:
            line["legs"] = legs
        emit(line)
    _finish(dist)


def _finish(dist):
    """Orderly exit.  Round 1 left through os._exit(0) because destroy_process_group could block after NCCL collectives had
    been captured into a CUDA graph: the captured graphs keep the communicator's streams and buffers referenced.  Now the
    graphs are destroyed first (measure() drops them), the device is synchronised, every rank meets at a barrier and the
    process group is destroyed.  A watchdog still guarantees that the launcher is released if teardown stalls (it says so
    on stderr)."""
    sys.stdout.flush()
    sys.stderr.flush()
    if not dist:
        return
    import gc
    gc.collect()
    torch.cuda.synchronize()

    def bail():
        sys.stderr.write("bench.py: process-group teardown did not finish in 30 s; leaving through os._exit\n")
        sys.stderr.flush()
        os._exit(0)
    timer = threading.Timer(30.0, bail)
    timer.daemon = True
    timer.start()
    try:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        sys.stderr.write("bench.py: process group destroyed cleanly\n")
    finally:
        timer.cancel()


if __name__ == "__main__":
    main()
