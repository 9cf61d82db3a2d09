#!/usr/bin/env python
"""bench.py -- images/s of the segmentation hot path on B200 (driver contract in the task brief).

Default workload = BASELINE.json configs[1]: ERFNet (19 classes), bf16, inference, batch 16 per GPU,
3x1024x2048 synthetic Cityscapes-shaped input, fused argmax head (uint8 mask out).
A "step" is one forward of one batch.  N>1 shards images across ranks with no data-path
collective (inference; "weak" scaling: 16 images per GPU).

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the gpu_eager_baseline leg (N=1)")
    ap.add_argument("--workload", default="erfnet_infer_bf16_b16_1024x2048", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-u8-leg", action="store_true", help="skip the extra e2e_u8 leg (inference, N=1)")
    ap.add_argument("--e2e-input", default="f32", choices=["f32", "u8"],
                    help="what crosses PCIe in the e2e leg: the reference's pre-processed fp32 NCHW batch (default), or the "
                         "decoded uint8 HWC BGR images with mean subtraction / BGR->RGB / CHW done on the device "
                         "(esn_image_u8hwc_to_f32nchw, SURVEY 8f-4)")
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
    config = {"workload": args.workload, "net": model_name, "classes": 19, "batch_per_gpu": batch,
              "input": "3x%dx%d fp32 NCHW" % (H, W),
              "mode": "training step: forward + weighted CE + backward + Adam" if train else "inference",
              "head": "bilinear -> fp32 logits -> fused weighted-CE kernel" if train else "argmax fused (uint8 mask)",
              "sharding": ("data parallel: flat fp32 gradient buckets all-reduced (NCCL) from inside the backward tape, "
                           "2-scalar loss all-reduce, per-GPU BatchNorm") if train else "images across ranks, no collective",
              "l2": "no flush: per-step working set (input %.0f MB + activations) >> 126 MB L2" % (batch * 3 * H * W * 4 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        base, mean, n = cpu_reference_leg(model_name, H, W, args.steps, max(args.warmup, 3), budget_s=90.0, train=train)
        line = {"impl": "reference", "metric": "images/s", "value": base["value"], "unit": "images/s",
                "n_gpus": args.gpus, "steps": n, "warmup": max(args.warmup, 3), "ms_per_step": mean * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, mode="inference (CPU, reference arithmetic)", batch_per_gpu=1),
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
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
              "config": config, "gpu_eager_baseline": g, "gpu_launches": 0})
        return
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from builders.model_builder import build_model
    from esn import ops
    from oracle import fixture

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
        opt = torch.optim.Adam(m.parameters(), lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, fused=True,
                               capturable=not args.no_graph)  # train.py:212-215
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

    # ---- warm-up (also builds the packed-weight caches), then optional CUDA-graph capture
    for _ in range(max(args.warmup, 3)):
        mask = step(x)
    torch.cuda.synchronize()
    ops.launch_count_reset()
    step(x)
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    graph = None
    gstep = None
    if train and not args.no_graph:
        # the whole iteration (forward, loss, backward, gradient all-reduce, Adam) as one CUDA graph (esn/graph.py)
        from esn.graph import GraphedTrainStep
        gstep = GraphedTrainStep(m, crit, opt, x, y, warmup=1)
        graph = gstep.graph
        mask = gstep.loss
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
    elif not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step(x)
            with torch.cuda.graph(graph, stream=s):
                mask = step(x)
        torch.cuda.current_stream().wait_stream(s)
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

    # ---- e2e: pinned host images -> H2D -> forward -> D2H uint8 masks, every step, double-buffered
    u8_in = args.e2e_input == "u8"
    mean_bgr = [72.3924, 82.90902, 73.158325]      # dataset/inform/cityscapes_inform.pkl['mean'] (BGR, fp32)
    u8 = {}

    def u8_buffers():
        if not u8:
            g8 = torch.Generator().manual_seed(1234 + rank)
            u8["host"] = torch.randint(0, 256, (batch, H, W, 3), dtype=torch.uint8, generator=g8).pin_memory()
            u8["dev"] = [torch.empty((batch, H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
        return u8["host"], u8["dev"]
    if u8_in:
        x_host_u8, xin_u8 = u8_buffers()
    h2d = (x_host_u8.numel() if u8_in else x_host.numel() * 4) + (y_host.numel() * 8 if train else 0)
    d2h = 4 if train else batch * H * W
    mask_host = [(torch.empty((), dtype=torch.float32) if train else torch.empty((batch, H, W), dtype=torch.uint8)).pin_memory()
                 for _ in range(2)]
    xin = [torch.empty_like(x), torch.empty_like(x)]
    yin = [torch.empty_like(y), torch.empty_like(y)] if train else [None, None]
    copy_s = torch.cuda.Stream()
    d2h_s = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()

    def e2e_loop(k, from_u8=u8_in):
        if from_u8:
            x_host_u8, xin_u8 = u8_buffers()
        ready = [None, None]
        done = [None, None]
        for i in range(k + 1):
            b = i & 1
            if i < k:
                with torch.cuda.stream(copy_s):
                    if done[b] is not None:
                        copy_s.wait_event(done[b])      # buffer b free (its compute finished)
                    if from_u8:
                        xin_u8[b].copy_(x_host_u8, non_blocking=True)
                    else:
                        xin[b].copy_(x_host, non_blocking=True)
                    if train:
                        yin[b].copy_(y_host, non_blocking=True)
                    ready[b] = torch.cuda.Event()
                    ready[b].record(copy_s)
            if i >= 1:
                pb = (i - 1) & 1
                main_s.wait_event(ready[pb])
                if from_u8:    # device half of the dataset class: uint8 HWC BGR -> fp32 NCHW RGB - mean (one launch)
                    ops.image_u8_to_f32(xin_u8[pb], mean_bgr, True, out=xin[pb])
                mk = gstep(xin[pb], yin[pb]) if gstep is not None else step(xin[pb], yin[pb])
                if from_u8 and gstep is None:
                    # with 3 bytes per pixel going up, the 1 byte per pixel coming down is no longer hidden behind the H2D
                    # copy: read the masks back on a third stream so the next step's kernels do not queue behind it
                    done[pb] = torch.cuda.Event()
                    done[pb].record(main_s)                 # compute finished: input buffer pb is free, the mask is ready
                    with torch.cuda.stream(d2h_s):
                        d2h_s.wait_event(done[pb])
                        mask_host[pb].copy_(mk, non_blocking=True)
                    mk.record_stream(d2h_s)                 # the allocator must not hand the mask's memory out before the copy ran
                    continue
                mask_host[pb].copy_(mk, non_blocking=True)
                done[pb] = torch.cuda.Event()
                done[pb].record(main_s)
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
    t1 = time.perf_counter()
    e2e_ms = (t1 - t0) * 1e3 / k_e2e
    if dist:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_value = world * batch / (e2e_ms / 1e3)

    # ---- per-kernel roofline: one instrumented eager step, CUDA events around every launch
    hbm_peak, tc_peak, peak_src = peaks()
    roofline, kernels = None, None
    # (every rank runs the step -- a training step contains collectives -- only rank 0 keeps the profile)
    for _ in range(2):
        ops.PROFILE = []
        step(x)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    if rank == 0:
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
        layers = {k: {"n": v[0], "ms": round(v[1], 3), "alg_GBps": round(v[2] / v[1] / 1e6, 1),
                      "TFLOPs": round(v[3] / v[1] / 1e9, 1)}
                  for k, v in sorted(by_tag.items(), key=lambda kv: -kv[1][1])[:24]}
        top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        achieved = top[1]["bytes"] / (top[1]["ms"] / 1e3) / 1e9
        # DRAM traffic of the same kernel family from the committed ncu launch list (separate run)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic_%s.json" % model_name.lower())
        fam_names = {"esn_conv2d_umma": "conv_umma_kernel", "esn_conv_pair_umma": "conv_pair_kernel", "esn_conv2d_direct": "conv_direct_kernel",
                     "esn_dab_dw_pair": "dab_dw_pair_kernel", "esn_affine_act": "pw_kernel"}
        if os.path.exists(tpath) and top[0] in fam_names:
            tj = json.load(open(tpath))
            if tj.get("workload") == args.workload:
                fams = [v for k, v in tj["families"].items() if k.startswith(fam_names[top[0]])]
                n_l = sum(v["launches"] for v in fams)
                if n_l:
                    traffic = {"dram_bytes_per_launch": int(sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in fams) / n_l),
                               "alg_bytes_per_launch": int(top[1]["bytes"] / top[1]["launches"]), "kernel_launches": n_l,
                               "source": tj["source"]}
        roofline = {"kernel": top[0], "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(achieved / hbm_peak, 4),
                    # contract: DRAM bytes (read + write) per launch of this kernel from the ncu capture, or null
                    "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_detail": traffic,
                    "peak_source": peak_src,
                    "launches_per_step": top[1]["launches"], "share_of_step": round(top[1]["ms"] / tot_ms, 4),
                    "tensor_TFLOPs": round(top[1]["flops"] / (top[1]["ms"] / 1e3) / 1e12, 1),
                    "tensor_frac_of_bf16_peak": round(top[1]["flops"] / (top[1]["ms"] / 1e3) / 1e12 / tc_peak, 4),
                    "definition": "sum over the step's launches of (|x|+|y|+|residual| bytes) / sum of their CUDA-event durations"}
        # whole-network figure against SURVEY 8(d)'s block-fused algorithmic bytes
        px = batch * H * W
        alg_bytes = px * ((ALG_ELEMS_PER_PX[model_name] - 19.0) * 2 + 1 + 3 * 4 - 3 * 2)
        flops = 2 * GMAC_512x1024[model_name] * 1e9 * (H * W) / (512 * 1024) * batch
        if train:   # ~3.5x the forward block-fused traffic (saved activations + gradients), 3x the FLOPs
            alg_bytes = px * 3.5 * ALG_ELEMS_PER_PX[model_name] * 2
            flops *= 3
        model_roof = {"alg_bytes_per_step": int(alg_bytes), "hbm_frac": round(alg_bytes / (ms_per_step / 1e3) / 1e9 / hbm_peak, 4),
                      "tensor_frac": round(flops / (ms_per_step / 1e3) / 1e12 / tc_peak, 4),
                      "note": "block-fused algorithmic bytes (SURVEY 8d) and conv FLOPs of the whole forward / step time"}

    # ---- extra leg (inference, N=1): the same e2e loop fed with decoded uint8 HWC BGR images, normalised on the device by
    # esn_image_u8hwc_to_f32nchw.  Reported next to `e2e`, never instead of it; guarded so that a failure of
    # this newer kernel cannot cost the line above.  Self-check: bit-equality with the same fp32 arithmetic done by torch.
    e2e_u8 = None
    if world == 1 and not train and not u8_in and not args.no_u8_leg:
        try:
            x_host_u8, xin_u8 = u8_buffers()
            xin_u8[0].copy_(x_host_u8)
            sub = xin_u8[0][:2]
            got = ops.image_u8_to_f32(sub, mean_bgr, True)
            want = (sub.float() - torch.tensor(mean_bgr, device="cuda")).flip(3).permute(0, 3, 1, 2)
            exact = bool(torch.equal(got, want))
            del got, want
            e2e_loop(2, True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_loop(k_e2e, True)
            torch.cuda.synchronize()
            u8_ms = (time.perf_counter() - t0) * 1e3 / k_e2e
            e2e_u8 = {"value": round(batch / (u8_ms / 1e3), 2), "unit": "images/s", "h2d_bytes_per_step": x_host_u8.numel(),
                      "d2h_bytes_per_step": d2h, "ms_per_step": round(u8_ms, 3), "steps": k_e2e, "bit_exact_vs_torch": exact,
                      "note": "pinned uint8 HWC BGR images -> H2D (side stream) -> esn_image_u8hwc_to_f32nchw -> model.predict_mask -> D2H uint8 masks (third stream)"}
        except Exception as exc:      # noqa: BLE001 -- report, do not lose the measured line
            e2e_u8 = {"error": repr(exc)[:300]}

    if rank != 0:
        _finish(dist)
        return
    line = {"metric": "images/s", "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(config, cuda_graph=graph is not None),
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "steps": k_e2e,
                    "note": ("pinned fp32 NCHW images + int64 labels -> H2D -> one training iteration (forward, weighted CE, backward, "
                             "all-reduce, Adam; CUDA graph: %s) -> D2H loss scalar; copies double-buffered on a side stream" % (gstep is not None))
                    if train else
                    ("pinned uint8 HWC BGR images -> H2D -> esn_image_u8hwc_to_f32nchw -> model.predict_mask -> D2H uint8 masks; "
                     "copies double-buffered on a side stream" if u8_in else
                     "pinned fp32 NCHW images -> H2D -> model.predict_mask -> D2H uint8 masks; copies double-buffered on a side stream"),
                    "input": args.e2e_input},
            "e2e_u8": e2e_u8,
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "model_roofline": model_roof, "kernels": kernels, "layers": layers}
    if not args.no_cpu_baseline and world == 1:      # contract: rank 0 at N=1 only
        base, _, _ = cpu_reference_leg(model_name, H, W, 4, 1, budget_s=20.0, train=train)
        line["cpu_baseline"] = base
    if not args.no_gpu_eager and world == 1:
        torch.cuda.empty_cache()
        try:
            g = gpu_eager_reference_leg(model_name, batch, H, W, train)
            g["speedup_vs_fastest"] = round(value / g["value"], 3) if g["value"] else None
            line["gpu_eager_baseline"] = g
        except Exception as exc:      # noqa: BLE001 -- report, do not lose the measured line
            line["gpu_eager_baseline"] = {"error": repr(exc)[:300]}
    emit(line)
    if e2e_u8 is not None and "error" in e2e_u8:
        # the guarded leg failed: if it left a sticky CUDA error, a normal interpreter shutdown could abort after the line
        # has been printed; leave directly (all results are out)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    _finish(dist)


def _finish(dist):
    """Leave without tearing NCCL down: communicators whose collectives were captured into a CUDA graph (the graphed
    training step) can block in destroy_process_group / interpreter shutdown, which would hang the launcher after the
    JSON line has been printed.  All device work is complete here (synchronised above)."""
    sys.stdout.flush()
    sys.stderr.flush()
    if dist:
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        os._exit(0)


if __name__ == "__main__":
    main()
