#!/usr/bin/env python
"""Layer-by-layer bf16 diagnosis of a net against the oracle run in fp32 ON THE GPU (oracle = torch functional code, so it
runs on CUDA tensors).  Development tool: for each top-level layer prints (a) the isolated error -- our layer and the
oracle layer fed with the SAME input (our previous bf16 output) -- and (b) the accumulated error against the oracle's own
fp32 chain.  Usage: python tools/diag_layers.py LEDNet 128 256"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    from builders.model_builder import build_model
    from esn import ops
    from oracle import fixture, nets
    name, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))
    proto = {k: torch.empty(shape, dtype=getattr(torch, dt.split(".")[1])) for k, shape, dt in spec[name]["keys"]}
    sd = fixture.randomize_state_dict(proto, 1234)
    m = build_model(name, 19)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    sdg = {k: v.cuda() for k, v in sd.items()}
    x = fixture.make_input(n, h, w).cuda()
    P = nets.SD(sdg, "", torch.float32)
    with torch.no_grad():
        if name == "LEDNet":
            stages = [("initial_block", m.initial_block, lambda t: nets.erf_downsampler(P.sub("initial_block"), t))]
            for i, d in enumerate(nets.LED_LAYERS):
                q = P.sub("layers.%d" % i)
                fn = (lambda t, q=q: nets.erf_downsampler(q, t)) if d is None else (lambda t, q=q, d=d: nets.led_ssnbt(q, t, d))
                stages.append(("layers.%d (d=%s)" % (i, d), m.layers[i], fn))
            stages.append(("apn", m.apn, lambda t: nets.led_apn(P.sub("apn"), t)))
        else:
            raise SystemExit("no stage table for " + name)
        ours, ref = x, x
        for label, mod, fn in stages:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = mod(ours)
            c = y.shape[1]
            yf = y.float()
            iso = fn(ours.float() if ours is not x else x)
            ref = fn(ref)
            print("%-22s C=%3d %4dx%-4d isolated %.3e  accumulated %.3e  |ref| %.3e" %
                  (label, c, y.shape[2], y.shape[3], rel(yf, iso), rel(yf, ref), ref.abs().mean().item()), flush=True)
            ours = y
        with torch.autocast("cuda", dtype=torch.bfloat16):
            full = m(x)
            sd_ac = nets.forward(name, sdg, x)
        want = nets.forward(name, sdg, x)
        print("full net: ours %.3e  torch-autocast %.3e" % (rel(full.float(), want), rel(sd_ac.float(), want)))
        if name == "LEDNet":
            # APN pieces on the fp32 oracle input: which stage of the single-channel pyramid carries the bf16 error
            feat = ref_feat = None
            t = x
            for label, mod, fn in stages[:-1]:
                t = fn(t)
            q = P.sub("apn")
            with torch.autocast("cuda", dtype=torch.bfloat16):
                got = m.apn(t)
            print("apn on the oracle's fp32 features: %.3e" % rel(got.float(), nets.led_apn(q, t)))


if __name__ == "__main__":
    main()
