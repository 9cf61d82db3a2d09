#!/usr/bin/env python
"""Per-tensor gradient-norm errors of a train-mode net against the reference's fp64 golden (development tool).
Usage: python tools/train_grad_diag.py NET [fp32|bf16]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from conftest import spec_state_dict  # noqa: E402
from oracle import fixture  # noqa: E402


def main():
    from builders.model_builder import build_model
    from utils.losses.loss import CrossEntropyLoss2d
    net = sys.argv[1]
    bf16 = len(sys.argv) > 2 and sys.argv[2] == "bf16"
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json")))
    g = np.load(os.path.join(ROOT, "tests", "golden", net + ".npz"))
    m = build_model(net, 19)
    m.load_state_dict(spec_state_dict(spec, net))
    m = m.cuda().train()
    for mod in m.modules():
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.p = 0.0
    x = fixture.make_input(2, 64, 128).cuda()
    lab = fixture.make_labels(2, 64, 128, 19).cuda()
    crit = CrossEntropyLoss2d(weight=torch.tensor(fixture.CLASS_WEIGHTS), ignore_label=255).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
        out = m(x)
        loss = crit(out, lab)
    loss.backward()
    ref = torch.from_numpy(g["train_2x64x128_logits_s4"])
    o = out.detach().float().cpu()[:, :, ::4, ::4]
    print("loss %.6f (ref %.6f)  logits rel-L2 %.3e" % (loss.item(), float(g["train_2x64x128_loss"][0]),
                                                        ((o - ref).norm() / ref.norm()).item()))
    stats = json.loads(bytes(g["train_2x64x128_gradstats"]).decode())
    named = dict(m.named_parameters())
    rows = []
    for k, (gn, gs, wn) in stats.items():
        gr = named[k].grad
        if gr is None:
            rows.append((float("inf"), k, gn, None, None))
            continue
        rows.append((abs(gr.double().norm().item() - gn) / max(gn, 1e-30), k, gn, gr.double().norm().item(), gr.double().sum().item() - gs))
    # torch's own fp32 autograd of the oracle graph on the GPU: the fp32 noise floor per tensor
    import torch.nn.functional as F
    from oracle import nets
    sd32 = {k: (v.cuda().requires_grad_(True) if v.is_floating_point() else v.cuda()) for k, v in spec_state_dict(spec, net).items()}
    y32 = nets.forward(net, sd32, x, train=True)
    F.cross_entropy(y32, lab, torch.tensor(fixture.CLASS_WEIGHTS, device="cuda"), ignore_index=255).backward()

    def tg(k):
        g_ = sd32[k].grad
        if g_ is None:
            g_ = sd32[k.split(".")[0] + ".out_prelu.weight"].grad
        return g_
    terr = {k: abs(tg(k).double().norm().item() - gn) / max(gn, 1e-30) for k, (gn, gs, wn) in stats.items()}
    order = {k: i for i, k in enumerate(stats)}
    print("in network order (err, name, ref norm, our norm, sum diff):")
    for r in sorted(rows, key=lambda r: order[r[1]]):
        flag = " <<<" if r[0] > (0.3 if bf16 else 2e-2) else ""
        print("  %.3e  (torch fp32 %.3e)  %-50s %.4e %s%s" % (r[0], terr[r[1]], r[1], r[2], "%.4e" % r[3] if r[3] is not None else "None", flag))
    for key in g.files:
        if key.startswith("train_2x64x128_grad::"):
            k = key.split("::")[1]
            gold = torch.from_numpy(g[key])
            print("full tensor", k, ((named[k].grad.cpu().double() - gold.double()).norm() / gold.double().norm()).item())


if __name__ == "__main__":
    main()
