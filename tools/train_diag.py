#!/usr/bin/env python
"""Per-tensor gradient error of a train step vs the reference fp64 golden (debug helper)."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-segmentation-networks_b200"), os.path.join(ROOT, "tests")]
from conftest import spec_state_dict  # noqa
from test_train_gpu import _train_step  # noqa
name = sys.argv[1] if len(sys.argv) > 1 else "DABNet"
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
spec = json.load(open(os.path.join(ROOT, "tests/golden/state_dict_spec.json")))
g = np.load(os.path.join(ROOT, "tests/golden/%s.npz" % name))
m, out, loss = _train_step(name, spec, dtype)
stats = json.loads(bytes(g["train_2x64x128_gradstats"]).decode())
named = dict(m.named_parameters())
rows = []
for k, (gnorm, gsum, wnorm) in stats.items():
    if gnorm < 1e-10 * max(wnorm, 1e-30):
        continue
    gr = named[k].grad.double()
    rows.append((abs(gr.norm().item() - gnorm) / gnorm, k, gnorm, gr.norm().item(), abs(gr.sum().item() - gsum) / (abs(gsum) + 1e-30)))
rows.sort(reverse=True)
print("loss", loss.item(), "ref", float(g["train_2x64x128_loss"][0]))
for r in rows[:25]:
    print("%.3e  %-60s ref|g| %.4e  ours %.4e  sum-err %.2e" % r)
print("median", rows[len(rows) // 2][0])

# ---- noise floor: the reference arithmetic itself (oracle port) under torch bf16 autocast on the GPU
if dtype == torch.bfloat16:
    from oracle import fixture, nets
    import torch.nn.functional as F
    sd = {k: (v.cuda().requires_grad_(True) if v.is_floating_point() else v.cuda()) for k, v in spec_state_dict(spec, name).items()}
    x = fixture.make_input(2, 64, 128).cuda()
    lab = fixture.make_labels(2, 64, 128, 19).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = nets.forward(name, sd, x, train=True)
        l = F.cross_entropy(y.float(), lab, torch.tensor(fixture.CLASS_WEIGHTS, device="cuda"), ignore_index=255)
    l.backward()
    rows2 = []
    for k, (gnorm, gsum, wnorm) in stats.items():
        if gnorm < 1e-10 * max(wnorm, 1e-30):
            continue
        gr = sd[k].grad.double()
        rows2.append((abs(gr.norm().item() - gnorm) / gnorm, k))
    rows2.sort(reverse=True)
    print("torch-autocast reference: loss", l.item(), " worst", rows2[0], " p90", rows2[len(rows2) // 10][0], " median", rows2[len(rows2) // 2][0])
    print("ours:                      worst", rows[0][0], " p90", rows[len(rows) // 10][0], " median", rows[len(rows) // 2][0])
