#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (build container only).

Imports /root/reference behind three import stubs (torchsummary, fvcore, thop
are absent here), gives each model the seeded fixture weights of
oracle/fixture.py, runs it on CPU in fp32 (fp64 for gradients) and stores
small outputs.  The reference cannot travel to the GPU box; these files can.

    python tools/make_golden.py            # regenerate everything
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ESN_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def import_reference():
    for name, attrs in (("torchsummary", {"summary": lambda *a, **k: None}),
                        ("fvcore", {}), ("fvcore.nn", {}),
                        ("fvcore.nn.flop_count", {"flop_count": lambda *a, **k: None}),
                        ("thop", {"profile": lambda *a, **k: None})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
    sys.path.insert(0, REF)
    from builders.model_builder import build_model  # noqa
    from utils.losses.loss import CrossEntropyLoss2d  # noqa
    return build_model, CrossEntropyLoss2d


def main():
    from oracle import fixture
    build_model, CE = import_reference()
    torch.set_num_threads(8)
    os.makedirs(GOLD, exist_ok=True)
    names = sys.argv[1:] or ["ERFNet", "DABNet"]
    spec = {}
    spec_path = os.path.join(GOLD, "state_dict_spec.json")
    if os.path.exists(spec_path):
        spec = json.load(open(spec_path))
    for name in names:
        torch.manual_seed(1234)
        m = build_model(name, 19)
        sd = fixture.randomize_state_dict(m.state_dict(), 1234)
        m.load_state_dict(sd)      # aliased parameters (ENet's shared PReLU): the last key loaded wins
        spec[name] = {"keys": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()],
                      "n_params": int(sum(p.numel() for p in m.parameters()))}
        out = {}
        # ---- eval forward, fp32, two sizes
        m.eval()
        with torch.no_grad():
            for (n, h, w) in ((1, 64, 128), (2, 128, 256)):
                x = fixture.make_input(n, h, w, 1234)
                y = m(x)
                tag = "eval_%dx%dx%d" % (n, h, w)
                if h == 64:
                    out[tag + "_logits"] = y.numpy()
                else:
                    out[tag + "_logits_s4"] = y[:, :, ::4, ::4].contiguous().numpy()
                    out[tag + "_sum"] = np.array([y.double().sum().item(), y.double().abs().sum().item()])
                out[tag + "_argmax"] = np.argmax(y.numpy(), axis=1).astype(np.uint8)
        if name not in ("ERFNet", "DABNet", "FastSCNN", "ESPNet_v2", "ENet", "CGNet", "ESPNet"):      # inference-only nets: no training golden
            np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
            print(name, "golden written:", {k: v.shape for k, v in out.items()})
            continue
        # ---- train-mode forward + weighted CE + backward, fp64 (dropout off)
        m64 = build_model(name, 19).double()
        m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
        m64.train()
        for mod in m64.modules():
            if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
                mod.p = 0.0
        x = fixture.make_input(2, 64, 128, 1234).double()
        lab = fixture.make_labels(2, 64, 128, 19, seed=1234)
        crit = CE(weight=torch.tensor(fixture.CLASS_WEIGHTS, dtype=torch.float64), ignore_label=255)
        y = m64(x)
        loss = crit(y, lab)
        loss.backward()
        out["train_2x64x128_loss"] = np.array([loss.item()])
        out["train_2x64x128_logits_s4"] = y.detach()[:, :, ::4, ::4].contiguous().numpy().astype(np.float32)
        gn = {}
        for k, p_ in m64.named_parameters():
            if p_.grad is not None:
                g = p_.grad
                gn[k] = [float(g.norm()), float(g.sum()), float(p_.detach().norm())]
        out["train_2x64x128_gradstats"] = np.frombuffer(json.dumps(gn).encode(), dtype=np.uint8)
        # a few full gradient tensors (first conv, a mid conv, last conv)
        named = dict(m64.named_parameters())
        picks = [k for k in named if named[k].grad is not None and named[k].dim() == 4]
        for k in (picks[0], picks[len(picks) // 2], picks[-1]):
            out["train_2x64x128_grad::" + k] = named[k].grad.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "golden written:", {k: v.shape for k, v in out.items() if not k.endswith("gradstats")})
    json.dump(spec, open(spec_path, "w"), indent=0)

    # ---- loss golden (reference CrossEntropyLoss2d on random logits)
    g = torch.Generator().manual_seed(99)
    logits = torch.randn(2, 19, 16, 32, generator=g, dtype=torch.float64) * 3
    lab = fixture.make_labels(2, 16, 32, 19, seed=99)
    wt = torch.tensor(fixture.CLASS_WEIGHTS, dtype=torch.float64)
    logits.requires_grad_(True)
    l = CE(weight=wt, ignore_label=255)(logits, lab)
    l.backward()
    # ---- focal loss golden (reference FocalLoss2d on the same logits)
    from utils.losses.loss import FocalLoss2d
    lf = logits.detach().clone().requires_grad_(True)
    fl = FocalLoss2d(alpha=0.5, gamma=2, weight=wt, ignore_index=255)(lf, lab)
    fl.backward()
    np.savez_compressed(os.path.join(GOLD, "loss.npz"), logits=logits.detach().numpy(), labels=lab.numpy(),
                        loss=np.array([l.item()]), grad=logits.grad.numpy(),
                        focal_loss=np.array([fl.item()]), focal_grad=lf.grad.numpy())
    print("loss golden written", l.item(), "focal", fl.item())


if __name__ == "__main__":
    main()
