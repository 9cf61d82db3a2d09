"""Shared measurement for the model-level parity tests and tools/parity_report.py: one net at one input size, CUDA path
(through the C ABI) against the CPU oracle (oracle/nets.py, fp32) -- fp32 logits, bf16 logits, fused argmax masks (raw and
margin-aware agreement) and, as the noise floor, torch's own bf16 autocast of the oracle graph on the GPU."""
import torch

from conftest import spec_state_dict
from oracle import fixture, nets


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def margin_mask(ref_logits, tol):
    top2 = torch.topk(ref_logits.double(), 2, dim=1).values
    return (top2[:, 0] - top2[:, 1]) > tol * top2[:, 0].abs().clamp_min(1e-6)


def measure(name, spec, n, h, w, fp32=True, bf16=True, autocast_floor=True, seed=1234):
    from builders.model_builder import build_model
    sd = spec_state_dict(spec, name)
    m = build_model(name, 19)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = fixture.make_input(n, h, w, seed=seed)
    out = {"net": name, "input": "%dx3x%dx%d" % (n, h, w)}
    with torch.no_grad():
        ref = nets.forward(name, sd, x)                    # CPU oracle, fp32
        ref_mask = torch.from_numpy(nets.argmax_mask(ref))
        safe = margin_mask(ref, 5e-2)
        out["margin_safe_fraction"] = safe.float().mean().item()
        xg = x.cuda()
        if fp32:
            logits, mask = m.predict_mask(xg, with_logits=True)
            out["fp32_rel_l2"] = rel_l2(logits.cpu(), ref)
            out["fp32_max_rel"] = ((logits.cpu() - ref).abs().max() / ref.abs().max()).item()
            out["fp32_argmax_raw"] = (mask.cpu() == ref_mask).float().mean().item()
            out["fp32_fused_argmax_equals_argmax_of_logits"] = bool(torch.equal(mask.long(), logits.argmax(1)))
            del logits, mask
        if bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits, mask = m.predict_mask(xg, with_logits=True)
            assert logits.dtype == torch.bfloat16
            agree = mask.cpu() == ref_mask
            out["bf16_rel_l2"] = rel_l2(logits.float().cpu(), ref)
            out["bf16_argmax_raw"] = agree.float().mean().item()
            out["bf16_argmax_margin_aware"] = agree[safe].float().mean().item()
            del logits, mask
        if autocast_floor:
            sdg = {k: v.cuda() for k, v in sd.items()}
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y_ac = nets.forward(name, sdg, xg).float().cpu()
            ac_agree = torch.from_numpy(nets.argmax_mask(y_ac)) == ref_mask
            out["torch_autocast_rel_l2"] = rel_l2(y_ac, ref)
            out["torch_autocast_argmax_raw"] = ac_agree.float().mean().item()
            out["torch_autocast_argmax_margin_aware"] = ac_agree[safe].float().mean().item()
    del m
    torch.cuda.empty_cache()
    return out
