"""Seeded fixture recipe shared by the golden generator, the tests and bench.py.

Mirrors the reference's own init path only in spirit: every
tensor of a model's ``state_dict`` is regenerated from a per-key seeded
generator so the reference (in the build container) and the CUDA modules (on
the GPU box) can be given bit-identical weights without shipping them.

* conv / linear weights: N(0, sqrt(2 / fan_in))   (kaiming_normal fan_in, as
  `train.py:120-122` applies through `utils/utils.py:10-30`)
* conv biases: N(0, 0.05)
* BN weight U(0.5, 1.5), bias N(0, 0.1), running_mean N(0, 0.1),
  running_var U(0.5, 1.5)  -- non-trivial so folded-BN paths are exercised
* PReLU alpha U(0.05, 0.45)
"""
import zlib

import numpy as np
import torch

MEAN_RGB = (73.158, 82.909, 72.392)  # dataset/cityscapes.py:74-78 applied to the pickle's BGR mean

# dataset/cityscapes.py:21-24 (documented class weights) -- config 3 of BASELINE.json
CLASS_WEIGHTS = (2.5959933, 6.7415504, 3.5354059, 9.8663225, 9.690899, 9.369352,
                 10.289121, 9.953208, 4.3097677, 9.490387, 7.674431, 9.396905,
                 10.347791, 6.3927646, 10.226669, 10.241062, 10.280587,
                 10.396974, 10.055647)


def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def randomize_state_dict(sd, seed=1234):
    """Return a new state_dict with the same keys/shapes/dtypes, seeded values."""
    out = {}
    keys = list(sd.keys())
    keyset = set(keys)
    for k in keys:
        v = sd[k]
        g = _gen(k, seed)
        prefix, _, leaf = k.rpartition(".")
        is_bn = (prefix + ".running_mean") in keyset
        if leaf == "num_batches_tracked":
            out[k] = torch.zeros_like(v)
        elif is_bn and leaf == "weight":
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif is_bn and leaf == "bias":
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif leaf == "running_mean":
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif leaf == "running_var":
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif leaf == "weight" and v.dim() >= 2:
            fan_in = int(np.prod(v.shape[1:]))
            out[k] = torch.randn(v.shape, generator=g) * float(np.sqrt(2.0 / fan_in))
        elif leaf == "weight" and v.dim() == 1:          # PReLU alpha
            out[k] = torch.rand(v.shape, generator=g) * 0.4 + 0.05
        elif leaf == "bias":
            out[k] = torch.randn(v.shape, generator=g) * 0.05
        else:
            raise KeyError("fixture: unclassified state_dict key %r" % k)
        out[k] = out[k].to(v.dtype)
    return out


def make_input(n, h, w, seed=1234):
    """Cityscapes-shaped synthetic batch: uint8 pixels minus the RGB mean, NCHW fp32."""
    g = torch.Generator()
    g.manual_seed(seed)
    x = torch.randint(0, 256, (n, 3, h, w), generator=g).float()
    return x - torch.tensor(MEAN_RGB).view(1, 3, 1, 1)


def make_labels(n, h, w, classes=19, ignore_frac=0.08, ignore_label=255, seed=1234):
    g = torch.Generator()
    g.manual_seed(seed + 7)
    y = torch.randint(0, classes, (n, h, w), generator=g)
    m = torch.rand((n, h, w), generator=g) < ignore_frac
    y[m] = ignore_label
    return y
