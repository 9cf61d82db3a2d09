"""Per-module cache of packed parameters (weights re-laid-out for the kernels, BN folded).

A block builds its "prep" lazily on first eval-mode forward and rebuilds it when
any parameter/buffer was modified in place (tensor `_version`), replaced
(`load_state_dict` keeps identity but bumps the version; `.to()` / `.cuda()`
replace the tensor), or when BN eps changed (`init_weight` rewrites eps:
utils/utils.py:16-19 of the reference).
"""
import torch

# Parameters can also change WITHOUT any ATen op being dispatched: a replay of a captured training iteration
# (esn/graph.py) updates weights, BN running statistics and num_batches_tracked on the device while every tensor's
# `_version` stays where it was at capture time.  Every cache keyed on `_version` therefore also carries this
# process-wide generation number, which GraphedTrainStep bumps on each replay.
_GENERATION = 0


def weights_generation():
    return _GENERATION


def bump_weights_generation():
    global _GENERATION
    _GENERATION += 1


def _signature(module):
    sig = [_GENERATION]
    for t in list(module.parameters()) + list(module.buffers()):
        sig.append((t.data_ptr(), t._version, t.dtype))
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            sig.append(m.eps)
    return tuple(sig)


class PrepMixin:
    """Adds `self.prep(device)`; subclasses implement `_build_prep(device)`."""

    def prep(self, device):
        sig = (str(device), _signature(self))
        cached = self.__dict__.get("_esn_prep")
        if cached is None or cached[0] != sig:
            with torch.no_grad():
                cached = (sig, self._build_prep(device))
            self.__dict__["_esn_prep"] = cached
        return cached[1]
