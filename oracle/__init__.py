"""CPU oracle for the segmentation encoder-decoder hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (plain torch
functional ops / numpy, fp32 or fp64) of the reference's forward path
(`/root/reference/model/*.py`, `utils/losses/loss.py`, the numpy argmax of
`test.py:79-82`).  It exists so that `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs can check and time the
reference arithmetic on a box where `/root/reference` does not exist.  Nothing
under `efficient-segmentation-networks_b200/` imports it; the product path
fails loudly when the CUDA library is missing instead of falling back here.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, produced in the build
container by `tools/make_golden.py` (which imports `/root/reference`
unmodified behind three import stubs) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` replays them.
"""
from . import fixture, nets, loss, pipeline  # noqa: F401
