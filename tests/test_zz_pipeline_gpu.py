"""Input pipeline kernel through the C-ABI entry esn_image_u8hwc_to_f32nchw: bit-exact against the
oracle (oracle/pipeline.py, pinned on the reference dataset classes) and the golden fixture.

Bit-exact on B200 (round-1 driver run)."""
import numpy as np
import pytest
import torch

from oracle import pipeline

pytestmark = pytest.mark.gpu


def _run(imgs, mean, reverse=True):
    from esn import ops
    return ops.image_u8_to_f32(torch.from_numpy(imgs).cuda(), mean, reverse).cpu().numpy()


def test_matches_reference_golden(golden):
    g = golden("pipeline")
    for i in range(3):
        y = _run(g["image%d" % i][None], g["mean"])
        assert np.array_equal(y[0], g["input%d" % i])


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 3, 5), (2, 17, 23), (3, 32, 32), (2, 33, 64), (5, 31, 33), (2, 64, 128)])
def test_ragged_sizes_bit_exact(shape):
    n, h, w = shape
    rng = np.random.RandomState(n * 1000 + h * 10 + w)
    imgs = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    assert np.array_equal(_run(imgs, pipeline.CITYSCAPES_MEAN_BGR), pipeline.batch_to_input(imgs))
    mean = np.array([128, 128, 128], dtype=np.float32)                # the dataset classes' default mean
    same_order = _run(imgs, mean, reverse=False)
    assert np.array_equal(same_order, (imgs.astype(np.float32) - mean).transpose(0, 3, 1, 2))


def test_empty_batch_and_errors():
    from esn import ops
    y = ops.image_u8_to_f32(torch.empty((0, 4, 4, 3), dtype=torch.uint8, device="cuda"), pipeline.CITYSCAPES_MEAN_BGR)
    assert y.shape == (0, 3, 4, 4)
    with pytest.raises(TypeError):
        ops.image_u8_to_f32(torch.zeros((1, 4, 4, 3), device="cuda"), pipeline.CITYSCAPES_MEAN_BGR)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.image_u8_to_f32(torch.zeros((1, 4, 4, 3), dtype=torch.uint8), pipeline.CITYSCAPES_MEAN_BGR)


def test_full_size_properties():
    """16 x 1024 x 2048 (BASELINE configs[1]), where the numpy oracle is slow: bit-equal to the same fp32 arithmetic done
    by torch on the device (one IEEE subtraction of exactly representable integers, so every implementation agrees)."""
    from esn import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    img = torch.randint(0, 256, (16, 1024, 2048, 3), dtype=torch.uint8, device="cuda", generator=g)
    mean = torch.tensor(pipeline.CITYSCAPES_MEAN_BGR, device="cuda")
    y = ops.image_u8_to_f32(img, pipeline.CITYSCAPES_MEAN_BGR)
    ref = (img.float() - mean).flip(3).permute(0, 3, 1, 2)
    assert torch.equal(y, ref)
    assert y.is_contiguous() and y.shape == (16, 3, 1024, 2048)
