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


# --------------------------------------------------------------------------- training-time augmentation (esn_augment_u8)
def test_train_augmentation_matches_reference_class(golden):
    """esn.augment.CityscapesTrainAugment against the UNMODIFIED CityscapesDataSet (tests/golden/augment.npz, 18 seeded
    samples: every scale factor, padding in no / one / both directions, both mirror states): with `random` / `np.random`
    seeded as the generator seeded them, the device crops and label crops are bit-identical."""
    import random
    from esn.augment import CityscapesTrainAugment
    g = golden("augment")
    for k in range(int(g["n_cases"][0])):
        i, ch, cw, seed = [int(v) for v in g["case%d" % k]]
        aug = CityscapesTrainAugment(crop_size=(ch, cw), mean=g["mean"], scale=True, mirror=True, ignore_label=255)
        random.seed(seed)
        np.random.seed(seed)
        x, y = aug([torch.from_numpy(g["image%d" % i]).cuda()], [torch.from_numpy(g["label%d" % i]).cuda()])
        assert np.array_equal(x[0].cpu().numpy(), g["x%d" % k]), k
        assert np.array_equal(y[0].cpu().numpy().astype(np.float32), g["y%d" % k]), k
    aug = CityscapesTrainAugment(crop_size=(32, 32), mean=g["mean"], scale=False, mirror=False)
    random.seed(5)
    np.random.seed(5)
    x, y = aug([torch.from_numpy(g["image0"]).cuda()], [torch.from_numpy(g["label0"]).cuda()])
    assert np.array_equal(x[0].cpu().numpy(), g["noscale_x"]) and np.array_equal(y[0].cpu().numpy().astype(np.float32), g["noscale_y"])


def test_train_augmentation_full_size_batch():
    """Cityscapes-sized inputs (1024 x 2048), crop 512 x 1024 (train.py's default input_size), a batch of 8 with mixed scale
    factors in ONE launch: every sample equals the oracle (oracle/pipeline.py, pinned on the reference class and on cv2)."""
    from esn.augment import CityscapesTrainAugment
    rng = np.random.RandomState(9)
    imgs = [rng.randint(0, 256, (1024, 2048, 3)).astype(np.uint8) for _ in range(2)]
    labs = [rng.randint(0, 19, (1024, 2048)).astype(np.uint8) for _ in range(2)]
    params = [(0.75, 0, 17, -1), (1.0, 300, 511, 1), (1.25, 768, 1536, -1), (1.5, 5, 2000, 1), (1.75, 1280, 0, -1), (2.0, 1536, 3072, 1),
              (0.75, 256, 512, 1), (None, 512, 1024, -1)]
    aug = CityscapesTrainAugment(crop_size=(512, 1024), mean=pipeline.CITYSCAPES_MEAN_BGR)
    gi = [torch.from_numpy(imgs[k % 2]).cuda() for k in range(8)]
    gl = [torch.from_numpy(labs[k % 2]).cuda() for k in range(8)]
    x, y = aug(gi, gl, params=params)
    assert x.shape == (8, 3, 512, 1024) and y.shape == (8, 512, 1024) and y.dtype == torch.int64
    for k, (f, ho, wo, fl) in enumerate(params):
        rx, ry = pipeline.train_item(imgs[k % 2], labs[k % 2], f, ho, wo, fl, (512, 1024))
        assert np.array_equal(x[k].cpu().numpy(), rx), k
        assert np.array_equal(y[k].cpu().numpy(), ry.astype(np.int64)), k
