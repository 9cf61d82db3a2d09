"""Input pipeline: the oracle against the unmodified reference dataset classes' outputs."""
import numpy as np

from oracle import pipeline


def test_oracle_matches_reference_dataset_classes(golden):
    g = golden("pipeline")
    assert g["mean"].dtype == np.float32 and np.array_equal(g["mean"], pipeline.CITYSCAPES_MEAN_BGR)
    for i in range(3):
        x = pipeline.image_to_input(g["image%d" % i], g["mean"])
        ref = g["input%d" % i]
        assert x.dtype == ref.dtype == np.float32 and x.shape == ref.shape
        assert np.array_equal(x, ref)                      # bit-exact
    # fixture.MEAN_RGB (bench / model tests) is this mean, reversed and rounded to 3 decimals
    from oracle import fixture
    assert np.allclose(np.asarray(fixture.MEAN_RGB), g["mean"][::-1], atol=1e-3)


def test_batch_layout():
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, (2, 6, 10, 3)).astype(np.uint8)
    x = pipeline.batch_to_input(imgs)
    assert x.shape == (2, 3, 6, 10)
    assert x[1, 0, 2, 3] == np.float32(imgs[1, 2, 3, 2]) - pipeline.CITYSCAPES_MEAN_BGR[2]     # R plane = BGR channel 2
