"""CPU restatement of the reference's image pre-processing tail (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The dataset classes of the reference finish every sample the same way (dataset/cityscapes.py:74-78 train, :164-170
val, :208-214 test; the CamVid twins are identical): the decoded uint8 BGR image becomes float32, the per-channel mean
(BGR order; float32 in dataset/inform/cityscapes_inform.pkl) is subtracted, the channel axis is reversed to RGB and
moved to the front.  Pinned on tests/golden/pipeline.npz (tools/make_golden_pipeline.py: the unmodified
CityscapesValDataSet / CityscapesTestDataSet run on PNG files written by the generator).
"""
import numpy as np

CITYSCAPES_MEAN_BGR = np.array([72.3924, 82.90902, 73.158325], dtype=np.float32)   # cityscapes_inform.pkl['mean']


def image_to_input(image_bgr_u8, mean_bgr=CITYSCAPES_MEAN_BGR):
    """(H,W,3) uint8 BGR -> (3,H,W) float32 RGB, mean-subtracted (float32 arithmetic, as numpy does for an fp32 mean)."""
    x = image_bgr_u8.astype(np.float32) - np.asarray(mean_bgr, dtype=np.float32).reshape(1, 1, 3)
    return np.ascontiguousarray(x[:, :, ::-1].transpose(2, 0, 1))


def batch_to_input(images_bgr_u8, mean_bgr=CITYSCAPES_MEAN_BGR):
    """(N,H,W,3) uint8 -> (N,3,H,W) float32: what the DataLoader's default collate stacks."""
    return np.stack([image_to_input(im, mean_bgr) for im in images_bgr_u8])
