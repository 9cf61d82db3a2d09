"""Evaluation metric on the device: esn_confusion_matrix (through esn.metric / the utils.metric.metric drop-in) against
the oracle -- integer work, so bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import metric as om

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "metric.npz")


@pytest.mark.parametrize("gt_dtype", [torch.int64, torch.uint8])
def test_confusion_matrix_matches_golden(gt_dtype):
    from esn.metric import ConfusionMatrix
    g = np.load(GOLD)
    nclass = int(g["nclass"])
    conf = ConfusionMatrix(nclass)
    for gt, pr in zip(g["gt"], g["pred"]):                   # accumulates over batches
        conf.add_batch(torch.from_numpy(pr).cuda(), torch.from_numpy(gt).to(gt_dtype).cuda())
    assert np.array_equal(conf.M, g["M"])
    mean, per_class, _ = conf.jaccard()
    assert abs(mean - float(g["meanIoU"])) < 1e-12
    assert np.allclose(per_class, g["per_class"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("shape,nclass", [((2, 512, 1024), 19), ((1, 37, 53), 19), ((3, 64, 64), 32), ((1, 1, 5), 2)])
def test_confusion_matrix_matches_oracle(shape, nclass):
    from esn.metric import ConfusionMatrix
    gen = torch.Generator().manual_seed(3)
    coarse = torch.randint(0, nclass, (shape[0], (shape[1] + 7) // 8, (shape[2] + 7) // 8), generator=gen)
    gt = coarse.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :shape[1], :shape[2]].contiguous()
    gt[torch.rand(shape, generator=gen) < 0.08] = 255
    pred = torch.where(torch.rand(shape, generator=gen) < 0.6, gt.clamp(max=nclass - 1),
                       torch.randint(0, nclass, shape, generator=gen)).to(torch.uint8)
    conf = ConfusionMatrix(nclass)
    conf.add_batch(pred.cuda(), gt.cuda())
    ref = om.confusion_matrix(gt.numpy(), pred.numpy(), nclass)
    assert np.array_equal(conf.M, ref)
    assert int(conf.M.sum()) == int((gt < nclass).sum())


def test_get_iou_drop_in():
    """utils.metric.metric.get_iou(data_list, classes) with the reference's [gt.flatten(), output.flatten()] items."""
    from utils.metric.metric import get_iou
    g = np.load(GOLD)
    data_list = [[gt.flatten(), pr.flatten()] for gt, pr in zip(g["gt"], g["pred"])]
    mean, per_class = get_iou(data_list, int(g["nclass"]))
    assert abs(mean - float(g["meanIoU"])) < 1e-12
    assert np.allclose(per_class, g["per_class"], rtol=0, atol=1e-12)
    data_list = [[torch.from_numpy(gt).cuda(), torch.from_numpy(pr).cuda()] for gt, pr in zip(g["gt"], g["pred"])]
    mean2, _ = get_iou(data_list, int(g["nclass"]))
    assert mean2 == mean
