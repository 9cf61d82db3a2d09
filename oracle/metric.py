"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of the reference's evaluation arithmetic:
`ConfusionMatrix.generateM` (utils/metric/metric.py:68-76) and `ConfusionMatrix.jaccard` (:58-66), as `get_iou` (:79-106)
combines them for test.py:90 / train.py:404.  Pinned against the unmodified reference class on seeded label / prediction
arrays (tools/make_golden_metric.py -> tests/golden/metric.npz)."""
import numpy as np


def confusion_matrix(gt, pred, nclass):
    """m[gt[i], pred[i]] += 1 for every i with gt[i] < nclass (metric.py:73-75), vectorised."""
    gt = np.asarray(gt).reshape(-1).astype(np.int64)
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    assert gt.shape == pred.shape
    keep = gt < nclass
    m = np.bincount(gt[keep] * nclass + pred[keep], minlength=nclass * nclass)
    return m.reshape(nclass, nclass).astype(np.float64)


def jaccard(M):
    """metric.py:58-66."""
    per_class = []
    for i in range(M.shape[0]):
        if not M[i, i] == 0:
            per_class.append(M[i, i] / (np.sum(M[i, :]) + np.sum(M[:, i]) - M[i, i]))
    return np.sum(per_class) / len(per_class), per_class
