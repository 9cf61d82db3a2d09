"""Evaluation metric, CPU side: the oracle restatement against golden vectors produced by the unmodified
reference class (tools/make_golden_metric.py)."""
import os

import numpy as np

from oracle import metric as om

GOLD = os.path.join(os.path.dirname(__file__), "golden", "metric.npz")


def test_oracle_confusion_matrix_and_miou_match_reference():
    g = np.load(GOLD)
    nclass = int(g["nclass"])
    M = sum(om.confusion_matrix(gt, pr, nclass) for gt, pr in zip(g["gt"], g["pred"]))
    assert np.array_equal(M, g["M"])
    mean, per_class = om.jaccard(M)
    assert len(per_class) == len(g["per_class"]) == nclass - 1          # class 17 never occurs
    assert abs(mean - float(g["meanIoU"])) < 1e-12
    assert np.allclose(per_class, g["per_class"], rtol=0, atol=1e-12)
    assert int(M.sum()) == int((g["gt"] < nclass).sum())                # ignore label skipped, nothing else


def test_host_jaccard_matches_reference():
    """esn.metric.jaccard_from_matrix (the host half of the device metric) on the reference's golden matrix."""
    from esn.metric import jaccard_from_matrix
    g = np.load(GOLD)
    mean, per_class, M = jaccard_from_matrix(g["M"])
    assert mean == float(g["meanIoU"])
    assert np.array_equal(np.asarray(per_class), g["per_class"])
    assert M is not None and M.shape == (int(g["nclass"]),) * 2
