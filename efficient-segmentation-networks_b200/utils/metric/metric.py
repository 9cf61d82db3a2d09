"""Drop-in for the reference's `utils/metric/metric.py` on the evaluation path (`test.py:90`, `train.py:404`):
`get_iou(data_list, class_num)` with the confusion matrix accumulated on the GPU (`esn.metric`).  Elements of
`data_list` are `[gt, output]` pairs as the reference builds them; CUDA tensors are used in place, numpy arrays are
uploaded (uint8 predictions, uint8 / int64 labels)."""
import numpy as np
import torch

from esn.metric import ConfusionMatrix


def _cuda(a, want_u8):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if want_u8 and a.dtype != torch.uint8:
        a = a.to(torch.uint8)
    elif not want_u8 and a.dtype not in (torch.uint8, torch.int64):
        a = a.to(torch.int64)
    return a.cuda(non_blocking=True)


def get_iou(data_list, class_num, save_path=None):
    conf = ConfusionMatrix(class_num)
    for gt, pred in data_list:
        conf.add_batch(_cuda(pred, True), _cuda(gt, False))
    aveJ, j_list, M = conf.jaccard()
    if save_path:
        with open(save_path, 'w') as f:
            f.write('meanIOU: ' + str(aveJ) + '\n')
            f.write(str(j_list) + '\n')
            f.write(str(M) + '\n')
    return aveJ, j_list
