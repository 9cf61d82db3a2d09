"""build_model(model_name, num_classes) -- same signature and names as the reference's
builders/model_builder.py:21-59.  Models on the B200 hot path are built from the CUDA-backed
modules under model/; names of the reference zoo that are outside the hot-path scope raise
NotImplementedError (the reference silently returns None for unknown names; there is no CPU or
eager fallback here, so the failure is explicit).
"""
from model.ERFNet import ERFNet
from model.DABNet import DABNet
from model.ENet import ENet
from model.CGNet import CGNet
from model.FastSCNN import FastSCNN
from model.ESPNet import ESPNet
from model.ESPNet_v2.SegmentationModel import EESPNet_Seg
from model.ESNet import ESNet
from model.ContextNet import ContextNet
from model.EDANet import EDANet
from model.LEDNet import LEDNet

_HOT_PATH = {
    "ERFNet": ERFNet,
    "DABNet": DABNet,
    "ENet": ENet,
    "CGNet": CGNet,
    "FastSCNN": FastSCNN,
    "ESPNet_v2": EESPNet_Seg,
    "ESPNet": ESPNet,
    # nets that reuse the ERFNet / Fast-SCNN kernels (inference)
    "ESNet": ESNet,
    "ContextNet": ContextNet,
    "EDANet": EDANet,
    "LEDNet": LEDNet,
}
_REFERENCE_NAMES = ("SQNet", "LinkNet", "SegNet", "UNet", "ENet", "ERFNet", "CGNet", "EDANet", "ESNet", "ESPNet",
                    "LEDNet", "ESPNet_v2", "ContextNet", "FastSCNN", "DABNet", "FSSNet", "FPENet", "DF1Seg", "DF1SegG")


def build_model(model_name, num_classes):
    if model_name in _HOT_PATH:
        return _HOT_PATH[model_name](classes=num_classes)
    if model_name in _REFERENCE_NAMES:
        raise NotImplementedError("%s is not (yet) on the B200 hot path; available: %s"
                                  % (model_name, ", ".join(sorted(_HOT_PATH))))
    raise NotImplementedError("unknown model name %r" % (model_name,))
