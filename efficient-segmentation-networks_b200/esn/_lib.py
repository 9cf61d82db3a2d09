"""ctypes binding of libesn_sm100.so (C ABI: include/esn.h).

The library is the product: if it is missing this module raises at import
time -- there is no CPU or eager-PyTorch fallback behind it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ESN_LIB_PATH: load another build of the same C ABI (the test suite points it at libesn_sm100_testing.so)
LIB_PATH = os.environ.get("ESN_LIB_PATH") or os.path.join(_HERE, "libesn_sm100.so")

ESN_F32, ESN_BF16, ESN_U8, ESN_I64, ESN_I32 = 0, 1, 2, 3, 4
ESN_NHWC, ESN_NCHW = 0, 1
ACT_NONE, ACT_RELU, ACT_PRELU = 0, 1, 2
ERR_UNSUPPORTED = -3
EP_ACT_BEFORE_RESIDUAL, EP_RESIDUAL_FIRST = 1, 2      # EsnEpilogue.flags (include/esn.h)


class EsnTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("layout", C.c_int32),
                ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("c_stride", C.c_int32), ("_pad", C.c_int32)]


class EsnEpilogue(C.Structure):
    _fields_ = [("scale", C.c_void_p), ("shift", C.c_void_p), ("alpha", C.c_void_p),
                ("act", C.c_int32), ("flags", C.c_int32), ("residual", EsnTensor)]


class EsnConv(C.Structure):
    _fields_ = [("x", EsnTensor), ("y", EsnTensor), ("w", C.c_void_p),
                ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32),
                ("pad_h", C.c_int32), ("pad_w", C.c_int32), ("dil_h", C.c_int32), ("dil_w", C.c_int32),
                ("groups", C.c_int32), ("transposed", C.c_int32), ("cout_pad", C.c_int32),
                ("ep", EsnEpilogue)]


class EsnConvDual(C.Structure):
    _fields_ = [("conv", EsnConv), ("y2", EsnTensor), ("scale2", C.c_void_p), ("shift2", C.c_void_p),
                ("alpha2", C.c_void_p), ("act2", C.c_int32), ("store_y", C.c_int32)]


class EsnConvPair(C.Structure):
    _fields_ = [("x", EsnTensor), ("y", EsnTensor), ("w1", C.c_void_p), ("w2", C.c_void_p),
                ("taps", C.c_int32), ("dilation", C.c_int32), ("ep1", EsnEpilogue), ("ep2", EsnEpilogue)]


class EsnPool(C.Structure):
    _fields_ = [("x", EsnTensor), ("y", EsnTensor), ("ep", EsnEpilogue)]


class EsnStem(C.Structure):
    _fields_ = [("x", EsnTensor), ("y", EsnTensor), ("w", C.c_void_p), ("cconv", C.c_int32),
                ("with_pool", C.c_int32), ("ep", EsnEpilogue)]


class EsnBnFinalize(C.Structure):
    _fields_ = [("sums", C.c_void_p), ("count", C.c_int64), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("eps", C.c_float), ("momentum", C.c_float), ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("channels", C.c_int32), ("_pad", C.c_int32)]


class EsnBnBwd(C.Structure):
    _fields_ = [("x", EsnTensor), ("dy", EsnTensor), ("dx", EsnTensor), ("extra", EsnTensor),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("alpha", C.c_void_p), ("mean", C.c_void_p),
                ("invstd", C.c_void_p), ("sums", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
                ("dalpha", C.c_void_p), ("act", C.c_int32), ("train_stats", C.c_int32)]


class EsnBnTrainFwd(C.Structure):
    _fields_ = [("x", EsnTensor), ("y", EsnTensor), ("fin", EsnBnFinalize), ("alpha", C.c_void_p), ("barrier", C.c_void_p),
                ("act", C.c_int32), ("_pad", C.c_int32)]


class EsnUnpool(C.Structure):
    _fields_ = [("v", EsnTensor), ("idx", C.c_void_p), ("ext", EsnTensor), ("y", EsnTensor), ("alpha", C.c_void_p),
                ("act", C.c_int32), ("_pad", C.c_int32)]


class EsnFGlo(C.Structure):
    _fields_ = [("sums", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("gate", C.c_void_p), ("n", C.c_int32), ("channels", C.c_int32), ("hidden", C.c_int32), ("hw", C.c_int32),
                ("chunks", C.c_int32), ("_pad", C.c_int32)]


class EsnDabPair(C.Structure):
    _fields_ = [("x", EsnTensor), ("y", EsnTensor), ("prm", C.c_void_p),
                ("dilation", C.c_int32), ("_pad", C.c_int32)]


class EsnHead(C.Structure):
    _fields_ = [("x", EsnTensor), ("w", C.c_void_p), ("bias", C.c_void_p), ("logits", EsnTensor),
                ("mask", C.c_void_p), ("classes", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
                ("align_corners", C.c_int32)]


class EsnBneck4(C.Structure):
    _fields_ = [("x", EsnTensor), ("y", EsnTensor), ("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p),
                ("scale1", C.c_void_p), ("shift1", C.c_void_p), ("alpha1", C.c_void_p),
                ("scale2", C.c_void_p), ("shift2", C.c_void_p), ("alpha2", C.c_void_p),
                ("scale3", C.c_void_p), ("shift3", C.c_void_p), ("alpha3", C.c_void_p),
                ("dilation", C.c_int32), ("act", C.c_int32)]


class EsnHeadT3(C.Structure):
    _fields_ = [("x", EsnTensor), ("wfrag", C.c_void_p), ("bias", C.c_void_p), ("mask", C.c_void_p),
                ("classes", C.c_int32), ("_pad", C.c_int32)]


class EsnAugItem(C.Structure):
    _fields_ = [("img", C.c_void_p), ("label", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("rh", C.c_int32), ("rw", C.c_int32),
                ("scale", C.c_double), ("h_off", C.c_int32), ("w_off", C.c_int32), ("flip", C.c_int32), ("do_scale", C.c_int32)]


class EsnCE(C.Structure):
    _fields_ = [("logits", EsnTensor), ("target", C.c_void_p), ("weight", C.c_void_p), ("sums", C.c_void_p),
                ("dlogits", EsnTensor), ("ignore_label", C.c_int32), ("_pad", C.c_int32),
                ("gnorm", C.c_void_p), ("gout", C.c_void_p), ("prob_out", C.c_void_p), ("keep_thresh", C.c_void_p)]


class EsnBilinearCE(C.Structure):
    _fields_ = [("scores", EsnTensor), ("target", C.c_void_p), ("weight", C.c_void_p), ("sums", C.c_void_p),
                ("dscores", EsnTensor), ("out_h", C.c_int32), ("out_w", C.c_int32), ("ignore_label", C.c_int32),
                ("align_corners", C.c_int32)]


# every symbol include/esn.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "esn_conv2d_direct": (C.c_int, [C.POINTER(EsnConv), C.c_void_p]),
    "esn_conv2d_umma": (C.c_int, [C.POINTER(EsnConv), C.c_void_p]),
    "esn_conv2d_umma_dual": (C.c_int, [C.POINTER(EsnConvDual), C.c_void_p]),
    "esn_conv_pair_umma": (C.c_int, [C.POINTER(EsnConvPair), C.c_void_p]),
    "esn_stem_conv3x3s2": (C.c_int, [C.POINTER(EsnStem), C.c_void_p]),
    "esn_maxpool2x2_affine_act": (C.c_int, [C.POINTER(EsnPool), C.c_void_p]),
    "esn_avgpool3x3s2_affine_act": (C.c_int, [C.POINTER(EsnPool), C.c_void_p]),
    "esn_affine_act": (C.c_int, [C.POINTER(EsnPool), C.c_void_p]),
    "esn_concat_tail": (C.c_int, [C.POINTER(EsnPool), C.c_int32, C.c_void_p]),
    "esn_convert_layout": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_void_p]),
    "esn_dab_dw_pair": (C.c_int, [C.POINTER(EsnDabPair), C.c_void_p]),
    "esn_head_convt2x2": (C.c_int, [C.POINTER(EsnHead), C.c_void_p]),
    "esn_head_bilinear": (C.c_int, [C.POINTER(EsnHead), C.c_void_p]),
    "esn_head_convt3x3s2_mask": (C.c_int, [C.POINTER(EsnHeadT3), C.c_void_p]),
    "esn_head_convt2x2_mask": (C.c_int, [C.POINTER(EsnHeadT3), C.c_void_p]),
    "esn_bottleneck4": (C.c_int, [C.POINTER(EsnBneck4), C.c_void_p]),
    "esn_weighted_ce": (C.c_int, [C.POINTER(EsnCE), C.c_void_p]),
    "esn_augment_max_batch": (C.c_int32, []),
    "esn_augment_u8": (C.c_int, [C.POINTER(EsnAugItem), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "esn_dot_nc": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_void_p, C.c_void_p]),
    "esn_scale_add_nc": (C.c_int, [C.POINTER(EsnTensor), C.c_void_p, C.c_void_p, C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_void_p]),
    "esn_maxpool3x3s2_idx_bwd": (C.c_int, [C.POINTER(EsnTensor), C.c_void_p, C.POINTER(EsnTensor), C.c_int32, C.c_void_p]),
    "esn_max_unpool2x2_bwd": (C.c_int, [C.POINTER(EsnTensor), C.c_void_p, C.POINTER(EsnTensor), C.c_void_p]),
    "esn_ohem_workspace_bytes": (C.c_int64, []),
    "esn_ohem_threshold": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esn_channel_stats": (C.c_int, [C.POINTER(EsnTensor), C.c_void_p, C.c_int32, C.c_void_p]),
    "esn_bn_finalize": (C.c_int, [C.POINTER(EsnBnFinalize), C.c_void_p]),
    "esn_bn_act_bwd_reduce": (C.c_int, [C.POINTER(EsnBnBwd), C.c_void_p]),
    "esn_bn_act_bwd_apply": (C.c_int, [C.POINTER(EsnBnBwd), C.c_void_p]),
    "esn_bn_act_train_fwd": (C.c_int, [C.POINTER(EsnBnTrainFwd), C.c_void_p]),
    "esn_bn_act_bwd_fused": (C.c_int, [C.POINTER(EsnBnBwd), C.c_void_p, C.c_void_p]),
    "esn_act_bwd": (C.c_int, [C.POINTER(EsnBnBwd), C.c_void_p]),
    "esn_conv2d_wgrad": (C.c_int, [C.POINTER(EsnConv), C.c_void_p]),
    "esn_wgrad_umma_supported": (C.c_int, [C.POINTER(EsnConv)]),
    "esn_maxpool2x2_bwd": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_int32, C.c_void_p]),
    "esn_bilinear_bwd": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_float, C.c_void_p]),
    "esn_bilinear_bwd_nhwc": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_int32, C.c_int32, C.c_void_p]),
    "esn_adaptive_avgpool_bwd": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_int32, C.c_void_p]),
    "esn_avgpool3x3s2_bwd": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_int32, C.c_void_p]),
    "esn_dropout": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_uint64, C.c_float, C.c_int32, C.c_void_p]),
    "esn_confusion_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "esn_gate_bcast": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_void_p]),
    "esn_image_u8hwc_to_f32nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float),
                                             C.c_int32, C.c_void_p]),
    "esn_dropout_step": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_uint64, C.c_void_p, C.c_float, C.c_int32,
                                   C.c_void_p]),
    "esn_dropout_mask_nc": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p, C.c_float, C.c_void_p]),
    "esn_maxpool3x3s2_idx": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_void_p, C.c_void_p]),
    "esn_max_unpool2x2": (C.c_int, [C.POINTER(EsnUnpool), C.c_void_p]),
    "esn_global_avgpool_chunks": (C.c_int, [C.POINTER(EsnTensor)]),
    "esn_global_avgpool": (C.c_int, [C.POINTER(EsnTensor), C.c_void_p, C.c_void_p]),
    "esn_fglo_gate": (C.c_int, [C.POINTER(EsnFGlo), C.c_void_p]),
    "esn_scale_nc": (C.c_int, [C.POINTER(EsnTensor), C.c_void_p, C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_void_p]),
    "esn_adaptive_avgpool": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_void_p]),
    "esn_bilinear_nhwc": (C.c_int, [C.POINTER(EsnTensor), C.POINTER(EsnTensor), C.c_int32, C.c_void_p]),
    "esn_bilinear_ce": (C.c_int, [C.POINTER(EsnBilinearCE), C.c_void_p]),
    "esn_adam_chunk": (C.c_int32, []),
    "esn_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_void_p]),
    "esn_version": (C.c_int, []),
    "esn_strerror": (C.c_char_p, [C.c_int]),
    "esn_launch_count": (C.c_int64, []),
    "esn_launch_count_reset": (None, []),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libesn_sm100.so not found at %s -- build it with `python __graft_entry__.py build` "
            "(or `make -C efficient-segmentation-networks_b200/csrc`). There is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class EsnError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise EsnError("%s failed: %s (code %d)" % (what, lib.esn_strerror(rc).decode(), rc))
