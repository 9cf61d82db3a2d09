"""Train-mode forward of ENet on the training kernels (esn/train.py): same graph as the reference's ENet.forward
(model/ENet.py:14-432) with batch-statistics BatchNorm and Dropout2d, recording the backward on a tape.

What is specific to ENet:
* one activation module per block is shared by all its call sites (ENet.py:53-89): for the encoder's nn.PReLU() that is ONE
  slope whose gradient is the sum over every use and every channel -- collected in a SharedParamGrad and handed to the tape
  once per block;
* MaxPool2d(3, 2, 1, return_indices=True) / MaxUnpool2d(2): forward kernels of the inference path (int32 indices, deterministic
  gather un-pool), backward = esn_maxpool3x3s2_idx_bwd / esn_max_unpool2x2_bwd;
* the zero channel padding of the down-sampling main branch (ENet.py:178-190) is never materialised: the add touches the first
  Cin channels, the rest of the extension branch is copied (forward) / passed through (backward).
"""
import torch
import torch.nn as nn

from esn import ops
from esn import train as T
from esn._lib import ACT_NONE, ACT_PRELU, ACT_RELU


def _convT(conv):
    t = conv.__dict__.get("_esn_T")
    if t is None:
        t = T.ConvTransposeT(conv) if isinstance(conv, nn.ConvTranspose2d) else T.ConvT(conv)
        conv.__dict__["_esn_T"] = t
    return t


class _Act:
    """The block's shared activation: (act code, PReLU module or None, gradient sink)."""

    def __init__(self, tape, activation):
        if isinstance(activation, nn.PReLU):
            self.code, self.prelu = ACT_PRELU, activation
            self.sink = T.SharedParamGrad(tape, activation.weight)        # pushed before the block's ops: flushes after them
        else:
            self.code, self.prelu, self.sink = ACT_RELU, None, None

    def bn_act(self, tape, bn, x, out=None):
        return T.BNActT(bn, self.code, self.prelu, self.sink).forward(tape, x, out=out)

    def act(self, tape, x):
        return T.BNActT(None, self.code, self.prelu, self.sink).forward(tape, x)


def _cba(tape, a, seq, i, x):
    """seq[i] conv -> seq[i+1] BatchNorm -> the block's activation."""
    return a.bn_act(tape, seq[i + 1], _convT(seq[i]).forward(tape, x))


def _dropout(tape, m, x):
    return T.dropout(tape, x, m.p, per_channel=True) if m.p > 0 else x


def _initial(tape, m, image, dt):
    """InitialBlock (ENet.py:14-44): cat[conv3x3 s2 (3 -> 13), MaxPool2d(3, 2, 1)] -> BN -> act; conv and pool in one pass of
    the stem kernel over the NCHW image (no gradient flows to the image)."""
    a = _Act(tape, m.out_prelu)
    ct = _convT(m.main_branch)
    fwd_prep, _ = ct.preps()
    n, _, h, w = image.shape
    nc = fwd_prep.cout
    cat = T.V(ops.new_act(n, nc + 3, (h - 1) // 2 + 1, (w - 1) // 2 + 1, dt, image.device))
    ones = torch.ones(nc + 3, device=image.device)
    shift = torch.zeros(nc + 3, device=image.device)
    shift[:nc] = fwd_prep.shift                       # the conv bias, if any
    ops.stem_conv3x3s2(image, fwd_prep.w_direct, nc, 2, cat.t, ones, shift, None, ACT_NONE)
    ct.forward(tape, T.V(image), out=cat.slice(0, nc), need_dx=False, dtype=dt, precomputed=True)
    return a.bn_act(tape, m.batch_norm, cat)


def _regular(tape, m, x):
    """RegularBottleneck (ENet.py:46-100)."""
    a = _Act(tape, m.out_prelu)
    e = _cba(tape, a, m.ext_conv1, 0, x)
    for i in range(0, len(m.ext_conv2), 3):
        e = _cba(tape, a, m.ext_conv2, i, e)
    e = _cba(tape, a, m.ext_conv3, 0, e)
    e = _dropout(tape, m.ext_regu1, e)
    return a.act(tape, T.add(tape, x, e))


def _down(tape, m, x):
    """DownsamplingBottleneck (ENet.py:102-197): act(cat(maxpool(x), zeros) + ext)."""
    a = _Act(tape, m.out_prelu)
    main, idx = T.maxpool3x3s2_idx(tape, x)
    e = _cba(tape, a, m.ext_conv1, 0, x)
    e = _cba(tape, a, m.ext_conv2, 0, e)
    e = _cba(tape, a, m.ext_conv3, 0, e)
    e = _dropout(tape, m.ext_regul, e)
    n, cout, h, w = e.t.shape
    cin = main.t.shape[1]
    s = T.V(ops.new_act(n, cout, h, w, e.t.dtype, e.t.device))
    T.add(tape, main, e.slice(0, cin), out=s.slice(0, cin))
    T.copy_into(tape, e.slice(cin, cout), s.slice(cin, cout))
    return a.act(tape, s), idx


def _up(tape, m, x, idx):
    """UpsamplingBottleneck (ENet.py:199-272): act(unpool(BN(conv1x1(x)), idx) + ext)."""
    a = _Act(tape, m.out_prelu)
    main = T.BNActT(m.main_conv1[1], ACT_NONE).forward(tape, _convT(m.main_conv1[0]).forward(tape, x))
    main = T.max_unpool2x2(tape, main, idx)
    e = _cba(tape, a, m.ext_conv1, 0, x)
    e = _cba(tape, a, m.ext_conv2, 0, e)
    e = _cba(tape, a, m.ext_conv3, 0, e)
    e = _dropout(tape, m.ext_regul, e)
    return a.act(tape, T.add(tape, main, e))


def enet_train_forward(model, input):
    ops.require_cuda(input, "ENet")
    if input.dtype != torch.float32 or not input.is_contiguous():
        input = input.float().contiguous()
    n, _, H, W = input.shape
    if (H | W) & 7:
        raise ValueError("ENet: input height and width must be multiples of 8, got %dx%d" % (H, W))
    dt = ops.compute_dtype(input)
    tape = T.Tape(model.__dict__.get("_esn_buckets"), device=input.device)
    m = model
    x = _initial(tape, m.initial_block, input, dt)
    x, i1 = _down(tape, m.downsample1_0, x)
    for blk in (m.regular1_1, m.regular1_2, m.regular1_3, m.regular1_4):
        x = _regular(tape, blk, x)
    x, i2 = _down(tape, m.downsample2_0, x)
    for blk in (m.regular2_1, m.dilated2_2, m.asymmetric2_3, m.dilated2_4, m.regular2_5, m.dilated2_6, m.asymmetric2_7,
                m.dilated2_8, m.regular3_0, m.dilated3_1, m.asymmetric3_2, m.dilated3_3, m.regular3_4, m.dilated3_5,
                m.asymmetric3_6, m.dilated3_7):
        x = _regular(tape, blk, x)
    x = _up(tape, m.upsample4_0, x, i2)
    x = _regular(tape, m.regular4_2, _regular(tape, m.regular4_1, x))
    x = _up(tape, m.upsample5_0, x, i1)
    x = _regular(tape, m.regular5_1, x)
    classes = m.transposed_conv.out_channels
    scores = T.V(ops.new_act(n, classes, H, W, dt, input.device, c_alloc=(classes + 7) // 8 * 8))
    _convT(m.transposed_conv).forward(tape, x, out=scores)              # (N, classes, H, W) NHWC, 8-channel-aligned pixels
    logits, holder = T.bilinear_logits(tape, scores, H, W, torch.float32)     # same size: NHWC -> NCHW fp32 logits
    return logits, tape, holder
