"""Functional CPU restatement of the reference forward passes (eval and train mode).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Each function takes a
``state_dict`` with the reference's key names and computes what the reference
``nn.Module`` computes, citing the file:line it follows (paths relative to
/root/reference).  dtype follows the input (fp32 or fp64).
"""
import torch
import torch.nn.functional as F


class SD:
    """state_dict view with a key prefix; casts to the working dtype."""

    def __init__(self, sd, prefix="", dtype=torch.float32, train=False, stats=None):
        self.sd, self.prefix, self.dtype, self.train = sd, prefix, dtype, train
        self.stats = stats if stats is not None else {}

    def sub(self, name):
        return SD(self.sd, self.prefix + name + ".", self.dtype, self.train, self.stats)

    def __getitem__(self, k):
        return self.sd[self.prefix + k].to(self.dtype)

    def has(self, k):
        return (self.prefix + k) in self.sd


# bench.py's GPU-eager arm sets this: BatchNorm and PReLU then go through F.batch_norm / F.prelu, the single fused ATen calls
# that the reference's nn.BatchNorm2d / nn.PReLU make (cuDNN on CUDA), instead of the written-out statements below -- same
# arithmetic, the reference's own launch count.  Parity tests leave it False.
REFERENCE_ATEN_CALLS = False


def bn(p, x, eps):
    """nn.BatchNorm2d: eval uses running stats; train uses biased batch var
    (torch semantics).  Batch stats are recorded in p.stats."""
    if REFERENCE_ATEN_CALLS:
        if p.train:
            return F.batch_norm(x, None, None, p["weight"], p["bias"], True, 0.1, eps)
        return F.batch_norm(x, p["running_mean"], p["running_var"], p["weight"], p["bias"], False, 0.1, eps)
    if p.train:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        p.stats[p.prefix] = (mean.detach(), var.detach())
    else:
        mean, var = p["running_mean"], p["running_var"]
    scale = p["weight"] / torch.sqrt(var + eps)
    return x * scale.view(1, -1, 1, 1) + (p["bias"] - mean * scale).view(1, -1, 1, 1)


def prelu(x, alpha):
    if REFERENCE_ATEN_CALLS:
        return F.prelu(x, alpha.reshape(-1))
    return torch.clamp(x, min=0) + alpha.view(1, -1, 1, 1) * torch.clamp(x, max=0)


# --------------------------------------------------------------------------- ERFNet
def erf_downsampler(p, x):
    """DownsamplerBlock, model/ERFNet.py:16-27; for odd sizes the pooled map is zero-padded at the bottom / right up to the
    conv's ceil(H/2) rows, as ESNet's and LEDNet's copies of the block do (model/ESNet.py:22-29; ERFNet's own copy raises)."""
    y2 = F.conv2d(x, p["conv.weight"], p["conv.bias"], stride=2, padding=1)
    y1 = F.max_pool2d(x, 2, 2)
    dy, dx = y2.shape[2] - y1.shape[2], y2.shape[3] - y1.shape[3]
    if dy or dx:
        y1 = F.pad(y1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    y = torch.cat([y2, y1], 1)
    return F.relu(bn(p.sub("bn"), y, 1e-3))


def erf_nb1d(p, x, d):
    """non_bottleneck_1d (dropout p=0 / eval), model/ERFNet.py:30-65."""
    y = F.relu(F.conv2d(x, p["conv3x1_1.weight"], p["conv3x1_1.bias"], padding=(1, 0)))
    y = F.conv2d(y, p["conv1x3_1.weight"], p["conv1x3_1.bias"], padding=(0, 1))
    y = F.relu(bn(p.sub("bn1"), y, 1e-3))
    y = F.relu(F.conv2d(y, p["conv3x1_2.weight"], p["conv3x1_2.bias"], padding=(d, 0), dilation=(d, 1)))
    y = F.conv2d(y, p["conv1x3_2.weight"], p["conv1x3_2.bias"], padding=(0, d), dilation=(1, d))
    y = bn(p.sub("bn2"), y, 1e-3)
    return F.relu(y + x)


def erf_upsampler(p, x):
    """UpsamplerBlock, model/ERFNet.py:103-112."""
    y = F.conv_transpose2d(x, p["conv.weight"], p["conv.bias"], stride=2, padding=1, output_padding=1)
    return F.relu(bn(p.sub("bn"), y, 1e-3))


ERF_ENC_DILATIONS = [None, 1, 1, 1, 1, 1, None, 2, 4, 8, 16, 2, 4, 8, 16]  # ERFNet.py:75-86
ERF_DEC_LAYERS = ["up", 1, 1, "up", 1, 1]                                   # ERFNet.py:120-126


def erfnet(sd, x, train=False, stats=None):
    """ERFNet.forward (only_encode=False), model/ERFNet.py:151-156."""
    p = SD(sd, "", x.dtype, train, stats)
    y = erf_downsampler(p.sub("encoder.initial_block"), x)
    for i, d in enumerate(ERF_ENC_DILATIONS):
        q = p.sub("encoder.layers.%d" % i)
        y = erf_downsampler(q, y) if d is None else erf_nb1d(q, y, d)
    for i, d in enumerate(ERF_DEC_LAYERS):
        q = p.sub("decoder.layers.%d" % i)
        y = erf_upsampler(q, y) if d == "up" else erf_nb1d(q, y, d)
    return F.conv_transpose2d(y, p["decoder.output_conv.weight"], p["decoder.output_conv.bias"], stride=2)


# --------------------------------------------------------------------------- DABNet
def dab_bnprelu(p, x):
    """BNPReLU, model/DABNet.py:38-48."""
    return prelu(bn(p.sub("bn"), x, 1e-3), p["acti.weight"])


def dab_conv(p, x, stride=1, padding=0, dilation=1, groups=1, bn_acti=False):
    """Conv (bias=False), model/DABNet.py:16-35."""
    y = F.conv2d(x, p["conv.weight"], None, stride, padding, dilation, groups)
    return dab_bnprelu(p.sub("bn_prelu"), y) if bn_acti else y


def dab_module(p, x, d):
    """DABModule, model/DABNet.py:51-83."""
    c = x.shape[1] // 2
    y = dab_bnprelu(p.sub("bn_relu_1"), x)
    y = dab_conv(p.sub("conv3x3"), y, 1, 1, bn_acti=True)
    b1 = dab_conv(p.sub("dconv3x1"), y, 1, (1, 0), 1, c, True)
    b1 = dab_conv(p.sub("dconv1x3"), b1, 1, (0, 1), 1, c, True)
    b2 = dab_conv(p.sub("ddconv3x1"), y, 1, (d, 0), (d, 1), c, True)
    b2 = dab_conv(p.sub("ddconv1x3"), b2, 1, (0, d), (1, d), c, True)
    y = dab_bnprelu(p.sub("bn_relu_2"), b1 + b2)
    y = dab_conv(p.sub("conv1x1"), y)
    return y + x


def dab_down(p, x, n_in, n_out):
    """DownSamplingBlock, model/DABNet.py:86-110."""
    y = dab_conv(p.sub("conv3x3"), x, 2, 1)
    if n_in < n_out:
        y = torch.cat([y, F.max_pool2d(x, 2, 2)], 1)
    return dab_bnprelu(p.sub("bn_prelu"), y)


def dab_inject(x, ratio):
    """InputInjection, model/DABNet.py:113-124 (AvgPool2d(3,2,1), count_include_pad)."""
    for _ in range(ratio):
        x = F.avg_pool2d(x, 3, 2, 1)
    return x


DAB_D1 = [2, 2, 2]                 # DABNet.py:145-147
DAB_D2 = [4, 4, 8, 8, 16, 16]      # DABNet.py:151


def dabnet(sd, x, train=False, stats=None, upsample=True):
    """DABNet.forward, model/DABNet.py:160-183."""
    p = SD(sd, "", x.dtype, train, stats)
    y = dab_conv(p.sub("init_conv.0"), x, 2, 1, bn_acti=True)
    y = dab_conv(p.sub("init_conv.1"), y, 1, 1, bn_acti=True)
    y = dab_conv(p.sub("init_conv.2"), y, 1, 1, bn_acti=True)
    d1, d2, d3 = dab_inject(x, 1), dab_inject(x, 2), dab_inject(x, 3)
    y0 = dab_bnprelu(p.sub("bn_prelu_1"), torch.cat([y, d1], 1))
    y10 = dab_down(p.sub("downsample_1"), y0, 35, 64)
    y = y10
    for i, d in enumerate(DAB_D1):
        y = dab_module(p.sub("DAB_Block_1.DAB_Module_1_%d" % i), y, d)
    y1 = dab_bnprelu(p.sub("bn_prelu_2"), torch.cat([y, y10, d2], 1))
    y20 = dab_down(p.sub("downsample_2"), y1, 131, 128)
    y = y20
    for i, d in enumerate(DAB_D2):
        y = dab_module(p.sub("DAB_Block_2.DAB_Module_2_%d" % i), y, d)
    y2 = dab_bnprelu(p.sub("bn_prelu_3"), torch.cat([y, y20, d3], 1))
    out = dab_conv(p.sub("classifier.0"), y2)
    if upsample:
        out = F.interpolate(out, x.shape[2:], mode="bilinear", align_corners=False)
    return out


FORWARD = {"ERFNet": erfnet, "DABNet": dabnet}


def forward(name, sd, x, **kw):
    return FORWARD[name](sd, x, **kw)


def argmax_mask(logits):
    """test.py:79-82 / predict.py:52-54: numpy argmax over classes, uint8 (first max wins)."""
    a = logits.detach().cpu().numpy()
    import numpy as np
    return np.asarray(np.argmax(a, axis=1), dtype=np.uint8)


# --------------------------------------------------------------------------- ENet
def _enet_act(p, x):
    """The block's shared activation module: nn.PReLU() (one alpha, encoder) or nn.ReLU (decoder),
    model/ENet.py:18-21,53-56.  All aliases of the shared PReLU load last from `out_prelu.weight`."""
    if p.has("out_prelu.weight"):
        a = p["out_prelu.weight"]
        return torch.clamp(x, min=0) + a.view(1, 1, 1, 1) * torch.clamp(x, max=0)
    return F.relu(x)


def _enet_cba(p, q, x, idx_conv, **kw):
    """conv -> BN(eps 1e-5) -> shared activation inside an nn.Sequential `q` (conv at idx_conv)."""
    y = F.conv2d(x, p["%s.%d.weight" % (q, idx_conv)], None, **kw)
    return _enet_act(p, bn(p.sub("%s.%d" % (q, idx_conv + 1)), y, 1e-5))


def enet_initial(p, x):
    """InitialBlock, model/ENet.py:14-44."""
    y = torch.cat([F.conv2d(x, p["main_branch.weight"], None, stride=2, padding=1), F.max_pool2d(x, 3, 2, 1)], 1)
    return _enet_act(p, bn(p.sub("batch_norm"), y, 1e-5))


def enet_regular(p, x, k=3, pad=0, dil=1, asym=False):
    """RegularBottleneck, model/ENet.py:46-100 (eval: Dropout2d is the identity)."""
    e = _enet_cba(p, "ext_conv1", x, 0)
    if asym:
        e = _enet_cba(p, "ext_conv2", e, 0, padding=(pad, 0), dilation=dil)
        e = _enet_cba(p, "ext_conv2", e, 3, padding=(0, pad), dilation=dil)
    else:
        e = _enet_cba(p, "ext_conv2", e, 0, padding=pad, dilation=dil)
    e = _enet_cba(p, "ext_conv3", e, 0)
    return _enet_act(p, x + e)


def enet_down(p, x):
    """DownsamplingBottleneck, model/ENet.py:102-197."""
    main, idx = F.max_pool2d(x, 3, 2, 1, return_indices=True)
    e = _enet_cba(p, "ext_conv1", x, 0, stride=2)
    e = _enet_cba(p, "ext_conv2", e, 0, padding=1)
    e = _enet_cba(p, "ext_conv3", e, 0)
    n, ce, h, w = e.shape
    main = torch.cat([main, torch.zeros(n, ce - main.shape[1], h, w, dtype=x.dtype, device=x.device)], 1)
    return _enet_act(p, main + e), idx


def enet_up(p, x, idx):
    """UpsamplingBottleneck, model/ENet.py:199-272 (MaxUnpool2d on the CPU: raster order, last writer wins)."""
    main = bn(p.sub("main_conv1.1"), F.conv2d(x, p["main_conv1.0.weight"]), 1e-5)
    main = F.max_unpool2d(main.contiguous(), idx, 2)
    e = _enet_cba(p, "ext_conv1", x, 0)
    e = F.conv_transpose2d(e, p["ext_conv2.0.weight"], None, stride=2, padding=1, output_padding=1)
    e = _enet_act(p, bn(p.sub("ext_conv2.1"), e, 1e-5))
    e = _enet_cba(p, "ext_conv3", e, 0)
    return _enet_act(p, main + e)


ENET_STAGE23 = [("regular", 1, 1), ("dilated", 2, 2), ("asymmetric", 2, 1), ("dilated", 4, 4),
                ("regular", 1, 1), ("dilated", 8, 8), ("asymmetric", 2, 1), ("dilated", 16, 16)]   # ENet.py:307-360


def enet(sd, x, train=False, stats=None):
    """ENet.forward, model/ENet.py:386-432."""
    p = SD(sd, "", x.dtype, train, stats)
    y = enet_initial(p.sub("initial_block"), x)
    y, i1 = enet_down(p.sub("downsample1_0"), y)
    for i in range(1, 5):
        y = enet_regular(p.sub("regular1_%d" % i), y, 3, 1, 1)
    y, i2 = enet_down(p.sub("downsample2_0"), y)
    for stage, first in ((2, 1), (3, 0)):
        for j, (kind, pad, dil) in enumerate(ENET_STAGE23):
            q = p.sub("%s%d_%d" % (kind, stage, first + j))
            y = enet_regular(q, y, 5 if kind == "asymmetric" else 3, pad, dil, kind == "asymmetric")
    y = enet_up(p.sub("upsample4_0"), y, i2)
    y = enet_regular(p.sub("regular4_1"), y, 3, 1, 1)
    y = enet_regular(p.sub("regular4_2"), y, 3, 1, 1)
    y = enet_up(p.sub("upsample5_0"), y, i1)
    y = enet_regular(p.sub("regular5_1"), y, 3, 1, 1)
    return F.conv_transpose2d(y, p["transposed_conv.weight"], None, stride=2, padding=1, output_padding=1)


FORWARD["ENet"] = enet


# --------------------------------------------------------------------------- CGNet
def _cg_bnprelu(p, x, bn_key="bn", act_key="act"):
    return prelu(bn(p.sub(bn_key), x, 1e-3), p[act_key + ".weight"])


def _cg_conv_bn_prelu(p, x, stride=1):
    """ConvBNPReLU, model/CGNet.py:13-36 (pad = (k-1)/2)."""
    w = p["conv.weight"]
    return _cg_bnprelu(p, F.conv2d(x, w, None, stride, (w.shape[2] - 1) // 2))


def _cg_fglo(p, x):
    """FGlo, model/CGNet.py:173-191: x * sigmoid(W2 relu(W1 avgpool(x) + b1) + b2)."""
    y = x.mean(dim=(2, 3))
    y = F.relu(F.linear(y, p["fc.0.weight"], p["fc.0.bias"]))
    y = torch.sigmoid(F.linear(y, p["fc.2.weight"], p["fc.2.bias"]))
    return x * y.view(y.shape[0], -1, 1, 1)


def _cg_joint(p, x, d):
    c = x.shape[1]
    loc = F.conv2d(x, p["F_loc.conv.weight"], None, 1, 1, 1, c)
    sur = F.conv2d(x, p["F_sur.conv.weight"], None, 1, d, d, c)
    return torch.cat([loc, sur], 1)


def cg_block_down(p, x, d):
    """ContextGuidedBlock_Down, model/CGNet.py:193-227."""
    y = _cg_conv_bn_prelu(p.sub("conv1x1"), x, 2)
    j = _cg_bnprelu(p, _cg_joint(p, y, d))
    j = F.conv2d(j, p["reduce.conv.weight"])
    return _cg_fglo(p.sub("F_glo"), j)


def cg_block(p, x, d):
    """ContextGuidedBlock (add=True), model/CGNet.py:230-260."""
    y = _cg_conv_bn_prelu(p.sub("conv1x1"), x)
    j = _cg_bnprelu(p.sub("bn_prelu"), _cg_joint(p, y, d))
    return x + _cg_fglo(p.sub("F_glo"), j)


def cgnet(sd, x, train=False, stats=None):
    """CGNet.forward (M=3, N=21), model/CGNet.py:274-367."""
    p = SD(sd, "", x.dtype, train, stats)
    y = _cg_conv_bn_prelu(p.sub("level1_0"), x, 2)
    y = _cg_conv_bn_prelu(p.sub("level1_1"), y)
    y = _cg_conv_bn_prelu(p.sub("level1_2"), y)
    i1, i2 = dab_inject(x, 1), dab_inject(x, 2)
    y0 = _cg_bnprelu(p.sub("b1"), torch.cat([y, i1], 1))
    y10 = cg_block_down(p.sub("level2_0"), y0, 2)
    y = y10
    for i in range(2):
        y = cg_block(p.sub("level2.%d" % i), y, 2)
    y1 = _cg_bnprelu(p.sub("bn_prelu_2"), torch.cat([y, y10, i2], 1))
    y20 = cg_block_down(p.sub("level3_0"), y1, 4)
    y = y20
    for i in range(20):
        y = cg_block(p.sub("level3.%d" % i), y, 4)
    y2 = _cg_bnprelu(p.sub("bn_prelu_3"), torch.cat([y20, y], 1))
    out = F.conv2d(y2, p["classifier.0.conv.weight"])
    return F.interpolate(out, x.shape[2:], mode="bilinear", align_corners=False)


FORWARD["CGNet"] = cgnet


# --------------------------------------------------------------------------- Fast-SCNN
def _fs_bn(p, key, x):
    return bn(p.sub(key), x, 1e-5)        # nn.BatchNorm2d default eps (FastSCNN.py:20)


def _fs_cbr(p, x, stride=1, padding=0):
    """_ConvBNReLU, model/FastSCNN.py:15-27."""
    return F.relu(_fs_bn(p, "conv.1", F.conv2d(x, p["conv.0.weight"], None, stride, padding)))


def _fs_dsconv(p, x, stride=1):
    """_DSConv, model/FastSCNN.py:30-45."""
    c = x.shape[1]
    y = F.relu(_fs_bn(p, "conv.1", F.conv2d(x, p["conv.0.weight"], None, stride, 1, 1, c)))
    return F.relu(_fs_bn(p, "conv.4", F.conv2d(y, p["conv.3.weight"])))


def _fs_dwconv(p, x, stride=1):
    """_DWConv, model/FastSCNN.py:48-59."""
    return F.relu(_fs_bn(p, "conv.1", F.conv2d(x, p["conv.0.weight"], None, stride, 1, 1, x.shape[1])))


def _fs_bottleneck(p, x, stride, shortcut):
    """LinearBottleneck, model/FastSCNN.py:62-82."""
    y = _fs_cbr(p.sub("block.0"), x)
    y = _fs_dwconv(p.sub("block.1"), y, stride)
    y = _fs_bn(p, "block.3", F.conv2d(y, p["block.2.weight"]))
    return x + y if shortcut else y


def _fs_ppm(p, x):
    """PyramidPooling, model/FastSCNN.py:85-112."""
    size = x.shape[2:]
    feats = [x]
    for i, s in enumerate((1, 2, 3, 6)):
        f = _fs_cbr(p.sub("conv%d" % (i + 1)), F.adaptive_avg_pool2d(x, s))
        feats.append(F.interpolate(f, size, mode="bilinear", align_corners=True))
    return _fs_cbr(p.sub("out"), torch.cat(feats, 1))


def _fs_ffm(p, hi, lo):
    """FeatureFusionModule, model/FastSCNN.py:157-182."""
    lo = F.interpolate(lo, hi.shape[2:], mode="bilinear", align_corners=True)
    lo = _fs_dwconv(p.sub("dwconv"), lo)
    lo = _fs_bn(p, "conv_lower_res.1", F.conv2d(lo, p["conv_lower_res.0.weight"], p["conv_lower_res.0.bias"]))
    hi = _fs_bn(p, "conv_higher_res.1", F.conv2d(hi, p["conv_higher_res.0.weight"], p["conv_higher_res.0.bias"]))
    return F.relu(hi + lo)


def fastscnn(sd, x, train=False, stats=None):
    """FastSCNN.forward (aux=False, eval: Dropout is the identity), model/FastSCNN.py:204-235."""
    p = SD(sd, "", x.dtype, train, stats)
    q = p.sub("learning_to_downsample")
    hi = _fs_cbr(q.sub("conv"), x, 2, 0)
    hi = _fs_dsconv(q.sub("dsconv1"), hi, 2)
    hi = _fs_dsconv(q.sub("dsconv2"), hi, 2)
    g = p.sub("global_feature_extractor")
    y = hi
    for name, cin, cout, stride in (("bottleneck1", 64, 64, 2), ("bottleneck2", 64, 96, 2), ("bottleneck3", 96, 128, 1)):
        for i in range(3):
            s = stride if i == 0 else 1
            ci = cin if i == 0 else cout
            y = _fs_bottleneck(g.sub("%s.%d" % (name, i)), y, s, s == 1 and ci == cout)
    y = _fs_ppm(g.sub("ppm"), y)
    y = _fs_ffm(p.sub("feature_fusion"), hi, y)
    c = p.sub("classifier")
    y = _fs_dsconv(c.sub("dsconv1"), y)
    y = _fs_dsconv(c.sub("dsconv2"), y)
    y = F.conv2d(y, c["conv.1.weight"], c["conv.1.bias"])
    return F.interpolate(y, x.shape[2:], mode="bilinear", align_corners=True)


FORWARD["FastSCNN"] = fastscnn


# --------------------------------------------------------------------------- ESPNetv2 (EESPNet_Seg, s=2)
def _e2_bn(p, key, x):
    return bn(p.sub(key), x, 1e-5)


def _e2_cbr(p, x, stride=1, groups=1):
    """CBR, model/ESPNet_v2/cnn_utils.py:27-49."""
    w = p["conv.weight"]
    y = F.conv2d(x, w, None, stride, (w.shape[2] - 1) // 2, 1, groups)
    return prelu(_e2_bn(p, "bn", y), p["act.weight"])


def _e2_cb(p, x, stride=1, groups=1):
    """CB, cnn_utils.py:66-89."""
    w = p["conv.weight"]
    return _e2_bn(p, "bn", F.conv2d(x, w, None, stride, (w.shape[2] - 1) // 2, 1, groups))


def _e2_br(p, x):
    """BR, cnn_utils.py:45-64."""
    return prelu(_e2_bn(p, "bn", x), p["act.weight"])


def _e2_dilations(k, r_lim):
    ks = sorted((3 + 2 * i) if (3 + 2 * i) <= r_lim else 3 for i in range(k))
    return [{3: 1, 5: 2, 7: 3, 9: 4, 11: 5, 13: 6, 15: 7, 17: 8}[s] for s in ks]   # Model.py:40-49


def e2_eesp(p, x, stride, r_lim, down_avg=False, k=4):
    """EESP, model/ESPNet_v2/Model.py:15-99."""
    o1 = _e2_cbr(p.sub("proj_1x1"), x, 1, k)
    n = o1.shape[1]
    outs = []
    for i, d in enumerate(_e2_dilations(k, r_lim)):
        o = F.conv2d(o1, p["spp_dw.%d.conv.weight" % i], None, stride, d, d, n)
        outs.append(o if i == 0 else o + outs[-1])
    e = _e2_cb(p.sub("conv_1x1_exp"), _e2_br(p.sub("br_after_cat"), torch.cat(outs, 1)), 1, k)
    if stride == 2 and down_avg:
        return e
    if e.shape == x.shape:
        e = e + x
    return prelu(e, p["module_act.weight"])


def e2_down(p, x, img, r_lim):
    """DownSampler (reinf=True), Model.py:102-147."""
    avg = F.avg_pool2d(x, 3, 2, 1)
    out = torch.cat([avg, e2_eesp(p.sub("eesp"), x, 2, r_lim, True)], 1)
    if img is not None:
        while True:
            img = F.avg_pool2d(img, 3, 2, 1)
            if img.shape[2] == avg.shape[2]:
                break
        out = out + _e2_cb(p.sub("inp_reinf.1"), _e2_cbr(p.sub("inp_reinf.0"), img))
    return prelu(out, p["act.weight"])


def _e2_psp(p, x):
    """PSPModule, cnn_utils.py:11-25."""
    h, w = x.shape[2:]
    out, f = [x], x
    for i in range(4):
        f = F.avg_pool2d(f, 3, 2, 1)
        s = F.conv2d(f, p["stages.%d.conv.weight" % i], None, 1, 1, 1, f.shape[1])
        out.append(F.interpolate(s, (h, w), mode="bilinear", align_corners=True))
    return _e2_cbr(p.sub("project"), torch.cat(out, 1))


def _e2_up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def espnetv2(sd, x, train=False, stats=None):
    """EESPNet_Seg.forward (s=2; eval: Dropout2d is the identity), SegmentationModel.py:20-77; encoder
    EESPNet.forward(seg=True), Model.py:236-287."""
    p = SD(sd, "", x.dtype, train, stats)
    q = p.sub("net")
    l1 = _e2_cbr(q.sub("level1"), x, 2)
    l2 = e2_down(q.sub("level2_0"), l1, x, 13)
    l3 = e2_down(q.sub("level3_0"), l2, x, 11)
    for i in range(3):
        l3 = e2_eesp(q.sub("level3.%d" % i), l3, 1, 9)
    l4 = e2_down(q.sub("level4_0"), l3, x, 9)
    for i in range(7):
        l4 = e2_eesp(q.sub("level4.%d" % i), l4, 1, 7)
    l4p = _e2_cbr(p.sub("proj_L4_C"), l4)
    up = F.interpolate(l4p, l3.shape[2:], mode="bilinear", align_corners=True)
    m3 = e2_eesp(p.sub("pspMod.0"), torch.cat([l3, up], 1), 1, 7)
    m3 = _e2_psp(p.sub("pspMod.1"), m3)
    s3 = _e2_br(p.sub("act_l3"), F.conv2d(m3, p["project_l3.1.conv.weight"]))
    m2 = _e2_cbr(p.sub("project_l2"), torch.cat([l2, _e2_up2(s3)], 1))
    m1 = F.conv2d(torch.cat([l1, _e2_up2(m2)], 1), p["project_l1.1.conv.weight"])
    return _e2_up2(m1)


FORWARD["ESPNet_v2"] = espnetv2


# --------------------------------------------------------------------------- ESPNet (v1, ESPNet-A decoder)
def _esp_br(p, x):
    """BR, model/ESPNet.py:43-62 (eps 1e-3)."""
    return prelu(bn(p.sub("bn"), x, 1e-3), p["act.weight"])


def _esp_branches(p, o1):
    """The five dilated 3x3 branches + hierarchical sums shared by DownSamplerB / the ESP block, ESPNet.py:155-169."""
    outs = []
    for name, d in (("d1", 1), ("d2", 2), ("d4", 4), ("d8", 8), ("d16", 16)):
        outs.append(F.conv2d(o1, p[name + ".conv.weight"], None, 1, d, d))
    d1, add = outs[0], outs[1]
    cat = [d1, add]
    for o in outs[2:]:
        add = add + o
        cat.append(add)
    return torch.cat(cat, 1)


def esp_down(p, x):
    """DownSamplerB, ESPNet.py:138-173."""
    o1 = F.conv2d(x, p["c1.conv.weight"], None, 2, 1)
    return prelu(bn(p.sub("bn"), _esp_branches(p, o1), 1e-3), p["act.weight"])


def esp_block(p, x, add=True):
    """DilatedParllelResidualBlockB, ESPNet.py:176-226."""
    c = _esp_branches(p, F.conv2d(x, p["c1.conv.weight"]))
    if add:
        c = x + c
    return _esp_br(p.sub("bn"), c)


def _esp_encoder(q, x, p_rep, q_rep):
    """ESPNet_Encoder.forward up to the three concat stages, ESPNet.py:283-318."""
    o0 = prelu(bn(q.sub("level1.bn"), F.conv2d(x, q["level1.conv.weight"], None, 2, 1), 1e-3), q["level1.act.weight"])
    inp1 = F.avg_pool2d(x, 3, 2, 1)
    inp2 = F.avg_pool2d(inp1, 3, 2, 1)
    o0_cat = _esp_br(q.sub("b1"), torch.cat([o0, inp1], 1))
    o1_0 = esp_down(q.sub("level2_0"), o0_cat)
    o1 = o1_0
    for i in range(p_rep):
        o1 = esp_block(q.sub("level2.%d" % i), o1)
    o1_cat = _esp_br(q.sub("b2"), torch.cat([o1, o1_0, inp2], 1))
    o2_0 = esp_down(q.sub("level3_0"), o1_cat)
    o2 = o2_0
    for i in range(q_rep):
        o2 = esp_block(q.sub("level3.%d" % i), o2)
    o2_cat = _esp_br(q.sub("b3"), torch.cat([o2_0, o2], 1))
    return o0_cat, o1_cat, o2_cat


def espnet(sd, x, train=False, stats=None):
    """ESPNet.forward (classes from the weights, p=2, q=3), ESPNet.py:355-385."""
    p = SD(sd, "", x.dtype, train, stats)
    q = p.sub("encoder")
    o0_cat, o1_cat, o2_cat = _esp_encoder(q, x, 2, 3)
    s = bn(p.sub("br"), F.conv2d(o2_cat, q["classifier.conv.weight"]), 1e-3)
    o2_c = F.conv_transpose2d(s, p["up_l3.0.weight"], None, 2)
    o1_c = F.conv2d(o1_cat, p["level3_C.conv.weight"])
    comb = esp_block(p.sub("combine_l2_l3.1"), _esp_br(p.sub("combine_l2_l3.0"), torch.cat([o1_c, o2_c], 1)), add=False)
    comb = _esp_br(p.sub("up_l2.1"), F.conv_transpose2d(comb, p["up_l2.0.weight"], None, 2))
    y = torch.cat([comb, o0_cat], 1)
    y = prelu(bn(p.sub("conv.bn"), F.conv2d(y, p["conv.conv.weight"], None, 1, 1), 1e-3), p["conv.act.weight"])
    return F.conv_transpose2d(y, p["classifier.weight"], None, 2)


FORWARD["ESPNet"] = espnet


# --------------------------------------------------------------------------- ESNet
def es_fcu(p, x, k, d):
    """FCU(chann, kernel_size=k, dropprob, dilated=d) (dropout p=0 / eval), model/ESNet.py:50-92."""
    h = (k - 1) // 2
    y = F.relu(F.conv2d(x, p["conv3x1_1.weight"], p["conv3x1_1.bias"], padding=(h, 0)))
    y = F.conv2d(y, p["conv1x3_1.weight"], p["conv1x3_1.bias"], padding=(0, h))
    y = F.relu(bn(p.sub("bn1"), y, 1e-3))
    y = F.relu(F.conv2d(y, p["conv3x1_2.weight"], p["conv3x1_2.bias"], padding=(h * d, 0), dilation=(d, 1)))
    y = F.conv2d(y, p["conv1x3_2.weight"], p["conv1x3_2.bias"], padding=(0, h * d), dilation=(1, d))
    return F.relu(x + bn(p.sub("bn2"), y, 1e-3))


def es_pfcu(p, x):
    """PFCU(chann) (eval), model/ESNet.py:95-151: three dilated branches (2, 5, 9) share bn2."""
    y = F.relu(F.conv2d(x, p["conv3x1_1.weight"], p["conv3x1_1.bias"], padding=(1, 0)))
    y = F.conv2d(y, p["conv1x3_1.weight"], p["conv1x3_1.bias"], padding=(0, 1))
    y = F.relu(bn(p.sub("bn1"), y, 1e-3))
    acc = x
    for d in (2, 5, 9):
        o = F.relu(F.conv2d(y, p["conv3x1_2%d.weight" % d], p["conv3x1_2%d.bias" % d], padding=(d, 0), dilation=(d, 1)))
        o = F.conv2d(o, p["conv1x3_2%d.weight" % d], p["conv1x3_2%d.bias" % d], padding=(0, d), dilation=(1, d))
        acc = acc + bn(p.sub("bn2"), o, 1e-3)
    return F.relu(acc)


# ESNet.py:162-180: ("down",) | ("fcu", k) | ("pfcu",) | ("up",)
ES_LAYERS = ([("fcu", 3)] * 3 + [("down",)] + [("fcu", 5)] * 2 + [("down",)] + [("pfcu",)] * 3 +
             [("up",), ("fcu", 5), ("fcu", 5), ("up",), ("fcu", 3), ("fcu", 3)])


def esnet(sd, x, train=False, stats=None):
    """ESNet.forward, model/ESNet.py:184-193 (any input size: erf_downsampler pads like ESNet.py:25-29)."""
    p = SD(sd, "", x.dtype, train, stats)
    y = erf_downsampler(p.sub("initial_block"), x)
    for i, spec in enumerate(ES_LAYERS):
        q = p.sub("layers.%d" % i)
        if spec[0] == "down":
            y = erf_downsampler(q, y)
        elif spec[0] == "up":
            y = erf_upsampler(q, y)
        elif spec[0] == "pfcu":
            y = es_pfcu(q, y)
        else:
            y = es_fcu(q, y, spec[1], 1)
    return F.conv_transpose2d(y, p["output_conv.weight"], p["output_conv.bias"], stride=2)


FORWARD["ESNet"] = esnet


# --------------------------------------------------------------------------- ContextNet
def _cx_bottlenecks(g, y):
    """Deep_net's six LinearBottleneck stages, model/ContextNet.py:101-133 (block = FastSCNN's, :58-74)."""
    cin = 32
    for idx, (cout, t, nblk, stride) in enumerate(zip((32, 32, 48, 64, 96, 128), (1, 6, 6, 6, 6, 6),
                                                      (1, 1, 3, 3, 2, 2), (1, 1, 2, 2, 1, 1))):
        for i in range(nblk):
            s = stride if i == 0 else 1
            ci = cin if i == 0 else cout
            y = _fs_bottleneck(g.sub("bottleneck%d.%d" % (idx + 1, i)), y, s, s == 1 and ci == cout)
        cin = cout
    return y


def contextnet(sd, x, train=False, stats=None):
    """ContextNet.forward (aux=False, eval), model/ContextNet.py:180-226: full-resolution shallow branch,
    quarter-resolution deep branch (bilinear align_corners=True), feature fusion, classifier, bilinear to input size."""
    p = SD(sd, "", x.dtype, train, stats)
    q = p.sub("spatial_detail")
    hi = _fs_cbr(q.sub("conv"), x, 2, 0)
    hi = _fs_dsconv(q.sub("dsconv1"), hi, 2)
    hi = _fs_dsconv(q.sub("dsconv2"), hi, 2)
    hi = _fs_dsconv(q.sub("dsconv3"), hi, 1)
    h, w = x.shape[2:]
    lo = F.interpolate(x, size=(h // 4, w // 4), mode="bilinear", align_corners=True)   # scale_factor=0.25
    g = p.sub("context_feature_extractor")
    lo = _fs_cbr(g.sub("conv_"), lo, 2, 0)
    lo = _cx_bottlenecks(g, lo)
    y = _fs_ffm(p.sub("feature_fusion"), hi, lo)
    c = p.sub("classifier")
    y = _fs_dsconv(c.sub("dsconv1"), y)
    y = _fs_dsconv(c.sub("dsconv2"), y)
    y = F.conv2d(y, c["conv.1.weight"], c["conv.1.bias"])
    return F.interpolate(y, x.shape[2:], mode="bilinear", align_corners=True)


FORWARD["ContextNet"] = contextnet


# --------------------------------------------------------------------------- EDANet
def eda_down(p, x, n_in, n_out):
    """DownsamplerBlock, model/EDANet.py:18-38 (BatchNorm2d default eps)."""
    y = F.conv2d(x, p["conv.weight"], p["conv.bias"], stride=2, padding=1)
    if n_in < n_out:
        y = torch.cat([y, F.max_pool2d(x, 2, 2)], 1)
    return F.relu(bn(p.sub("bn"), y, 1e-5))


def eda_module(p, x, d):
    """EDAModule (eval: Dropout2d is the identity), model/EDANet.py:41-85: no activation between a 3x1 and its 1x3;
    ``dilation=d`` on a (3,1) / (1,3) kernel only acts along the kernel's long axis."""
    y = F.relu(bn(p.sub("bn0"), F.conv2d(x, p["conv1x1.weight"], p["conv1x1.bias"]), 1e-5))
    y = F.conv2d(y, p["conv3x1_1.weight"], p["conv3x1_1.bias"], padding=(1, 0))
    y = F.relu(bn(p.sub("bn1"), F.conv2d(y, p["conv1x3_1.weight"], p["conv1x3_1.bias"], padding=(0, 1)), 1e-5))
    y = F.conv2d(y, p["conv3x1_2.weight"], p["conv3x1_2.bias"], padding=(d, 0), dilation=(d, d))
    y = F.relu(bn(p.sub("bn2"), F.conv2d(y, p["conv1x3_2.weight"], p["conv1x3_2.bias"], padding=(0, d), dilation=(d, d)), 1e-5))
    return torch.cat([y, x], 1)


EDA_BLOCKS = {2: (1, 1, 1, 2, 2), 4: (2, 2, 4, 4, 8, 8, 16, 16)}       # EDANet.py:127-131


def edanet(sd, x, train=False, stats=None):
    """EDANet.forward, model/EDANet.py:148-157."""
    p = SD(sd, "", x.dtype, train, stats)
    y = eda_down(p.sub("layers.0"), x, 3, 15)
    y = eda_down(p.sub("layers.1"), y, 15, 60)
    for i, d in enumerate(EDA_BLOCKS[2]):
        y = eda_module(p.sub("layers.2.residual_dense_layers.%d" % i), y, d)
    y = eda_down(p.sub("layers.3"), y, 260, 130)
    for i, d in enumerate(EDA_BLOCKS[4]):
        y = eda_module(p.sub("layers.4.residual_dense_layers.%d" % i), y, d)
    y = F.conv2d(y, p["project_layer.weight"], p["project_layer.bias"])
    h, w = y.shape[2:]
    return F.interpolate(y, size=(8 * h, 8 * w), mode="bilinear", align_corners=True)    # scale_factor=8


FORWARD["EDANet"] = edanet


# --------------------------------------------------------------------------- LEDNet (oracle only so far)
def _led_shuffle(x, groups=2):
    """Channel_shuffle, model/LEDNet.py:27-39: out[:, j*groups + g] = x[:, g*(C/groups) + j]."""
    n, c, h, w = x.shape
    return x.view(n, groups, c // groups, h, w).transpose(1, 2).reshape(n, c, h, w)


def led_ssnbt(p, x, d):
    """SS_nbt_module_paper (eval), model/LEDNet.py:108-186: split in halves; left 3x1 -> 1x3, right 1x3 -> 3x1 (twice,
    second time dilated), merge, + input, ReLU, channel shuffle."""
    c1 = x.shape[1] // 2          # round(c * 0.5) for even c
    x1, x2 = x[:, :c1], x[:, c1:]

    def conv(t, key, vertical, dil):
        pad = (dil, 0) if vertical else (0, dil)
        dl = (dil, 1) if vertical else (1, dil)
        return F.conv2d(t, p[key + ".weight"], p[key + ".bias"], padding=pad, dilation=dl)
    o1 = F.relu(conv(x1, "conv3x1_1_l", True, 1))
    o1 = F.relu(bn(p.sub("bn1_l"), conv(o1, "conv1x3_1_l", False, 1), 1e-3))
    o2 = F.relu(conv(x2, "conv1x3_1_r", False, 1))
    o2 = F.relu(bn(p.sub("bn1_r"), conv(o2, "conv3x1_1_r", True, 1), 1e-3))
    o1 = F.relu(conv(o1, "conv3x1_2_l", True, d))
    o1 = bn(p.sub("bn2_l"), conv(o1, "conv1x3_2_l", False, d), 1e-3)
    o2 = F.relu(conv(o2, "conv1x3_2_r", False, d))
    o2 = bn(p.sub("bn2_r"), conv(o2, "conv3x1_2_r", True, d), 1e-3)
    return _led_shuffle(F.relu(x + torch.cat([o1, o2], 1)))


def led_apn(p, x):
    """APNModule, model/LEDNet.py:189-283: global-pool branch + 1x1 mid branch gated by a 3-level single-channel pyramid
    (7/5/3-tap asymmetric convs with strides (2,1) then (1,2)); every resize is bilinear, align_corners=True."""
    h, w = x.shape[2:]
    up = lambda t, size: F.interpolate(t, size=size, mode="bilinear", align_corners=True)

    def cbr(q, t):      # Conv2dBnRelu :46-56
        return F.relu(bn(q.sub("conv.1"), F.conv2d(t, q["conv.0.weight"], q["conv.0.bias"]), 1e-3))

    def pair(q, t, i, k, stride):       # (k,1) conv stride (s,1), (1,k) conv stride (1,s), BN, ReLU
        t = F.conv2d(t, q["%d.weight" % i], q["%d.bias" % i], stride=(stride, 1), padding=(k // 2, 0))
        t = F.conv2d(t, q["%d.weight" % (i + 1)], q["%d.bias" % (i + 1)], stride=(1, stride), padding=(0, k // 2))
        return F.relu(bn(q.sub("%d" % (i + 2)), t, 1e-3))
    b1 = up(cbr(p.sub("branch1.1"), F.adaptive_avg_pool2d(x, 1)), (h, w))
    mid = cbr(p.sub("mid.0"), x)
    x1 = pair(p.sub("down1"), x, 0, 7, 2)
    x2 = pair(p.sub("down2"), x1, 0, 5, 2)
    x3 = pair(p.sub("down3"), pair(p.sub("down3"), x2, 0, 3, 2), 4, 3, 1)
    x3 = up(x3, ((h + 3) // 4, (w + 3) // 4))
    y = up(pair(p.sub("conv2"), x2, 0, 5, 1) + x3, ((h + 1) // 2, (w + 1) // 2))
    y = up(y + pair(p.sub("conv1"), x1, 0, 7, 1), (h, w))
    return y * mid + b1


LED_LAYERS = [1, 1, 1, None, 1, 1, None, 1, 2, 5, 9, 2, 5, 9, 17]       # LEDNet.py:291-313 (None = DownsamplerBlock)


def lednet(sd, x, train=False, stats=None):
    """LEDNet.forward, model/LEDNet.py:316-327 (even sizes: the F.pad of the downsampler is a no-op)."""
    p = SD(sd, "", x.dtype, train, stats)
    y = erf_downsampler(p.sub("initial_block"), x)
    for i, d in enumerate(LED_LAYERS):
        q = p.sub("layers.%d" % i)
        y = erf_downsampler(q, y) if d is None else led_ssnbt(q, y, d)
    y = led_apn(p.sub("apn"), y)
    return F.interpolate(y, x.shape[2:], mode="bilinear", align_corners=True)


FORWARD["LEDNet"] = lednet
