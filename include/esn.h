/*
 * esn.h -- C ABI of libesn_sm100.so: B200 (sm_100a) kernels for the
 * convolutional encoder-decoder hot path of the lightweight segmentation zoo.
 *
 * The reference (Ethan-ye/Efficient-Segmentation-Networks) has no FFI boundary
 * of its own: every op on the path is an ATen call made from the nn.Module
 * forwards under model/ (SURVEY.md section 8b).  Each entry point below therefore
 * cites the reference nn.Module lines whose ATen calls it replaces.  The host
 * side (efficient-segmentation-networks_b200/esn/_lib.py) binds these with
 * ctypes; INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *    the name ends in _host; `stream` is a cudaStream_t passed as void*.
 *  - no allocation, no synchronisation, no exceptions across the ABI.
 *  - return 0 on success, a negative ESN_ERR_* otherwise (esn_strerror()).
 *  - activations are NHWC ("channels last"): element (n,h,w,c) of a tensor
 *    view lives at ptr[((n*H + h)*W + w)*c_stride + c]; c_stride >= c lets a
 *    kernel read or write a channel slice of a wider concat buffer.  The only
 *    NCHW tensors are the network input (fp32, as the reference's Dataset
 *    yields it, dataset/cityscapes.py:58-106) and the logits returned to the
 *    caller (N,classes,H,W), as `model(images)` returns them (train.py:351).
 */
#ifndef ESN_H_
#define ESN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESN_VERSION 100

enum {
  ESN_OK = 0,
  ESN_ERR_BAD_ARG = -1,      /* null pointer, bad dtype/layout combination      */
  ESN_ERR_BAD_SHAPE = -2,    /* shapes inconsistent with the op                 */
  ESN_ERR_UNSUPPORTED = -3,  /* valid op, but this kernel family cannot run it  */
  ESN_ERR_CUDA = -4,         /* a CUDA runtime/driver call failed               */
  ESN_ERR_ALIGN = -5         /* pointer / stride alignment requirement not met  */
};

enum { ESN_F32 = 0, ESN_BF16 = 1, ESN_U8 = 2, ESN_I64 = 3, ESN_I32 = 4 };
enum { ESN_NHWC = 0, ESN_NCHW = 1 };
enum { ESN_ACT_NONE = 0, ESN_ACT_RELU = 1, ESN_ACT_PRELU = 2 };

/* A 4-D activation view. */
typedef struct EsnTensor {
  void* ptr;        /* device pointer to element (0,0,0,0) of the view */
  int32_t dtype;    /* ESN_F32 / ESN_BF16 / ...                          */
  int32_t layout;   /* ESN_NHWC (c_stride used) or ESN_NCHW (contiguous) */
  int32_t n, h, w, c;
  int32_t c_stride; /* NHWC only: elements between consecutive pixels    */
  int32_t _pad;
} EsnTensor;

/* Fused epilogue shared by the conv kernels:
 *   v = acc * scale[c] + shift[c]        (conv bias and eval-mode BN folded here)
 *   v += residual[n,h,w,c]               (optional)
 *   v = act(v)                           (none / ReLU / per-channel PReLU)
 */
typedef struct EsnEpilogue {
  const float* scale;  /* [Cout] or NULL (=1) */
  const float* shift;  /* [Cout] or NULL (=0) */
  const float* alpha;  /* [Cout] PReLU slopes, used when act == ESN_ACT_PRELU */
  int32_t act;
  int32_t flags;       /* ESN_EP_ACT_BEFORE_RESIDUAL: v = act(act(acc*scale+shift) + residual), the
                          "ext = act(BN(conv)); out = act(main + ext)" pattern of ENet.py:93-100 */
  EsnTensor residual;  /* ptr == NULL: none; else same N,H,W,C as the output */
} EsnEpilogue;
enum {
  ESN_EP_ACT_BEFORE_RESIDUAL = 1,
  ESN_EP_RESIDUAL_FIRST = 2 /* v = act((x + residual)*scale + shift): "BN(input + combine)" of ESPNet.py:221-224;
                               esn_affine_act only, the conv entry points answer ESN_ERR_UNSUPPORTED */
};
enum { ESN_STEM_PAD0 = 256 };

/* 2-D convolution / transposed convolution, NHWC, stride/dilation/groups as
 * torch.nn.Conv2d / ConvTranspose2d (cross-correlation, zero padding).
 * Replaces the aten::convolution (+ native_batch_norm eval + relu/prelu + add)
 * sequences of:
 *   ERFNet.py:24-27,49-65,109-112,128,136   (DownsamplerBlock conv, non_bottleneck_1d,
 *                                            UpsamplerBlock, output_conv)
 *   DABNet.py:27-35,69-83,101-110,132-136,158,179  (Conv, DABModule, DownSamplingBlock, stem, classifier)
 * Weight packing (done once by the host at plan-build time):
 *   direct kernels : fp32 [kh*kw][Cin/groups][Cout]      (Cout contiguous)
 *   umma kernels   : bf16 [kh*kw][Cout_pad][Cin]         (Cin contiguous, K-major B operand)
 * For transposed convs the packed tap (kh,kw) holds W[ci][co][kh][kw].
 */
typedef struct EsnConv {
  EsnTensor x;   /* input  view (NHWC f32/bf16, or NCHW f32 for the network input) */
  EsnTensor y;   /* output view (NHWC f32/bf16)                                    */
  const void* w; /* packed weights, see above                                      */
  int32_t kh, kw;
  int32_t stride;
  int32_t pad_h, pad_w;
  int32_t dil_h, dil_w;
  int32_t groups;      /* 1 or Cin (depthwise) */
  int32_t transposed;  /* 0 / 1 (ConvTranspose2d; output_padding implied by y.h/y.w) */
  int32_t cout_pad;    /* umma only: padded Cout rows per tap in w                    */
  EsnEpilogue ep;
} EsnConv;

/* CUDA-core direct convolution (fp32 accumulate; f32 or bf16 activations).
 * The exact-arithmetic path (fp32 parity) and the path for shapes the tensor
 * core kernel does not take (Cin = 3 stem, depthwise). */
int esn_conv2d_direct(const EsnConv* p, void* stream);

/* tcgen05 implicit-GEMM convolution: bf16 operands staged by TMA, fp32
 * accumulators in TMEM, fused epilogue.  Requires bf16 NHWC x/y, groups == 1,
 * Cin in {16,32,64,128,...multiple of 64}, Cout <= 256.  Returns
 * ESN_ERR_UNSUPPORTED for anything else (the host then calls esn_conv2d_direct). */
int esn_conv2d_umma(const EsnConv* p, void* stream);

/* esn_conv2d_umma with a second epilogue stage computed on the value as it is stored (bf16):
 *   v  = act(acc*scale + shift (+ residual))            -> conv.y   (store_y != 0)
 *   v2 = act2(bf16(v) * scale2[c] + shift2[c])          -> y2
 * One launch for "conv1x1 + input" followed by the next DABModule's bn_relu_1 (DABNet.py:69-83 then :70 of the next
 * module: the sum is needed as that module's residual, its BNPReLU as the module's conv input), and, with store_y == 0,
 * for a conv whose only consumer is a BNPReLU over a concat (init_conv[2] -> bn_prelu_1, DABNet.py:158-166; the last
 * DABModule of a block -> bn_prelu_2 / bn_prelu_3, DABNet.py:171,176).  y2 has conv.y's shape (its own channel stride);
 * results are bit-identical to esn_conv2d_umma followed by esn_affine_act.  Stride-1/2 forward convs whose Cout the
 * staged epilogue takes (a multiple of 8, <= 64 or a multiple of 64); ESN_ERR_UNSUPPORTED otherwise. */
typedef struct EsnConvDual {
  EsnConv conv;
  EsnTensor y2;
  const float* scale2; /* [Cout] or NULL (=1) */
  const float* shift2; /* [Cout] or NULL (=0) */
  const float* alpha2; /* [Cout] PReLU slopes when act2 == ESN_ACT_PRELU */
  int32_t act2;
  int32_t store_y;     /* 0: conv.y is not written (it still carries the output shape) */
} EsnConvDual;
int esn_conv2d_umma_dual(const EsnConvDual* p, void* stream);

/* Fused factorized pair: y = act2( conv_1xk( act1( conv_kx1(x)*s1 + b1 ) )*s2 + b2 (+ residual) ), both convs
 * dense C -> C with `taps` taps, padding (taps-1)/2 * dilation, the same dilation.  One half of the reference's
 * non_bottleneck_1d (ERFNet.py:44-65: conv3x1 -> ReLU -> conv1x3 -> BN -> ReLU, and the dilated second pair
 * + residual); the intermediate tensor stays in shared memory.  bf16 NHWC, C in {16, 64}, taps == 3,
 * dilation <= 8, W a multiple of 128 (C=64) / 512 (C=16), shared memory for one intermediate row; anything
 * else answers ESN_ERR_UNSUPPORTED and the host runs two esn_conv2d_umma calls.
 * w1 / w2: bf16 [tap][C][C] (tap-major, Cin contiguous), as for esn_conv2d_umma. */
typedef struct EsnConvPair {
  EsnTensor x, y;
  const void* w1;
  const void* w2;
  int32_t taps, dilation;
  EsnEpilogue ep1;     /* scale/shift + none|ReLU; no residual */
  EsnEpilogue ep2;     /* scale/shift, optional residual, none|ReLU|PReLU */
} EsnConvPair;
int esn_conv_pair_umma(const EsnConvPair* p, void* stream);

/* Network stem on the caller's NCHW fp32 image (Cin = 3): Conv2d(3, cconv, 3, stride 2, pad 1)
 * [ || MaxPool2d(2,2) -> concat ] -> per-channel affine (bias + eval BN) -> activation -> NHWC.
 * Replaces ERFNet.py:24-27 for DownsamplerBlock(3,16) (cconv 13 + 3 pooled channels) and
 * DABNet.py:132 Conv(3,32,3,2)+BNPReLU.  w: fp32 [9][3][cconv]; ep.scale/shift/alpha: [y.c]. */
typedef struct EsnStem {
  EsnTensor x, y;
  const float* w;
  int32_t cconv;
  int32_t with_pool;   /* bits 0-1: 0 none, 1 MaxPool2d(2,2) (ERFNet), 2 MaxPool2d(3, stride 2, padding 1) (ENet.py:33);
                          bit 8 (ESN_STEM_PAD0): the conv has padding 0 instead of 1 (FastSCNN.py:120) */
  EsnEpilogue ep;
} EsnStem;
int esn_stem_conv3x3s2(const EsnStem* p, void* stream);

/* MaxPool2d(2, stride 2) followed by the per-channel affine + activation of
 * the BatchNorm slice it is concatenated into.
 * Replaces the pool branch of ERFNet.py:24-27 (DownsamplerBlock) and
 * DABNet.py:104-108 (DownSamplingBlock): y[..., c] = act(max2x2(x)[c]*scale[c]+shift[c]).
 * y is floor(h/2) x floor(w/2); for an odd h or w it may instead be ceil(h/2) x ceil(w/2): the pooled map then sits in the
 * upper-left corner and the last row / column is the zero padding ESNet's DownsamplerBlock inserts before the concat
 * (ESNet.py:25-29: F.pad(x1, [0, diffX, 0, diffY])), i.e. y = act(shift) there. */
typedef struct EsnPool {
  EsnTensor x, y;
  EsnEpilogue ep;
} EsnPool;
int esn_maxpool2x2_affine_act(const EsnPool* p, void* stream);

/* AvgPool2d(3, stride 2, padding 1, count_include_pad=True) applied `ratio`
 * times (DABNet.py:113-124 InputInjection), then affine+act, written into a
 * channel slice.  x may be NCHW f32 (network input) or NHWC. */
int esn_avgpool3x3s2_affine_act(const EsnPool* p, void* stream);

/* ENet pooling pair (ENet.py:126-130,225,262).  esn_maxpool3x3s2_idx: MaxPool2d(3,2,1,return_indices) on
 * NHWC, idx int32 [N,Ho,Wo,C] = h*W+w of the first maximum.  esn_max_unpool2x2: MaxUnpool2d(2) as a
 * deterministic gather (last writer in raster order wins, as the CPU reference), fused with
 * y = act(unpool(v, idx) + ext). */
int esn_maxpool3x3s2_idx(const EsnTensor* x, const EsnTensor* y, int32_t* idx, void* stream);
typedef struct EsnUnpool {
  EsnTensor v;          /* pooled-resolution values */
  const int32_t* idx;   /* [N,Hp,Wp,C] */
  EsnTensor ext;        /* optional addend at output resolution (ptr NULL = none) */
  EsnTensor y;          /* output [N,2Hp,2Wp,C] */
  const float* alpha;
  int32_t act;
  int32_t _pad;
} EsnUnpool;
int esn_max_unpool2x2(const EsnUnpool* p, void* stream);
/* Their backward passes (training; autograd of the same modules reached from loss.backward(), train.py:353):
 * esn_maxpool3x3s2_idx_bwd: dx[n,h,w,c] (+)= sum of dy over the windows whose recorded arg-max is (h,w) -- a gather over the
 * input pixels (deterministic, no atomics although the windows overlap); accumulate != 0 adds to dx.
 * esn_max_unpool2x2_bwd: dv[n,i,j,c] = dy at the position idx[n,i,j,c] of the (2Hp x 2Wp) output plane -- every pooled cell,
 * including cells that lost a collision in the forward scatter, as torch's max_unpool2d backward does. */
int esn_maxpool3x3s2_idx_bwd(const EsnTensor* dy, const int32_t* idx, const EsnTensor* dx, int32_t accumulate, void* stream);
int esn_max_unpool2x2_bwd(const EsnTensor* dy, const int32_t* idx, const EsnTensor* dv, void* stream);

/* CGNet global-context gate FGlo (CGNet.py:173-191): x * sigmoid(W2 relu(W1 avgpool(x) + b1) + b2), and the
 * block's residual (CGNet.py:258-260).  esn_global_avgpool: per-chunk partial sums over H*W;
 * esn_fglo_gate: gate[n][c] from the sums and the two nn.Linear layers (row-major weights);
 * esn_scale_nc: y = x * gate[n][c] (+ residual). */
/* Backward of the gate (training): esn_dot_nc: out[n][c] += sum_hw a*b (fp32 atomics, out zeroed by the caller) = d gate;
 * esn_scale_add_nc: y = a * s[n][c] + t[n][c] (+ extra) = d x (t may be NULL: plain per-(n,c) scaling, the forward gate). */
int esn_dot_nc(const EsnTensor* a, const EsnTensor* b, float* out, void* stream);
int esn_scale_add_nc(const EsnTensor* a, const float* s, const float* t, const EsnTensor* extra, const EsnTensor* y, void* stream);
int esn_global_avgpool_chunks(const EsnTensor* x);   /* host query: number of partial-sum chunks K for this view */
int esn_global_avgpool(const EsnTensor* x, float* sums /* [K][N][C] partial sums, no atomics */, void* stream);
typedef struct EsnFGlo {
  const float* sums;   /* [chunks][N][C] partial sums, reduced in fixed order (deterministic) */
  const float* w1;     /* [hidden][C]  (fc.0.weight) */
  const float* b1;     /* [hidden] */
  const float* w2;     /* [C][hidden]  (fc.2.weight) */
  const float* b2;     /* [C] */
  float* gate;         /* out [N][C] */
  int32_t n, channels, hidden, hw;
  int32_t chunks, _pad;
} EsnFGlo;
int esn_fglo_gate(const EsnFGlo* p, void* stream);
int esn_scale_nc(const EsnTensor* x, const float* gate, const EsnTensor* residual, const EsnTensor* y, void* stream);

/* nn.AdaptiveAvgPool2d (FastSCNN.py:97-99) and NHWC->NHWC F.interpolate(bilinear, align_corners 0/1)
 * (FastSCNN.py:101-102,174), y may be a channel slice of a concat buffer. */
int esn_adaptive_avgpool(const EsnTensor* x, const EsnTensor* y, void* stream);
int esn_bilinear_nhwc(const EsnTensor* x, const EsnTensor* y, int32_t align_corners, void* stream);

/* Elementwise per-channel affine + activation (+ residual) on an NHWC view:
 * standalone BNPReLU on concat tensors (DABNet.py:38-48,166,171,176). */
int esn_affine_act(const EsnPool* p, void* stream);

/* Tail of a channel concat: y[..., 0:c) = act(x*scale + shift) for the c <= 4 injected channels of x (fp32 NHWC, pixel
 * stride 4: what the InputInjection average pools write, DABNet.py:113-124) and y[..., c:tail_c) = 0, where y is the view of
 * the concat buffer that starts at the first injected channel and tail_c (a power of two >= 8, <= y.c_stride) reaches the end of
 * the padded pixel.  With the concat's BNPReLU slice in (scale, shift, alpha) this is "torch.cat([..., down_k], 1)" +
 * bn_prelu_k for those channels (DABNet.py:166,171,176) in whole-sector writes; the concat buffer needs no zero fill. */
int esn_concat_tail(const EsnPool* p, int32_t tail_c, void* stream);

/* Layout / dtype conversion between NCHW f32 and NHWC f32|bf16 views. */
int esn_convert_layout(const EsnTensor* x, const EsnTensor* y, void* stream);

/* DABNet depthwise asymmetric pair, both branches fused (DABNet.py:73-78):
 *   br1 = BNPReLU(dw1x3(BNPReLU(dw3x1(x))))           dilation 1
 *   br2 = BNPReLU(dw1x3_d(BNPReLU(dw3x1_d(x))))       dilation d
 *   y   = PReLU(BN(br1 + br2))                        (bn_relu_2)
 * Per-channel parameter block `prm`, fp32 [27][C]:
 *   rows  0-11  three taps each of dw3x1, dw1x3 (branch 1) and dw3x1, dw1x3 (branch 2, dilated)
 *   rows 12-23  (scale, shift, PReLU alpha) of the BNPReLU after each of those four convs, same order
 *   rows 24-26  (scale, shift, alpha) of bn_relu_2
 * (packed by DABModule._build_prep in model/DABNet.py; read in csrc/esn_stencil.cu) */
typedef struct EsnDabPair {
  EsnTensor x, y;
  const float* prm;    /* [27][C] fp32 */
  int32_t dilation;
  int32_t _pad;
} EsnDabPair;
int esn_dab_dw_pair(const EsnDabPair* p, void* stream);

/* Segmentation heads.
 * esn_head_convt2x2: ConvTranspose2d(Cin, classes, 2, stride 2, bias) (ERFNet.py:128,136)
 *   fused with either NCHW logits output (f32/bf16) and/or the uint8 argmax mask
 *   (test.py:79-82: np.argmax over classes, first maximum wins).
 * esn_head_bilinear: F.interpolate(bilinear, align_corners=False) of NHWC low-res
 *   class scores (DABNet.py:181) fused the same way. */
typedef struct EsnHead {
  EsnTensor x;          /* NHWC input features (convt) or low-res scores (bilinear) */
  const float* w;       /* convt: fp32 [2][2][Cin][classes_pad(32)]; bilinear: unused */
  const float* bias;    /* convt: [classes] */
  EsnTensor logits;     /* optional (ptr may be NULL): NCHW f32/bf16 (N,classes,H,W) */
  uint8_t* mask;        /* optional: (N,H,W) uint8 argmax */
  int32_t classes;
  int32_t out_h, out_w;
  int32_t align_corners; /* bilinear head only: 0 (DABNet/CGNet) or 1 (FastSCNN.py:229, ESPNetv2) */
} EsnHead;
int esn_head_convt2x2(const EsnHead* p, void* stream);
int esn_head_bilinear(const EsnHead* p, void* stream);

/* ENet's head as ONE launch: ConvTranspose2d(16, classes <= 24, 3, stride 2, padding 1, output_padding 1, bias optional)
 * (ENet.py:229-236) fused with the argmax over classes (test.py:79-82: first maximum wins) on the warp-level tensor cores;
 * the full-resolution scores are never written.  x: bf16 NHWC with 16 channels, width a multiple of 16; mask: (N, 2h, 2w)
 * uint8.  wfrag: the bf16 weights in mma.m16n8k16 B-fragment order, [9 pairs][3 class tiles][32 lanes][2] 32-bit words:
 * pair p = (output position (a, b), neighbour (dy, dx)) in the order (0,0|0,0) (0,1|0,0) (0,1|0,1) (1,0|0,0) (1,0|1,0)
 * (1,1|0,0) (1,1|0,1) (1,1|1,0) (1,1|1,1) uses W[:, :, a+1-2dy, b+1-2dx]; lane (g = lane / 4, t = lane % 4) holds
 * {W[2t, n], W[2t+1, n]} and {W[2t+8, n], W[2t+9, n]} for class n = 8*tile + g (zero for n >= classes), low half first.
 * Anything else answers ESN_ERR_UNSUPPORTED and the host runs esn_conv2d_umma + esn_head_bilinear. */
/* ENet's RegularBottleneck with four internal channels (channels = 16, regular 3x3: ENet.py:46-100, `regular5_1`) as ONE launch:
 *   y = act(x + act(BN3(W3 . act(BN2(W2 * act(BN1(W1 . x)))))))      (eval-mode BN as scale / shift, Dropout2d = identity)
 * W1 [16][4], W2 [9][4][4], W3 [4][16] fp32 in the direct kernels' (tap, cin, cout) order; act = ReLU or PReLU (alpha1/2: [4],
 * alpha3: [16], also the slope of the outer activation).  bf16 NHWC, 16 channels, dilation 1..4; anything else answers
 * ESN_ERR_UNSUPPORTED and the host runs the three convs. */
typedef struct EsnBneck4 {
  EsnTensor x, y;
  const float *w1, *w2, *w3;
  const float *scale1, *shift1, *alpha1;
  const float *scale2, *shift2, *alpha2;
  const float *scale3, *shift3, *alpha3;
  int32_t dilation;
  int32_t act;
} EsnBneck4;
int esn_bottleneck4(const EsnBneck4* p, void* stream);

typedef struct EsnHeadT3 {
  EsnTensor x;
  const uint32_t* wfrag;
  const float* bias;      /* [classes] or NULL */
  uint8_t* mask;
  int32_t classes;
  int32_t _pad;
} EsnHeadT3;
int esn_head_convt3x3s2_mask(const EsnHeadT3* p, void* stream);

/* ConvTranspose2d(16, classes <= 24, 2, stride 2) + argmax over classes in one tensor-core launch: ERFNet's / ESNet's
 * output_conv (ERFNet.py:112,128; ESNet.py:182) fused with the CPU argmax of test.py:79-82, mask only -- the mask-only form of
 * esn_head_convt2x2.  x: bf16 NHWC, 16 channels, w % 16 == 0, 8-byte aligned pixels; mask: uint8 (N, 2h, 2w).
 * wfrag: uint32 [2][4][3][32][2] = the weights as bf16 B fragments of mma.m16n8k16, hi part ([0]) and lo part ([1]) of the fp32
 * value (w = hi + lo), per output position a * 2 + b, per 8-class tile nt, per lane (g = lane / 4, t = lane % 4):
 * register r holds (low half, high half) = W[4t + 2r, nt * 8 + g, a, b], W[4t + 2r + 1, nt * 8 + g, a, b] (zero for padded
 * classes) -- the K slots of the MMA carry the channels in the order that lets a lane load its four channels in one piece.
 * Ties resolve to the lower class, as numpy's argmax does. */
int esn_head_convt2x2_mask(const EsnHeadT3* p, void* stream);

/* Weighted cross-entropy over NCHW logits (utils/losses/loss.py:15-32):
 *   sums[0] += sum_i w[y_i]*nll_i,  sums[1] += sum_i w[y_i]   (fp32 atomics per CTA)
 * and, if dlogits != NULL, the gradient w[y_i]*(softmax - onehot) * (*gout) / (*gnorm); gnorm is the
 * device scalar holding sum_i w[y_i] AFTER the cross-rank all-reduce (so data-parallel training
 * reproduces the reference's gathered-batch weighted mean), NULL = unnormalised. */
typedef struct EsnCE {
  EsnTensor logits;       /* NCHW f32/bf16 */
  const int64_t* target;  /* (N,H,W) */
  const float* weight;    /* [classes] or NULL */
  float* sums;            /* [2], zeroed by the caller */
  EsnTensor dlogits;      /* optional NCHW, same dtype as logits */
  int32_t ignore_label;
  int32_t _pad;
  const float* gnorm;     /* optional device scalar: dlogits are divided by *gnorm (the global sum of weights) */
  const float* gout;      /* optional device scalar: upstream gradient of the loss */
  /* Online hard example mining, ProbOhemCrossEntropy2d (utils/losses/loss.py:163-216): */
  float* prob_out;          /* optional (N,H,W) f32: softmax probability of the labelled class per pixel, 1 where ignored
                               (mask_prob, loss.py:195-198) */
  const float* keep_thresh; /* optional device scalar: pixels whose labelled-class probability exceeds *keep_thresh are
                               treated as ignored (kept_mask, loss.py:203-206) */
} EsnCE;
int esn_weighted_ce(const EsnCE* p, void* stream);

/* OHEM threshold on the device (loss.py:199-203), no host synchronisation:
 *   if min_kept > *num_valid            -> *out = +inf                (nothing is filtered)
 *   else kth = the min(n, min_kept)-th smallest of prob[0..n)  (what mask_prob.argsort()[min_kept - 1] selects)
 *                                       -> *out = max(thresh, kth)
 * prob holds non-negative floats (esn_weighted_ce's prob_out), so the order of their bit patterns is their numeric order:
 * the k-th value is found exactly by a three-pass radix select (12 + 12 + 8 bits) over shared-memory histograms.
 * num_valid: device scalar (f32 count of non-ignored pixels = sums[1] of an unweighted esn_weighted_ce pass).
 * workspace: esn_ohem_workspace_bytes() bytes, zeroed by the caller before the call. */
int64_t esn_ohem_workspace_bytes(void);
int esn_ohem_threshold(const float* prob, int64_t n, int64_t min_kept, float thresh, const float* num_valid, float* out,
                       void* workspace, void* stream);

/* ------------------------------------------------------------------ training path
 * Train-mode BatchNorm2d (torch semantics: biased batch variance normalises, running_var is updated
 * with the unbiased one, momentum as given) is split the way the fused inference epilogue wants it:
 *   esn_channel_stats  : sums[0..C) += sum_x, sums[C..2C) += sum_x^2 over N*H*W (fp64 atomics)
 *   esn_bn_finalize    : scale = gamma*invstd, shift = beta - mean*scale (consumed by esn_affine_act /
 *                        the conv epilogues), saves mean / invstd, updates the running statistics
 * Replaces aten::native_batch_norm(training=True) of every nn.BatchNorm2d on the path
 * (ERFNet.py:21,38,45,107; DABNet.py:41). */
int esn_channel_stats(const EsnTensor* x, double* sums, int32_t with_squares, void* stream);

typedef struct EsnBnFinalize {
  const double* sums;      /* [2][C] from esn_channel_stats */
  int64_t count;           /* N*H*W */
  const float* gamma;      /* [C] or NULL */
  const float* beta;       /* [C] or NULL */
  float eps, momentum;
  float* running_mean;     /* [C] updated in place, or NULL */
  float* running_var;      /* [C] updated in place, or NULL */
  float* scale;            /* out [C] */
  float* shift;            /* out [C] */
  float* mean;             /* out [C] */
  float* invstd;           /* out [C] */
  int32_t channels;
  int32_t _pad;
} EsnBnFinalize;
int esn_bn_finalize(const EsnBnFinalize* p, void* stream);

/* Backward of y = act(x*scale + shift) where (scale, shift) come from train-mode BN (train_stats=1)
 * or are constants (train_stats=0: conv bias + ReLU, eval BN).  dz = dy * act'(z).
 *   esn_bn_act_bwd_reduce: sums[0..C) += sum dz, [C..2C) += sum dz*xhat, [2C..3C) += sum dy*z*[z<0]
 *   esn_bn_act_bwd_apply : dx = scale*(dz - mean(dz) - xhat*mean(dz*xhat)) (+ extra), and writes
 *                          dbeta = sum dz, dgamma = sum dz*xhat, dalpha = sum dy*z*[z<0]
 * Replaces native_batch_norm_backward + _prelu_kernel_backward / threshold_backward (+ the add of a
 * second consumer's gradient, passed as `extra`). */
typedef struct EsnBnBwd {
  EsnTensor x;        /* BN input (= conv output), saved by the forward */
  EsnTensor dy;       /* gradient w.r.t. the activation output */
  EsnTensor dx;       /* out: gradient w.r.t. x (same dtype as dy) */
  EsnTensor extra;    /* optional: added to dx (ptr NULL = none) */
  const float* scale; /* [C] as used by the forward (NULL = 1) */
  const float* shift; /* [C] (NULL = 0) */
  const float* alpha; /* [C] PReLU slopes */
  const float* mean;  /* [C] batch mean   (train_stats) */
  const float* invstd;/* [C] 1/sqrt(var+eps) (train_stats) */
  double* sums;       /* [3][C], zeroed by the caller before the reduce pass */
  float* dgamma;      /* out [C] or NULL */
  float* dbeta;       /* out [C] or NULL */
  float* dalpha;      /* out [C] or NULL */
  int32_t act;
  int32_t train_stats;
} EsnBnBwd;
int esn_bn_act_bwd_reduce(const EsnBnBwd* p, void* stream);
int esn_bn_act_bwd_apply(const EsnBnBwd* p, void* stream);

/* The same two layers as ONE launch each (bf16 NHWC operands with 16-byte-aligned channel vectors; anything else returns
 * ESN_ERR_UNSUPPORTED with nothing launched and the caller uses the calls above): a co-resident grid (cooperative launch)
 * reduces over its pixel chunks, meets at one grid-wide barrier and then normalises / back-propagates the SAME chunks out of
 * L2, so a BatchNorm layer reads its operands from DRAM once and costs one launch instead of three (forward: statistics,
 * finalize, affine + activation) or two (backward).  The reduction scratch holds ESN_BN_FUSED_REPLICAS copies of the sums (a
 * CTA adds into one of them, so fewer atomics meet at one address); it and `barrier` must be zero on entry.
 *   esn_bn_act_train_fwd: y = act(BN_train(x)); fin.scale / shift / mean / invstd and the running statistics are written as
 *                         by esn_bn_finalize; fin.sums: [ESN_BN_FUSED_REPLICAS][2][C] doubles of scratch.
 *   esn_bn_act_bwd_fused: esn_bn_act_bwd_reduce + esn_bn_act_bwd_apply with train_stats = 1; p->sums:
 *                         [ESN_BN_FUSED_REPLICAS][3][C] doubles of scratch.
 * Replaces aten::native_batch_norm(training=True) + _prelu_kernel / threshold and native_batch_norm_backward +
 * _prelu_kernel_backward / threshold_backward of every BatchNorm2d on the path (DABNet.py:41, ERFNet.py:21,38,45,107;
 * train.py:351-356). */
#define ESN_BN_FUSED_REPLICAS 8
typedef struct EsnBnTrainFwd {
  EsnTensor x;          /* BN input (= conv output) */
  EsnTensor y;          /* out: act(x*scale + shift), may be a channel slice of a wider buffer; must not alias x */
  EsnBnFinalize fin;    /* as for esn_bn_finalize; fin.sums is written (scratch), fin.channels == x.c */
  const float* alpha;   /* [C] PReLU slopes (act == ESN_ACT_PRELU) */
  uint32_t* barrier;    /* one zeroed 32-bit word of device memory, private to this call */
  int32_t act;
  int32_t _pad;
} EsnBnTrainFwd;
int esn_bn_act_train_fwd(const EsnBnTrainFwd* p, void* stream);
int esn_bn_act_bwd_fused(const EsnBnBwd* p, uint32_t* barrier, void* stream);
/* Backward of a bare activation (train_stats = 0, ReLU or none: conv bias + ReLU, ERFNet.py:49,55): dx = dy * act'(x*scale +
 * shift) * scale (+ extra) in one streaming pass -- nothing is reduced, p->sums / dgamma / dbeta / dalpha are not touched.
 * bf16 NHWC, channel counts multiples of 8; otherwise ESN_ERR_UNSUPPORTED and the caller uses esn_bn_act_bwd_apply. */
int esn_act_bwd(const EsnBnBwd* p, void* stream);

/* Weight gradient of a dense or depthwise Conv2d: p->x = forward input, p->y = gradient of the conv
 * output, p->w = fp32 accumulator [kh*kw][Cin/groups][Cout] (zeroed by the caller; atomics).
 * Replaces the weight branch of aten::convolution_backward.  The input gradient is the transposed /
 * flipped convolution and runs through esn_conv2d_umma / esn_conv2d_direct. */
int esn_conv2d_wgrad(const EsnConv* p, void* stream);
/* 1 when esn_conv2d_wgrad runs this problem on the tcgen05 kernel (bf16 NHWC operands, dense, stride 1, k <= 3:
 * both operands MN-major straight from NHWC through TMA, all taps of a filter row per pass), else 0. */
int esn_wgrad_umma_supported(const EsnConv* p);

/* MaxPool2d(2,2) backward (first maximum wins, as max_pool2d_with_indices), optionally accumulating
 * into dx; bilinear (align_corners=False) backward from NCHW d logits to NHWC low-res scores,
 * scaled by gscale (1 / sum of class weights: the loss normalisation). */
int esn_maxpool2x2_bwd(const EsnTensor* x, const EsnTensor* dy, const EsnTensor* dx, int32_t accumulate, void* stream);
int esn_bilinear_bwd(const EsnTensor* dlogits, const EsnTensor* dlow, float gscale, void* stream);

/* Backward of F.interpolate(bilinear, align_corners 0/1) with the upstream gradient in NCHW (logits) or
 * NHWC, of F.adaptive_avg_pool2d (FastSCNN.py:96-105: dy is (N,C,S,S)), and nn.Dropout / nn.Dropout2d
 * (forward and backward are the same call: the keep mask is a counter-based hash of (seed, element) or
 * (seed, n, c), scaled by 1/(1-p); FastSCNN.py:193, SegmentationModel.py:50-53).  `accumulate` adds to dx. */
int esn_bilinear_bwd_nhwc(const EsnTensor* dy, const EsnTensor* dx, int32_t align_corners, int32_t accumulate, void* stream);
int esn_adaptive_avgpool_bwd(const EsnTensor* dy, const EsnTensor* dx, int32_t accumulate, void* stream);
int esn_avgpool3x3s2_bwd(const EsnTensor* dy, const EsnTensor* dx, int32_t accumulate, void* stream);   /* AvgPool2d(3,2,1): /9 always */
int esn_dropout(const EsnTensor* x, const EsnTensor* y, uint64_t seed, float p, int32_t per_channel, void* stream);
/* Same, with a device-resident iteration counter mixed into the seed (NULL = none): a training step captured in a
 * CUDA graph (esn/graph.py) advances the counter inside the graph, so every replay draws fresh masks and the
 * backward of the same replay regenerates them. */
int esn_dropout_step(const EsnTensor* x, const EsnTensor* y, uint64_t seed, const uint64_t* step, float p,
                     int32_t per_channel, void* stream);
/* nn.Dropout2d as a table: mask[n*C + c] = 0 or 1 / (1 - p), the same keys, hash and device-side iteration counter as
 * esn_dropout_step(per_channel = 1); forward and backward are then one esn_scale_nc each (y = x * mask[n][c] (+ residual)).
 * Replaces aten::feature_dropout of ERFNet.py:62 / ENet.py:86 / ESPNet_v2 in train mode. */
int esn_dropout_mask_nc(float* mask, int64_t count, uint64_t seed, const uint64_t* step, float p, void* stream);

/* Confusion matrix of predicted masks against labels, accumulated on the device: M[gt * nclass + pred] += 1 for every
 * pixel with 0 <= gt < nclass (ignore label 255 skipped) -- ConfusionMatrix.generateM, utils/metric/metric.py:68-76, as
 * called by get_iou from test.py:90 / train.py:404.  pred: uint8 [n_pixels] (the fused argmax masks); gt: uint8 or int64
 * [n_pixels]; M: uint64 [nclass * nclass], zeroed by the caller before the first batch; nclass <= 32. */
int esn_confusion_matrix(const uint8_t* pred, const void* gt, int32_t gt_is_int64, int64_t n_pixels, int32_t nclass,
                         uint64_t* M, void* stream);

/* Input pipeline on the device: the arithmetic tail of the reference's dataset classes
 * (dataset/cityscapes.py:74-78, 164-170, 208-214): uint8 HWC image batch img[n][h][w][3] (BGR as cv2.imread returns it)
 * -> out[n][3][h][w] fp32, out[co] = (float)img[ci] - mean3[ci] with ci = 2 - co when reverse_channels (BGR -> RGB) else
 * co.  mean3 is a HOST pointer to three fp32 values in the INPUT channel order (the pickle's dtype,
 * dataset/inform/cityscapes_inform.pkl); it is read before the call returns.  Bit-identical to the numpy float32 result.
 * img needs no alignment (16-byte aligned image starts take the vector path); out must be 4-byte aligned. */
int esn_image_u8hwc_to_f32nchw(const uint8_t* img, float* out, int32_t n, int32_t h, int32_t w, const float* mean3,
                               int32_t reverse_channels, void* stream);

/* Training-time augmentation of the reference's CityscapesDataSet.__getitem__ (dataset/cityscapes.py:67-104) on the device,
 * one launch per batch: random scale (cv2.resize: INTER_LINEAR on the uint8 image in OpenCV's 11-bit fixed-point arithmetic,
 * bit-identical to cv2; INTER_NEAREST on the labels), fp32 mean subtraction, BGR -> RGB, zero / ignore padding up to the crop
 * size, crop, CHW, mirror.  The random draws (scale factor, offsets, mirror) are made by the caller -- the reference makes
 * them with Python's random / numpy.random -- and passed per image; the resized image is never materialised.
 * items: HOST array of n <= esn_augment_max_batch() entries, read before the call returns.  mean3: HOST, input channel order. */
typedef struct EsnAugItem {
  const uint8_t* img;    /* device, (h, w, 3) uint8 BGR */
  const uint8_t* label;  /* device, (h, w) uint8 */
  int32_t h, w;          /* decoded size */
  int32_t rh, rw;        /* size after the resize: cvRound(h * f), cvRound(w * f) (= h, w when do_scale == 0) */
  double scale;          /* 1 / f */
  int32_t h_off, w_off;  /* crop offset inside the padded resized image */
  int32_t flip;          /* 1 = mirror the columns of the crop */
  int32_t do_scale;      /* 0 = no resize step (scale=False) */
} EsnAugItem;
int32_t esn_augment_max_batch(void);
int esn_augment_u8(const EsnAugItem* items, int32_t n, int32_t crop_h, int32_t crop_w, const float* mean3, int32_t ignore_label,
                   float* out_img /* (n,3,crop_h,crop_w) */, int64_t* out_label /* (n,crop_h,crop_w) */, void* stream);

/* Per-pixel gate with a per-image bias: y[n,h,w,c] = g[n,h,w] * x[n,h,w,c] + b[n,c] -- the close of LEDNet's attention
 * pyramid (APNModule.forward, model/LEDNet.py:279-281: torch.mul(x, mid) + the global-pooling branch, whose bilinear
 * upsampling from 1x1 with align_corners=True is a constant per image and class).  g (N,1,H,W), x and y (N,C,H,W),
 * b (N,C,1,1) or NULL; all NHWC; x, b and y share one dtype (f32 or bf16), g has the same dtype or is f32 next to bf16
 * scores (LEDNet keeps its single-channel pyramid in fp32). */
int esn_gate_bcast(const EsnTensor* g, const EsnTensor* x, const EsnTensor* b, const EsnTensor* y, void* stream);

/* Bilinear up-sampling of low-resolution class scores + weighted cross-entropy + the gradient of the scores, in one pass that
 * never writes a full-resolution tensor: the close of a training iteration of the nets whose head is
 * F.interpolate(scores, input.size()[2:], mode='bilinear', align_corners=False / True) (DABNet.py:181, CGNet.py:332;
 * FastSCNN.py:233, ESPNet_v2/SegmentationModel.py:76) under CrossEntropyLoss2d (utils/losses/loss.py:15-32; train.py:351-353).  Replaces esn_head_bilinear ->
 * esn_weighted_ce (forward, backward) -> esn_bilinear_bwd.
 *   scores: NHWC (n, h, w, c <= 32), f32 or bf16;  target: int64 (n, out_h, out_w);  weight: [c] f32 or NULL;
 *   any out_h, out_w >= 1 and both align_corners modes: source coordinates and weights as ATen's upsample_bilinear2d computes
 *   them in fp32.
 *   sums[0] += sum_p w[t_p] (lse(logits_p) - logits_p[t_p]),  sums[1] += sum_p w[t_p]   over pixels with t_p in [0, c) and
 *   t_p != ignore_label (fp32 atomics; zeroed by the caller);
 *   dscores: f32 NHWC (n, h, w, c), every lane of its pixel stride written (zeros behind the classes):
 *   d sums[0] / d scores -- the caller scales it by (upstream gradient) / sums[1] for the mean reduction. */
typedef struct EsnBilinearCE {
  EsnTensor scores;
  const int64_t* target;
  const float* weight;
  float* sums;
  EsnTensor dscores;
  int32_t out_h, out_w;
  int32_t ignore_label;
  int32_t align_corners;
} EsnBilinearCE;
int esn_bilinear_ce(const EsnBilinearCE* p, void* stream);

/* Optimizer step of the training iteration (train.py:355 `optimizer.step()` on the torch.optim.Adam of train.py:212-215):
 * Adam with L2 weight decay (torch.optim.Adam semantics, amsgrad = False, maximize = False) over every parameter tensor of a
 * param group in ONE launch.  table: DEVICE array of fp32 tensors (parameter, gradient, exp_avg, exp_avg_sq: n elements each,
 * 4-byte aligned; 16-byte aligned chunks take the vector path).  blocks: DEVICE array of n_blocks (tensor index, chunk index)
 * int32 pairs, one per CTA, chunk = esn_adam_chunk() elements -- together they cover every element once.  lr, step: DEVICE
 * scalars (fp32); *step is the number of updates done so far: the launch uses step + 1 for the bias corrections and stores it
 * back (the last CTA to finish does, through the zero-initialised DEVICE counter done, which it re-arms).  For every element:
 *   g += weight_decay * p;  m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;
 *   p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
 * in fp32, with 1 - beta and the bias corrections formed in double (as torch's Python scalars are). */
typedef struct EsnAdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
} EsnAdamTensor;
int32_t esn_adam_chunk(void);
int esn_adam_step(const EsnAdamTensor* table, const int32_t* blocks, int32_t n_blocks, const float* lr, float* step,
                  uint32_t* done, double beta1, double beta2, double eps, double weight_decay, void* stream);

/* Library / device queries (host-side, no stream). */
int esn_version(void);
const char* esn_strerror(int code);
/* Number of kernels this library has launched in this process (bench.py's
 * `gpu_launches`), and reset. */
int64_t esn_launch_count(void);
void esn_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* ESN_H_ */
