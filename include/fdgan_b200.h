/*
 * fdgan_b200 -- C ABI of the B200-native FD-GAN hot path.
 *
 * The reference (WeilanAnnn/FD-GAN) has no FFI: its boundary is the torch.nn.Module
 * API (demo.py:73,132).  The Python modules in fdgan_b200/ mirror that surface and
 * call the entry points below through ctypes; every entry point replaces a PyTorch
 * library call the reference makes on the path.  Citations are into /root/reference.
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative FDG_E* code otherwise; fdg_last_error() gives
 *     the message (thread-local).  Never throws, never allocates, never synchronises.
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch), fp32 unless
 *     stated; work is enqueued on the cudaStream_t passed last (as void*).
 *   - activations are addressed with explicit element strides (n, h, w, c) so that
 *     NCHW images, NHWC feature maps and channel slices of a dense-block concat
 *     buffer are all expressible; the fast paths need c-stride 1 and 16-byte alignment.
 */
#ifndef FDGAN_B200_H_
#define FDGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDG_OK 0
#define FDG_EINVAL (-1)   /* bad argument / unsupported shape */
#define FDG_ECUDA (-2)    /* CUDA runtime error at launch */
#define FDG_ENOSUPPORT (-3)

typedef void* fdg_stream_t; /* cudaStream_t */

/* A strided 4-D activation view: element (n,h,w,c) lives at p[n*sn + h*sh + w*sw + c*sc]. */
typedef struct FdgTensor {
  float* p;
  int64_t sn, sh, sw, sc;
} FdgTensor;

/* gather modes of the conv A-operand (what the loader does instead of a separate pass) */
#define FDG_GATHER_DIRECT 0
#define FDG_GATHER_AVGPOOL2 1 /* logical pixel = mean of the 2x2 physical block, AFTER the prologue: replaces
                                 F.avg_pool2d(x,2) (models/dehaze1113.py:763,780) and the AvgPool2d of the
                                 torchvision transition (models/densenet.py:221), commuted in front of its 1x1 conv */
#define FDG_GATHER_UP2 2      /* logical pixel (y,x) reads physical (y/2,x/2): F.upsample_nearest (dehaze1113.py:370) */

/* activations of the conv epilogue */
#define FDG_ACT_NONE 0
#define FDG_ACT_RELU 1    /* relu0 (dehaze1113.py:760), F.relu (vgg16.py:28-47) */
#define FDG_ACT_TANH 2    /* dehaze1113.py:799 */
#define FDG_ACT_SIGMOID 3 /* dehaze1113.py:223 */

/* store modes */
#define FDG_STORE_NORMAL 0
#define FDG_STORE_UP2 1   /* write each output pixel to its 2x2 nearest-neighbour block (TransitionBlockdy, dehaze1113.py:370) */
#define FDG_STORE_ACCUM 2 /* y += result (gradient accumulation into a dense-block gradient buffer) */

/*
 * fdg_conv2d: implicit-GEMM convolution y = act(conv(prologue(gather(x)), W) + bias).
 * Replaces nn.Conv2d / nn.ConvTranspose2d(1x1) / F.conv2d on the path (dehaze1113.py:744-755,
 * 262-266,363; torchvision _DenseLayer conv1/conv2 restated at models/densenet.py:193-198;
 * D convs dehaze1113.py:196-222; vgg16.py:9-22) and, with flipped weights, their data gradients.
 *   prologue:  v = has_affine ? x*scale[c] + shift[c] : x ;  a = v > 0 ? v : slope*v
 *              (slope 1 = identity, 0 = ReLU, 0.2 = LeakyReLU; affine = BatchNorm apply, dehaze1113.py:40 /
 *              torchvision norm1/norm2).  Zero padding is applied AFTER the prologue.
 *   epilogue:  *alpha, +bias, activation, optional multiply by (e > 0 ? 1 : eslope) with e a second
 *              tensor of the output's shape (ReLU/LeakyReLU backward mask), optional per-channel
 *              sum / sum-of-squares of the stored values accumulated in fp64 (BatchNorm batch statistics).
 */
typedef struct FdgConv {
  FdgTensor x;          /* physical input */
  int N, H, W, Cin;     /* logical input extent seen by the filter (after gather) */
  int gather;
  int has_affine;
  const float* scale;   /* [Cin] */
  const float* shift;   /* [Cin] */
  float slope;
  const float* w;       /* packed weights [R*S*Cin][w_ld], k = (r*S + s)*Cin + ci */
  int w_ld;
  int R, S, stride, pad;
  int Cout, OH, OW;
  const float* bias;    /* [Cout] or NULL */
  int act;
  FdgTensor e;          /* epilogue mask source or p == NULL */
  float eslope;
  FdgTensor y;          /* output (for FDG_STORE_UP2 the physical output is 2OH x 2OW) */
  int store;
  double* stats;        /* sums at stats[c], sums of squares at stats[stats_ld + c] (c < Cout); or NULL */
  int stats_ld;
  float alpha;          /* result scale applied before the bias (4 = adjoint of nearest x2 through FDG_GATHER_AVGPOOL2) */
  int impl;             /* 0 auto, 1 force SIMT fp32, 2 force tcgen05 (error if unsupported) */
  const void* w_umma;   /* tcgen05 operand image of the weights (fdg_pack_weight_umma) or NULL */
  /* BatchNorm-backward epilogue (tcgen05 path, 128-bit views only), enabled by e_scale != NULL: with z = e_scale*e + e_shift
   * (the BatchNorm output whose (Leaky)ReLU mask applies) and dz = acc * (z > 0 ? 1 : eslope), the kernel stores
   * y (=|+=) e_scale * dz and reduces stats[c] += sum dz, stats[stats_ld + c] += sum dz * e.  Together with
   * fdg_bn_bwd_finalize (dx = alpha dz + beta x + delta) this leaves only the per-channel affine term beta x + delta,
   * which fdg_affine_accum applies later, once per channel, for ALL consumers of that channel (the dense-block layers
   * share the batch statistics of a concat channel, so their beta / delta simply add up). */
  const float* e_scale; /* [Cout] or NULL */
  const float* e_shift; /* [Cout] */
  const void* w_k1;     /* operand image for the 3x3 / stride 1 / pad 1 / Cout <= 32 kernel (fdg_pack_weight_k1) or NULL */
  /* Split-bf16 input (or NULL): the input as two dense bf16 planes [N*H*W][Cin], hi at x_split and lo right behind it
   * (value = hi + lo, exactly the operand split the tensor-core kernels apply to fp32 inputs).  Produced by
   * fdg_ew_bwd(out_split); lets the kernel feed its A operand with bulk tensor loads instead of loading, converting and
   * re-storing it.  1x1 / stride 1 / direct gather / no prologue / Cin % 64 == 0 / BatchNorm-backward epilogue only. */
  const void* x_split;
} FdgConv;

int fdg_conv2d(const FdgConv* p, fdg_stream_t stream);

/*
 * fdg_conv2d_wgrad: weight gradient dW[k][co] (+)= sum_pixels a[pixel][k] * g[pixel][co] with the same
 * A-operand (gather + prologue) as fdg_conv2d and g = gradient w.r.t. the conv output.  Written straight
 * into the PyTorch parameter layout: OIHW (nn.Conv2d) or, for transposed=1, [Cin][Cout] (ConvTranspose2d 1x1).
 * Replaces autograd's conv backward-weight for every conv listed above.  dw must be zeroed by the caller
 * unless it is accumulating (split-K partial sums are added atomically).
 */
typedef struct FdgWgrad {
  FdgTensor x;
  int N, H, W, Cin;
  int gather;
  int has_affine;
  const float* scale;
  const float* shift;
  float slope;
  FdgTensor g;          /* [N, OH, OW, Cout] */
  int R, S, stride, pad;
  int Cout, OH, OW;
  float* dw;
  int transposed;
  float* dbias;         /* [Cout] (+)= sum_pixels g, or NULL */
  int impl;             /* 0 auto, 1 force SIMT fp32, 2 force tcgen05 (error if unsupported) */
  const void* g_split;  /* the gradient as split-bf16 planes [N*OH*OW][Cout] (see FdgConv.x_split) or NULL; Cout % 64 == 0 */
  const void* x_split;  /* the conv input AFTER its prologue as split-bf16 planes [N*H*W][Cin] (hi plane, then lo plane) or NULL:
                         * wide stride-1 RxS weight gradients feed both operands through the bulk-tensor engine (needs g_split,
                         * OW % 32 == 0, Cin % 8 == 0, 64 < Cout; x / scale / shift / slope are then not read) */
} FdgWgrad;

int fdg_conv2d_wgrad(const FdgWgrad* p, fdg_stream_t stream);

/* Weight repacking (PyTorch layout -> GEMM operand [K][N]).
 *   mode 0: OIHW          -> [(r,s,ci)][co]           forward of nn.Conv2d
 *   mode 1: OIHW          -> [(r,s,co)][ci], flipped  data gradient of a stride-1 nn.Conv2d
 *   mode 2: [Cin][Cout]   -> [(co)][ci]               data gradient of ConvTranspose2d 1x1
 * (forward of ConvTranspose2d 1x1 and data gradient of a 1x1 nn.Conv2d use the parameter as is.) */
int fdg_pack_weight(const float* w, int Cout, int Cin, int R, int S, int mode, float* out, int out_ld,
                    fdg_stream_t stream);

/* tcgen05 operand image of a GEMM-form weight w[K = taps*Cin][w_ld] (the output of fdg_pack_weight, or a 1x1 parameter
 * used as is): bf16 hi/lo split, pre-swizzled (SWIZZLE_128B, K-major) per (N tile, 64-deep K chunk) so that the conv
 * kernel fetches each stage with one bulk TMA copy.  fdg_umma_weight_bytes gives the size of `out` (16-byte aligned). */
int64_t fdg_umma_weight_bytes(int taps, int Cin, int Cout);
int fdg_pack_weight_umma(const float* w, int w_ld, int taps, int Cin, int Cout, void* out, fdg_stream_t stream);

/* Operand image for the growth-convolution kernel (3x3, stride 1, pad 1, Cout <= 32; conv_k1.cu): per (64-channel chunk,
 * filter column kx) one [192 x 64] bf16 tile whose rows are the three filter rows' output channels, hi then lo. */
int64_t fdg_k1_weight_bytes(int Cin);
int fdg_pack_weight_k1(const float* w, int w_ld, int Cin, int Cout, void* out, fdg_stream_t stream);

/*
 * BatchNorm2d training-mode bookkeeping (nn.BatchNorm2d as used at README.md:38: always batch statistics).
 * fdg_bn_finalize: stats (fp64 sum, sumsq over count elements per channel) -> scale = gamma*invstd,
 * shift = beta - mean*scale, saved mean/invstd; updates running_mean/var (momentum, unbiased variance)
 * when running_mean != NULL.  With training == 0 scale/shift come from the running statistics.
 */
typedef struct FdgBnFinalize {
  const double* stats; /* [2*stats_ld]: sums at [c], sums of squares at [stats_ld + c] */
  int stats_ld;
  int C;
  double count;
  const float* gamma;
  const float* beta;
  float eps;
  float momentum;
  float* running_mean; /* may be NULL */
  float* running_var;
  int training;
  float* scale;        /* out [C] */
  float* shift;        /* out [C] */
  float* mean;         /* out [C] or NULL */
  float* invstd;       /* out [C] or NULL */
} FdgBnFinalize;

int fdg_bn_finalize(const FdgBnFinalize* p, fdg_stream_t stream);

/*
 * fdg_ew_bwd: the element-wise half of BatchNorm+ReLU / ReLU / LeakyReLU backward.
 *   v  = has_affine ? x*scale[c] + shift[c] : x
 *   dz = gscale * g * (v > 0 ? 1 : slope)            (g optionally read through FDG_GATHER_UP2: avg-pool adjoint)
 *   if stats != NULL:  stats[c] += dz, stats[C + c] += dz*x      (fp64; BatchNorm backward reductions), no store
 *   else:              out (=|+=) coef ? coef[c]*dz + coef[C+c]*x + coef[2C+c] : dz
 */
typedef struct FdgEwBwd {
  FdgTensor g;
  int g_gather;        /* FDG_GATHER_DIRECT or FDG_GATHER_UP2 */
  float gscale;
  FdgTensor x;
  int N, H, W, C;      /* extent of x / out */
  int has_affine;
  const float* scale;
  const float* shift;
  float slope;
  const float* coef;   /* [3*C] alpha, beta, delta or NULL */
  FdgTensor out;
  int accumulate;
  double* stats;       /* [2*C] or NULL */
  void* out_split;     /* apply pass: write the result as split-bf16 planes [N*H*W][C] (hi, then lo), instead of `out` when out.p is NULL,
                          in addition to it otherwise; or NULL */
} FdgEwBwd;

int fdg_ew_bwd(const FdgEwBwd* p, fdg_stream_t stream);

/* BatchNorm backward finalize: from sum(dz), sum(dz*x) -> coef (alpha, beta, delta) such that
 * dx = alpha*dz + beta*x + delta, and dgamma/dbeta (+)=. */
typedef struct FdgBnBwdFinalize {
  const double* stats; /* [2*C] */
  int C;
  double count;
  const float* gamma;
  const float* mean;
  const float* invstd;
  float* coef;         /* out [3*C] */
  float* dgamma;       /* (+)= or NULL */
  float* dbeta;
  int accumulate;
  int unit_alpha;      /* write coef[0..C) = 1: the producer of dz already multiplied it by alpha (FdgConv.e_scale epilogue, normal store) */
  float* acc_beta;     /* += beta  [C], or NULL: running sums of the deferred affine term over the consumers of a concat channel */
  float* acc_delta;    /* += delta [C], or NULL */
} FdgBnBwdFinalize;

int fdg_bn_bwd_finalize(const FdgBnBwdFinalize* p, fdg_stream_t stream);

/* 2x2 max pooling (F.max_pool2d(h, 2, 2), vgg16.py:31,36,42) and its gradient (routed to the first maximum).
 * Backward `accumulate` is a bit set: 1 = add to gx (otherwise every element of the 2x2 blocks is written), 2 = multiply the routed
 * gradient by [max > 0], the ReLU mask of a post-ReLU input (saves the separate mask pass over gx). */
int fdg_maxpool2_fwd(const FdgTensor* x, const FdgTensor* y, int N, int OH, int OW, int C, fdg_stream_t stream);
int fdg_maxpool2_bwd(const FdgTensor* x, const FdgTensor* gy, const FdgTensor* gx, int N, int OH, int OW, int C,
                     int accumulate, fdg_stream_t stream);

/* Strided gather-copy y (=|+=) scale * gather(leaky(x, slope)); N,H,W,C are the extents of y.  Places trans_block2's
 * output into the decoder concat (dehaze1113.py:786, slope 0: the in-place ReLU of BottleneckBlockdy), and is the adjoint
 * of nearest x2 (gather AVGPOOL2, scale 4) and of avg-pool (gather UP2, scale 0.25) in the backward pass. */
int fdg_copy4d(const FdgTensor* x, const FdgTensor* y, int N, int H, int W, int C, int gather, float slope, float scale,
               int accumulate, fdg_stream_t stream);

/* y(n,h,w,c) = mean over the 2x2 block of leaky(x * scale[c] + shift[c], slope): BatchNorm + ReLU + AvgPool2d(2) of a torchvision
 * transition (models/densenet.py:214-221) with the pool commuted in front of the 1x1 convolution, materialised once so that the
 * convolution and its weight gradient run on a quarter of the pixels without gather.  scale / shift may both be NULL.
 * x: [N,2OH,2OW,C], y: [N,OH,OW,C]; unit channel stride, 16-byte aligned, C % 4 == 0 (else FDG_ENOSUPPORT). */
int fdg_pool2_bn_act(const FdgTensor* x, const FdgTensor* y, int N, int OH, int OW, int C, const float* scale, const float* shift, float slope,
                     fdg_stream_t stream);

/* Single-output-channel stride-1 convolutions by taps (Fusion-D layer 5, dehaze1113.py:222: 8nf -> 1, 4x4): the convolution is run as
 * a 1x1 convolution Cin -> R*S (column t = filter tap t; the OIHW weight [1][Cin][R][S] IS that [Cin][R*S] operand) followed by
 *   fdg_tap_sum:     out(n,oy,ox) = act( sum_t s(n, oy+ky-pad, ox+kx-pad)[t] ),  s: [N,H,W,R*S], out: [N,OH,OW,1]
 * and its weight gradient as a 1x1 weight gradient (transposed layout) against
 *   fdg_tap_spread:  gs(n,y,x)[t] = g(n, y-ky+pad, x-kx+pad) (0 outside),          g: [N,OH,OW,1], gs: [N,H,W,R*S]
 * with OH = H + 2 pad - R + 1, OW = W + 2 pad - S + 1. */
int fdg_tap_sum(const FdgTensor* s, const FdgTensor* out, int N, int H, int W, int R, int S, int pad, int act, fdg_stream_t stream);
int fdg_tap_spread(const FdgTensor* g, const FdgTensor* gs, int N, int H, int W, int R, int S, int pad, fdg_stream_t stream);

/* out = g * act'(y) for contiguous arrays: act = FDG_ACT_TANH (1 - y^2) or FDG_ACT_SIGMOID (y (1 - y)). */
int fdg_act_bwd(const float* g, const float* y, float* out, int64_t n, int act, fdg_stream_t stream);

/* Data gradient of a strided nn.Conv2d (the 4x4 stride-2 first layer of D, dehaze1113.py:196), i.e. a transposed
 * convolution evaluated as a gather: dx[n,y,x,ci] (=|+=) sum_{r,s,co} g[n,(y+pad-r)/stride,(x+pad-s)/stride,co] W[co,ci,r,s]. */
typedef struct FdgDgradStrided {
  FdgTensor g;
  int N, OH, OW, Cout;
  const float* w;      /* OIHW parameter */
  int Cin, R, S, stride, pad;
  FdgTensor dx;
  int H, W;
  int accumulate;
} FdgDgradStrided;

int fdg_conv2d_dgrad_strided(const FdgDgradStrided* p, fdg_stream_t stream);

/* out[n,h,w,c] += cb[c] * x[n,h,w,c] + cd[c]: the deferred affine part of the BatchNorm backward (see FdgConv.e_scale). */
int fdg_affine_accum(const FdgTensor* x, const FdgTensor* out, int N, int H, int W, int C, const float* cb, const float* cd,
                     fdg_stream_t stream);

/* per-channel column sum: out[c] (+)= sum over (n,h,w) of x (bias gradients). */
int fdg_colsum(const FdgTensor* x, int N, int H, int W, int C, float* out, int accumulate, fdg_stream_t stream);

/*
 * Frequency decomposition feeding the Fusion-discriminator (loss.pyc Blur@L122-151 with the
 * 15x15 sigma=3 kernel of @L153-162, Laplacian@L245-301; concat per facades/network.png):
 *   z[:, 0:3] = x ; z[:, 3:6] = gauss15x15(reflect_pad7((x - mean)/std)) ; z[:, 6:9] = 3x3 [1..;1 -8 1;..] zero pad.
 * x is [N,3,H,W] (any strides), z is [N,9,H,W] (any strides).  bwd is the exact adjoint: dx = dz0 + Blur^T dz1 + Lap dz2.
 */
int fdg_freq_concat_fwd(const FdgTensor* x, const FdgTensor* z, int N, int H, int W, fdg_stream_t stream);
int fdg_freq_concat_bwd(const FdgTensor* dz, const FdgTensor* dx, float* scratch /* N*3*H*W floats */, int N, int H,
                        int W, fdg_stream_t stream);

/* Generic depth-wise filter: ONE l x l kernel applied to every (image, channel) plane -- the reference's Laplacian on any
 * channel count (loss.pyc@L286-301: kernel.repeat(c,1,1,1), F.conv2d(groups=c), zero padding (k-1)/2) and Blur with a
 * non-default size / kernel / normalisation (loss.pyc@L123-151: optional (x-mean)/std, ReflectionPad2d(l//2), conv per plane).
 *   fwd: y (=|+=) correlate(pad((x - mean[c]) * inv_std[c]), kernel)      mean / inv_std may be NULL (no normalisation)
 *   bwd: y += adjoint applied to x, i.e. x is the incoming gradient dL/d(fwd output) and y accumulates dL/d(fwd input)
 *        (y must be initialised: zeros for a plain gradient).  l odd, 1 <= l <= 31; pad_mode 0 = zero, 1 = reflect. */
typedef struct FdgDepthwise {
  FdgTensor x, y;
  int32_t n, h, w, c;
  const float* kernel; /* l*l taps, row-major (device) */
  int32_t l, pad_mode;
  const float* mean;    /* [c] or NULL (device) */
  const float* inv_std; /* [c] or NULL (device) */
  int32_t accumulate;   /* fwd only */
} FdgDepthwise;
int fdg_depthwise2d_fwd(const FdgDepthwise* d, fdg_stream_t stream);
int fdg_depthwise2d_bwd(const FdgDepthwise* d, fdg_stream_t stream);

/* SSIM term of the generator loss and its gradient (reference arithmetic pytorch_ssim._ssim,
 * models/pytorch_ssim/__init__.py:17-37: window 11, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2):
 *   loss[0] += lscale * sum over (n,c,h,w) of ssim_map(x, y);   grad (=|+=) gscale * d(sum ssim_map)/dx   (grad may be NULL)
 * x, y, grad: strided [N,H,W,C] views of any strides; scratch: 3*N*C*H*W floats.  For the loss w*(1 - mean(ssim_map)) pass
 * lscale = gscale = -w / (N*C*H*W) and add the constant w. */
int fdg_ssim_loss_grad(const FdgTensor* x, const FdgTensor* y, int N, int H, int W, int C, float lscale, float gscale,
                       const FdgTensor* grad, int accumulate, double* loss, float* scratch, fdg_stream_t stream);

/* Output path and quality metrics.
 * fdg_image_minmax: out2 = {min, max} over the whole [N,H,W,C] view; fdg_image_pack_u8: the bytes
 * torchvision.utils.save_image(normalize=True, scale_each=False) would write (demo.py:151), dense uint8 [N][H][W][C].
 * fdg_psnr_ssim_u8: PSNRSSIM.py:201-240 on two uint8 [H][W][3] images: sums4[0] = sum of squared /255 differences over the
 * 1-pixel-cropped images, sums4[1..3] = sum of the 5-pixel-cropped Gaussian-weighted SSIM map of each channel;
 * PSNR = 10 log10(3 (H-2)(W-2) / sums4[0]), SSIM = mean_c sums4[1+c] / ((H-12)(W-12)). */
int fdg_image_minmax(const FdgTensor* x, int N, int H, int W, int C, float* out2, fdg_stream_t stream);
int fdg_image_pack_u8(const FdgTensor* x, int N, int H, int W, int C, const float* minmax, uint8_t* out, fdg_stream_t stream);
int fdg_psnr_ssim_u8(const uint8_t* ref, const uint8_t* res, int H, int W, double* sums4, fdg_stream_t stream);

/* Fused Adam over a flat fp32 buffer (torch.optim.Adam arithmetic; --lrG/--lrD 2e-4, --beta1 0.5: demo.py:43-46).
 * grad_scale multiplies the gradient first (1/world_size after the NCCL sum). */
int fdg_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, fdg_stream_t stream);

/* fdg_adam_flat with the step counter and bias corrections on the device (state: 3 floats {step, 1-b1^t, sqrt(1-b2^t)},
 * zero-initialised by the caller): lets a captured CUDA graph of the training step be replayed unchanged. */
int fdg_adam_flat_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, float* state, float grad_scale, fdg_stream_t stream);

/* Fused scalar loss + its gradient over flat arrays (the training step never builds an autograd graph for these):
 *   kind FDG_LOSS_L1:  loss += scale * sum |a-b|        grad (=|+=) scale * sign(a-b)       (F.l1_loss, mean folded into scale)
 *   kind FDG_LOSS_MSE: loss += scale * sum (a-b)^2      grad (=|+=) 2 scale (a-b)           (F.mse_loss on VGG features)
 *   kind FDG_LOSS_BCE: loss += scale * sum -[t log a + (1-t) log(1-a)], logs clamped at -100 like torch;
 *                      grad (=|+=) scale * (a-t) / max(a (1-a), 1e-12)                      (nn.BCELoss against the constant target t;
 *                      D ends in a Sigmoid, dehaze1113.py:223)
 * loss is a device double accumulated atomically; grad may be NULL (value only). */
#define FDG_LOSS_L1 0
#define FDG_LOSS_MSE 1
#define FDG_LOSS_BCE 2
int fdg_loss_grad(const float* a, const float* b, float target, int kind, int64_t n, float scale, float* grad,
                  int accumulate, double* loss, fdg_stream_t stream);

/* All weight-operand repacks of one network pass in one launch (replaces the per-layer fdg_pack_weight /
 * fdg_pack_weight_umma / fdg_pack_weight_k1 launches; the reference has no counterpart -- cuDNN re-reads the PyTorch
 * layout).  jobs_dev: device array of njobs descriptors sorted by first_block; job j owns blocks
 * [first_block, first_block + nblocks) of 256 threads, total_blocks = sum of nblocks.  Jobs of one launch must not
 * depend on each other (an image made from a packed operand goes into a second launch).
 *   kind 0..2         fdg_pack_weight mode: src = OIHW parameter, dst = fp32 [K][ld];  cout, cin, r, s = filter dims
 *   FDG_PACK_UMMA     fdg_pack_weight_umma: src = fp32 [K][ld] operand, r = taps, s = fdg_umma_tile_code(taps, cin, cout)
 *   FDG_PACK_K1       fdg_pack_weight_k1:   src = fp32 [9*cin][ld] operand                                         */
#define FDG_PACK_UMMA 3
#define FDG_PACK_K1 4
typedef struct FdgPackJob {
  const float* src;
  void* dst;
  int32_t kind, cout, cin, r, s, ld;
  int32_t first_block, nblocks;
  int64_t total; /* work items = fdg_pack_job_items(job) */
} FdgPackJob;
int64_t fdg_pack_job_items(const FdgPackJob* job);
int fdg_umma_tile_code(int taps, int Cin, int Cout); /* fdg_umma_ntile | (two taps per 64-deep chunk: filters of >= 4 taps over <= 32 channels) << 16 */
int fdg_umma_ntile(int taps, int Cout); /* output-channel tile of the packed image: 32 / 64 / 128, or 80 / 96 for filters of >= 4 taps (halo-tile kernel) */
int fdg_pack_batch(const FdgPackJob* jobs_dev, int njobs, int total_blocks, fdg_stream_t stream);

/* Live per-launch timing for bench.py's roofline: when enabled every entry point brackets its kernel with CUDA events
 * on the launching stream; collect() sums elapsed ms, algorithmic flops / bytes and launches per kernel family. */
#define FDG_PROF_FAMILIES 6 /* 0 conv fp32 SIMT, 1 conv tcgen05, 2 weight gradient, 3 element-wise backward, 4 frequency, 5 other */
int fdg_profile_enable(int on);
int fdg_profile_collect(double* ms, double* flops, double* bytes, int64_t* launches);

/* Runtime options: "halo" (default 1) = use the halo-tile tcgen05 kernel for stride-1 RxS convolutions. */
int fdg_set_option(const char* name, int value);

/* Stream fork / join for the host executors (side-stream weight gradients): record a point on `stream` (returns a handle >= 0,
 * valid for the next 1024 records on the device, or a negative FDG_E* code) and make another stream wait for it.  Legal under
 * CUDA-graph stream capture. */
int fdg_event_record(fdg_stream_t stream);
int fdg_stream_wait(fdg_stream_t stream, int handle);

/* diagnostics */
const char* fdg_last_error(void);
int fdg_version(void);
/* number of kernels this library has launched from this process (all threads); bench.py reports the delta */
int64_t fdg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FDGAN_B200_H_ */
