/*
 * xmc.h — C ABI of libxmc.so: the B200-native (sm_100a) kernels behind XMC-GAN's train_step hot path.
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes; the caller owns every buffer (device memory unless stated otherwise);
 *   - no allocation, no synchronisation and no global mutable state inside the library;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the call returns at once;
 *   - returns 0 on success or a negative XMC_E* code; xmc_strerror() gives the message;
 *   - activations are NHWC bf16 ("pixels x channels", channel pitch `ld*` in elements), parameters fp32,
 *     convolution kernels HWIO, dense kernels [in,out] — the layouts of the reference's Flax variables.
 *
 * The reference has no FFI layer: it reaches its device code through XLA. Each entry point therefore cites the
 * reference *Python* call site whose arithmetic it replaces (paths relative to the reference root).
 */
#ifndef XMC_H_
#define XMC_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XMC_OK 0
#define XMC_EINVAL (-1)   /* bad descriptor / unsupported shape */
#define XMC_ECUDA (-2)    /* a CUDA runtime/driver call failed */
#define XMC_EALIGN (-3)   /* pointer or pitch not 16-byte aligned */

const char* xmc_strerror(int code);
/* Last CUDA error string seen by this thread inside the library (diagnostics only). */
const char* xmc_last_cuda_error(void);
int xmc_version(void);
/* sizeof of the ABI structs as compiled into the library: 0 XmcConvDesc, 1 XmcWgradDesc, 2 XmcBnDesc, 3 XmcPrepEntry,
 * 4 XmcSnEntry (-1 otherwise). A binding checks its own mirror of a struct against this before the first call. */
int xmc_sizeof(int which);
int xmc_num_sms(void);
/* Caps the grid of the persistent tensor-core kernels (xmc_conv2d_fwd, xmc_conv2d_wgrad) at `sms` thread blocks so that
 * a collective kernel (NCCL all-reduce of the gradients, reference: jax.lax.pmean in xmc_gan.py:170-171) launched on
 * another stream finds free SMs instead of queueing behind 148 resident one-per-SM blocks. 0 = no limit. Results do not
 * depend on the limit (work is planned on the full SM count). Host-side state, read at launch time. */
int xmc_set_sm_limit(int sms);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / dense / batched GEMM on tcgen05 tensor cores (bf16 operands, fp32 accumulate).
 *   y[n,h,w,co] = epilogue( alpha * sum_{kh,kw,ci} x[n,h+kh-pad_h,w+kw-pad_w,ci] * wk[co][(kh*KW+kw)*C+ci] )
 * x is read through a 4-D TMA tensor map with zero fill outside the image (SAME padding), wk is the K-major bf16
 * copy of the kernel produced by xmc_prep_weights. Replaces flax nn.Conv / lax.conv_general_dilated
 * (xmcgan/libml/layers.py:223-240, xmcgan/nets/common.py:71-75,127-132,153-159,179-185), nn.Dense / dot_general
 * (layers.py:104-112, xmcgan/nets/xmc_net.py:99-100,213-215) and the jnp.matmul calls of attention_lib.py:120,126.
 * Epilogue order: v = alpha*acc + bias[co]; if mask: v = mask>0 ? v : 0; if residual: v += residual; if relu: max(v,0).
 */
typedef struct {
  int N, H, W, C;          /* activation dims; C = input channels (GEMM K per tap) */
  int ldA;                 /* pixel pitch of x in elements (>= C, multiple of 8) */
  int KH, KW, pad_h, pad_w;
  int Cout;                /* GEMM N */
  int ldB;                 /* row pitch of wk in elements (>= KH*KW*C, multiple of 8) */
  long long strideB_batch; /* elements between per-batch B matrices when batched */
  int batched;             /* 1: N is a batch index, B matrix n is used for image n */
  int out_dtype;           /* 0 bf16, 1 fp32 */
  int ldOut;               /* pixel pitch of y in elements */
  float alpha;
  int relu;
  int res_shift;           /* residual is an (N, H>>s, W>>s) tensor read at (n, h>>s, w>>s): fused nearest upsample */
  int ldRes, ldMask;
  /* Optional strided / re-pitched input view (all 0 = dense stride-1 input of the output's size). The input is read
   * through a tensor map of extents (C, Win, Hin, N) with element pitches (pitchW, pitchH, pitchN) and traversal
   * strides (strideW, strideH): output pixel (h, w), tap (kh, kw) reads input position
   * (h*strideH + kh - pad_h, w*strideW + kw - pad_w). Used for the stride-2 convolutions of the frozen ResNet-50
   * (xmcgan/utils/resnet_v1.py:64,79,146-151; XLA SAME padding low=floor(p/2)). */
  int strideH, strideW, Hin, Win;
  long long pitchW, pitchH, pitchN;
  int mask_last;           /* 1: apply the mask after the residual add: v = (alpha*acc + bias + residual) masked */
  /* 1: sub-pixel mode = conv3x3(nearest_upsample2x(x)) computed as four 2x2 convolutions on x (one per output
   * parity) with pre-summed weights: KH=KW=2, pad=1, wk = [4*Cout][4*C] from xmc_subpixel_prep, y is the
   * [N,2H,2W,Cout] tensor and output pixel (2h+a, 2w+b) belongs to parity a*2+b. 2.25x fewer FLOPs than the
   * reference's upsample-then-conv (xmcgan/nets/common.py:151-153,178-179), identical in exact arithmetic. */
  int subpixel;
  /* 1: fp32-activation mode (config.dtype = "float32"; the frozen ResNet branch, which the reference always runs in
   * fp32, pretrained_model_utils.py:87-91). x is the bf16 [hi | lo | hi] split of the fp32 input made by xmc_split3
   * (C = 3 x the real channel count), wk the matching [hi | hi | lo] weight copy of xmc_prep_weights' split mode:
   * the three bf16 products hi*hi + lo*hi + hi*lo accumulated in fp32 carry 16 mantissa bits per operand (SURVEY.md
   * 8c(4) "3 x bf16 split"). residual / mask are fp32 tensors, out_dtype must be 1.
   * 2: the same with the two-part operand: x is bf16 [.., hi(Cr) | lo(Cr)] (Cr = C / 3 real channels, Cr % 64 == 0,
   * from xmc_split3 mode 2 or from a previous launch's y_pair) and the kernel reads the hi part twice. */
  int act_f32;
  int ldPair;              /* pixel pitch of y_pair (>= 2 * Cout), see xmc_conv2d_fwd */
  /* fp32-activation mode: residual / mask given as two-part bf16 tensors [.., hi(Cout) | lo(Cout)] (ldRes / ldMask in
   * bf16 elements) instead of fp32 ones: residual = hi + lo, mask = [hi > 0]. With y_pair as the only output (y NULL) a
   * chain of convolutions never materialises fp32 activations. */
  int res_pair, mask_pair;
} XmcConvDesc;

/* y_pair: optional (NULL) second output in fp32-activation mode: the result once more as bf16 [.., hi(Cout) |
 * lo(Cout)] with pixel pitch d->ldPair — the operand form the next launch consumes with act_f32 = 2, so that a chain of
 * convolutions needs no separate split pass. Needs Cout % 16 == 0 and 32-byte aligned pointers / pitches. */
int xmc_conv2d_fwd(const XmcConvDesc* d, const void* x, const void* wk, const float* bias, const void* residual,
                   const void* mask, void* y, void* y_pair, void* stream);

/* Weight gradient / "A^T B" GEMM with both operands pixel-major (MN-major UMMA descriptors), split-K over pixels:
 *   dw[b][tap][ca][cb] (+)= alpha * sum_{pixels p of batch b} xa[p + shift(tap)][ca] * xb[p][cb]
 * Replaces the kernel cotangent of lax.conv_general_dilated / dot_general that jax.vjp derives
 * (xmcgan/xmc_gan.py:162-167,247-248) and jnp.matmul(alpha^T, region_feat) of attention_lib.py:126.
 */
typedef struct {
  int N, H, W;
  int Ca, ldA;             /* xa channels (GEMM M) and pixel pitch */
  int Cb, ldB;             /* xb channels (GEMM N) and pixel pitch */
  int KH, KW, pad_h, pad_w;
  int batched;             /* 1: one output matrix per image n (no reduction over n) */
  int out_mode;            /* 0 fp32 accumulate (dw += ...), 1 fp32 store (dw = ...; saves the read of dw), 2 bf16 store
                            * (one K split). 0 and 1 are deterministic and take every decomposition (K split, tap groups) */
  int ldOut;
  long long out_tap_stride, out_batch_stride;
  float alpha;
  /* 1: weight gradient of conv3x3(nearest_upsample2x(xa)): xa is the low-resolution [N,H,W,Ca] input, xb the
   * [N,2H,2W,Cb] output gradient (read with stride 2, one parity per tap); the 16 parity/tap products are added to
   * the 9 taps of dw ([3][3][Ca][Cb]) they belong to. KH=KW=3, out_mode 0 or 1.
   * 2: weight gradient of dsample(conv3x3(xa)) (pool-fused form, xmc_poolconv_prep): xa is the full-resolution
   * [N,2H,2W,Ca] input (read with stride 2, one 4x4 tap per work item), xb the LOW-resolution [N,H,W,Cb] output
   * gradient; the 16 tap products are folded into the 9 taps of dw (pass alpha = 0.25 for the mean). */
  int subpixel;
  /* Optional re-pitched view of xa (all 0 = dense [N,H,W,ldA]): xa is read through a tensor map of extents
   * (Ca, W, HinA, N) with element pitches (pitchWA, pitchHA, pitchNA); tap (kh, kw) reads position
   * (h + kh - pad_h, w + kw - pad_w). Used for the packed-window form of the 3-channel image convolutions (an
   * 8-channel zero-bordered image whose 3 kw taps x 8 channels are one contiguous 24-element run: Ca=24, KH=3, KW=1). */
  int HinA;
  long long pitchWA, pitchHA, pitchNA;
} XmcWgradDesc;

/* out_mode 0 is DETERMINISTIC: when the reduction over pixels is split across CTAs (or sub-pixel parity taps share a
 * destination) every CTA stores its partial tile into the caller's workspace and a second kernel adds the partials
 * to dw in a fixed order — two launches on the same inputs give bit-identical results (XLA's behaviour; round 1 used
 * red.global.add). xmc_conv2d_wgrad_workspace_bytes reports the bytes needed (0 when no second stage is needed);
 * workspace must be 16-byte aligned and may be reused by the next launch on the same stream. */
int xmc_conv2d_wgrad_workspace_bytes(const XmcWgradDesc* d, long long* bytes);
int xmc_conv2d_wgrad(const XmcWgradDesc* d, const void* xa, const void* xb, void* dw, void* workspace,
                     long long workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Conditional batch normalisation (train mode), fused with relu and nearest 2x upsampling.
 * Replaces flax nn.BatchNorm(use_scale=False,use_bias=False) + `x*(gamma+1)+beta` + relu + common.upsample
 * (xmcgan/libml/layers.py:256-257,271-272; xmcgan/nets/common.py:48-51,148-151,175-178).
 * gamma/beta for pixel (n,h,w) are read from row (n*Hc + h/(H/Hc))*Hc + w/(W/Hc) of the bf16 matrix gb at column
 * offsets goff / boff: Hc=1 is ConditionalBatchNorm, Hc=cond_size(16) is LocalConditionalBatchNorm with its 1x1
 * convolutions evaluated once at 16x16 (a 1x1 conv commutes with nearest upsampling).
 */
typedef struct {
  int N, H, W, C;  /* dims of x (before upsampling); H == W, H = Hc * 2^s */
  int Hc;
  int ldG, goff, boff;
  int relu, upsample;
  /* cross-replica BatchNorm (nn.BatchNorm axis_name="batch" with axis_index_groups, xmc_net.py:192-201): number of
   * replicas whose statistics were summed into `sums` by the caller's all-reduce; 0 or 1 = replica-local. Only
   * xmc_bn_bwd_apply reads it (its means divide by N*H*W*replicas). */
  int replicas;
  int act_f32;             /* 0: x / gb / y / dy / dx are bf16 tensors, 1: fp32 (config.dtype = "float32") */
} XmcBnDesc;

/* Reductions are DETERMINISTIC (no atomics): a kernel with `rows` thread blocks leaves one partial row per block in
 * the caller's `partials` buffer (rows x width floats) and a second kernel adds the rows in index order. `rows` is
 * the caller's choice (any value >= 1; a few per SM for large tensors, see xmc_num_sms). */
int xmc_sum_partials(const float* partials, int rows, int width, float* out, int accumulate, void* stream);
/* sums[2C] = per-channel (sum x, sum x^2) over P pixels (overwritten); partials: rows x 2C floats of scratch */
int xmc_bn_stats(const void* x, int act_f32, long long P, int C, int ld, float* sums, float* partials, int rows,
                 void* stream);
int xmc_bn_finalize(const float* sums, long long P, int C, float eps, float momentum, const float* ra_mean,
                    const float* ra_var, float* new_ra_mean, float* new_ra_var, float* mean_rstd, void* stream);
int xmc_bn_eval_stats(const float* ra_mean, const float* ra_var, int C, float eps, float* mean_rstd, void* stream);
int xmc_bn_apply(const XmcBnDesc* d, const void* x, const float* mean_rstd, const void* gb, void* y, void* stream);
/* backward, pass 1: dgb (fp32, same row/column layout as gb) = d(gamma), d(beta) (overwritten); sums[2C] =
 * per-channel sum(dxhat), sum(dxhat*xhat) (overwritten). dy has the upsampled shape when d->upsample.
 * partials: rows x 2C floats of scratch (see xmc_sum_partials). */
int xmc_bn_bwd_reduce(const XmcBnDesc* d, const void* dy, const void* x, const float* mean_rstd, const void* gb,
                      float* dgb, float* sums, float* partials, int rows, void* stream);
/* backward, pass 2: dx = rstd*(dxhat - mean(dxhat) - xhat*mean(dxhat*xhat)) */
int xmc_bn_bwd_apply(const XmcBnDesc* d, const void* dy, const void* x, const float* mean_rstd, const void* gb,
                     const float* sums, void* dx, void* stream);

/* 2x2/2 pooling: out = scale * sum_{2x2}(a (+ b)) (+ low); optional out_relu = relu(out). scale=0.25 is
 * common.dsample (xmcgan/nets/common.py:23-45,54-55); scale=1 is the transpose of common.upsample.
 * a, b: [N,2*Hout,2*Wout,C]; low: [N,Hout,Wout,C]. */
int xmc_pool2(const void* a, const void* b, const void* low, int act_f32, int N, int Hout, int Wout, int C,
              float scale, void* out, void* out_relu, void* stream);
/* transpose of pool2: g[n,2h+i,2w+j,c] = scale * dout[n,h,w,c] */
int xmc_unpool2(const void* dout, int act_f32, int N, int Hin, int Win, int C, float scale, void* g, void* stream);
/* out[c] += sum_p x[p][c] (bias gradients); partials: rows x C floats of scratch, may be NULL when rows == 1 */
int xmc_colsum(const void* x, int act_f32, long long P, int C, int ld, float* out, float* partials, int rows,
               void* stream);
/* x_pool[n][c] = sum_hw relu(x[n,hw,c]) (xmcgan/nets/xmc_net.py:97-98) and its backward */
int xmc_relu_sumhw(const void* x, int act_f32, int N, int HW, int C, float* out, void* stream);
int xmc_relu_sumhw_bwd(const void* x, int act_f32, const float* dout, int N, int HW, int C, void* dx, void* stream);
/* y = relu(a) when b is NULL (flax nn.relu), else y = a + b (the residual add of nets/common.py's blocks): the
 * stand-alone elementwise ops of the module-level API. n elements (multiple of 8), contiguous. */
int xmc_relu_or_add(const void* a, const void* b, int act_f32, long long n, void* y, void* stream);
/* fp32 [rows][C] (pitch ld_src) -> bf16 [rows][hi(C) | lo(C) | hi(C)] (pitch ld_dst >= 3C): the A operand of
 * xmc_conv2d_fwd in fp32-activation mode (XmcConvDesc.act_f32); xmc_conv2d_wgrad reads the hi / lo parts as views.
 * mode 0: [hi | lo | hi]; mode 1: the B-operand order [hi | hi | lo] (an activation used as the second GEMM
 * operand); mode 2: the two-part form [hi | lo] (ld_dst >= 2C) for XmcConvDesc.act_f32 = 2. */
int xmc_split3(const float* src, long long rows, int C, long long ld_src, int mode, void* dst, long long ld_dst,
               void* stream);
int xmc_cast_f32_to_bf16(const float* src, long long rows, int cols, long long ld_src, void* dst, long long ld_dst,
                         void* stream);
int xmc_cast_bf16_to_f32(const void* src, long long rows, int cols, long long ld_src, float* dst, long long ld_dst,
                         int accumulate, void* stream);
/* dst[b*reps + r][c] = src[b][c]  (jnp.tile of global_cond, xmc_net.py:233-234) and its transpose */
/* Input-contract producer for decoded examples (COCODataset.preprocess, xmcgan/libml/coco_dataset.py:127-167):
 * xmc_prep_image:   out[n] = clip(flip[n] ? fliplr(img[n]) : img[n], 0, 1), fp32 [N,H,W,3]
 * xmc_prep_caption: picks caption idx[n] of M: emb_out [N,L,E] = emb[n][idx[n]], len_out [N] = (float)len[n][idx[n]],
 *                   sent_out [N,E] = sum over all L word slots of emb_out / len_out  (coco_dataset.py:139-142,156-158) */
int xmc_prep_image(const float* img, const unsigned char* flip, int N, int H, int W, float* out, void* stream);
int xmc_prep_caption(const float* emb, const int* len, const int* idx, int N, int M, int L, int E, float* emb_out,
                     float* len_out, float* sent_out, void* stream);
int xmc_bcast_rows(const void* src, int act_f32, int B, int reps, int cols, int ld_src, void* dst, int ld_dst,
                   void* stream);
int xmc_sum_rows(const void* src, int act_f32, int B, int reps, int cols, int ld_src, float* dst, int ld_dst,
                 int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Parameter-side multi-tensor kernels. Tables live in device memory; offsets index flat fp32 buffers.
 */
typedef struct {
  long long w_off;        /* kernel [taps][cin][cout] (HWIO / [in,out]) inside the flat fp32 parameter buffer */
  long long wk_fwd_off;   /* bf16 arena offset of the [cout][taps*cin] forward copy, or -1 */
  long long wk_dg_off;    /* bf16 arena offset of the [cin][flip(tap)*cout+co] dgrad copy, or -1 */
  long long bias_off;     /* bias inside the parameter buffer, or -1 */
  long long bias_dst_off; /* destination inside the fp32 bias arena, or -1 */
  int taps, cin, cout;
  int ld_fwd, ld_dg;
  int sn;                 /* spectral-norm slot, or -1 */
  int tile_begin;         /* prefix sum of ceil(taps*cin/64)*ceil(cout/64) */
  /* 1: split copies for the fp32-activation mode (XmcConvDesc.act_f32). Every weight w is written as hi = bf16(w) and
   * lo = bf16(w - hi) in the part order [hi | hi | lo] that pairs with the activations' [hi | lo | hi]:
   *   forward copy [cout][tap][part][cin]  (ld_fwd >= 3*taps*cin),
   *   dgrad copy   [cin][flip(tap)][part][dg_part_stride]  (ld_dg >= 3*taps*cout). */
  int split;
  long long cscale_off;   /* per-output-channel scale inside the fp32 `cscale` buffer (folded eval BatchNorm), or -1 */
  /* elements between the parts of the dgrad copy; 0 = cout. Layers whose dgrad copies are column slices of one
   * concatenated matrix (the gamma / beta layers of the conditional BatchNorms) give the full matrix width here. */
  long long dg_part_stride;
} XmcPrepEntry;

typedef struct {
  long long w_off;        /* kernel viewed as [rows][cols] = [taps*cin][cout] */
  long long t_off;        /* row-vector workspace offset (t = W u0^T) */
  long long s_off;        /* workspace offset of the [ceil(rows/256)][cols] row-tile partials of s = t W */
  long long u_off;        /* offset of u0 inside the spectral_norm_stats buffers */
  int rows, cols;
  int row_block_begin;    /* prefix sum of ceil(rows/8) */
  int col_tile_begin;     /* prefix sum of ceil(rows/256)*ceil(cols/32) */
  int elem_block_begin;   /* prefix sum of ceil(rows*cols/2048) */
  int reserved;
} XmcSnEntry;

/* One power-iteration step for every layer of the table (layers.py:94-101 / :211-221). scalars: float[4*n] =
 * unused | norm_t | 1/(sigma+eps) | <dW,W>. Writes the new u0 to u0_new (old state is left untouched). Deterministic:
 * s_ws holds per-row-tile partial sums (sum over layers of ceil(rows/256)*cols floats) that are added in tile order. */
int xmc_sn_forward(const XmcSnEntry* table_dev, int n, float eps, const float* params, const float* u0, float* u0_new,
                   float* t_ws, float* s_ws, long long s_ws_floats, float* scalars, int total_row_blocks,
                   int total_col_tiles, void* stream);
/* grads (holding d/dW~) -> d/dW in place: dW = dW~/s' - <dW~,W>/s'^2 * v0^T u1 (u1, v0 stop-gradient).
 * dot_partials: total_elem_blocks floats of scratch (block partials of <dW~,W>, added in a fixed order). */
int xmc_sn_backward(const XmcSnEntry* table_dev, int n, const float* params, float* grads, const float* t_ws,
                    const float* u0_new, float* scalars, float* dot_partials, int total_elem_blocks, void* stream);
int xmc_prep_weights(const XmcPrepEntry* table_dev, int n, int total_tiles, const float* params,
                     const float* sn_scalars, int n_sn, void* arena, float* bias_arena, const float* cscale,
                     void* stream);
/* Weights of the sub-pixel form of conv3x3(nearest_upsample2x(x)) (xmcgan/nets/common.py:151-153,178-179) from the
 * fp32 HWIO kernel w [3][3][Cin][Cout]: wf bf16 [4*Cout][4*Cin] for XmcConvDesc.subpixel, vd bf16 [Cin][16*Cout] for
 * the input gradient (= xmc_conv2d_fwd with KH=KW=4, stride 2, pad 1 over the [N,2H,2W,Cout] output gradient).
 * scale: optional device scalar 1/(sigma+eps) of a spectrally normalised kernel (layers.py:221), NULL = 1. */
/* Weights of the pool-fused form of dsample(conv3x3(x)) (xmcgan/nets/common.py:66-78,125-132) from the fp32 HWIO
 * kernel w [3][3][Cin][Cout]: wf4 bf16 [Cout][16*Cin] for xmc_conv2d_fwd with KH=KW=4, stride 2, pad 1 on the
 * full-resolution input (the 2x2 mean and its 1/4 are folded into the 16 summed taps: 2.25x fewer FLOPs than
 * conv-then-pool, no full-resolution output), wdg bf16 [4*Cin][4*Cout] for its input gradient = xmc_conv2d_fwd in
 * sub-pixel mode on the low-resolution output gradient. scale / split as xmc_subpixel_prep. The weight gradient is
 * XmcWgradDesc.subpixel = 2. */
int xmc_poolconv_prep(const float* w, const float* scale, int Cin, int Cout, int split, void* wf4, void* wdg,
                      void* stream);
int xmc_subpixel_prep(const float* w, const float* scale, int Cin, int Cout, int split, void* wf, void* vd,
                      void* stream);
/* flax.optim.Adam.apply_gradient (weight_decay 0) in place on flat buffers, gradient pre-scaled by grad_scale
 * (1/world after a sum all-reduce == lax.pmean); optional polyak EMA of the updated parameters
 * (xmcgan/xmc_gan.py:172-177,252). bias_corr = 1 - beta^t. n must be a multiple of 4.
 * step_dev: optional device int holding the number of steps taken so far; when given, the bias corrections are
 * computed on the device from t = *step_dev + 1 (bias_corr1/2 are ignored) and, with advance_step != 0, *step_dev is
 * incremented afterwards, so that the launch can be replayed from a CUDA graph. A step applied in several calls over
 * sub-ranges of the buffers (each as soon as its slice of the gradient all-reduce has arrived) advances on the last.
 * The hyper-parameters are doubles: flax forms (1 - beta) and the bias corrections in Python double precision before
 * they meet the fp32 arrays; 1.f - 0.999f is 1.3e-5 off, which is visible in the second moment. */
int xmc_adam(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
             double eps, double bias_corr1, double bias_corr2, double grad_scale, float* ema, double ema_decay,
             int* step_dev, int advance_step, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Row l2-normalisation xhat = x*rsqrt(max(sum x^2, eps)) (attention_lib.l2_normalize, attention_lib.py:30-33), one
 * warp per row, and its backward dx = (dxhat - xhat<dxhat,xhat>)*invnorm. *_f32 flags select fp32 (1) or bf16 (0).
 */
int xmc_l2norm_rows(const void* x, int in_f32, long long rows, int D, long long ld_in, void* y, int out_f32,
                    long long ld_out, float* invnorm, float eps, void* stream);
int xmc_l2norm_rows_bwd(const void* dxhat, int g_f32, long long ld_g, const void* xhat, int x_f32, long long ld_x,
                        const float* invnorm, long long rows, int D, void* dx, int out_f32, long long ld_o,
                        int accumulate, void* stream);

/* attention_lib.attention_for_g (attention_lib.py:194-219), fused QK^T - masked softmax over words - V, one warp per
 * region: q bf16 [B*R][ld_q] (un-normalised regions), what fp32 [B][L][D] (l2-normalised words), max_len fp32 [B].
 * Writes ctx bf16 [B*R][ld_ctx] and the attention weights attn fp32 [B*R][L]. bwd returns dq (words are data). */
int xmc_attention_g_fwd(const void* q, int act_f32, int ld_q, const float* what, const float* max_len, int B, int R,
                        int L, int D, float gamma, void* ctx, int ld_ctx, float* attn, void* stream);
int xmc_attention_g_bwd(const void* dctx, int act_f32, int ld_dctx, const void* q, int ld_q, const float* what,
                        const float* attn, int B, int R, int L, int D, float gamma, void* dq, int ld_dq,
                        void* stream);

/* Elementwise / reduction stages of attention_lib.word_loss (attention_lib.py:105-191); i image, j sentence, w word,
 * r region, jw = j*L+w, BL = B*L, ldS = BL rounded up to 8. The three GEMM stages use xmc_conv2d_fwd/_wgrad.
 *   wl_softmax : alpha[i][r][jw] = softmax_r(gamma1*S[i*R+r][jw]) (+ alphaT[i][jw][r]), both bf16
 *   wl_cos     : cos[i][jw] = <W_jw, ctx[i][jw]>/(|W_jw||ctx|), cnorm = |ctx|     (winv = 1/|W_jw|)
 *   wl_sim     : sim[j][i] = gamma3*LSE_{w<max_len[j]}(gamma2*cos)/gamma2, pw = the softmax weights of that LSE
 *   wl_cos_bwd : dctx (bf16) from dsim[j][i];  wl_softmax_bwd: dS (bf16) from dalpha (fp32)
 */
int xmc_wl_softmax(const float* S, int B, int R, int BL, int ldS, float gamma1, void* alpha, void* alphaT,
                   int act_f32, void* stream);
int xmc_wl_softmax_bwd(const void* alpha, const float* dalpha, int B, int R, int BL, int ldS, float gamma1, void* dS,
                       int act_f32, void* stream);
int xmc_wl_cos(const float* ctx, long long ctx_batch_stride, const float* words, const float* winv, int B, int L,
               int D, float* cosv, float* cnorm, void* stream);
int xmc_wl_sim(const float* cosv, const float* max_len, int B, int L, float gamma2, float gamma3, float* sim,
               float* pw, void* stream);
int xmc_wl_cos_bwd(const float* dsim, const float* pw, const float* cosv, const float* cnorm, const float* ctx,
                   long long ctx_batch_stride, const float* words, const float* winv, int B, int L, int D,
                   float gamma3, void* dctx, long long dctx_batch_stride, int act_f32, void* stream);
/* dst[c][r] = src[r][c] for r<rows, zero for rows <= r < ld_dst */
int xmc_transpose_bf16(const void* src, int act_f32, int rows, int cols, int ld_src, void* dst, int ld_dst,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loss heads. InfoNCE = l2norm_rows + small_gemm_nt (logits = scale*A B^T) + ce_sym, local-batch negatives only
 * (attention_lib.contrastive_loss, attention_lib.py:46-79).
 */
int xmc_small_gemm_nt(const float* A, const float* B, int n, int m, int D, float scale, float* C, void* stream);
/* out[i][f] (+)= scale * sum_j G(i,j) X[j][f], G(i,j) = transposed ? G[j][i] : G[i][j];  G is [n][m] ([m][n] if transposed) */
int xmc_small_gemm_nn(const float* G, int transposed, const float* X, int n, int m, int D, float scale, float* out,
                      int accumulate, void* stream);
/* *loss_out = mean_i CE(row i) + mean_j CE(col j) with identity labels (losses.py:47-51); dlogits = weight*dloss/dlogits */
int xmc_ce_sym(const float* logits, int n, float weight, float* loss_out, float* dlogits, void* stream);
/* attention_lib.get_statistics (attention_lib.py:36-43) for both directions of an [n][n] logit matrix with identity
 * labels: out[0] = accuracy (argmax == label, first maximum wins; bit-exact index op), out[1] = entropy. These side
 * statistics are dead on the train path (XLA removes them); they exist for the functional API. */
/* losses.tf_cross_entropy_loss_with_logits (losses.py:47-51) for arbitrary labels [rows][n]: one loss per row */
int xmc_softmax_xent(const float* labels, const float* logits, long long rows, int n, float* out, void* stream);
int xmc_ce_stats(const float* logits, int n, float* out, void* stream);
/* losses.hinge_loss (losses.py:30-35) on logit = [real(B); fake(B)] and the two cotangents */
int xmc_hinge(const float* logit, int B, float* d_loss, float* g_loss, float* dlogit_d, float* dlogit_g,
              void* stream);
/* projection-discriminator logit (xmc_net.py:97-104): out[n] = <xpool[n], w1*inv_sigma + emb[n%B]> + b1 */
int xmc_proj_logit(const float* xpool, const float* w1, const float* inv_sigma, const float* b1, const float* emb,
                   int N2, int B, int C, float* out, void* stream);
int xmc_proj_logit_bwd(const float* dlogit, const float* xpool, const float* w1, const float* inv_sigma,
                       const float* emb, int n0, int cnt, int B, int C, float* dxpool, int accumulate_x, float* dw1,
                       float* db1, float* demb, void* stream);
int xmc_colsum_f32(const float* x, int rows, int cols, int ld, float* out, void* stream);
int xmc_axpy_f32(float* y, const float* x, float a, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * 3-channel (image-side) convolutions: first discriminator block (common.py:125-133) and generator head
 * (xmc_net.py:245-247). w is a bf16 [rows][ldw] K-major matrix from xmc_prep_weights.
 */
/* Packed-window form of the 3x3 convolutions with 3 image channels on the input side (first discriminator conv,
 * xmc_net.py:88 / common.py:125-127; input gradient of the generator's output conv): they run on the tcgen05 GEMM
 * kernels over an 8-channel zero-bordered copy of the image.
 *   xmc_pad_c3_to_c8:   x bf16 [N,H,W,3] -> out bf16 [N,H+2,W+2,8] (border and channels 3..7 zero; writes all of out)
 *   xmc_pack_c3_weights: w bf16 [Cout][ldw] with k = (kh*3+kw)*3+c -> out bf16 [Cout][72] with k = kh*24 + kw*8 + c
 *   xmc_unpack_c3_wgrad: tmp fp32 [3 kh][24 = kw*8+c][C] (from xmc_conv2d_wgrad on the packed view) is ADDED to
 *                        out[tap_o*s_tap + c3*s_c3 + c*s_c], tap_o = flip ? 8-tap : tap (as xmc_wgrad_c3) */
int xmc_pad_c3_to_c8(const void* x, int N, int H, int W, void* out, void* stream);
int xmc_pack_c3_weights(const void* w, int ldw, int Cout, void* out, void* stream);
int xmc_unpack_c3_wgrad(const float* tmp, int C, int flip, long long s_tap, int s_c3, int s_c, float* out,
                        void* stream);
int xmc_conv_c3_in(const void* x, int act_f32, const void* w, int ldw, const float* bias, int N, int H, int W,
                   int Cout, int KH, int KW, int relu, void* y, void* stream);
/* mode 0: y fp32 (+= if accumulate); mode 1: y = (tanh(v)+1)/2 fp32 plus bf16 copy y_bf16 */
int xmc_conv_c3_out(const void* x, int act_f32, const void* w, int ldw, const float* bias, int N, int H, int W,
                    int Cin, int KH, int KW, int mode, int accumulate, float* y, void* y_bf16, void* stream);
/* out[tap_o*s_tap + c3*s_c3 + c*s_c] += sum_p x3[p+shift(tap)][c3]*y[p][c], tap_o = flip ? taps-1-tap : tap.
 * Deterministic: partials = N*ceil(H/8)*ceil(W/64) x (KH*KW*3*C) floats of scratch, added in block order. */
int xmc_wgrad_c3(const void* x3, const void* y, int act_f32, int N, int H, int W, int C, int KH, int KW, int flip,
                 long long s_tap, int s_c3, int s_c, float* out, float* partials, void* stream);
int xmc_pool2_small(const void* a, int act_f32, int N, int Hout, int Wout, int C, float scale, void* out,
                    void* stream);
int xmc_unpool2_add_f32(const float* d, int N, int Hin, int Win, int C, float scale, float* g, void* stream);
int xmc_tanh01_bwd(const float* dimg, const float* img, long long n, void* dpre, int act_f32, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Frozen ResNet-50 feature branch (xmcgan/xmc_gan.py:74-90, xmcgan/utils/resnet_v1.py:129-172,
 * xmcgan/utils/pretrained_model_utils.py:102-127). Its convolutions use xmc_conv2d_fwd (BatchNorm folded).
 */
/* jax.image.resize(img, (T,T), "bilinear") of fp32 [N,S,S,3] into a zero-bordered bf16 [N,Tp,Tp,8] buffer */
int xmc_resize_bilinear_pad(const float* img, int N, int S, int T, int Tp, int pad_lo, int split, void* out,
                            void* stream);
/* transpose of the resize: dimg[N,S,S,3] += R^T dout[N,T,T,3] */
int xmc_resize_bilinear_bwd(const float* dout, int N, int S, int T, float* dimg, void* stream);
/* Input gradient of the 7x7 / stride-2 stem, col2im half: cols fp32 [N,Ho,Ho,ldc] holds, per stem output pixel, the
 * 147 products (kh, kw, c) of dy with the transposed stem matrix (made by xmc_conv2d_fwd as a 1x1 convolution over dy);
 * dimg [N,T,T,3] = the gather over the taps that hit each input pixel (overwritten). */
int xmc_stem_col2im(const float* cols, int N, int T, int Ho, int ldc, int pad_lo, float* dimg, void* stream);
/* nn.max_pool 3x3/2 SAME (resnet_v1.py:154) on bf16 [N,H,H,C] and its transpose (first-max tie rule) */
int xmc_maxpool3s2(const void* x, int act_f32, int N, int H, int C, void* y, void* stream);
int xmc_maxpool3s2_bwd(const void* dy, const void* x, const void* y, int act_f32, int N, int H, int C, void* dx,
                       void* stream);
/* z[n,2h,2w,:] = dy[n,h,w,:], zero elsewhere (transpose of stride-2 sampling) */
int xmc_zero_insert2(const void* dy, int act_f32, int N, int H, int W, int C, void* z, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMC_H_ */
