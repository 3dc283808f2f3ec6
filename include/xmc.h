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
int xmc_num_sms(void);

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
} XmcConvDesc;

int xmc_conv2d_fwd(const XmcConvDesc* d, const void* x, const void* wk, const float* bias, const void* residual,
                   const void* mask, void* y, void* stream);

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
  int out_mode;            /* 0 fp32 atomic accumulate, 1 fp32 store, 2 bf16 store (1,2 force a single K split) */
  int ldOut;
  long long out_tap_stride, out_batch_stride;
  float alpha;
} XmcWgradDesc;

int xmc_conv2d_wgrad(const XmcWgradDesc* d, const void* xa, const void* xb, void* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMC_H_ */
