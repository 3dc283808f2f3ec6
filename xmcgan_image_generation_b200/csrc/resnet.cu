// Kernels that only the frozen ResNet-50 feature branch needs (xmcgan/xmc_gan.py:74-90, xmcgan/utils/resnet_v1.py,
// xmcgan/utils/pretrained_model_utils.py:102-127): bilinear resize 128->224 (+ transpose), the 7x7/2 stem's input
// gradient, 3x3/2 max-pool (+ transpose) and zero insertion (transpose of a stride-2 sampling). All convolutions of
// the network itself run on the tcgen05 implicit-GEMM kernel (gemm.cu) with eval-mode BatchNorm folded into weights.
#include "common.h"
#include "devutil.cuh"

namespace xmc {

// jax.image.resize(..., "bilinear") (pretrained_model_utils.py:118-121): half-pixel centres, triangle kernel widened by
// the down-scaling factor (anti-aliasing: kernel scale ks = max(S/T, 1)), taps outside the image dropped and the rest
// renormalised. Up-sampling (128 -> 224) has 2 taps per axis, the 256 -> 224 case up to 3; kMaxTaps bounds S/T <= 1.5.
// One axis: output index o -> taps j in [lo, lo+cnt) with normalised weights w[].
constexpr int kMaxTaps = 4;
__device__ __forceinline__ int resize_taps(int o, int S, float inv_scale, float ks, float* w) {
  const float center = (o + 0.5f) * inv_scale - 0.5f;
  int lo = (int)ceilf(center - ks), hi = (int)floorf(center + ks);
  lo = max(lo, 0);
  hi = min(hi, S - 1);
  int cnt = hi - lo + 1;
  if (cnt > kMaxTaps) cnt = kMaxTaps;
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < kMaxTaps; ++t) {
    const float v = (t < cnt) ? fmaxf(0.f, 1.f - fabsf((float)(lo + t) - center) / ks) : 0.f;
    w[t] = v;
    sum += v;
  }
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
  for (int t = 0; t < kMaxTaps; ++t) w[t] *= inv;
  return lo;
}

// Writes a zero-bordered, 8-channel bf16 buffer [N, Tp, Tp, 8] (image at offset pad_lo) so that the stride-2 7x7 stem
// can read (kw, c) as one contiguous 56-element run per output pixel.
// split = 1 (fp32 mode): channels 3..5 hold the bf16 remainders lo = x - hi of channels 0..2, which the stem GEMM
// multiplies with the same weights (see ResNetEngine: pass 1 [w_hi, w_hi], pass 2 [w_lo, 0]).
__global__ void resize_bilinear_pad_kernel(const float* __restrict__ img, int N, int S, int T, int Tp, int pad_lo,
                                           int split, bf16* __restrict__ out) {
  const long long total = (long long)N * Tp * Tp;
  const float inv_scale = (float)S / (float)T;
  const float ks = fmaxf(inv_scale, 1.f);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xp = idx % Tp, yp = (idx / Tp) % Tp;
    const long long n = idx / ((long long)Tp * Tp);
    const int x = xp - pad_lo, y = yp - pad_lo;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    if (x >= 0 && x < T && y >= 0 && y < T) {
      float wx[kMaxTaps], wy[kMaxTaps];
      const int x0 = resize_taps(x, S, inv_scale, ks, wx);
      const int y0 = resize_taps(y, S, inv_scale, ks, wy);
      const float* b = img + n * S * S * 3;
#pragma unroll
      for (int ty = 0; ty < kMaxTaps; ++ty) {
        if (wy[ty] == 0.f) continue;
#pragma unroll
        for (int tx = 0; tx < kMaxTaps; ++tx) {
          if (wx[tx] == 0.f) continue;
          const float wgt = wy[ty] * wx[tx];
          const float* px = b + ((long long)(y0 + ty) * S + x0 + tx) * 3;
          o[0] += wgt * px[0]; o[1] += wgt * px[1]; o[2] += wgt * px[2];
        }
      }
    }
    if (split) {
#pragma unroll
      for (int c = 0; c < 3; ++c) o[3 + c] = o[c] - __bfloat162float(__float2bfloat16(o[c]));
    }
    store8(out + idx * 8, o);
  }
}

// transpose of the resize as a GATHER (one thread per source pixel, fixed summation order -> deterministic; round 1
// scattered with atomics): dimg[n, sy, sx] += sum over the destination pixels (oy, ox) whose taps cover (sy, sx) of
// wy * wx * dout[n, oy, ox]. Per axis the candidates are the outputs whose centre lies within ks (+ a safety margin
// of one output) of the source index; each candidate's exact normalised weight comes from resize_taps.
constexpr int kMaxCand = 10;
__device__ __forceinline__ int resize_candidates(int s, int S, int T, float inv_scale, float ks, float* w) {
  int lo_o = (int)floorf((s - ks + 0.5f) / inv_scale - 0.5f) - 1;
  int hi_o = (int)ceilf((s + ks + 0.5f) / inv_scale - 0.5f) + 1;
  lo_o = max(lo_o, 0);
  hi_o = min(hi_o, T - 1);
  if (hi_o - lo_o + 1 > kMaxCand) hi_o = lo_o + kMaxCand - 1;
#pragma unroll
  for (int k = 0; k < kMaxCand; ++k) {
    float wv = 0.f;
    const int o = lo_o + k;
    if (o <= hi_o) {
      float t[kMaxTaps];
      const int lo = resize_taps(o, S, inv_scale, ks, t);
      const int d = s - lo;
#pragma unroll
      for (int j = 0; j < kMaxTaps; ++j)
        if (j == d) wv = t[j];
    }
    w[k] = wv;
  }
  return lo_o;
}

__global__ void resize_bilinear_bwd_kernel(const float* __restrict__ dout, int N, int S, int T,
                                           float* __restrict__ dimg) {
  const long long total = (long long)N * S * S;
  const float inv_scale = (float)S / (float)T;
  const float ks = fmaxf(inv_scale, 1.f);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int sx = idx % S, sy = (idx / S) % S;
    const long long n = idx / ((long long)S * S);
    float wx[kMaxCand], wy[kMaxCand];
    const int ox0 = resize_candidates(sx, S, T, inv_scale, ks, wx);
    const int oy0 = resize_candidates(sy, S, T, inv_scale, ks, wy);
    const float* b = dout + n * T * T * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int ty = 0; ty < kMaxCand; ++ty) {
      if (wy[ty] == 0.f) continue;
#pragma unroll
      for (int tx = 0; tx < kMaxCand; ++tx) {
        if (wx[tx] == 0.f) continue;
        const float wgt = wy[ty] * wx[tx];
        const float* px = b + ((long long)(oy0 + ty) * T + ox0 + tx) * 3;
        a0 += wgt * px[0]; a1 += wgt * px[1]; a2 += wgt * px[2];
      }
    }
    float* o = dimg + idx * 3;
    o[0] += a0; o[1] += a1; o[2] += a2;
  }
}

// Input gradient of the 7x7 stride-2 stem (SAME padding low = pad_lo), second half. The first half is a tensor-core
// GEMM (xmc_conv2d_fwd as a 1x1 convolution): cols[n, ho, wo, (kh*7+kw)*3 + c] = sum_co dy[n, ho, wo, co] * W[kh][kw][c][co]
// (147 useful columns, pitch ldc). This kernel is the col2im gather: every input pixel (y, x) adds the <= 4 x 4 taps
// (kh, kw) whose output position ((y + pad_lo - kh) / 2, (x + pad_lo - kw) / 2) is integral and inside the map. Each
// cols element is read exactly once; fixed summation order (deterministic). Round 1 ran the whole transposed
// convolution on CUDA cores (1.3 ms bf16 / 1.95 ms fp32 per step); GEMM + gather: see DESIGN.md.
__global__ void stem_col2im_kernel(const float* __restrict__ cols, int N, int T, int Ho, int ldc, int pad_lo,
                                   float* __restrict__ dimg) {
  const long long total = (long long)N * T * T;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = idx % T, y = (idx / T) % T;
    const long long n = idx / ((long long)T * T);
    const int yp = y + pad_lo, xp = x + pad_lo;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int kh = yp & 1; kh < 7; kh += 2) {
      const int ho = (yp - kh) >> 1;
      if (ho < 0 || ho >= Ho) continue;
      for (int kw = xp & 1; kw < 7; kw += 2) {
        const int wo = (xp - kw) >> 1;
        if (wo < 0 || wo >= Ho) continue;
        const float* c = cols + ((n * Ho + ho) * Ho + wo) * ldc + (kh * 7 + kw) * 3;
        a0 += c[0]; a1 += c[1]; a2 += c[2];
      }
    }
    float* o = dimg + idx * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
  }
}

// flax nn.max_pool(x, (3,3), strides=(2,2), padding="SAME") on an even-sized map: window rows 2ho..2ho+2 (pad high)
template <typename T>
__global__ void maxpool3s2_kernel(const T* __restrict__ x, int N, int H, int C, T* __restrict__ y) {
  const int Ho = H / 2, cv = C >> 3;
  const long long total = (long long)N * Ho * Ho * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = idx % cv;
    long long pix = idx / cv;
    const int wo = pix % Ho, ho = (pix / Ho) % Ho;
    const long long n = pix / ((long long)Ho * Ho);
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -3.0e38f;
    for (int kh = 0; kh < 3; ++kh) {
      const int h = 2 * ho + kh;
      if (h >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int w = 2 * wo + kw;
        if (w >= H) continue;
        float f[8];
        load8(x + ((n * H + h) * H + w) * C + v * 8, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
      }
    }
    store8(y + pix * C + v * 8, m);
  }
}

// transpose: the gradient of a window goes to its FIRST maximal element in row-major window order
// (XLA select_and_scatter with a >= selector)
template <typename T>
__global__ void maxpool3s2_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                      const T* __restrict__ y, int N, int H, int C, T* __restrict__ dx) {
  const int Ho = H / 2, cv = C >> 3;
  const long long total = (long long)N * H * H * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = idx % cv;
    long long pix = idx / cv;
    const int w = pix % H, h = (pix / H) % H;
    const long long n = pix / ((long long)H * H);
    const int c = v * 8;
    float xv[8], acc[8];
    load8(x + pix * C + c, xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int ho = max(0, (h - 1) >> 1); ho <= min(Ho - 1, h >> 1); ++ho) {
      for (int wo = max(0, (w - 1) >> 1); wo <= min(Ho - 1, w >> 1); ++wo) {
        if (2 * ho > h || 2 * ho + 2 < h || 2 * wo > w || 2 * wo + 2 < w) continue;
        float yv[8], g[8];
        load8(y + ((n * Ho + ho) * Ho + wo) * C + c, yv);
        load8(dy + ((n * Ho + ho) * Ho + wo) * C + c, g);
        bool first[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) first[i] = (xv[i] == yv[i]);
        // an earlier window element with the same (maximal) value takes the gradient instead
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const int hh = 2 * ho + kh, ww = 2 * wo + kw;
            if (hh >= H || ww >= H) continue;
            if (hh > h || (hh == h && ww >= w)) continue;
            float e[8];
            load8(x + ((n * H + hh) * H + ww) * C + c, e);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (e[i] == yv[i]) first[i] = false;
          }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (first[i]) acc[i] += g[i];
      }
    }
    store8(dx + pix * C + c, acc);
  }
}

// z[n,2h,2w,:] = dy[n,h,w,:], zero elsewhere (transpose of stride-2 sampling)
template <typename T>
__global__ void zero_insert2_kernel(const T* __restrict__ dy, int N, int H, int W, int C, T* __restrict__ z) {
  const int cv = C >> 3;
  const long long total = (long long)N * H * W * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = idx % cv;
    long long pix = idx / cv;
    const int w = pix % W, h = (pix / W) % H;
    const long long n = pix / ((long long)W * H);
    float val[8];
    load8(dy + pix * C + v * 8, val);
    const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int W2 = 2 * W;
    const long long base = ((n * 2 * H + 2 * h) * W2 + 2 * w) * C + v * 8;
    store8(z + base, val);   // bf16 -> fp32 -> bf16 is exact
    store8(z + base + C, zero);
    store8(z + base + (long long)W2 * C, zero);
    store8(z + base + (long long)W2 * C + C, zero);
  }
}

static int grid1(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_resize_bilinear_pad(const float* img, int N, int S, int T, int Tp, int pad_lo, int split, void* out,
                                       void* stream) {
  if (!img || !out || N < 1 || S < 1 || T < 1 || 2 * S > 3 * T || Tp < T + pad_lo) return XMC_EINVAL;
  resize_bilinear_pad_kernel<<<grid1((long long)N * Tp * Tp, 256), 256, 0, (cudaStream_t)stream>>>(
      img, N, S, T, Tp, pad_lo, split, (bf16*)out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_resize_bilinear_bwd(const float* dout, int N, int S, int T, float* dimg, void* stream) {
  // 2S <= 3T bounds the taps per output (kMaxTaps), T <= 3S the outputs per source pixel (kMaxCand)
  if (!dout || !dimg || N < 1 || S < 1 || T < 1 || 2 * S > 3 * T || T > 3 * S) return XMC_EINVAL;
  resize_bilinear_bwd_kernel<<<grid1((long long)N * S * S, 256), 256, 0, (cudaStream_t)stream>>>(dout, N, S, T, dimg);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_stem_col2im(const float* cols, int N, int T, int Ho, int ldc, int pad_lo, float* dimg, void* stream) {
  if (!cols || !dimg || N < 1 || T < 1 || Ho < 1 || ldc < 147) return XMC_EINVAL;
  stem_col2im_kernel<<<grid1((long long)N * T * T, 256), 256, 0, (cudaStream_t)stream>>>(cols, N, T, Ho, ldc, pad_lo, dimg);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_maxpool3s2(const void* x, int act_f32, int N, int H, int C, void* y, void* stream) {
  if (!x || !y || N < 1 || H < 2 || (H % 2) || C < 8 || (C % 8)) return XMC_EINVAL;
  XMC_ACT(act_f32, maxpool3s2_kernel<T><<<grid1((long long)N * (H / 2) * (H / 2) * (C / 8), 256), 256, 0,
                                       (cudaStream_t)stream>>>((const T*)x, N, H, C, (T*)y));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_maxpool3s2_bwd(const void* dy, const void* x, const void* y, int act_f32, int N, int H, int C,
                                  void* dx, void* stream) {
  if (!dy || !x || !y || !dx || N < 1 || H < 2 || (H % 2) || C < 8 || (C % 8)) return XMC_EINVAL;
  XMC_ACT(act_f32, maxpool3s2_bwd_kernel<T><<<grid1((long long)N * H * H * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
                       (const T*)dy, (const T*)x, (const T*)y, N, H, C, (T*)dx));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_zero_insert2(const void* dy, int act_f32, int N, int H, int W, int C, void* z, void* stream) {
  if (!dy || !z || N < 1 || H < 1 || W < 1 || C < 8 || (C % 8)) return XMC_EINVAL;
  XMC_ACT(act_f32, zero_insert2_kernel<T><<<grid1((long long)N * H * W * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
                       (const T*)dy, N, H, W, C, (T*)z));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}
