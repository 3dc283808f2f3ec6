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

// Input gradient of the 7x7 stride-2 stem (SAME padding low=2). One thread computes 4 input pixels of the same column
// parity in one row (x, x+2, x+4, x+6): they use the same filter taps, so every weight vector read from shared memory
// feeds 4 pixels. wk: bf16 [Cout][7*56] with k = kh*56 + kw*8 + c (the packed forward weights), dy: [N,Ho,Ho,Cout].
template <typename TA>
__global__ void __launch_bounds__(128)
stem_dgrad_kernel(const TA* __restrict__ dy, const bf16* __restrict__ wk, const bf16* __restrict__ wk_lo, int N, int T,
                  int Ho, int Cout, int pad_lo, float* __restrict__ dimg) {
  extern __shared__ float ws[];  // [7][7][3][Cout]
  for (int t = threadIdx.x; t < 49 * 3 * Cout; t += blockDim.x) {
    const int co = t % Cout, c = (t / Cout) % 3, kw = (t / (Cout * 3)) % 7, kh = t / (Cout * 21);
    const long long wi = (long long)co * 392 + kh * 56 + kw * 8 + c;
    ws[t] = __bfloat162float(wk[wi]) + (wk_lo ? __bfloat162float(wk_lo[wi]) : 0.f);  // fp32 mode: hi + lo
  }
  __syncthreads();
  const int groups_x = (T + 7) / 8;             // 8 consecutive pixels = 2 parities x 4 pixels
  const long long total = (long long)N * T * groups_x * 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int par = idx & 1;
  const int gx = (idx >> 1) % groups_x;
  const int y = (idx / (2 * groups_x)) % T;
  const long long n = idx / ((long long)2 * groups_x * T);
  const int x0 = gx * 8 + par;                  // pixels x0, x0+2, x0+4, x0+6
  const int yp = y + pad_lo;
  const int xp0 = x0 + pad_lo;
  float acc[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
  for (int ho = max(0, (yp - 5) >> 1); ho <= min(Ho - 1, yp >> 1); ++ho) {
    const int kh = yp - 2 * ho;
    if (kh < 0 || kh > 6) continue;
    // taps for this column parity: kw = (xp0 & 1) + 2j, output column of pixel i: wo = (xp0 >> 1) + i - j
    for (int j = 0; j < 4; ++j) {
      const int kw = (xp0 & 1) + 2 * j;
      if (kw > 6) continue;
      const float* w = ws + ((kh * 7 + kw) * 3) * Cout;
      const TA* grow = dy + (n * Ho + ho) * (long long)Ho * Cout;
      for (int co = 0; co < Cout; co += 8) {
        const float4 wa0 = *reinterpret_cast<const float4*>(w + co), wb0 = *reinterpret_cast<const float4*>(w + co + 4);
        const float4 wa1 = *reinterpret_cast<const float4*>(w + Cout + co);
        const float4 wb1 = *reinterpret_cast<const float4*>(w + Cout + co + 4);
        const float4 wa2 = *reinterpret_cast<const float4*>(w + 2 * Cout + co);
        const float4 wb2 = *reinterpret_cast<const float4*>(w + 2 * Cout + co + 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int wo = (xp0 >> 1) + i - j;
          if (wo < 0 || wo >= Ho) continue;
          float f[8];
          load8(grow + (long long)wo * Cout + co, f);
          acc[i][0] += f[0] * wa0.x + f[1] * wa0.y + f[2] * wa0.z + f[3] * wa0.w + f[4] * wb0.x + f[5] * wb0.y +
                       f[6] * wb0.z + f[7] * wb0.w;
          acc[i][1] += f[0] * wa1.x + f[1] * wa1.y + f[2] * wa1.z + f[3] * wa1.w + f[4] * wb1.x + f[5] * wb1.y +
                       f[6] * wb1.z + f[7] * wb1.w;
          acc[i][2] += f[0] * wa2.x + f[1] * wa2.y + f[2] * wa2.z + f[3] * wa2.w + f[4] * wb2.x + f[5] * wb2.y +
                       f[6] * wb2.z + f[7] * wb2.w;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x0 + 2 * i;
    if (x < T) {
      float* o = dimg + ((n * T + y) * T + x) * 3;
      o[0] = acc[i][0]; o[1] = acc[i][1]; o[2] = acc[i][2];
    }
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

extern "C" int xmc_stem_dgrad(const void* dy, int act_f32, const void* wk, const void* wk_lo, int N, int T, int Ho,
                              int Cout, int pad_lo, float* dimg, void* stream) {
  if (!dy || !wk || !dimg || N < 1 || Cout < 8 || (Cout % 8)) return XMC_EINVAL;
  const size_t smem = (size_t)49 * 3 * Cout * sizeof(float);
  if (smem > 48 * 1024) return XMC_EINVAL;
  const long long total = (long long)N * T * ((T + 7) / 8) * 2;
  const int size = T;  // the dispatch macro names the activation type T
  XMC_ACT(act_f32, stem_dgrad_kernel<T><<<(unsigned)ceil_div_ll(total, 128), 128, smem, (cudaStream_t)stream>>>(
                       (const T*)dy, (const bf16*)wk, (const bf16*)wk_lo, N, size, Ho, Cout, pad_lo, dimg));
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
