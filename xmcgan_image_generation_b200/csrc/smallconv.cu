// Convolutions with 3 image channels on one side (first discriminator block, generator output head) and their
// gradients. K=27 or N=3 is hostile to tensor cores and these layers are HBM-bound (a [.,128,128,96] bf16 tensor on
// the other side), so they run as direct CUDA-core kernels with weights staged in shared memory.
#include "common.h"
#include "devutil.cuh"

namespace xmc {

constexpr int kImgC = 3;

// y[p][co] = relu?( sum_{tap,c3} x[p+d(tap)][c3] * w[co][tap*3+c3] + bias[co] ),  x: bf16 [N,H,W,3], y: bf16 [.,Cout]
__global__ void conv_c3_in_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, int ldw,
                                  const float* __restrict__ bias, int N, int H, int W, int Cout, int KH, int KW,
                                  int relu, bf16* __restrict__ y) {
  extern __shared__ float ws[];  // [Cout][K] + bias[Cout]
  const int K = KH * KW * kImgC;
  for (int t = threadIdx.x; t < Cout * K; t += blockDim.x) ws[t] = __bfloat162float(w[(t / K) * ldw + (t % K)]);
  float* bs = ws + Cout * K;
  for (int t = threadIdx.x; t < Cout; t += blockDim.x) bs[t] = bias ? bias[t] : 0.f;
  __syncthreads();
  const long long P = (long long)N * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int wq = p % W, hq = (p / W) % H;
  const long long nbase = p - (long long)hq * W - wq;
  float in[27];
  const int ph = KH / 2, pw = KW / 2;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int hh = hq + kh - ph, ww = wq + kw - pw;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      const bf16* src = x + (nbase + (long long)hh * W + ww) * kImgC;
#pragma unroll
      for (int c = 0; c < kImgC; ++c) in[(kh * KW + kw) * kImgC + c] = ok ? __bfloat162float(src[c]) : 0.f;
    }
  bf16* yo = y + p * Cout;
  for (int c0 = 0; c0 < Cout; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bs[c0 + i];
    for (int k = 0; k < K; ++k) {
      const float v = in[k];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v * ws[(c0 + i) * K + k];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    store8(yo + c0, acc);
  }
}

// y[p][c3] = sum_{tap,ci} x[p+d(tap)][ci] * w[c3][tap*Cin+ci] + bias[c3];  x: bf16 [N,H,W,Cin] (Cin % 8 == 0)
// mode 0: fp32 out (accumulate optional); mode 1: img = (tanh(v)+1)/2 -> fp32 out + bf16 copy (xmc_net.py:245-247)
__global__ void conv_c3_out_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, int ldw,
                                   const float* __restrict__ bias, int N, int H, int W, int Cin, int KH, int KW,
                                   int mode, int accumulate, float* __restrict__ y, bf16* __restrict__ y_bf16) {
  extern __shared__ float ws[];  // [3][K]
  const int K = KH * KW * Cin;
  for (int t = threadIdx.x; t < kImgC * K; t += blockDim.x) ws[t] = __bfloat162float(w[(t / K) * ldw + (t % K)]);
  __syncthreads();
  const long long P = (long long)N * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int wq = p % W, hq = (p / W) % H;
  const long long nbase = p - (long long)hq * W - wq;
  const int ph = KH / 2, pw = KW / 2;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int hh = hq + kh - ph, ww = wq + kw - pw;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const bf16* src = x + (nbase + (long long)hh * W + ww) * Cin;
      const float* wk = ws + (kh * KW + kw) * Cin;
      for (int c = 0; c < Cin; c += 8) {
        float f[8];
        load8(src + c, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a0 += f[i] * wk[c + i];
          a1 += f[i] * wk[K + c + i];
          a2 += f[i] * wk[2 * K + c + i];
        }
      }
    }
  if (bias) { a0 += bias[0]; a1 += bias[1]; a2 += bias[2]; }
  float* yo = y + p * kImgC;
  if (mode == 1) {
    a0 = (tanhf(a0) + 1.f) * 0.5f; a1 = (tanhf(a1) + 1.f) * 0.5f; a2 = (tanhf(a2) + 1.f) * 0.5f;
    yo[0] = a0; yo[1] = a1; yo[2] = a2;
    if (y_bf16) {
      bf16* yb = y_bf16 + p * kImgC;
      yb[0] = __float2bfloat16(a0); yb[1] = __float2bfloat16(a1); yb[2] = __float2bfloat16(a2);
    }
  } else if (accumulate) {
    yo[0] += a0; yo[1] += a1; yo[2] += a2;
  } else {
    yo[0] = a0; yo[1] = a1; yo[2] = a2;
  }
}

// out[tap_o][..] += sum_p x3[p + d(tap)][c3] * y[p][c];  x3: bf16 [N,H,W,3], y: bf16 [N,H,W,C]
// out index = tap_o*s_tap + c3*s_c3 + c*s_c with tap_o = flip ? taps-1-tap : tap. One block = 8 image rows x 64 px.
__global__ void wgrad_c3_kernel(const bf16* __restrict__ x3, const bf16* __restrict__ y, int N, int H, int W, int C,
                                int KH, int KW, int flip, long long s_tap, int s_c3, int s_c,
                                float* __restrict__ out) {
  extern __shared__ float xs[];  // [(rows+2ph)][64+2pw][3]
  const int ph = KH / 2, pw = KW / 2;
  const int rows = 8, seg = 64;
  const int wsegs = (W + seg - 1) / seg;
  const int hblocks = (H + rows - 1) / rows;
  int b = blockIdx.x;
  const int ws_i = b % wsegs; b /= wsegs;
  const int hb = b % hblocks; b /= hblocks;
  const int n = b;
  const int h0 = hb * rows, w0 = ws_i * seg;
  const int xw = seg + 2 * pw, xh = rows + 2 * ph;
  for (int t = threadIdx.x; t < xh * xw * kImgC; t += blockDim.x) {
    const int c = t % kImgC, ww = (t / kImgC) % xw, hh = t / (kImgC * xw);
    const int gh = h0 + hh - ph, gw = w0 + ww - pw;
    float v = 0.f;
    if (gh >= 0 && gh < H && gw >= 0 && gw < W) v = __bfloat162float(x3[(((long long)n * H + gh) * W + gw) * kImgC + c]);
    xs[t] = v;
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= C) return;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  const int taps = KH * KW;
  for (int r = 0; r < rows; ++r) {
    const int gh = h0 + r;
    if (gh >= H) break;
    for (int q = 0; q < seg; ++q) {
      const int gw = w0 + q;
      if (gw >= W) break;
      const float yv = __bfloat162float(y[(((long long)n * H + gh) * W + gw) * C + c]);
      if (taps == 9) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float* xp = xs + ((r + kh) * xw + (q + kw)) * kImgC;
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) acc[(kh * 3 + kw) * 3 + c3] += xp[c3] * yv;
          }
      } else {
        const float* xp = xs + (r * xw + q) * kImgC;
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) acc[c3] += xp[c3] * yv;
      }
    }
  }
  for (int tap = 0; tap < taps; ++tap) {
    const int tap_o = flip ? taps - 1 - tap : tap;
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) atomicAdd(out + tap_o * s_tap + c3 * s_c3 + (long long)c * s_c, acc[tap * 3 + c3]);
  }
}

// scalar-channel 2x2 mean pool of a bf16 [N,2H,2W,C] tensor (the 3-channel image, common.py:131)
__global__ void pool2_small_kernel(const bf16* __restrict__ a, int N, int H, int W, int C, float scale,
                                   bf16* __restrict__ out) {
  const long long total = (long long)N * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % C;
    long long pix = idx / C;
    const int w = pix % W, h = (pix / W) % H;
    const long long n = pix / ((long long)W * H);
    const long long base = ((n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
    const float s = __bfloat162float(a[base]) + __bfloat162float(a[base + C]) +
                    __bfloat162float(a[base + 2LL * W * C]) + __bfloat162float(a[base + 2LL * W * C + C]);
    out[idx] = __float2bfloat16(s * scale);
  }
}

// g[n,2h+i,2w+j,c] += scale * d[n,h,w,c]   (fp32)
__global__ void unpool2_add_f32_kernel(const float* __restrict__ d, int N, int H, int W, int C, float scale,
                                       float* __restrict__ g) {
  const long long total = (long long)N * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % C;
    long long pix = idx / C;
    const int w = pix % W, h = (pix / W) % H;
    const long long n = pix / ((long long)W * H);
    const long long base = ((n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
    const float v = d[idx] * scale;
    g[base] += v; g[base + C] += v; g[base + 2LL * W * C] += v; g[base + 2LL * W * C + C] += v;
  }
}

// d(pre-tanh) = dimg * 0.5 * (1 - t^2), t = 2*img - 1   -> bf16
__global__ void tanh01_bwd_kernel(const float* __restrict__ dimg, const float* __restrict__ img, long long n,
                                  bf16* __restrict__ dpre) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float t = 2.f * img[i] - 1.f;
    dpre[i] = __float2bfloat16(dimg[i] * 0.5f * (1.f - t * t));
  }
}

static int grid1d(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_conv_c3_in(const void* x, const void* w, int ldw, const float* bias, int N, int H, int W, int Cout,
                              int KH, int KW, int relu, void* y, void* stream) {
  if (!x || !w || !y || Cout < 8 || (Cout % 8) || KH * KW > 9) return XMC_EINVAL;
  const int K = KH * KW * kImgC;
  const size_t smem = (size_t)(Cout * K + Cout) * sizeof(float);
  if (smem > 48 * 1024) return XMC_EINVAL;
  const long long P = (long long)N * H * W;
  conv_c3_in_kernel<<<(unsigned)ceil_div_ll(P, 128), 128, smem, (cudaStream_t)stream>>>(
      (const bf16*)x, (const bf16*)w, ldw, bias, N, H, W, Cout, KH, KW, relu, (bf16*)y);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_conv_c3_out(const void* x, const void* w, int ldw, const float* bias, int N, int H, int W, int Cin,
                               int KH, int KW, int mode, int accumulate, float* y, void* y_bf16, void* stream) {
  if (!x || !w || !y || Cin < 8 || (Cin % 8)) return XMC_EINVAL;
  const size_t smem = (size_t)kImgC * KH * KW * Cin * sizeof(float);
  if (smem > 48 * 1024) return XMC_EINVAL;
  const long long P = (long long)N * H * W;
  conv_c3_out_kernel<<<(unsigned)ceil_div_ll(P, 128), 128, smem, (cudaStream_t)stream>>>(
      (const bf16*)x, (const bf16*)w, ldw, bias, N, H, W, Cin, KH, KW, mode, accumulate, y, (bf16*)y_bf16);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_wgrad_c3(const void* x3, const void* y, int N, int H, int W, int C, int KH, int KW, int flip,
                            long long s_tap, int s_c3, int s_c, float* out, void* stream) {
  if (!x3 || !y || !out || C < 1 || C > 1024 || KH * KW > 9) return XMC_EINVAL;
  const int ph = KH / 2, pw = KW / 2;
  const size_t smem = (size_t)(8 + 2 * ph) * (64 + 2 * pw) * kImgC * sizeof(float);
  const int blocks = N * ceil_div(H, 8) * ceil_div(W, 64);
  const int threads = ceil_div(C, 32) * 32;
  wgrad_c3_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>((const bf16*)x3, (const bf16*)y, N, H, W, C, KH, KW,
                                                                  flip, s_tap, s_c3, s_c, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_pool2_small(const void* a, int N, int Hout, int Wout, int C, float scale, void* out, void* stream) {
  if (!a || !out) return XMC_EINVAL;
  pool2_small_kernel<<<grid1d((long long)N * Hout * Wout * C, 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)a, N, Hout, Wout, C, scale, (bf16*)out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_unpool2_add_f32(const float* d, int N, int Hin, int Win, int C, float scale, float* g,
                                   void* stream) {
  if (!d || !g) return XMC_EINVAL;
  unpool2_add_f32_kernel<<<grid1d((long long)N * Hin * Win * C, 256), 256, 0, (cudaStream_t)stream>>>(d, N, Hin, Win, C,
                                                                                                     scale, g);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_tanh01_bwd(const float* dimg, const float* img, long long n, void* dpre, void* stream) {
  if (!dimg || !img || !dpre || n < 1) return XMC_EINVAL;
  tanh01_bwd_kernel<<<grid1d(n, 256), 256, 0, (cudaStream_t)stream>>>(dimg, img, n, (bf16*)dpre);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}
