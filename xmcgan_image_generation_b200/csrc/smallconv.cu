// Convolutions with 3 image channels on one side (first discriminator block, generator output head) and their
// gradients. K=27 or N=3 is hostile to tensor cores and these layers are HBM-bound (a [.,128,128,96] bf16 tensor on
// the other side), so they run as direct CUDA-core kernels with weights staged in shared memory.
#include "common.h"
#include "devutil.cuh"

namespace xmc {

constexpr int kImgC = 3;

// two adjacent channels as floats
__device__ __forceinline__ void load2(const bf16* p, float& a, float& b) {
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(p);
  a = __low2float(v); b = __high2float(v);
}
__device__ __forceinline__ void load2(const float* p, float& a, float& b) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  a = v.x; b = v.y;
}

// Weight element k of row `row` of a K-major bf16 weight matrix. wsplit = 0: plain [row][k]. wsplit = 1 (fp32 mode):
// the matrix is stored in the split form the tensor-core GEMMs consume, [row][tap][hi | hi | lo][inner], and the
// weight is hi + lo (16 mantissa bits), k = tap*inner + i.
__device__ __forceinline__ float weight_at(const bf16* w, int ldw, int row, int k, int inner, int wsplit) {
  if (!wsplit) return __bfloat162float(w[(long long)row * ldw + k]);
  const int tap = k / inner, i = k - tap * inner;
  const bf16* q = w + (long long)row * ldw + (long long)tap * 3 * inner + i;
  return __bfloat162float(q[0]) + __bfloat162float(q[2 * inner]);
}

// y[p][co] = relu?( sum_{tap,c3} x[p+d(tap)][c3] * w[co][tap*3+c3] + bias[co] ),  x: bf16 [N,H,W,3], y: bf16 [.,Cout]
// Each thread computes 4 horizontally adjacent pixels x 8 channels at a time: the 3x6x3 input window lives in
// registers, every weight read from shared memory (one LDS.128 broadcast per 4 weights) feeds 4 FMAs.
constexpr int kPx = 4;
constexpr int kPxIn = 4;   // conv_c3_in: pixels per thread. ncu (r02_c3in): issue slots 55 % busy, 40 % of the stalls wait on the
                           // weight LDS.128 stream; 2 (more warps) and 8 (fewer LDS per FFMA, 3 blocks per SM) were no faster
template <int KS, typename T>
__global__ void __launch_bounds__(128)
conv_c3_in_kernel(const T* __restrict__ x, const bf16* __restrict__ w, int ldw, int wsplit,
                  const float* __restrict__ bias, int N, int H, int W, int Cout, int relu, T* __restrict__ y) {
  extern __shared__ float ws[];  // [K][Cout] (transposed for vector reads) + bias[Cout]
  constexpr int KH = KS, KW = KS;
  constexpr int K = KH * KW * kImgC;
  for (int t = threadIdx.x; t < Cout * K; t += blockDim.x) {
    const int co = t / K, k = t - co * K;
    ws[k * Cout + co] = weight_at(w, ldw, co, k, kImgC, wsplit);
  }
  float* bs = ws + Cout * K;
  for (int t = threadIdx.x; t < Cout; t += blockDim.x) bs[t] = bias ? bias[t] : 0.f;
  __syncthreads();
  // persistent blocks (a few per SM): the weight staging above is paid once per block, not once per 512 pixels
  const int wq4 = (W + kPxIn - 1) / kPxIn;
  const long long groups = (long long)N * H * wq4;
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups;
       gidx += (long long)gridDim.x * blockDim.x) {
  const int wg = gidx % wq4;
  const int hq = (gidx / wq4) % H;
  const long long n = gidx / ((long long)wq4 * H);
  const int w0 = wg * kPxIn;
  constexpr int ph = KH / 2, pw = KW / 2;
  constexpr int cols = kPxIn + KW - 1;
  float win[KH * cols * kImgC];  // [kh][col][c]
#pragma unroll
  for (int kh = 0; kh < KH; ++kh) {
    const int hh = hq + kh - ph;
#pragma unroll
    for (int cc = 0; cc < cols; ++cc) {
      const int ww = w0 + cc - pw;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      const T* src = x + ((n * H + hh) * W + ww) * kImgC;
#pragma unroll
      for (int c = 0; c < kImgC; ++c) win[(kh * cols + cc) * kImgC + c] = ok ? to_f(src[c]) : 0.f;
    }
  }
  const long long pbase = (n * H + hq) * W + w0;
  for (int c0 = 0; c0 < Cout; c0 += 8) {
    float acc[kPxIn][8];
#pragma unroll
    for (int px = 0; px < kPxIn; ++px)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[px][i] = bs[c0 + i];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh)
#pragma unroll
      for (int kw = 0; kw < KW; ++kw)
#pragma unroll
        for (int c = 0; c < kImgC; ++c) {
          const float* wp = ws + ((kh * KW + kw) * kImgC + c) * Cout + c0;
          const float4 wa = *reinterpret_cast<const float4*>(wp);
          const float4 wb = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
          for (int px = 0; px < kPxIn; ++px) {
            const float v = win[(kh * cols + px + kw) * kImgC + c];
            acc[px][0] += v * wa.x; acc[px][1] += v * wa.y; acc[px][2] += v * wa.z; acc[px][3] += v * wa.w;
            acc[px][4] += v * wb.x; acc[px][5] += v * wb.y; acc[px][6] += v * wb.z; acc[px][7] += v * wb.w;
          }
        }
#pragma unroll
    for (int px = 0; px < kPxIn; ++px) {
      if (w0 + px < W) {
        if (relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[px][i] = fmaxf(acc[px][i], 0.f);
        }
        store8(y + (pbase + px) * Cout + c0, acc[px]);
      }
    }
  }
  }
}

// y[p][c3] = sum_{tap,ci} x[p+d(tap)][ci] * w[c3][tap*Cin+ci] + bias[c3];  x: bf16 [N,H,W,Cin] (Cin % 8 == 0)
// mode 0: fp32 out (accumulate optional); mode 1: img = (tanh(v)+1)/2 -> fp32 out + bf16 copy (xmc_net.py:245-247)
// Each thread computes 4 horizontally adjacent pixels: per (kh, 8-channel vector) it loads the 6 input vectors once
// and reads each weight vector from shared memory once for all 4 pixels.
template <int KS, typename T>
__global__ void __launch_bounds__(128)
conv_c3_out_kernel(const T* __restrict__ x, const bf16* __restrict__ w, int ldw, int wsplit,
                   const float* __restrict__ bias, int N, int H, int W, int Cin, int mode, int accumulate,
                   float* __restrict__ y, T* __restrict__ y_bf16) {
  extern __shared__ float ws[];  // [3][K]
  constexpr int KH = KS, KW = KS;
  const int K = KH * KW * Cin;
  for (int t = threadIdx.x; t < kImgC * K; t += blockDim.x) ws[t] = weight_at(w, ldw, t / K, t % K, Cin, wsplit);
  __syncthreads();
  // persistent blocks: the weight staging above is paid once per block
  const int wq4 = (W + kPx - 1) / kPx;
  const long long groups = (long long)N * H * wq4;
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups;
       gidx += (long long)gridDim.x * blockDim.x) {
  const int wg = gidx % wq4;
  const int hq = (gidx / wq4) % H;
  const long long n = gidx / ((long long)wq4 * H);
  const int w0 = wg * kPx;
  constexpr int ph = KH / 2, pw = KW / 2;
  float acc[kPx][kImgC];
#pragma unroll
  for (int px = 0; px < kPx; ++px)
#pragma unroll
    for (int c = 0; c < kImgC; ++c) acc[px][c] = 0.f;
  for (int kh = 0; kh < KH; ++kh) {
    const int hh = hq + kh - ph;
    if (hh < 0 || hh >= H) continue;
    const T* row = x + (n * H + hh) * (long long)W * Cin;
    for (int c = 0; c < Cin; c += 8) {
      float in[kPx + KW - 1][8];
#pragma unroll
      for (int cc = 0; cc < kPx + KW - 1; ++cc) {
        const int ww = w0 + cc - pw;
        if (ww >= 0 && ww < W) {
          load8(row + (long long)ww * Cin + c, in[cc]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) in[cc][i] = 0.f;
        }
      }
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
#pragma unroll
        for (int c3 = 0; c3 < kImgC; ++c3) {
          const float* wp = ws + c3 * K + (kh * KW + kw) * Cin + c;
          const float4 wa = *reinterpret_cast<const float4*>(wp);
          const float4 wb = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
          for (int px = 0; px < kPx; ++px) {
            const float* v = in[px + kw];
            acc[px][c3] += v[0] * wa.x + v[1] * wa.y + v[2] * wa.z + v[3] * wa.w + v[4] * wb.x + v[5] * wb.y +
                           v[6] * wb.z + v[7] * wb.w;
          }
        }
      }
    }
  }
  const long long pbase = (n * H + hq) * W + w0;
#pragma unroll
  for (int px = 0; px < kPx; ++px) {
    if (w0 + px >= W) continue;
    float a0 = acc[px][0], a1 = acc[px][1], a2 = acc[px][2];
    if (bias) { a0 += bias[0]; a1 += bias[1]; a2 += bias[2]; }
    float* yo = y + (pbase + px) * kImgC;
    if (mode == 1) {
      a0 = (tanhf(a0) + 1.f) * 0.5f; a1 = (tanhf(a1) + 1.f) * 0.5f; a2 = (tanhf(a2) + 1.f) * 0.5f;
      yo[0] = a0; yo[1] = a1; yo[2] = a2;
      if (y_bf16) {
        T* yb = y_bf16 + (pbase + px) * kImgC;
        yb[0] = from_f<T>(a0); yb[1] = from_f<T>(a1); yb[2] = from_f<T>(a2);
      }
    } else if (accumulate) {
      yo[0] += a0; yo[1] += a1; yo[2] += a2;
    } else {
      yo[0] = a0; yo[1] = a1; yo[2] = a2;
    }
  }
  }
}

// out[tap_o][..] += sum_p x3[p + d(tap)][c3] * y[p][c];  x3: bf16 [N,H,W,3], y: bf16 [N,H,W,C]
// out index = tap_o*s_tap + c3*s_c3 + c*s_c with tap_o = flip ? taps-1-tap : tap. One block = 8 image rows x 64 px;
// one thread = 2 adjacent channels; the 3x3x3 input window slides along the row in registers (9 shared-memory reads
// per pixel feed 54 FMAs).
template <int KS, typename T>
__global__ void wgrad_c3_kernel(const T* __restrict__ x3, const T* __restrict__ y, int N, int H, int W, int C,
                                float* __restrict__ partials) {
  extern __shared__ float xs[];  // [(rows+2ph)][64+2pw][3]
  constexpr int KH = KS, KW = KS;
  constexpr int ph = KH / 2, pw = KW / 2;
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
    if (gh >= 0 && gh < H && gw >= 0 && gw < W) v = to_f(x3[(((long long)n * H + gh) * W + gw) * kImgC + c]);
    xs[t] = v;
  }
  __syncthreads();
  const int c = threadIdx.x * 2;
  if (c >= C) return;
  constexpr int taps = KH * KW;
  float acc[2][taps * 3];
#pragma unroll
  for (int i = 0; i < taps * 3; ++i) acc[0][i] = acc[1][i] = 0.f;
  for (int r = 0; r < rows; ++r) {
    const int gh = h0 + r;
    if (gh >= H) break;
    const T* yrow = y + (((long long)n * H + gh) * W + w0) * C + c;
    const int qn = min(seg, W - w0);
    if constexpr (taps == 9) {
      // win[slot][kh][c3]: slot (col % 3) holds input column `col` of the halo tile
      float win[3][3][3];
#pragma unroll
      for (int col = 0; col < 2; ++col)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int c3 = 0; c3 < 3; ++c3) win[col][kh][c3] = xs[((r + kh) * xw + col) * kImgC + c3];
      // 12 pixels per outer iteration: all 12 y loads are issued first (the loop is latency-bound otherwise)
      for (int q0 = 0; q0 < qn; q0 += 12) {
        float yv0[12], yv1[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          yv0[j] = yv1[j] = 0.f;
          if (q0 + j < qn) load2(yrow + (long long)(q0 + j) * C, yv0[j], yv1[j]);
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const int q = q0 + j;
          if (q < qn) {
            // bring in column q+2 into slot (j+2)%3
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int c3 = 0; c3 < 3; ++c3) win[(j + 2) % 3][kh][c3] = xs[((r + kh) * xw + q + 2) * kImgC + c3];
            const float y0 = yv0[j], y1 = yv1[j];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int c3 = 0; c3 < 3; ++c3) {
                  const float xv = win[(j + kw) % 3][kh][c3];
                  acc[0][(kh * 3 + kw) * 3 + c3] += xv * y0;
                  acc[1][(kh * 3 + kw) * 3 + c3] += xv * y1;
                }
          }
        }
      }
    } else {
      for (int q = 0; q < qn; ++q) {
        float y0, y1;
        load2(yrow + (long long)q * C, y0, y1);
        const float* xp = xs + (r * xw + q) * kImgC;
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          acc[0][c3] += xp[c3] * y0;
          acc[1][c3] += xp[c3] * y1;
        }
      }
    }
  }
  // block partial [tap*3 + c3][C] (plain stores); wgrad_c3_finish_kernel adds the blocks in index order
  float* part = partials + (long long)blockIdx.x * taps * 3 * C;
#pragma unroll
  for (int i = 0; i < taps * 3; ++i) {
    part[(long long)i * C + c] = acc[0][i];
    if (c + 1 < C) part[(long long)i * C + c + 1] = acc[1][i];
  }
}

// out[tap_o*s_tap + c3*s_c3 + c*s_c] += sum_blocks partials[block][tap*3 + c3][c] in a fixed order: 32 outputs x 8 block
// lanes per thread block, lane l adds blocks l, l+8, ... in index order, the lane sums are added in lane order.
__global__ void wgrad_c3_finish_kernel(const float* __restrict__ partials, int blocks, int taps, int C, int flip,
                                       long long s_tap, int s_c3, int s_c, float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int idx = blockIdx.x * 32 + threadIdx.x;
  const int width = taps * 3 * C;
  float a = 0.f;
  if (idx < width)
    for (int b = threadIdx.y; b < blocks; b += 8) a += partials[(long long)b * width + idx];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y != 0 || idx >= width) return;
  float t = 0.f;
#pragma unroll
  for (int l = 0; l < 8; ++l) t += sm[l][threadIdx.x];
  const int c = idx % C, i = idx / C, c3 = i % 3, tap = i / 3;
  const int tap_o = flip ? taps - 1 - tap : tap;
  out[tap_o * s_tap + c3 * s_c3 + (long long)c * s_c] += t;
}

// scalar-channel 2x2 mean pool of a bf16 [N,2H,2W,C] tensor (the 3-channel image, common.py:131)
template <typename T>
__global__ void pool2_small_kernel(const T* __restrict__ a, int N, int H, int W, int C, float scale,
                                   T* __restrict__ out) {
  const long long total = (long long)N * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % C;
    long long pix = idx / C;
    const int w = pix % W, h = (pix / W) % H;
    const long long n = pix / ((long long)W * H);
    const long long base = ((n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
    const float s = to_f(a[base]) + to_f(a[base + C]) + to_f(a[base + 2LL * W * C]) + to_f(a[base + 2LL * W * C + C]);
    out[idx] = from_f<T>(s * scale);
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
template <typename T>
__global__ void tanh01_bwd_kernel(const float* __restrict__ dimg, const float* __restrict__ img, long long n,
                                  T* __restrict__ dpre) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float t = 2.f * img[i] - 1.f;
    dpre[i] = from_f<T>(dimg[i] * 0.5f * (1.f - t * t));
  }
}

// ---- packed-window helpers: the 3-channel 3x3 convolutions on the tensor-core GEMM kernels ------------------------
// out[n][h+1][w+1][0..2] = x[n][h][w][0..2]; everything else (1-pixel border, channels 3..7) zero. One thread per
// output pixel writes one 16-byte vector.
__global__ void pad_c3_to_c8_kernel(const bf16* __restrict__ x, int N, int H, int W, bf16* __restrict__ out) {
  const int HP = H + 2, WP = W + 2;
  const long long total = (long long)N * HP * WP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int wp = idx % WP, hp = (idx / WP) % HP;
    const long long n = idx / ((long long)WP * HP);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (hp >= 1 && hp <= H && wp >= 1 && wp <= W) {
      const unsigned short* src =
          reinterpret_cast<const unsigned short*>(x) + ((n * H + hp - 1) * W + wp - 1) * kImgC;
      v.x = (uint32_t)src[0] | ((uint32_t)src[1] << 16);
      v.y = (uint32_t)src[2];
    }
    reinterpret_cast<uint4*>(out)[idx] = v;
  }
}

__global__ void pack_c3_weights_kernel(const bf16* __restrict__ w, int ldw, int Cout, bf16* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Cout * 72) return;
  const int co = t / 72, k = t - co * 72;
  const int kh = k / 24, r = k - kh * 24, kw = r >> 3, c = r & 7;
  out[t] = c < kImgC ? w[co * ldw + (kh * 3 + kw) * kImgC + c] : __float2bfloat16(0.f);
}

__global__ void unpack_c3_wgrad_kernel(const float* __restrict__ tmp, int C, int flip, long long s_tap, int s_c3,
                                       int s_c, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 27 * C) return;
  const int c = t % C, k = t / C;          // k = (kh*3+kw)*3 + c3
  const int c3 = k % 3, tap = k / 3, kh = tap / 3, kw = tap - kh * 3;
  const int tap_o = flip ? 8 - tap : tap;
  out[tap_o * s_tap + c3 * s_c3 + (long long)c * s_c] += tmp[(kh * 24 + kw * 8 + c3) * C + c];
}

static int grid1d(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_pad_c3_to_c8(const void* x, int N, int H, int W, void* out, void* stream) {
  if (!x || !out || N < 1 || H < 1 || W < 1 || !aligned16(out)) return XMC_EINVAL;
  pad_c3_to_c8_kernel<<<grid1d((long long)N * (H + 2) * (W + 2), 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, N, H, W, (bf16*)out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_pack_c3_weights(const void* w, int ldw, int Cout, void* out, void* stream) {
  if (!w || !out || Cout < 1 || ldw < 27) return XMC_EINVAL;
  pack_c3_weights_kernel<<<ceil_div(Cout * 72, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)w, ldw, Cout,
                                                                                    (bf16*)out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_unpack_c3_wgrad(const float* tmp, int C, int flip, long long s_tap, int s_c3, int s_c, float* out,
                                   void* stream) {
  if (!tmp || !out || C < 1) return XMC_EINVAL;
  unpack_c3_wgrad_kernel<<<ceil_div(27 * C, 256), 256, 0, (cudaStream_t)stream>>>(tmp, C, flip, s_tap, s_c3, s_c, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_conv_c3_in(const void* x, int act_f32, const void* w, int ldw, const float* bias, int N, int H, int W,
                              int Cout, int KH, int KW, int relu, void* y, void* stream) {
  if (!x || !w || !y || Cout < 8 || (Cout % 8) || KH * KW > 9) return XMC_EINVAL;
  const int K = KH * KW * kImgC;
  const size_t smem = (size_t)(Cout * K + Cout) * sizeof(float);
  if (smem > 48 * 1024) return XMC_EINVAL;
  const long long P = (long long)N * H * ceil_div(W, kPxIn);
  if (KH != KW || (KH != 1 && KH != 3)) return XMC_EINVAL;
  // fp32 activations come with split-form weights ([row][tap][hi|hi|lo][3], see weight_at)
  long long blocks = ceil_div_ll(P, 128);
  if (blocks > (long long)num_sms() * 4) blocks = (long long)num_sms() * 4;   // 126 registers x 128 threads: 4 per SM
  if (KH == 3)
    XMC_ACT(act_f32, conv_c3_in_kernel<3, T><<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(
                         (const T*)x, (const bf16*)w, ldw, act_f32, bias, N, H, W, Cout, relu, (T*)y));
  else
    XMC_ACT(act_f32, conv_c3_in_kernel<1, T><<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(
                         (const T*)x, (const bf16*)w, ldw, act_f32, bias, N, H, W, Cout, relu, (T*)y));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_conv_c3_out(const void* x, int act_f32, const void* w, int ldw, const float* bias, int N, int H, int W,
                               int Cin, int KH, int KW, int mode, int accumulate, float* y, void* y_bf16,
                               void* stream) {
  if (!x || !w || !y || Cin < 8 || (Cin % 8)) return XMC_EINVAL;
  const size_t smem = (size_t)kImgC * KH * KW * Cin * sizeof(float);
  if (smem > 48 * 1024) return XMC_EINVAL;
  const long long P = (long long)N * H * ceil_div(W, kPx);
  if (KH != KW || (KH != 1 && KH != 3)) return XMC_EINVAL;
  long long blocks = ceil_div_ll(P, 128);
  if (blocks > (long long)num_sms() * 4) blocks = (long long)num_sms() * 4;
  if (KH == 3)
    XMC_ACT(act_f32, conv_c3_out_kernel<3, T><<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(
                         (const T*)x, (const bf16*)w, ldw, act_f32, bias, N, H, W, Cin, mode, accumulate, y, (T*)y_bf16));
  else
    XMC_ACT(act_f32, conv_c3_out_kernel<1, T><<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(
                         (const T*)x, (const bf16*)w, ldw, act_f32, bias, N, H, W, Cin, mode, accumulate, y, (T*)y_bf16));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_wgrad_c3(const void* x3, const void* y, int act_f32, int N, int H, int W, int C, int KH, int KW,
                            int flip, long long s_tap, int s_c3, int s_c, float* out, float* partials, void* stream) {
  if (!x3 || !y || !out || !partials || C < 2 || (C % 2) || C > 2048 || KH * KW > 9) return XMC_EINVAL;
  const int ph = KH / 2, pw = KW / 2;
  const size_t smem = (size_t)(8 + 2 * ph) * (64 + 2 * pw) * kImgC * sizeof(float);
  const int blocks = N * ceil_div(H, 8) * ceil_div(W, 64);
  const int threads = ceil_div(C / 2, 32) * 32;
  if (KH != KW || (KH != 1 && KH != 3)) return XMC_EINVAL;
  if (KH == 3)
    XMC_ACT(act_f32, wgrad_c3_kernel<3, T><<<blocks, threads, smem, (cudaStream_t)stream>>>((const T*)x3, (const T*)y, N, H,
                                                                                       W, C, partials));
  else
    XMC_ACT(act_f32, wgrad_c3_kernel<1, T><<<blocks, threads, smem, (cudaStream_t)stream>>>((const T*)x3, (const T*)y, N, H,
                                                                                       W, C, partials));
  XMC_LAUNCH_CHECK();
  const int width = KH * KW * 3 * C;
  wgrad_c3_finish_kernel<<<ceil_div(width, 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(partials, blocks, KH * KW, C,
                                                                                       flip, s_tap, s_c3, s_c, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_pool2_small(const void* a, int act_f32, int N, int Hout, int Wout, int C, float scale, void* out,
                               void* stream) {
  if (!a || !out) return XMC_EINVAL;
  XMC_ACT(act_f32, pool2_small_kernel<T><<<grid1d((long long)N * Hout * Wout * C, 256), 256, 0, (cudaStream_t)stream>>>(
                       (const T*)a, N, Hout, Wout, C, scale, (T*)out));
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

extern "C" int xmc_tanh01_bwd(const float* dimg, const float* img, long long n, void* dpre, int act_f32, void* stream) {
  if (!dimg || !img || !dpre || n < 1) return XMC_EINVAL;
  XMC_ACT(act_f32, tanh01_bwd_kernel<T><<<grid1d(n, 256), 256, 0, (cudaStream_t)stream>>>(dimg, img, n, (T*)dpre));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}
