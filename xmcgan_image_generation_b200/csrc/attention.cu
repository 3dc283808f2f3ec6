// Row l2-normalisation, the generator's word-region attention (attention_lib.attention_for_g) and the elementwise /
// reduction stages of the discriminator's word-level matching loss (attention_lib.word_loss). The GEMM stages of
// word_loss run on the tcgen05 kernels in gemm.cu; nothing of shape [B,B,R,L,D] is ever materialised.
#include "common.h"
#include "devutil.cuh"

namespace xmc {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

// xhat = x * rsqrt(max(sum x^2, eps))   (attention_lib.l2_normalize, attention_lib.py:30-33). One warp per row.
template <typename TI, typename TO>
__global__ void l2norm_rows_kernel(const TI* __restrict__ x, long long rows, int D, long long ld_in, TO* __restrict__ y,
                                   long long ld_out, float* __restrict__ invnorm, float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const TI* xr = x + row * ld_in;
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) { const float v = ldf(xr + i); ss += v * v; }
  ss = warp_sum(ss);
  const float inv = rsqrtf(fmaxf(ss, eps));
  if (y) {
    TO* yr = y + row * ld_out;
    for (int i = lane; i < D; i += 32) stf(yr + i, ldf(xr + i) * inv);
  }
  if (invnorm && lane == 0) invnorm[row] = inv;
}

// dx = (dxhat - xhat <dxhat,xhat>) * invnorm
template <typename TG, typename TX, typename TO>
__global__ void l2norm_rows_bwd_kernel(const TG* __restrict__ dxh, long long ld_g, const TX* __restrict__ xh,
                                       long long ld_x, const float* __restrict__ invnorm, long long rows, int D,
                                       TO* __restrict__ dx, long long ld_o, int accumulate) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const TG* g = dxh + row * ld_g;
  const TX* xr = xh + row * ld_x;
  float dot = 0.f;
  for (int i = lane; i < D; i += 32) dot += ldf(g + i) * ldf(xr + i);
  dot = warp_sum(dot);
  const float inv = invnorm[row];
  TO* o = dx + row * ld_o;
  for (int i = lane; i < D; i += 32) {
    float v = (ldf(g + i) - ldf(xr + i) * dot) * inv;
    if (accumulate) v += ldf(o + i);
    stf(o + i, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention_for_g (attention_lib.py:194-219): one warp per region.
//   qhat = l2n(q); s_w = gamma*<qhat, what_w> - 1e9*[w >= max_len]; a = softmax_w(s); ctx = sum_w a_w what_w
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxWords = 32;
constexpr int kMaxVec = 4;  // D <= 1024

template <typename T>
__global__ void attn_g_fwd_kernel(const T* __restrict__ q, int ld_q, const float* __restrict__ what,
                                  const float* __restrict__ max_len, int B, int R, int L, int D, float gamma,
                                  T* __restrict__ ctx, int ld_ctx, float* __restrict__ attn) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * R) return;
  const int b = row / R;
  const int nvec = D >> 3;
  float qv[kMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int vi = j * 32 + lane;
    if (vi < nvec) {
      load8(q + row * ld_q + vi * 8, qv[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += qv[j][i] * qv[j][i];
    }
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  const float ml = max_len[b];
  float s[kMaxWords];
  float mx = -3.0e38f;
  const float* wb = what + (long long)b * L * D;
#pragma unroll 1
  for (int w = 0; w < L; ++w) {
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = j * 32 + lane;
      if (vi < nvec) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8 + 4);
        d += qv[j][0] * w0.x + qv[j][1] * w0.y + qv[j][2] * w0.z + qv[j][3] * w0.w + qv[j][4] * w1.x +
             qv[j][5] * w1.y + qv[j][6] * w1.z + qv[j][7] * w1.w;
      }
    }
    d = warp_sum(d) * inv * gamma;
    if ((float)w >= ml) d += -1e9f;
    s[w] = d;
    mx = fmaxf(mx, d);
  }
  float sum = 0.f;
#pragma unroll 1
  for (int w = 0; w < L; ++w) { s[w] = __expf(s[w] - mx); sum += s[w]; }
  const float rs = 1.f / sum;
  float acc[kMaxVec][8];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
#pragma unroll 1
  for (int w = 0; w < L; ++w) {
    const float a = s[w] * rs;
    if (lane == 0) attn[row * L + w] = a;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = j * 32 + lane;
      if (vi < nvec) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8 + 4);
        acc[j][0] += a * w0.x; acc[j][1] += a * w0.y; acc[j][2] += a * w0.z; acc[j][3] += a * w0.w;
        acc[j][4] += a * w1.x; acc[j][5] += a * w1.y; acc[j][6] += a * w1.z; acc[j][7] += a * w1.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int vi = j * 32 + lane;
    if (vi < nvec) store8(ctx + row * ld_ctx + vi * 8, acc[j]);
  }
}

template <typename T>
__global__ void attn_g_bwd_kernel(const T* __restrict__ dctx, int ld_dctx, const T* __restrict__ q, int ld_q,
                                  const float* __restrict__ what, const float* __restrict__ attn, int B, int R, int L,
                                  int D, float gamma, T* __restrict__ dq, int ld_dq) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * R) return;
  const int b = row / R;
  const int nvec = D >> 3;
  float qv[kMaxVec][8], gv[kMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int vi = j * 32 + lane;
    if (vi < nvec) {
      load8(q + row * ld_q + vi * 8, qv[j]);
      load8(dctx + row * ld_dctx + vi * 8, gv[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += qv[j][i] * qv[j][i];
    }
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  const float* wb = what + (long long)b * L * D;
  float dA[kMaxWords];
  float t = 0.f;
#pragma unroll 1
  for (int w = 0; w < L; ++w) {
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = j * 32 + lane;
      if (vi < nvec) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8 + 4);
        d += gv[j][0] * w0.x + gv[j][1] * w0.y + gv[j][2] * w0.z + gv[j][3] * w0.w + gv[j][4] * w1.x +
             gv[j][5] * w1.y + gv[j][6] * w1.z + gv[j][7] * w1.w;
      }
    }
    d = warp_sum(d);
    dA[w] = d;
    t += attn[row * L + w] * d;
  }
  float dqh[kMaxVec][8];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) dqh[j][i] = 0.f;
#pragma unroll 1
  for (int w = 0; w < L; ++w) {
    const float ds = attn[row * L + w] * (dA[w] - t) * gamma;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = j * 32 + lane;
      if (vi < nvec) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + (long long)w * D + vi * 8 + 4);
        dqh[j][0] += ds * w0.x; dqh[j][1] += ds * w0.y; dqh[j][2] += ds * w0.z; dqh[j][3] += ds * w0.w;
        dqh[j][4] += ds * w1.x; dqh[j][5] += ds * w1.y; dqh[j][6] += ds * w1.z; dqh[j][7] += ds * w1.w;
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int vi = j * 32 + lane;
    if (vi < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dot += dqh[j][i] * qv[j][i] * inv;
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int vi = j * 32 + lane;
    if (vi < nvec) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (dqh[j][i] - qv[j][i] * inv * dot) * inv;
      store8(dq + row * ld_dq + vi * 8, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// word_loss stages (attention_lib.py:105-191). Index names: i image, j sentence, w word, r region; jw = j*L + w.
// ---------------------------------------------------------------------------------------------------------------------
// alpha[i][r][jw] = softmax_r(gamma1 * S[i*R+r][jw]) (softmax over REGIONS, attention_lib.py:125); also alpha^T.
// The reference adds -1e9 to padded-word columns before the region softmax; those columns are dropped again by the
// word mask of the LSE (attention_lib.py:164-165), so they are computed unmasked here.
// One block = image i x 32 columns jw, 256 threads as (32 columns) x (8 row groups): every global access is a run of 32
// consecutive jw (S, alpha) or 32 consecutive r (alpha^T); the [R][32] tile lives in shared memory between the passes.
constexpr int kWlMaxR = 256;
template <typename T>
__global__ void __launch_bounds__(256)
wl_softmax_kernel(const float* __restrict__ S, int B, int R, int BL, int ldS, float gamma1,
                  T* __restrict__ alpha, T* __restrict__ alphaT) {
  __shared__ float tile[kWlMaxR][33];
  __shared__ float red[8][33];
  const int i = blockIdx.y;
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int jw = blockIdx.x * 32 + cx;
  const bool live = jw < BL;
  const long long base = (long long)i * R * ldS + jw;
  float mx = -3.0e38f;
  for (int r = ry; r < R; r += 8) {
    const float v = live ? S[base + (long long)r * ldS] : 0.f;
    tile[r][cx] = v;
    mx = fmaxf(mx, v);
  }
  red[ry][cx] = mx;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) mx = fmaxf(mx, red[k][cx]);
  __syncthreads();
  float sum = 0.f;
  for (int r = ry; r < R; r += 8) {
    const float e = __expf(gamma1 * (tile[r][cx] - mx));
    tile[r][cx] = e;
    sum += e;
  }
  red[ry][cx] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += red[k][cx];
  const float rs = 1.f / sum;
  for (int r = ry; r < R; r += 8) {
    const T v = from_f<T>(live ? tile[r][cx] * rs : 0.f);
    tile[r][cx] = to_f(v);
    if (jw < ldS) alpha[base + (long long)r * ldS] = v;   // padded columns [BL, ldS) are written as zeros
  }
  __syncthreads();
  // alpha^T[i][jw][r]: one warp per column, lanes along r
  for (int c = ry; c < 32; c += 8) {
    const int jc = blockIdx.x * 32 + c;
    if (jc >= BL) continue;
    T* at = alphaT + ((long long)i * ldS + jc) * R;
    for (int r = cx; r < R; r += 32) at[r] = from_f<T>(tile[r][c]);
  }
}

// dS = gamma1 * alpha * (dalpha - sum_r alpha*dalpha); same block shape as the forward kernel
template <typename T>
__global__ void __launch_bounds__(256)
wl_softmax_bwd_kernel(const T* __restrict__ alpha, const float* __restrict__ dalpha, int B, int R,
                      int BL, int ldS, float gamma1, T* __restrict__ dS) {
  __shared__ float red[8][33];
  const int i = blockIdx.y;
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int jw = blockIdx.x * 32 + cx;
  const bool live = jw < BL;
  const long long base = (long long)i * R * ldS + jw;
  float t = 0.f;
  if (live)
    for (int r = ry; r < R; r += 8)
      t += to_f(alpha[base + (long long)r * ldS]) * dalpha[base + (long long)r * ldS];
  red[ry][cx] = t;
  __syncthreads();
  t = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += red[k][cx];
  if (jw >= ldS) return;
  for (int r = ry; r < R; r += 8) {
    float v = 0.f;
    if (live) {
      const float a = to_f(alpha[base + (long long)r * ldS]);
      v = gamma1 * a * (dalpha[base + (long long)r * ldS] - t);
    }
    dS[base + (long long)r * ldS] = from_f<T>(v);
  }
}

// cos[i][jw] = <W_jw, ctx_i,jw> / (|W_jw| |ctx_i,jw|)   (attention_lib.cosine_similarity with the UN-normalised
// words, attention_lib.py:23-27,162). One warp per (i, jw).
__global__ void wl_cos_kernel(const float* __restrict__ ctx, long long ctx_batch_stride, const float* __restrict__ words,
                              const float* __restrict__ winv, int B, int BL, int D, float* __restrict__ cosv,
                              float* __restrict__ cnorm) {
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= (long long)B * BL) return;
  const int i = idx / BL, jw = idx - (long long)i * BL;
  const float* c = ctx + (long long)i * ctx_batch_stride + (long long)jw * D;
  const float* w = words + (long long)jw * D;
  float dot = 0.f, cc = 0.f;
  for (int f = lane * 4; f < D; f += 128) {
    const float4 a = *reinterpret_cast<const float4*>(c + f);
    const float4 b = *reinterpret_cast<const float4*>(w + f);
    dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    cc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  dot = warp_sum(dot);
  cc = warp_sum(cc);
  if (lane == 0) {
    const float cn = sqrtf(cc);
    cnorm[idx] = cn;
    cosv[idx] = dot * winv[jw] / cn;
  }
}

// sim[j][i] = gamma3 * LSE_{w<max_len_j}(gamma2*cos[i][j,w]) / gamma2 ; pw = softmax weights (0 for padded words)
__global__ void wl_sim_kernel(const float* __restrict__ cosv, const float* __restrict__ max_len, int B, int L,
                              float gamma2, float gamma3, float* __restrict__ sim, float* __restrict__ pw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * B) return;
  const int i = idx / B, j = idx - i * B;
  const float* c = cosv + (long long)i * B * L + (long long)j * L;
  float* p = pw + (long long)i * B * L + (long long)j * L;
  const float ml = max_len[j];
  float mx = -3.0e38f;
  for (int w = 0; w < L; ++w)
    if ((float)w < ml) mx = fmaxf(mx, gamma2 * c[w]);
  float sum = 0.f;
  for (int w = 0; w < L; ++w)
    if ((float)w < ml) sum += __expf(gamma2 * c[w] - mx);
  for (int w = 0; w < L; ++w) p[w] = ((float)w < ml) ? __expf(gamma2 * c[w] - mx) / sum : 0.f;
  sim[(long long)j * B + i] = gamma3 * (mx + logf(sum)) / gamma2;
}

// dctx[i][jw][:] = dsim[j][i]*gamma3*pw * ( W/(|W||ctx|) - cos*ctx/|ctx|^2 )
template <typename T>
__global__ void wl_cos_bwd_kernel(const float* __restrict__ dsim, const float* __restrict__ pw,
                                  const float* __restrict__ cosv, const float* __restrict__ cnorm,
                                  const float* __restrict__ ctx, long long ctx_batch_stride,
                                  const float* __restrict__ words, const float* __restrict__ winv, int B, int L, int D,
                                  float gamma3, T* __restrict__ dctx, long long dctx_batch_stride) {
  const int BL = B * L;
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= (long long)B * BL) return;
  const int i = idx / BL, jw = idx - (long long)i * BL;
  const int j = jw / L;
  const float dcos = dsim[(long long)j * B + i] * gamma3 * pw[idx];
  const float cn = cnorm[idx], cs = cosv[idx];
  const float k1 = dcos * winv[jw] / cn, k2 = dcos * cs / (cn * cn);
  const float* c = ctx + (long long)i * ctx_batch_stride + (long long)jw * D;
  const float* w = words + (long long)jw * D;
  T* o = dctx + (long long)i * dctx_batch_stride + (long long)jw * D;
  for (int f = lane * 8; f < D; f += 256) {
    float v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = (dcos == 0.f) ? 0.f : (k1 * w[f + t] - k2 * c[f + t]);
    store8(o + f, v);
  }
}

template <typename T>
__global__ void transpose_bf16_kernel(const T* __restrict__ src, int rows, int cols, int ld_src,
                                      T* __restrict__ dst, int ld_dst) {
  __shared__ T tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int rr = r0 + r, cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (rr < rows && cc < cols) ? src[(long long)rr * ld_src + cc] : from_f<T>(0.f);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int cc = c0 + r, rr = r0 + threadIdx.x;
    if (cc < cols && rr < ld_dst) dst[(long long)cc * ld_dst + rr] = tile[threadIdx.x][r];
  }
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_l2norm_rows(const void* x, int in_f32, long long rows, int D, long long ld_in, void* y, int out_f32,
                               long long ld_out, float* invnorm, float eps, void* stream) {
  if (!x || rows < 1 || D < 1) return XMC_EINVAL;
  const unsigned grid = (unsigned)ceil_div_ll(rows, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_f32 && out_f32)
    l2norm_rows_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, rows, D, ld_in, (float*)y, ld_out, invnorm, eps);
  else if (in_f32 && !out_f32)
    l2norm_rows_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)x, rows, D, ld_in, (bf16*)y, ld_out, invnorm, eps);
  else if (!in_f32 && out_f32)
    l2norm_rows_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)x, rows, D, ld_in, (float*)y, ld_out, invnorm, eps);
  else
    l2norm_rows_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)x, rows, D, ld_in, (bf16*)y, ld_out, invnorm, eps);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_l2norm_rows_bwd(const void* dxhat, int g_f32, long long ld_g, const void* xhat, int x_f32,
                                   long long ld_x, const float* invnorm, long long rows, int D, void* dx, int out_f32,
                                   long long ld_o, int accumulate, void* stream) {
  if (!dxhat || !xhat || !invnorm || !dx || rows < 1 || D < 1) return XMC_EINVAL;
  const unsigned grid = (unsigned)ceil_div_ll(rows, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (g_f32 && x_f32 && out_f32)
    l2norm_rows_bwd_kernel<float, float, float><<<grid, 256, 0, st>>>((const float*)dxhat, ld_g, (const float*)xhat, ld_x, invnorm, rows, D, (float*)dx, ld_o, accumulate);
  else if (!g_f32 && !x_f32 && !out_f32)
    l2norm_rows_bwd_kernel<bf16, bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)dxhat, ld_g, (const bf16*)xhat, ld_x, invnorm, rows, D, (bf16*)dx, ld_o, accumulate);
  else
    return XMC_EINVAL;
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_attention_g_fwd(const void* q, int act_f32, int ld_q, const float* what, const float* max_len, int B,
                                   int R, int L, int D, float gamma, void* ctx, int ld_ctx, float* attn, void* stream) {
  if (!q || !what || !max_len || !ctx || !attn) return XMC_EINVAL;
  if (L > kMaxWords || D > kMaxVec * 256 || (D % 8) || (ld_q % 8) || (ld_ctx % 8)) return XMC_EINVAL;
  const unsigned grid = (unsigned)ceil_div_ll((long long)B * R, 8);
  XMC_ACT(act_f32, attn_g_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)q, ld_q, what, max_len, B, R, L, D,
                                                                            gamma, (T*)ctx, ld_ctx, attn));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_attention_g_bwd(const void* dctx, int act_f32, int ld_dctx, const void* q, int ld_q, const float* what,
                                   const float* attn, int B, int R, int L, int D, float gamma, void* dq, int ld_dq,
                                   void* stream) {
  if (!dctx || !q || !what || !attn || !dq) return XMC_EINVAL;
  if (L > kMaxWords || D > kMaxVec * 256 || (D % 8) || (ld_q % 8) || (ld_dctx % 8) || (ld_dq % 8)) return XMC_EINVAL;
  const unsigned grid = (unsigned)ceil_div_ll((long long)B * R, 8);
  XMC_ACT(act_f32, attn_g_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)dctx, ld_dctx, (const T*)q, ld_q,
                                                                            what, attn, B, R, L, D, gamma, (T*)dq, ld_dq));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_wl_softmax(const float* S, int B, int R, int BL, int ldS, float gamma1, void* alpha, void* alphaT,
                              int act_f32, void* stream) {
  if (!S || !alpha || !alphaT || B < 1 || R < 1 || R > kWlMaxR || BL < 1 || ldS < BL) return XMC_EINVAL;
  XMC_ACT(act_f32, wl_softmax_kernel<T><<<dim3(ceil_div(ldS, 32), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(
                       S, B, R, BL, ldS, gamma1, (T*)alpha, (T*)alphaT));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_wl_softmax_bwd(const void* alpha, const float* dalpha, int B, int R, int BL, int ldS, float gamma1,
                                  void* dS, int act_f32, void* stream) {
  if (!alpha || !dalpha || !dS || B < 1 || R < 1 || BL < 1 || ldS < BL) return XMC_EINVAL;
  XMC_ACT(act_f32, wl_softmax_bwd_kernel<T><<<dim3(ceil_div(ldS, 32), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(
                       (const T*)alpha, dalpha, B, R, BL, ldS, gamma1, (T*)dS));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_wl_cos(const float* ctx, long long ctx_batch_stride, const float* words, const float* winv, int B,
                          int L, int D, float* cosv, float* cnorm, void* stream) {
  if (!ctx || !words || !winv || !cosv || !cnorm || (D % 4)) return XMC_EINVAL;
  const long long n = (long long)B * B * L;
  wl_cos_kernel<<<(unsigned)ceil_div_ll(n, 8), 256, 0, (cudaStream_t)stream>>>(ctx, ctx_batch_stride, words, winv, B,
                                                                              B * L, D, cosv, cnorm);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_wl_sim(const float* cosv, const float* max_len, int B, int L, float gamma2, float gamma3, float* sim,
                          float* pw, void* stream) {
  if (!cosv || !max_len || !sim || !pw) return XMC_EINVAL;
  wl_sim_kernel<<<ceil_div(B * B, 128), 128, 0, (cudaStream_t)stream>>>(cosv, max_len, B, L, gamma2, gamma3, sim, pw);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_wl_cos_bwd(const float* dsim, const float* pw, const float* cosv, const float* cnorm,
                              const float* ctx, long long ctx_batch_stride, const float* words, const float* winv,
                              int B, int L, int D, float gamma3, void* dctx, long long dctx_batch_stride,
                              int act_f32, void* stream) {
  if (!dsim || !pw || !cosv || !cnorm || !ctx || !words || !winv || !dctx || (D % 8)) return XMC_EINVAL;
  const long long n = (long long)B * B * L;
  XMC_ACT(act_f32, wl_cos_bwd_kernel<T><<<(unsigned)ceil_div_ll(n, 8), 256, 0, (cudaStream_t)stream>>>(
                       dsim, pw, cosv, cnorm, ctx, ctx_batch_stride, words, winv, B, L, D, gamma3, (T*)dctx,
                       dctx_batch_stride));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_transpose_bf16(const void* src, int act_f32, int rows, int cols, int ld_src, void* dst, int ld_dst,
                                  void* stream) {
  if (!src || !dst || rows < 1 || cols < 1) return XMC_EINVAL;
  // also zero-fills dst columns [rows, ld_dst)
  XMC_ACT(act_f32, transpose_bf16_kernel<T><<<dim3(ceil_div(cols, 32), ceil_div(ld_dst, 32)), dim3(32, 8), 0,
                                           (cudaStream_t)stream>>>((const T*)src, rows, cols, ld_src, (T*)dst, ld_dst));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}
