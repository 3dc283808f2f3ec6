// Parameter-side kernels (all multi-tensor: one launch walks a device table of layers):
//   * spectral normalisation: one power-iteration step, sigma, and the sigma term of its backward
//     (xmcgan/libml/layers.py:94-101 and :211-221);
//   * weight preparation: fp32 HWIO kernel -> bf16 K-major copies for the forward ([Cout][tap*Cin+ci]) and dgrad
//     ([Cin][flip(tap)*Cout+co]) tcgen05 GEMMs, scaled by 1/(sigma+eps) when spectrally normalised;
//   * Adam (+ polyak EMA) on flat fp32 buffers (flax.optim.Adam as applied at xmcgan/xmc_gan.py:172-177,252).
#include "common.h"
#include "devutil.cuh"

namespace xmc {

template <typename E>
__device__ __forceinline__ int find_entry(const E* tab, int n, int block, int E::*field) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].*field <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------- spectral norm
// scalars layout (SoA, n = number of SN layers): [0,n) unused | [n,2n) norm_t | [2n,3n) inv_sigma | [3n,4n) dot
// All reductions are two-stage with a fixed summation order (no atomics): bit-identical from run to run.
__global__ void sn_rowdot_kernel(const XmcSnEntry* __restrict__ tab, int n, const float* __restrict__ params,
                                 const float* __restrict__ u0, float* __restrict__ t_ws) {
  const int e = find_entry(tab, n, (int)blockIdx.x, &XmcSnEntry::row_block_begin);
  const XmcSnEntry en = tab[e];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = ((int)blockIdx.x - en.row_block_begin) * 8 + warp;
  float t = 0.f;
  if (k < en.rows) {
    const float* w = params + en.w_off + (long long)k * en.cols;
    const float* u = u0 + en.u_off;
    for (int c = lane; c < en.cols; c += 32) t += w[c] * u[c];
    t = warp_sum(t);
    if (lane == 0) t_ws[en.t_off + k] = t;
  }
}

__global__ void sn_colsum_kernel(const XmcSnEntry* __restrict__ tab, int n, const float* __restrict__ params,
                                 const float* __restrict__ t_ws, float* __restrict__ s_ws) {
  const int e = find_entry(tab, n, (int)blockIdx.x, &XmcSnEntry::col_tile_begin);
  const XmcSnEntry en = tab[e];
  const int local = (int)blockIdx.x - en.col_tile_begin;
  const int ctiles = (en.cols + 31) / 32;
  const int rt = local / ctiles, ct = local - rt * ctiles;
  const int c = ct * 32 + threadIdx.x;
  const int k0 = rt * 256;
  float acc = 0.f;
  if (c < en.cols) {
    const int k1 = min(en.rows, k0 + 256);
    for (int k = k0 + threadIdx.y; k < k1; k += 8)
      acc += t_ws[en.t_off + k] * params[en.w_off + (long long)k * en.cols + c];
  }
  __shared__ float sm[8][33];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < en.cols) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += sm[i][threadIdx.x];
    s_ws[en.s_off + (long long)rt * en.cols + c] = a;  // partial of row tile rt; sn_finalize adds the tiles in order
  }
}

__global__ void sn_finalize_kernel(const XmcSnEntry* __restrict__ tab, int n, float eps, const float* __restrict__ t_ws,
                                   const float* __restrict__ s_ws, float* __restrict__ u_new,
                                   float* __restrict__ scalars) {
  __shared__ float sm[32];
  const int e = blockIdx.x;
  const XmcSnEntry en = tab[e];
  float tt = 0.f;
  for (int k = threadIdx.x; k < en.rows; k += blockDim.x) {
    const float t = t_ws[en.t_off + k];
    tt += t * t;
  }
  tt = block_sum(tt, sm);
  const float norm_t = rsqrtf(tt + eps);  // v0 = t * norm_t   (layers.py:213 / :96)
  const int rtiles = (en.rows + 255) / 256;
  auto col = [&](int c) {                 // (t W)[c]: the row-tile partials of sn_colsum_kernel, added in tile order
    float a = 0.f;
    for (int rt = 0; rt < rtiles; ++rt) a += s_ws[en.s_off + (long long)rt * en.cols + c];
    return a;
  };
  float ss = 0.f;
  for (int c = threadIdx.x; c < en.cols; c += blockDim.x) {
    const float s_ = col(c) * norm_t;  // (v0 W)[c]
    ss += s_ * s_;
  }
  ss = block_sum(ss, sm);
  const float norm_s = rsqrtf(ss + eps);          // u1 = (v0 W) * norm_s  (layers.py:214 / :97)
  for (int c = threadIdx.x; c < en.cols; c += blockDim.x) u_new[en.u_off + c] = col(c) * norm_t * norm_s;
  if (threadIdx.x == 0) {
    const float sigma = ss * norm_s;              // v0 W u1^T
    scalars[n + e] = norm_t;
    scalars[2 * n + e] = 1.f / (sigma + eps);     // kernel / (sigma + eps)  (layers.py:101 / :221)
  }
}

// dot[e] = <dWtilde, W>
__global__ void sn_bwd_dot_kernel(const XmcSnEntry* __restrict__ tab, int n, const float* __restrict__ params,
                                  const float* __restrict__ grads, float* __restrict__ dot_partials) {
  __shared__ float sm[32];
  const int e = find_entry(tab, n, (int)blockIdx.x, &XmcSnEntry::elem_block_begin);
  const XmcSnEntry en = tab[e];
  const long long total = (long long)en.rows * en.cols;
  const long long base = (long long)((int)blockIdx.x - en.elem_block_begin) * 2048;
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long idx = base + i * 256 + threadIdx.x;
    if (idx < total) a += grads[en.w_off + idx] * params[en.w_off + idx];
  }
  a = block_sum(a, sm);
  if (threadIdx.x == 0) dot_partials[blockIdx.x] = a;
}

// scalars[3n + e] = sum of entry e's block partials (fixed order: thread-strided, then the block tree)
__global__ void sn_bwd_dot_finish_kernel(const XmcSnEntry* __restrict__ tab, int n, int total_elem_blocks,
                                         const float* __restrict__ dot_partials, float* __restrict__ scalars) {
  __shared__ float sm[32];
  const int e = blockIdx.x;
  const int b0 = tab[e].elem_block_begin;
  const int b1 = (e + 1 < n) ? tab[e + 1].elem_block_begin : total_elem_blocks;
  float a = 0.f;
  for (int b = b0 + threadIdx.x; b < b1; b += blockDim.x) a += dot_partials[b];
  a = block_sum(a, sm);
  if (threadIdx.x == 0) scalars[3 * n + e] = a;
}

// dW = dWtilde/sigma' - <dWtilde,W>/sigma'^2 * v0^T u1     (sigma' = sigma + eps; u1, v0 are stop-gradient)
__global__ void sn_bwd_apply_kernel(const XmcSnEntry* __restrict__ tab, int n, float* __restrict__ grads,
                                    const float* __restrict__ t_ws, const float* __restrict__ u_new,
                                    const float* __restrict__ scalars) {
  const int e = find_entry(tab, n, (int)blockIdx.x, &XmcSnEntry::elem_block_begin);
  const XmcSnEntry en = tab[e];
  const long long total = (long long)en.rows * en.cols;
  const long long base = (long long)((int)blockIdx.x - en.elem_block_begin) * 2048;
  const float norm_t = scalars[n + e], inv = scalars[2 * n + e], dot = scalars[3 * n + e];
  const float coef = dot * inv * inv * norm_t;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long idx = base + i * 256 + threadIdx.x;
    if (idx < total) {
      const int k = idx / en.cols, c = idx - (long long)k * en.cols;
      grads[en.w_off + idx] = grads[en.w_off + idx] * inv - coef * t_ws[en.t_off + k] * u_new[en.u_off + c];
    }
  }
}

// ------------------------------------------------------------------------------------------------- weight prep
// One block = a 64 (k = tap*Cin+ci) x 64 (cout) tile, 256 threads as 64 x 4: 16 independent row loads per thread, the
// dgrad copy is written in the read orientation (128-byte runs along cout), the forward copy after a transpose through
// shared memory (128-byte runs along k).
constexpr int kPrepTile = 64;
__global__ void __launch_bounds__(256)
prep_weights_kernel(const XmcPrepEntry* __restrict__ tab, int n, const float* __restrict__ params,
                    const float* __restrict__ sn_scalars, int n_sn, bf16* __restrict__ arena,
                    float* __restrict__ bias_arena, const float* __restrict__ cscale) {
  __shared__ float tile[kPrepTile][kPrepTile + 1];
  const int e = find_entry(tab, n, (int)blockIdx.x, &XmcPrepEntry::tile_begin);
  const XmcPrepEntry en = tab[e];
  const int local = (int)blockIdx.x - en.tile_begin;
  const int K = en.taps * en.cin;
  const int ctiles = (en.cout + kPrepTile - 1) / kPrepTile;
  const int kt = local / ctiles, ct = local - kt * ctiles;
  const float scale = (en.sn >= 0) ? sn_scalars[2 * n_sn + en.sn] : 1.f;
  const int c = ct * kPrepTile + threadIdx.x;
  const float cs = (en.cscale_off >= 0 && c < en.cout) ? cscale[en.cscale_off + c] : 1.f;
  // split mode (fp32 activations): every weight is stored as hi = bf16(w), lo = bf16(w - hi) in the parts [hi | hi | lo]
  const int S = en.split ? 3 : 1;
  const long long dgps = en.dg_part_stride > 0 ? en.dg_part_stride : en.cout;
#pragma unroll 4
  for (int r = threadIdx.y; r < kPrepTile; r += 4) {
    const int k = kt * kPrepTile + r;
    float v = 0.f;
    if (k < K && c < en.cout) {
      v = params[en.w_off + (long long)k * en.cout + c] * scale * cs;
      if (en.wk_dg_off >= 0) {
        const int tap = k / en.cin, ci = k - tap * en.cin;
        bf16* o = arena + en.wk_dg_off + (long long)ci * en.ld_dg + (long long)(en.taps - 1 - tap) * S * en.cout + c;
        const bf16 hi = __float2bfloat16(v);
        o[0] = hi;
        if (en.split) {
          o[dgps] = hi;
          o[2 * dgps] = __float2bfloat16(v - __bfloat162float(hi));
        }
      }
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  if (en.wk_fwd_off >= 0) {
    const int k = kt * kPrepTile + threadIdx.x;
    const int tap = k / en.cin, ci = k - tap * en.cin;
    for (int r = threadIdx.y; r < kPrepTile; r += 4) {
      const int cc = ct * kPrepTile + r;
      if (k < K && cc < en.cout) {
        const float v = tile[threadIdx.x][r];
        bf16* o = arena + en.wk_fwd_off + (long long)cc * en.ld_fwd + (long long)tap * S * en.cin + ci;
        const bf16 hi = __float2bfloat16(v);
        o[0] = hi;
        if (en.split) {
          o[en.cin] = hi;
          o[2 * en.cin] = __float2bfloat16(v - __bfloat162float(hi));
        }
      }
    }
  }
  if (local == 0 && en.bias_off >= 0 && en.bias_dst_off >= 0) {
    for (int i = threadIdx.y * kPrepTile + threadIdx.x; i < en.cout; i += 256)
      bias_arena[en.bias_dst_off + i] = params[en.bias_off + i];
  }
}

// ------------------------------------------------------------------------------------------------- sub-pixel prep
// conv3x3(nearest_upsample2x(x)) == four 2x2 convolutions on x, one per output parity (a,b), whose weights are sums of
// the 3x3 taps: W_a[dh] = sum_{kh in S(a,dh)} W[kh], S(0,0)={0}, S(0,1)={1,2}, S(1,0)={0,1}, S(1,1)={2} (same for
// columns). Writes the forward matrix wf[(a*2+b)*Cout + co][(dh*2+dw)*Cin + ci] and the input-gradient matrix
// vd[ci][(r*4+s)*Cout + co] of the equivalent 4x4 / stride-2 / pad-1 convolution over the output gradient, where
// row offset r-1 in {-1,0,1,2} <-> (a,dh) = (1,1),(0,1),(1,0),(0,0).
// One block = a 32(ci) x 32(co) tile. The 9 taps are read with co fastest (coalesced fp32 reads of the HWIO kernel), the
// 16 parity/tap sums are staged in shared memory; vd (co contiguous) is written in the same orientation, wf (ci
// contiguous) after a transpose through the staging tile, so both bf16 matrices are written in full 64-byte runs.
__global__ void __launch_bounds__(256)
subpixel_prep_kernel(const float* __restrict__ w, const float* __restrict__ scale, int Cin, int Cout, int split,
                     bf16* __restrict__ wf, bf16* __restrict__ vd) {
  // split = 1 (fp32 activations): wf [4*Cout][(dh*2+dw)][hi | hi | lo][Cin], vd [Cin][(r*4+s)][hi | hi | lo][Cout]
  __shared__ float tile[8][32][33];  // [dh,b,dw][ci][co] of one output-row parity a, fp32 sums
  const int S = split ? 3 : 1;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const float sc = scale ? *scale : 1.f;
  const int r_of[2][2] = {{3, 1}, {2, 0}};  // r_of[a][dh]
  for (int a = 0; a < 2; ++a) {  // the two output-row parities one after the other (the staging tile holds one)
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    if (ci >= Cin || co >= Cout) continue;
    float k[3][3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) k[kh][kw] = w[((long long)(kh * 3 + kw) * Cin + ci) * Cout + co] * sc;
    float rc[2][2][3];  // row-combined: rc[a][dh][kw]
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      rc[0][0][kw] = k[0][kw];
      rc[0][1][kw] = k[1][kw] + k[2][kw];
      rc[1][0][kw] = k[0][kw] + k[1][kw];
      rc[1][1][kw] = k[2][kw];
    }
#pragma unroll
      for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int dw = 0; dw < 2; ++dw) {
            const float* q = rc[a][dh];
            const float v = (b == 0) ? (dw == 0 ? q[0] : q[1] + q[2]) : (dw == 0 ? q[0] + q[1] : q[2]);
            const bf16 qv = __float2bfloat16(v);
            tile[(dh * 2 + b) * 2 + dw][r][threadIdx.x] = v;
            bf16* o = vd + (long long)ci * (16 * S * Cout) + (long long)(r_of[a][dh] * 4 + r_of[b][dw]) * S * Cout + co;
            o[0] = qv;
            if (split) {
              o[Cout] = qv;
              o[2 * Cout] = __float2bfloat16(v - __bfloat162float(qv));
            }
          }
  }
  __syncthreads();
  // wf[(a*2+b)*Cout + co][(dh*2+dw)*Cin + ci]: ci fastest across the warp
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    if (ci >= Cin || co >= Cout) continue;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int dh = (t >> 2) & 1, b = (t >> 1) & 1, dw = t & 1;
      const float v = tile[t][threadIdx.x][r];
      const bf16 qv = __float2bfloat16(v);
      bf16* o = wf + ((long long)((a * 2 + b) * Cout + co)) * (4 * S * Cin) + (long long)(dh * 2 + dw) * S * Cin + ci;
      o[0] = qv;
      if (split) {
        o[Cin] = qv;
        o[2 * Cin] = __float2bfloat16(v - __bfloat162float(qv));
      }
    }
  }
  __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------- pool-fused prep
// dsample(conv3x3(x)) (the tail of DiscBlock / DiscOptimizedBlock, common.py:66-78,125-132: a 3x3 convolution whose
// output is 2x2-mean-pooled) == ONE 4x4 / stride-2 / pad-1 convolution with the summed weights
//   W4[r][s] = 0.25 * sum_{kh in K(r)} sum_{kw in K(s)} W[kh][kw],   K(0)={0}, K(1)={0,1}, K(2)={1,2}, K(3)={2}
// (output pixel q averages the conv outputs at 2q+a, a in {0,1}, which read input 2q+a+kh-1 = 2q+r-1 with r = a+kh):
// 16 taps per low-resolution output instead of 4 x 9 = 2.25x fewer FLOPs, and the full-resolution conv output is
// never written — the mirror image of the generator's sub-pixel convolution. Writes
//   wf4 [Cout][(r*4+s)*Cin + ci]            forward: xmc_conv2d_fwd KH=KW=4, stride 2, pad 1 on the full-res input
//   wdg [(a*2+b)*Cin + ci][(dh*2+dw)*Cout + co] input gradient: xmc_conv2d_fwd in sub-pixel mode on the LOW-res output
//       gradient (parity (a,b) of the full-res input pixel, window row dh <-> r = {3,1} for a = 0, {2,0} for a = 1)
// split = 1: every K group stored as [hi | hi | lo] (fp32-activation mode). One block = 32 (ci) x 32 (co); the four
// taps of one row r are staged in shared memory so that both matrices are written in contiguous runs.
__global__ void __launch_bounds__(256)
poolconv_prep_kernel(const float* __restrict__ w, const float* __restrict__ scale, int Cin, int Cout, int split,
                     bf16* __restrict__ wf4, bf16* __restrict__ wdg) {
  __shared__ float tile[4][32][33];  // [s][ci][co] of one tap row r
  const int S = split ? 3 : 1;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const float sc = 0.25f * (scale ? *scale : 1.f);
  auto put = [&](bf16* o, long long part_stride, float v) {
    const bf16 hi = __float2bfloat16(v);
    o[0] = hi;
    if (split) {
      o[part_stride] = hi;
      o[2 * part_stride] = __float2bfloat16(v - __bfloat162float(hi));
    }
  };
  for (int r = 0; r < 4; ++r) {
    const int kh0 = r == 0 ? 0 : r - 1, kh1 = r == 3 ? 2 : r;   // K(r) = [kh0, kh1] clipped to 0..2
    const int a = (r == 3 || r == 1) ? 0 : 1, dh = (r == 3 || r == 2) ? 0 : 1;  // r = 3,1 -> a = 0; r = 2,0 -> a = 1
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
      const int ci = ci0 + rr, co = co0 + threadIdx.x;
      if (ci >= Cin || co >= Cout) continue;
      float col[3] = {0.f, 0.f, 0.f};  // sum over kh in K(r) for each kw
      for (int kh = kh0; kh <= min(kh1, 2); ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) col[kw] += w[((long long)(kh * 3 + kw) * Cin + ci) * Cout + co];
      const float v4[4] = {col[0] * sc, (col[0] + col[1]) * sc, (col[1] + col[2]) * sc, col[2] * sc};
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        tile[s2][rr][threadIdx.x] = v4[s2];
        const int b = (s2 == 3 || s2 == 1) ? 0 : 1, dw = (s2 == 3 || s2 == 2) ? 0 : 1;
        put(wdg + ((long long)((a * 2 + b) * Cin + ci)) * (4 * S * Cout) + (long long)(dh * 2 + dw) * S * Cout + co, Cout,
            v4[s2]);
      }
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
      const int co = co0 + rr, ci = ci0 + threadIdx.x;
      if (ci >= Cin || co >= Cout) continue;
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2)
        put(wf4 + (long long)co * (16 * S * Cin) + (long long)(r * 4 + s2) * S * Cin + ci, Cin, tile[s2][threadIdx.x][rr]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------- Adam (+EMA)
// Scalars of one Adam step, formed in double on the host (or from the device step count) and rounded to fp32 once:
// flax.optim.Adam multiplies fp32 arrays by the Python doubles (1 - beta), 1 / (1 - beta^t).
struct AdamScalars {
  float lr, b1, b2, omb1, omb2, eps, c1, c2, gscale, decay, omdecay;
  double b1d, b2d;
};

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, const AdamScalars sc, float* __restrict__ ema,
                            const int* __restrict__ step_dev) {
  const float lr = sc.lr, b1 = sc.b1, b2 = sc.b2, omb1 = sc.omb1, omb2 = sc.omb2, eps = sc.eps, gscale = sc.gscale;
  const float decay = sc.decay, omdecay = sc.omdecay;
  float c1 = sc.c1, c2 = sc.c2;  // bias corrections 1 - beta^t
  if (step_dev) {  // step count t kept on the device (CUDA-graph replays cannot take new host scalars)
    const double t = (double)(*step_dev + 1);
    c1 = (float)(1.0 - pow(sc.b1d, t));
    c2 = (float)(1.0 - pow(sc.b2d, t));
  }
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * gscale;
      ma[k] = b1 * ma[k] + omb1 * gk;
      va[k] = b2 * va[k] + omb2 * gk * gk;
      pa[k] = pa[k] - lr * (ma[k] / c1) / (sqrtf(va[k] / c2) + eps);  // the operation order of flax.optim.Adam
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (ema) {
      float4 ee = reinterpret_cast<float4*>(ema)[i];
      ee.x = ee.x * decay + omdecay * pp.x;
      ee.y = ee.y * decay + omdecay * pp.y;
      ee.z = ee.z * decay + omdecay * pp.z;
      ee.w = ee.w * decay + omdecay * pp.w;
      reinterpret_cast<float4*>(ema)[i] = ee;
    }
  }
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_sn_forward(const XmcSnEntry* table_dev, int n, float eps, const float* params, const float* u0,
                              float* u0_new, float* t_ws, float* s_ws, long long s_ws_floats, float* scalars,
                              int total_row_blocks, int total_col_tiles, void* stream) {
  if (!table_dev || n < 1 || !params || !u0 || !u0_new || !t_ws || !s_ws || !scalars) return XMC_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  (void)s_ws_floats;
  sn_rowdot_kernel<<<total_row_blocks, 256, 0, st>>>(table_dev, n, params, u0, t_ws);
  XMC_LAUNCH_CHECK();
  sn_colsum_kernel<<<total_col_tiles, dim3(32, 8), 0, st>>>(table_dev, n, params, t_ws, s_ws);
  XMC_LAUNCH_CHECK();
  sn_finalize_kernel<<<n, 256, 0, st>>>(table_dev, n, eps, t_ws, s_ws, u0_new, scalars);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_sn_backward(const XmcSnEntry* table_dev, int n, const float* params, float* grads,
                               const float* t_ws, const float* u0_new, float* scalars, float* dot_partials,
                               int total_elem_blocks, void* stream) {
  if (!table_dev || n < 1 || !params || !grads || !t_ws || !u0_new || !scalars || !dot_partials) return XMC_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  sn_bwd_dot_kernel<<<total_elem_blocks, 256, 0, st>>>(table_dev, n, params, grads, dot_partials);
  XMC_LAUNCH_CHECK();
  sn_bwd_dot_finish_kernel<<<n, 256, 0, st>>>(table_dev, n, total_elem_blocks, dot_partials, scalars);
  XMC_LAUNCH_CHECK();
  sn_bwd_apply_kernel<<<total_elem_blocks, 256, 0, st>>>(table_dev, n, grads, t_ws, u0_new, scalars);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_prep_weights(const XmcPrepEntry* table_dev, int n, int total_tiles, const float* params,
                                const float* sn_scalars, int n_sn, void* arena, float* bias_arena,
                                const float* cscale, void* stream) {
  if (!table_dev || n < 1 || total_tiles < 1 || !params || !arena) return XMC_EINVAL;
  prep_weights_kernel<<<total_tiles, dim3(64, 4), 0, (cudaStream_t)stream>>>(table_dev, n, params, sn_scalars, n_sn,
                                                                            (bf16*)arena, bias_arena, cscale);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_subpixel_prep(const float* w, const float* scale, int Cin, int Cout, int split, void* wf, void* vd,
                                 void* stream) {
  if (!w || !wf || !vd || Cin < 8 || Cout < 8 || (Cin % 8) || (Cout % 8)) return XMC_EINVAL;
  subpixel_prep_kernel<<<dim3(ceil_div(Cout, 32), ceil_div(Cin, 32)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      w, scale, Cin, Cout, split, (bf16*)wf, (bf16*)vd);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_poolconv_prep(const float* w, const float* scale, int Cin, int Cout, int split, void* wf4, void* wdg,
                                 void* stream) {
  if (!w || !wf4 || !wdg || Cin < 8 || Cout < 8 || (Cin % 8) || (Cout % 8)) return XMC_EINVAL;
  poolconv_prep_kernel<<<dim3(ceil_div(Cout, 32), ceil_div(Cin, 32)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      w, scale, Cin, Cout, split, (bf16*)wf4, (bf16*)wdg);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

__global__ void inc_i32_kernel(int* x) { *x += 1; }

extern "C" int xmc_adam(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                        double beta2, double eps, double bias_corr1, double bias_corr2, double grad_scale, float* ema,
                        double ema_decay, int* step_dev, int advance_step, void* stream) {
  if (!p || !g || !m || !v || n < 4 || (n & 3)) return XMC_EINVAL;
  if (!aligned16(p) || !aligned16(g) || !aligned16(m) || !aligned16(v) || (ema && !aligned16(ema))) return XMC_EALIGN;
  long long blocks = ceil_div_ll(n / 4, 256);
  // 4 resident blocks per SM (1024 threads, ~100 KB of loads in flight per SM) keep the kernel HBM-bound
  long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  AdamScalars sc;
  sc.lr = (float)lr; sc.b1 = (float)beta1; sc.b2 = (float)beta2;
  sc.omb1 = (float)(1.0 - beta1); sc.omb2 = (float)(1.0 - beta2);
  sc.eps = (float)eps;
  sc.c1 = (float)bias_corr1; sc.c2 = (float)bias_corr2;
  sc.gscale = (float)grad_scale; sc.decay = (float)ema_decay; sc.omdecay = (float)(1.0 - ema_decay);
  sc.b1d = beta1; sc.b2d = beta2;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sc, ema, step_dev);
  XMC_LAUNCH_CHECK();
  if (step_dev && advance_step) {
    inc_i32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    XMC_LAUNCH_CHECK();
  }
  return XMC_OK;
}
