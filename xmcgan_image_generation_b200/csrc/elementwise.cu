// HBM-bound kernels of the XMC-GAN hot path: batch-norm statistics / conditional modulation (+relu, +nearest
// upsample) and their backward, 2x2 pooling, bias-gradient column sums, casts and broadcasts.
// All activation accesses are 16-byte vectors (8 bf16) with the channel index fastest -> fully coalesced.
#include <stdlib.h>
#include "common.h"
#include "devutil.cuh"

namespace xmc {

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm statistics (flax nn.BatchNorm, fp32 stats): block b leaves its partial sums in row b of
// partials[gridDim.x][2C] ([c] = sum_p x[p][c], [C+c] = sum_p x[p][c]^2); sum_partials_kernel adds the rows in a fixed
// order. No atomics anywhere: two runs on the same input are bit-identical.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void bn_stats_kernel(const T* __restrict__ x, long long P, int C, int ld, float* __restrict__ partials) {
  extern __shared__ float sm[];  // [lanes][cv*16]
  const int cv = C >> 3;
  const int cvb = min(cv - blockIdx.y * 256, 256);  // channel vectors handled by this blockIdx.y
  const int lanes = blockDim.x / cvb;
  const int v = threadIdx.x % cvb, pl = threadIdx.x / cvb;
  const int cvec = blockIdx.y * 256 + v;
  float s[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = s2[i] = 0.f;
  if (pl < lanes) {
    // four independent, unpredicated 16-byte loads in flight per thread (pixels p, p+st, p+2st, p+3st), added in pixel
    // order; the tail runs one pixel at a time
    const long long st = (long long)gridDim.x * lanes;
    long long p = (long long)blockIdx.x * lanes + pl;
    for (; p + 3 * st < P; p += 4 * st) {
      Raw8<T> r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k].ld(x + (p + k * st) * ld + cvec * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        r[k].cvt(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; s2[i] += f[i] * f[i]; }
      }
    }
    for (; p < P; p += st) {
      float f[8];
      load8(x + p * ld + cvec * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; s2[i] += f[i] * f[i]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sm[(pl * cvb + v) * 16 + i] = s[i];
      sm[(pl * cvb + v) * 16 + 8 + i] = s2[i];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < cvb * 16; t += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += sm[l * cvb * 16 + t];
    const int vv = t >> 4, i = t & 15;
    const int c = (blockIdx.y * 256 + vv) * 8 + (i & 7);
    partials[(long long)blockIdx.x * 2 * C + (i >> 3) * C + c] = a;
  }
}

// out[j] (+)= sum_r partials[r][j] in a FIXED order (the deterministic second stage of every reduction here): block =
// 32 columns x 8 row lanes; lane l adds rows l, l+8, l+16, ... in index order, the 8 lane sums are then added in lane
// order. The order depends only on (rows, width), never on scheduling.
__global__ void sum_partials_kernel(const float* __restrict__ partials, int rows, int width, float* __restrict__ out,
                                    int accumulate) {
  __shared__ float sm[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (j < width)
    for (int r = threadIdx.y; r < rows; r += 8) a += partials[(long long)r * width + j];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && j < width) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += sm[l][threadIdx.x];
    out[j] = accumulate ? out[j] + t : t;
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, float invP, int C, float eps, float momentum,
                                   const float* ra_mean, const float* ra_var, float* new_ra_mean, float* new_ra_var,
                                   float* mean_rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = sums[c] * invP;
  const float var = sums[C + c] * invP - mean * mean;   // flax: E[x^2] - E[x]^2, no clamp
  mean_rstd[c] = mean;
  mean_rstd[C + c] = rsqrtf(var + eps);
  if (new_ra_mean) {
    new_ra_mean[c] = momentum * ra_mean[c] + (1.f - momentum) * mean;
    new_ra_var[c] = momentum * ra_var[c] + (1.f - momentum) * var;
  }
}

// eval-mode: mean_rstd from running statistics
__global__ void bn_eval_kernel(const float* ra_mean, const float* ra_var, int C, float eps, float* mean_rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean_rstd[c] = ra_mean[c];
  mean_rstd[C + c] = rsqrtf(ra_var[c] + eps);
}

// ---------------------------------------------------------------------------------------------------------------------
// y = relu?( (x-mean)*rstd*(gamma+1)+beta ), optionally written to the 2x2 nearest-upsampled positions.
// gamma/beta come from row (n*Hc + (h>>s))*Hc + (w>>s) of gb: Hc=1 -> ConditionalBatchNorm (layers.py:244-258),
// Hc=16 -> LocalConditionalBatchNorm (layers.py:261-273) with the 1x1 convs evaluated once at 16x16.
// ---------------------------------------------------------------------------------------------------------------------
struct BnP {
  int N, H, W, C, Hc, s, ldG, goff, boff, relu, upsample;
};

__device__ __forceinline__ long long cond_row(const BnP& p, int n, int h, int w) {
  return ((long long)n * p.Hc + (h >> p.s)) * p.Hc + (w >> p.s);
}

// blockDim = (C/8 channel vectors, units per block). A unit is a run of `seg` = min(cell side, 8) consecutive pixels of
// one image row inside one conditioning cell: gamma / beta are loaded once per unit (per pixel they were 2 of the 3
// loads of every 16-byte output vector — L1 traffic, not HBM, bound the kernel at half the copy bandwidth), the pixels
// of the unit are loaded raw four at a time. mean / rstd stay in registers for the whole loop.
__device__ __forceinline__ int bn_seg_shift(const BnP& p) { return p.s < 3 ? p.s : 3; }
template <int V> struct IntC { static constexpr int value = V; };

template <typename T>
__global__ void bn_apply_kernel(BnP p, const T* __restrict__ x, const float* __restrict__ mr,
                                const T* __restrict__ gb, T* __restrict__ y) {
  const int c = threadIdx.x * 8;
  float mean[8], rstd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mean[i] = mr[c + i]; rstd[i] = mr[p.C + c + i]; }
  const int sg = bn_seg_shift(p), seg = 1 << sg;
  const int spr = p.W >> sg;   // units per image row
  const long long units = (long long)p.N * p.H * spr;
  constexpr int kB = 4;
  for (long long u = (long long)blockIdx.x * blockDim.y + threadIdx.y; u < units;
       u += (long long)gridDim.x * blockDim.y) {
    const int ws = (int)(u % spr);
    const long long t = u / spr;
    const int h = (int)(t % p.H), n = (int)(t / p.H);
    const int w0 = ws << sg;
    float g[8], b[8];
    const long long row = cond_row(p, n, h, w0);
    load8(gb + row * p.ldG + p.goff + c, g);
    load8(gb + row * p.ldG + p.boff + c, b);
    const long long pix0 = ((long long)n * p.H + h) * p.W + w0;
    // KB pixels from q0: KB unpredicated raw loads issued back to back, then the arithmetic and the stores
    auto run = [&](auto kb, int q0) {
      constexpr int KB = decltype(kb)::value;
      Raw8<T> xr[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) xr[k].ld(x + (pix0 + q0 + k) * p.C + c);
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        float f[8], o[8];
        xr[k].cvt(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = (f[i] - mean[i]) * rstd[i] * (g[i] + 1.f) + b[i];
          o[i] = p.relu ? fmaxf(v, 0.f) : v;
        }
        if (p.upsample) {
          const int W2 = p.W * 2;
          const long long base = (((long long)n * p.H * 2 + 2 * h) * W2 + 2 * (w0 + q0 + k)) * p.C + c;
          store8(y + base, o);
          store8(y + base + p.C, o);
          store8(y + base + (long long)W2 * p.C, o);
          store8(y + base + (long long)W2 * p.C + p.C, o);
        } else {
          store8(y + (pix0 + q0 + k) * p.C + c, o);
        }
      }
    };
    if (seg >= kB)
      for (int q0 = 0; q0 < seg; q0 += kB) run(IntC<kB>{}, q0);
    else
      for (int q0 = 0; q0 < seg; ++q0) run(IntC<1>{}, q0);
  }
}

// gradient wrt the modulated/normalised tensor for one (pixel, 8 channels): returns g (post relu-mask, upsample-summed)
template <typename T>
__device__ __forceinline__ void bn_load_grad(const BnP& p, const T* dy, int n, int h, int w, int c, float* g) {
  if (p.upsample) {
    const int W2 = p.W * 2;
    const long long base = (((long long)n * p.H * 2 + 2 * h) * W2 + 2 * w) * p.C + c;
    float a[8], b[8], cc[8], d[8];
    load8(dy + base, a);
    load8(dy + base + p.C, b);
    load8(dy + base + (long long)W2 * p.C, cc);
    load8(dy + base + (long long)W2 * p.C + p.C, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = (a[i] + b[i]) + (cc[i] + d[i]);
  } else {
    load8(dy + (((long long)n * p.H + h) * p.W + w) * p.C + c, g);
  }
}

// One thread owns one 8-channel vector v and a stream of (cond row, pixel-split) units: for each unit it walks the
// (H/Hc)^2 pixels of that cond cell (or every psplit-th pixel row of it), so dgamma/dbeta need no cross-block reduction.
// The psplit parts of one cond row sit in consecutive `ri` slots of the same block (rpi % psplit == 0) and are combined
// through shared memory in a fixed order. The per-channel BN terms S1 = sum dxhat, S2 = sum dxhat*xhat stay in
// registers over ALL units of the thread, are combined per block in shared memory (fixed order) and leave as row
// blockIdx.x of partials[gridDim.x][2C]; sum_partials_kernel finishes. No atomics: bit-identical from run to run.
template <typename T>
__global__ void __launch_bounds__(512)
bn_bwd_reduce_kernel(BnP p, const T* __restrict__ dy, const T* __restrict__ x,
                                     const float* __restrict__ mr, const T* __restrict__ gb,
                                     float* __restrict__ dgb, float* __restrict__ partials, long long total_rows,
                                     int psplit, int rpi) {
  extern __shared__ float sm[];  // [rpi][cv][16] combine slots, then mean[C], rstd[C]
  const int cv = p.C >> 3;
  const int v = threadIdx.x % cv, ri = threadIdx.x / cv;  // blockDim.x == cv * rpi
  const int c = v * 8;
  const int side = 1 << p.s;
  // per-channel mean / rstd live in shared memory (16 registers less per thread: the raw loads in flight need them)
  float* s_mr = sm + (long long)blockDim.x * 16;
  for (int t = threadIdx.x; t < 2 * p.C; t += blockDim.x) s_mr[t] = mr[t];
  __syncthreads();
  const float* mean = s_mr + c;
  const float* rstd = s_mr + p.C + c;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  float* slot = sm + ((long long)ri * cv + v) * 16;
  const long long units = total_rows * psplit;
  // every thread of the block runs the same number of iterations (the shared-memory combine below needs barriers)
  const long long iters = (units + (long long)gridDim.x * rpi - 1) / ((long long)gridDim.x * rpi);
  for (long long it = 0; it < iters; ++it) {
    const long long u = (it * gridDim.x + blockIdx.x) * rpi + ri;
    const bool live = u < units;
    const int sp = (int)(u % psplit);  // pixel-split index (psplit > 1 only for per-image cells, Hc == 1)
    const long long row = u / psplit;
    float dg[8], db[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dg[i] = db[i] = 0.f;
    if (live) {
      const int wc = row % p.Hc, hc = (row / p.Hc) % p.Hc, n = row / (p.Hc * p.Hc);
      float gm[8], bt[8];
      load8(gb + row * p.ldG + p.goff + c, gm);
      load8(gb + row * p.ldG + p.boff + c, bt);
      // this thread's pixels of the cell, flattened: q -> (a = sp + (q >> s) * psplit, b = q & (side-1)); kB of them
      // are loaded raw before the first is used (kB x 2 or kB x 5 independent 16-byte loads in flight)
      const int cnt = ((side - sp + psplit - 1) / psplit) << p.s;
      const long long img = (long long)n * p.H;
      auto acc = [&](const float* f, const float* g) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (f[i] - mean[i]) * rstd[i];
          float gi = g[i];
          if (p.relu && (xh * (gm[i] + 1.f) + bt[i]) <= 0.f) gi = 0.f;
          dg[i] += gi * xh;
          db[i] += gi;
        }
      };
      if (!p.upsample) {
        auto run = [&](auto kb, int q0) {   // KB x 2 independent, unpredicated loads, then the sums in pixel order
          constexpr int KB = decltype(kb)::value;
          Raw8<T> xr[KB], gr[KB];
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            const int q = q0 + k;
            const int h = hc * side + sp + (q >> p.s) * psplit, w = wc * side + (q & (side - 1));
            const long long off = ((img + h) * p.W + w) * p.C + c;
            xr[k].ld(x + off);
            gr[k].ld(dy + off);
          }
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            float f[8], g[8];
            xr[k].cvt(f);
            gr[k].cvt(g);
            acc(f, g);
          }
        };
        constexpr int kB = sizeof(T) == 2 ? 4 : 2;
        int q0 = 0;
        for (; q0 + kB <= cnt; q0 += kB) run(IntC<kB>{}, q0);
        for (; q0 < cnt; ++q0) run(IntC<1>{}, q0);
      } else {
        const int W2 = p.W * 2;
        auto run = [&](auto kb, int q0) {   // KB x 5 independent, unpredicated loads
          constexpr int KB = decltype(kb)::value;
          Raw8<T> xr[KB], gr[KB][4];
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            const int q = q0 + k;
            const int h = hc * side + sp + (q >> p.s) * psplit, w = wc * side + (q & (side - 1));
            xr[k].ld(x + ((img + h) * p.W + w) * p.C + c);
            const long long base = ((2 * img + 2 * h) * W2 + 2 * w) * p.C + c;
            gr[k][0].ld(dy + base);
            gr[k][1].ld(dy + base + p.C);
            gr[k][2].ld(dy + base + (long long)W2 * p.C);
            gr[k][3].ld(dy + base + (long long)W2 * p.C + p.C);
          }
#pragma unroll
          for (int k = 0; k < KB; ++k) {
            float f[8], g[8], a1[8], a2[8], a3[8];
            xr[k].cvt(f);
            gr[k][0].cvt(g);
            gr[k][1].cvt(a1);
            gr[k][2].cvt(a2);
            gr[k][3].cvt(a3);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = (g[i] + a1[i]) + (a2[i] + a3[i]);
            acc(f, g);
          }
        };
        constexpr int kB = sizeof(T) == 2 ? 2 : 1;
        int q0 = 0;
        for (; q0 + kB <= cnt; q0 += kB) run(IntC<kB>{}, q0);
        for (; q0 < cnt; ++q0) run(IntC<1>{}, q0);
      }
      // BN terms of this cell: sum dxhat = (gamma+1) * sum g, sum dxhat*xhat = (gamma+1) * sum g*xhat
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += (gm[i] + 1.f) * db[i];
        s2[i] += (gm[i] + 1.f) * dg[i];
      }
    }
    if (psplit == 1) {
      if (live) {
        float* og = dgb + row * p.ldG + p.goff + c;
        float* ob = dgb + row * p.ldG + p.boff + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) { og[i] = dg[i]; ob[i] = db[i]; }
      }
    } else {
      // the psplit parts of a row occupy slots ri = k*psplit .. k*psplit + psplit-1 of this block: part 0 adds them up
#pragma unroll
      for (int i = 0; i < 8; ++i) { slot[i] = dg[i]; slot[8 + i] = db[i]; }
      __syncthreads();
      if (live && sp == 0) {
        float* og = dgb + row * p.ldG + p.goff + c;
        float* ob = dgb + row * p.ldG + p.boff + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = 0.f, b = 0.f;
          for (int k = 0; k < psplit; ++k) {
            a += slot[(long long)k * cv * 16 + i];
            b += slot[(long long)k * cv * 16 + 8 + i];
          }
          og[i] = a; ob[i] = b;
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { slot[i] = s1[i]; slot[8 + i] = s2[i]; }
  __syncthreads();
  for (int t = threadIdx.x; t < cv * 16; t += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < rpi; ++r) a += sm[(long long)r * cv * 16 + t];
    const int vv = t >> 4, k = (t >> 3) & 1, i = t & 7;
    partials[(long long)blockIdx.x * 2 * p.C + k * p.C + vv * 8 + i] = a;
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(BnP p, const T* __restrict__ dy, const T* __restrict__ x,
                                    const float* __restrict__ mr, const T* __restrict__ gb,
                                    const float* __restrict__ sums, float invP, T* __restrict__ dx) {
  const int c = threadIdx.x * 8;
  float mean[8], rstd[8], m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean[i] = mr[c + i];
    rstd[i] = mr[p.C + c + i];
    m1[i] = sums[c + i] * invP;
    m2[i] = sums[p.C + c + i] * invP;
  }
  // units of `seg` consecutive pixels inside one conditioning cell, as in bn_apply_kernel: gamma / beta once per unit
  const int sg = bn_seg_shift(p), seg = 1 << sg;
  const int spr = p.W >> sg;
  const long long units = (long long)p.N * p.H * spr;
  constexpr int kB = sizeof(T) == 2 ? 4 : 2;
  for (long long u = (long long)blockIdx.x * blockDim.y + threadIdx.y; u < units;
       u += (long long)gridDim.x * blockDim.y) {
    const int ws = (int)(u % spr);
    const long long t = u / spr;
    const int h = (int)(t % p.H), n = (int)(t / p.H);
    const int w0 = ws << sg;
    float gm[8], bt[8];
    const long long row = cond_row(p, n, h, w0);
    load8(gb + row * p.ldG + p.goff + c, gm);
    load8(gb + row * p.ldG + p.boff + c, bt);
    const long long pix0 = ((long long)n * p.H + h) * p.W + w0;
    // dx of one pixel from its x vector f and (upsample-summed) gradient g
    auto emit = [&](int q, const float* f, const float* g) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (f[i] - mean[i]) * rstd[i];
        float gi = g[i];
        if (p.relu && (xh * (gm[i] + 1.f) + bt[i]) <= 0.f) gi = 0.f;
        const float dxh = gi * (gm[i] + 1.f);
        o[i] = rstd[i] * (dxh - m1[i] - xh * m2[i]);
      }
      store8(dx + (pix0 + q) * p.C + c, o);
    };
    if (!p.upsample) {
      auto run = [&](auto kb, int q0) {        // KB x 2 independent, unpredicated loads in flight
        constexpr int KB = decltype(kb)::value;
        Raw8<T> xr[KB], gr[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          xr[k].ld(x + (pix0 + q0 + k) * p.C + c);
          gr[k].ld(dy + (pix0 + q0 + k) * p.C + c);
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          float f[8], g[8];
          xr[k].cvt(f);
          gr[k].cvt(g);
          emit(q0 + k, f, g);
        }
      };
      if (seg >= kB)
        for (int q0 = 0; q0 < seg; q0 += kB) run(IntC<kB>{}, q0);
      else
        for (int q0 = 0; q0 < seg; ++q0) run(IntC<1>{}, q0);
    } else {
      const int W2 = p.W * 2;
      for (int q = 0; q < seg; ++q) {          // 5 independent loads in flight
        Raw8<T> xr, gr[4];
        xr.ld(x + (pix0 + q) * p.C + c);
        const long long base = (((long long)n * p.H * 2 + 2 * h) * W2 + 2 * (w0 + q)) * p.C + c;
        gr[0].ld(dy + base);
        gr[1].ld(dy + base + p.C);
        gr[2].ld(dy + base + (long long)W2 * p.C);
        gr[3].ld(dy + base + (long long)W2 * p.C + p.C);
        float f[8], g[8], a1[8], a2[8], a3[8];
        xr.cvt(f);
        gr[0].cvt(g);
        gr[1].cvt(a1);
        gr[2].cvt(a2);
        gr[3].cvt(a3);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = (g[i] + a1[i]) + (a2[i] + a3[i]);
        emit(q, f, g);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2x2 pooling (dsample, common.py:54-55 — scale 0.25; scale 1 gives the transpose of nearest upsample) and its
// transpose.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pool2_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ low,
                             int N, int H, int W, int C, float scale, T* __restrict__ out,
                             T* __restrict__ out_relu) {
  const int cv = C >> 3;
  const long long total = (long long)N * H * W * cv;  // H,W are OUTPUT dims
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = idx % cv;
    long long pix = idx / cv;
    const int w = pix % W;
    const int h = (pix / W) % H;
    const int n = pix / ((long long)W * H);
    const int c = v * 8;
    const int W2 = 2 * W;
    const long long base = (((long long)n * H * 2 + 2 * h) * W2 + 2 * w) * C + c;
    const long long offs[4] = {0, C, (long long)W2 * C, (long long)W2 * C + C};
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      load8(a + base + offs[k], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
      if (b) {
        load8(b + base + offs[k], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= scale;
    if (low) {
      float f[8];
      load8(low + pix * C + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
    store8(out + pix * C + c, acc);
    if (out_relu) {
      // relu of the value as stored (bf16-rounded), so mask and activation agree bit-for-bit
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
      store8(out_relu + pix * C + c, acc);
    }
  }
}

template <typename T>
__global__ void unpool2_kernel(const T* __restrict__ dout, int N, int H, int W, int C, float scale,
                               T* __restrict__ g) {
  const int cv = C >> 3;
  const long long total = (long long)N * H * W * cv;  // H,W are the LOW-res dims of dout
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = idx % cv;
    long long pix = idx / cv;
    const int w = pix % W;
    const int h = (pix / W) % H;
    const int n = pix / ((long long)W * H);
    const int c = v * 8;
    float f[8];
    load8(dout + pix * C + c, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] *= scale;
    const int W2 = 2 * W;
    const long long base = (((long long)n * H * 2 + 2 * h) * W2 + 2 * w) * C + c;
    store8(g + base, f);
    store8(g + base + C, f);
    store8(g + base + (long long)W2 * C, f);
    store8(g + base + (long long)W2 * C + C, f);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// out[c] += sum_p x[p][c]  (bias gradients). gridDim.x == 1: the block adds its sums to out directly; otherwise block b
// stores them in row b of out (= partials[gridDim.x][C]) and sum_partials_kernel finishes in a fixed order.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int kU = 4>
__global__ void colsum_kernel(const T* __restrict__ x, long long P, int C, int ld, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int cv = C >> 3;
  const int cvb = min(cv - blockIdx.y * 256, 256);
  const int lanes = blockDim.x / cvb;
  const int v = threadIdx.x % cvb, pl = threadIdx.x / cvb;
  const int cvec = blockIdx.y * 256 + v;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (pl < lanes) {
    // kU independent 16-byte loads in flight per thread (unpredicated, so that they are issued back to back), added in
    // pixel order; the tail runs one pixel at a time
    const long long st = (long long)gridDim.x * lanes;
    long long p = (long long)blockIdx.x * lanes + pl;
    for (; p + (kU - 1) * st < P; p += kU * st) {
      Raw8<T> r[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) r[k].ld(x + (p + k * st) * ld + cvec * 8);
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        float f[8];
        r[k].cvt(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += f[i];
      }
    }
    for (; p < P; p += st) {
      float f[8];
      load8(x + p * ld + cvec * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[(pl * cvb + v) * 8 + i] = s[i];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < cvb * 8; t += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += sm[l * cvb * 8 + t];
    const int c = (blockIdx.y * 256 + (t >> 3)) * 8 + (t & 7);
    if (gridDim.x == 1) out[c] += a;
    else out[(long long)blockIdx.x * C + c] = a;
  }
}

// scalar fallback for channel counts that are not a multiple of 8 (the 3-channel image gradient)
template <typename T>
__global__ void colsum_scalar_kernel(const T* __restrict__ x, long long P, int C, int ld, float* __restrict__ out) {
  __shared__ float sm[256];
  const int c = blockIdx.y;
  float s = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x)
    s += to_f(x[p * ld + c]);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (gridDim.x == 1) out[c] += sm[0];
    else out[(long long)blockIdx.x * C + c] = sm[0];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// x_pool[n][c] = sum_hw relu(x[n][hw][c])   (xmc_net.py:97-98) and its backward
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void relu_sumhw_kernel(const T* __restrict__ x, int HW, int C, float* __restrict__ out) {
  const int n = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c >= C) return;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (int q = 0; q < HW; ++q) {
    float f[8];
    load8(x + ((long long)n * HW + q) * C + c, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += fmaxf(f[i], 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) out[(long long)n * C + c + i] = s[i];
}

template <typename T>
__global__ void relu_sumhw_bwd_kernel(const T* __restrict__ x, const float* __restrict__ dout, int HW, int C,
                                      T* __restrict__ dx) {
  const int n = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (c >= C) return;
  float d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = dout[(long long)n * C + c + i];
  for (int q = 0; q < HW; ++q) {
    float f[8], o[8];
    load8(x + ((long long)n * HW + q) * C + c, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = f[i] > 0.f ? d[i] : 0.f;
    store8(dx + ((long long)n * HW + q) * C + c, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// casts / broadcasts / grouped row sums (2-D, pitched)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, long long rows, int cols, long long ld_src,
                                     bf16* __restrict__ dst, long long ld_dst) {
  const long long total = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols;
    const int c = idx % cols;
    dst[r * ld_dst + c] = __float2bfloat16(src[r * ld_src + c]);
  }
}

__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, long long rows, int cols, long long ld_src,
                                     float* __restrict__ dst, long long ld_dst, int accumulate) {
  const long long total = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols;
    const int c = idx % cols;
    const float v = __bfloat162float(src[r * ld_src + c]);
    if (accumulate) dst[r * ld_dst + c] += v; else dst[r * ld_dst + c] = v;
  }
}

// dst[(b*reps + r)][c] = src[b][c]
template <typename T>
__global__ void bcast_rows_kernel(const T* __restrict__ src, int B, int reps, int cols, int ld_src,
                                  T* __restrict__ dst, int ld_dst) {
  const long long total = (long long)B * reps * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % cols;
    const long long br = idx / cols;
    const int b = br / reps;
    dst[br * ld_dst + c] = src[(long long)b * ld_src + c];
  }
}

// dst[b][c] (+)= sum_r src[(b*reps + r)][c]
template <typename T>
__global__ void sum_rows_kernel(const T* __restrict__ src, int B, int reps, int cols, int ld_src,
                                float* __restrict__ dst, int ld_dst, int accumulate) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < reps; ++r) s += to_f(src[((long long)b * reps + r) * ld_src + c]);
  if (accumulate) dst[(long long)b * ld_dst + c] += s; else dst[(long long)b * ld_dst + c] = s;
}

// fp32 -> bf16 [hi | lo | hi] split along the channel axis: dst[r][0:C] = hi = bf16(x), dst[r][C:2C] = lo = bf16(x - hi),
// dst[r][2C:3C] = hi. This is the A operand of the tensor-core GEMMs in fp32-activation mode (XmcConvDesc.act_f32): with
// the weights stored as [hi | hi | lo] the K-concatenated product is hi*hi + lo*hi + hi*lo, fp32-accumulated, i.e. 16
// mantissa bits per operand. The weight-gradient GEMM reads the hi and lo parts as two pitched views of the same buffer.
// mode 1 gives the B-operand order [hi | hi | lo] instead; mode 2 the two-part "pair" form [hi | lo] (the forward GEMM
// reads its hi part twice through the tensor map, XmcConvDesc.act_f32 = 2; the weight-gradient GEMM only needs views).
__global__ void split3_kernel(const float* __restrict__ src, long long rows, int C, long long ld_src, int mode,
                              bf16* __restrict__ dst, long long ld_dst) {
  const int cv = C >> 3;
  const long long total = rows * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cv;
    const int c = (int)(idx - r * cv) * 8;
    float f[8], lo[8];
    load8(src + r * ld_src + c, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) lo[i] = f[i] - __bfloat162float(__float2bfloat16(f[i]));
    bf16* o = dst + r * ld_dst + c;
    store8(o, f);
    store8(o + C, mode == 1 ? f : lo);
    if (mode != 2) store8(o + 2 * C, mode == 1 ? lo : f);
  }
}

// y = relu(a) (b == nullptr) or y = a + b: the two stand-alone elementwise ops of the module-level API (nn.relu and the
// residual `x + x0` of the blocks in nets/common.py); the fused engine folds both into GEMM epilogues instead.
template <typename T>
__global__ void relu_or_add_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n8, T* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8], g[8];
    load8(a + i * 8, f);
    if (b) {
      load8(b + i * 8, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += g[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
    }
    store8(y + i * 8, f);
  }
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// (C/8, units-per-block) thread block and a grid of at most `per_sm` blocks per SM — the blocks that are resident at the
// same time, so that the per-block prologue (mean / rstd from global memory) is paid once, not once per wave
static void bn_launch_dims(const BnP& p, dim3* grid, dim3* block, int per_sm) {
  const int cv = p.C / 8;
  const int py = 256 / cv > 0 ? 256 / cv : 1;
  const long long P = ((long long)p.N * p.H * p.W) >> (p.s < 3 ? p.s : 3);   // units of min(cell side, 8) pixels
  long long g = (P + py - 1) / py;
  const long long cap = (long long)num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  *grid = dim3((unsigned)g);
  *block = dim3(cv, py);
}

static int fill_bnp(const XmcBnDesc* d, BnP* p) {
  if (!d || d->N < 1 || d->H < 1 || d->W < 1 || d->C < 8 || (d->C % 8) || d->Hc < 1) return XMC_EINVAL;
  if (d->H != d->W || d->H % d->Hc) return XMC_EINVAL;
  int s = 0;
  while ((d->Hc << s) < d->H) ++s;
  if ((d->Hc << s) != d->H) return XMC_EINVAL;
  if ((d->ldG % 8) || (d->goff % 8) || (d->boff % 8)) return XMC_EINVAL;
  p->N = d->N; p->H = d->H; p->W = d->W; p->C = d->C; p->Hc = d->Hc; p->s = s;
  p->ldG = d->ldG; p->goff = d->goff; p->boff = d->boff; p->relu = d->relu; p->upsample = d->upsample;
  return XMC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Input contract producer (COCODataset.preprocess, xmcgan/libml/coco_dataset.py:127-167) for already decoded / resized
// examples: per-example left-right flip + clip to [0,1]; pick sentence idx[n] of the M captions: embedding[n] =
// emb[n][idx], max_len[n] = len[n][idx], sentence_embedding[n] = sum_words(emb[n][idx]) / len[n][idx] (the sum runs
// over ALL L word slots, padded ones included, exactly as tf.reduce_sum(embedding, axis=-2) does).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void prep_image_kernel(const float* __restrict__ img, const unsigned char* __restrict__ flip, int N, int H,
                                  int W, float* __restrict__ out) {
  const long long total = (long long)N * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int w = idx % W;
    const long long row = idx / W;           // n*H + h
    const int n = (int)(row / H);
    const int ws = flip[n] ? W - 1 - w : w;
    const float* s = img + (row * W + ws) * 3;
    float* o = out + idx * 3;
    o[0] = fminf(fmaxf(s[0], 0.f), 1.f);
    o[1] = fminf(fmaxf(s[1], 0.f), 1.f);
    o[2] = fminf(fmaxf(s[2], 0.f), 1.f);
  }
}

// one block per example; threads along the feature dimension
__global__ void prep_caption_kernel(const float* __restrict__ emb, const int* __restrict__ len,
                                    const int* __restrict__ idx, int M, int L, int E, float* __restrict__ emb_out,
                                    float* __restrict__ len_out, float* __restrict__ sent_out) {
  const int n = blockIdx.x;
  const int j = idx[n];
  const float* src = emb + ((long long)n * M + j) * L * E;
  const float ml = (float)len[(long long)n * M + j];
  if (threadIdx.x == 0) len_out[n] = ml;
  for (int f = threadIdx.x; f < E; f += blockDim.x) {
    float acc = 0.f;
    for (int w = 0; w < L; ++w) {
      const float v = src[(long long)w * E + f];
      emb_out[((long long)n * L + w) * E + f] = v;
      acc += v;
    }
    sent_out[(long long)n * E + f] = acc / ml;
  }
}

}  // namespace xmc

using namespace xmc;

static int launch_sum_partials(const float* partials, int rows, int width, float* out, int accumulate,
                               cudaStream_t stream) {
  sum_partials_kernel<<<ceil_div(width, 32), dim3(32, 8), 0, stream>>>(partials, rows, width, out, accumulate);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_sum_partials(const float* partials, int rows, int width, float* out, int accumulate, void* stream) {
  if (!partials || !out || rows < 1 || width < 1) return XMC_EINVAL;
  return launch_sum_partials(partials, rows, width, out, accumulate, (cudaStream_t)stream);
}

extern "C" int xmc_bn_stats(const void* x, int act_f32, long long P, int C, int ld, float* sums, float* partials,
                            int rows, void* stream) {
  if (!x || !sums || !partials || rows < 1 || P < 1 || C < 8 || (C % 8) || (ld % 8)) return XMC_EINVAL;
  const int cv = C / 8;
  const int ny = ceil_div(cv, 256);
  const int cvb = cv < 256 ? cv : 256;
  const int lanes = 256 / cvb;
  const size_t smem = (size_t)lanes * cvb * 16 * sizeof(float);
  XMC_ACT(act_f32, bn_stats_kernel<T><<<dim3((unsigned)rows, ny), 256, smem, (cudaStream_t)stream>>>((const T*)x, P, C, ld,
                                                                                               partials));
  XMC_LAUNCH_CHECK();
  return launch_sum_partials(partials, rows, 2 * C, sums, 0, (cudaStream_t)stream);
}

extern "C" int xmc_bn_finalize(const float* sums, long long P, int C, float eps, float momentum, const float* ra_mean,
                               const float* ra_var, float* new_ra_mean, float* new_ra_var, float* mean_rstd,
                               void* stream) {
  if (!sums || !mean_rstd || P < 1 || C < 1) return XMC_EINVAL;
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, 1.f / (float)P, C, eps, momentum,
                                                                        ra_mean, ra_var, new_ra_mean, new_ra_var,
                                                                        mean_rstd);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_bn_eval_stats(const float* ra_mean, const float* ra_var, int C, float eps, float* mean_rstd,
                                 void* stream) {
  if (!ra_mean || !ra_var || !mean_rstd || C < 1) return XMC_EINVAL;
  bn_eval_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(ra_mean, ra_var, C, eps, mean_rstd);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_bn_apply(const XmcBnDesc* d, const void* x, const float* mean_rstd, const void* gb, void* y,
                            void* stream) {
  BnP p;
  int r = fill_bnp(d, &p);
  if (r) return r;
  if (!x || !mean_rstd || !gb || !y) return XMC_EINVAL;
  if (p.C / 8 > 1024) return XMC_EINVAL;
  dim3 grid, block;
  bn_launch_dims(p, &grid, &block, d->act_f32 ? 2 : 3);   // 81 / 96 registers x <= 256 threads
  XMC_ACT(d->act_f32, bn_apply_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(p, (const T*)x, mean_rstd,
                                                                                (const T*)gb, (T*)y));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_bn_bwd_reduce(const XmcBnDesc* d, const void* dy, const void* x, const float* mean_rstd,
                                 const void* gb, float* dgb, float* sums, float* partials, int rows, void* stream) {
  BnP p;
  int r = fill_bnp(d, &p);
  if (r) return r;
  if (!dy || !x || !mean_rstd || !gb || !dgb || !sums || !partials || rows < 1) return XMC_EINVAL;
  const int cv = p.C / 8;
  const long long cond_rows = (long long)p.N * p.Hc * p.Hc;
  // per-image cells (ConditionalBatchNorm) have few rows and many pixels: the pixel rows of a cell are split over
  // psplit threads of the SAME block, whose dgamma/dbeta parts are combined through shared memory in a fixed order
  int psplit = 1;
  if (p.Hc == 1) {
    const int side = 1 << p.s;
    psplit = side < 8 ? side : 8;
    while (psplit > 1 && cv * psplit > 512) psplit >>= 1;
  }
  // block = rpi units x cv channel vectors (rpi a multiple of psplit); every thread streams over its units
  int rpi = cv >= 256 ? 1 : 256 / cv;
  rpi = rpi / psplit * psplit;
  if (rpi < psplit) rpi = psplit;
  const int threads = cv * rpi;
  const size_t smem = ((size_t)threads * 16 + 2 * (size_t)p.C) * sizeof(float);
  if (threads > 512 || smem > 48 * 1024) return XMC_EINVAL;
  long long gx = ceil_div_ll(cond_rows * psplit, rpi);
  if (gx > rows) gx = rows;
  XMC_ACT(d->act_f32, bn_bwd_reduce_kernel<T><<<(unsigned)gx, threads, smem, (cudaStream_t)stream>>>(
                          p, (const T*)dy, (const T*)x, mean_rstd, (const T*)gb, dgb, partials, cond_rows, psplit, rpi));
  XMC_LAUNCH_CHECK();
  return launch_sum_partials(partials, (int)gx, 2 * p.C, sums, 0, (cudaStream_t)stream);
}

extern "C" int xmc_bn_bwd_apply(const XmcBnDesc* d, const void* dy, const void* x, const float* mean_rstd,
                                const void* gb, const float* sums, void* dx, void* stream) {
  BnP p;
  int r = fill_bnp(d, &p);
  if (r) return r;
  if (!dy || !x || !mean_rstd || !gb || !sums || !dx) return XMC_EINVAL;
  if (p.C / 8 > 1024) return XMC_EINVAL;
  const float invP = 1.f / (float)((long long)p.N * p.H * p.W * (d->replicas > 1 ? d->replicas : 1));
  dim3 grid, block;
  bn_launch_dims(p, &grid, &block, 2);                    // __launch_bounds__(256, 2)
  XMC_ACT(d->act_f32, bn_bwd_apply_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(
                          p, (const T*)dy, (const T*)x, mean_rstd, (const T*)gb, sums, invP, (T*)dx));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_pool2(const void* a, const void* b, const void* low, int act_f32, int N, int Hout, int Wout, int C,
                         float scale, void* out, void* out_relu, void* stream) {
  if (!a || !out || N < 1 || Hout < 1 || Wout < 1 || C < 8 || (C % 8)) return XMC_EINVAL;
  const long long total = (long long)N * Hout * Wout * (C / 8);
  XMC_ACT(act_f32, pool2_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                       (const T*)a, (const T*)b, (const T*)low, N, Hout, Wout, C, scale, (T*)out, (T*)out_relu));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_unpool2(const void* dout, int act_f32, int N, int Hin, int Win, int C, float scale, void* g,
                           void* stream) {
  if (!dout || !g || N < 1 || Hin < 1 || Win < 1 || C < 8 || (C % 8)) return XMC_EINVAL;
  const long long total = (long long)N * Hin * Win * (C / 8);
  XMC_ACT(act_f32, unpool2_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)dout, N, Hin, Win,
                                                                                        C, scale, (T*)g));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_colsum(const void* x, int act_f32, long long P, int C, int ld, float* out, float* partials, int rows,
                          void* stream) {
  if (!x || !out || P < 1 || C < 1 || rows < 1 || (rows > 1 && !partials)) return XMC_EINVAL;
  // one block (per channel group) adds straight into out; several blocks leave partial rows that are added in order
  if (C % 8 || ld % 8) {
    long long gx = ceil_div_ll(P, 256 * 16);
    if (gx > rows) gx = rows;
    if (gx < 1) gx = 1;
    XMC_ACT(act_f32, colsum_scalar_kernel<T><<<dim3((unsigned)gx, C), 256, 0, (cudaStream_t)stream>>>(
                         (const T*)x, P, C, ld, gx == 1 ? out : partials));
    XMC_LAUNCH_CHECK();
    return gx == 1 ? XMC_OK : launch_sum_partials(partials, (int)gx, C, out, 1, (cudaStream_t)stream);
  }
  const int cv = C / 8;
  const int ny = ceil_div(cv, 256);
  const int cvb = cv < 256 ? cv : 256;
  const int lanes = 256 / cvb;
  long long gx = ceil_div_ll(P, (long long)lanes * 8);
  if (gx > rows) gx = rows;
  if (gx < 1) gx = 1;
  const size_t smem = (size_t)lanes * cvb * 8 * sizeof(float);
  XMC_ACT(act_f32, colsum_kernel<T, 4><<<dim3((unsigned)gx, ny), 256, smem, (cudaStream_t)stream>>>(
                       (const T*)x, P, C, ld, gx == 1 ? out : partials));
  XMC_LAUNCH_CHECK();
  return gx == 1 ? XMC_OK : launch_sum_partials(partials, (int)gx, C, out, 1, (cudaStream_t)stream);
}

extern "C" int xmc_relu_sumhw(const void* x, int act_f32, int N, int HW, int C, float* out, void* stream) {
  if (!x || !out || N < 1 || HW < 1 || C < 8 || (C % 8)) return XMC_EINVAL;
  XMC_ACT(act_f32, relu_sumhw_kernel<T><<<dim3(ceil_div(C / 8, 64), N), 64, 0, (cudaStream_t)stream>>>((const T*)x, HW, C,
                                                                                                  out));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_relu_sumhw_bwd(const void* x, int act_f32, const float* dout, int N, int HW, int C, void* dx,
                                  void* stream) {
  if (!x || !dout || !dx || N < 1 || HW < 1 || C < 8 || (C % 8)) return XMC_EINVAL;
  XMC_ACT(act_f32, relu_sumhw_bwd_kernel<T><<<dim3(ceil_div(C / 8, 64), N), 64, 0, (cudaStream_t)stream>>>(
                       (const T*)x, dout, HW, C, (T*)dx));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_relu_or_add(const void* a, const void* b, int act_f32, long long n, void* y, void* stream) {
  if (!a || !y || n < 8 || (n % 8)) return XMC_EINVAL;
  if (!aligned16(a) || !aligned16(y) || (b && !aligned16(b))) return XMC_EALIGN;
  XMC_ACT(act_f32, relu_or_add_kernel<T><<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
                       (const T*)a, (const T*)b, n / 8, (T*)y));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_split3(const float* src, long long rows, int C, long long ld_src, int mode, void* dst,
                          long long ld_dst, void* stream) {
  if (!src || !dst || rows < 1 || C < 8 || (C % 8) || (ld_src % 4) || (ld_dst % 8) || mode < 0 || mode > 2 ||
      ld_dst < (mode == 2 ? 2 : 3) * C)
    return XMC_EINVAL;
  if (!aligned16(src) || !aligned16(dst)) return XMC_EALIGN;
  split3_kernel<<<grid_for(rows * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(src, rows, C, ld_src, mode,
                                                                                (bf16*)dst, ld_dst);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_cast_f32_to_bf16(const float* src, long long rows, int cols, long long ld_src, void* dst,
                                    long long ld_dst, void* stream) {
  if (!src || !dst || rows < 1 || cols < 1) return XMC_EINVAL;
  cast_f32_bf16_kernel<<<grid_for(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld_src,
                                                                                    (bf16*)dst, ld_dst);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_cast_bf16_to_f32(const void* src, long long rows, int cols, long long ld_src, float* dst,
                                    long long ld_dst, int accumulate, void* stream) {
  if (!src || !dst || rows < 1 || cols < 1) return XMC_EINVAL;
  cast_bf16_f32_kernel<<<grid_for(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, rows, cols,
                                                                                    ld_src, dst, ld_dst, accumulate);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_bcast_rows(const void* src, int act_f32, int B, int reps, int cols, int ld_src, void* dst,
                              int ld_dst, void* stream) {
  if (!src || !dst || B < 1 || reps < 1 || cols < 1) return XMC_EINVAL;
  XMC_ACT(act_f32, bcast_rows_kernel<T><<<grid_for((long long)B * reps * cols, 256), 256, 0, (cudaStream_t)stream>>>(
                       (const T*)src, B, reps, cols, ld_src, (T*)dst, ld_dst));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_sum_rows(const void* src, int act_f32, int B, int reps, int cols, int ld_src, float* dst,
                            int ld_dst, int accumulate, void* stream) {
  if (!src || !dst || B < 1 || reps < 1 || cols < 1) return XMC_EINVAL;
  XMC_ACT(act_f32, sum_rows_kernel<T><<<dim3(ceil_div(cols, 128), B), 128, 0, (cudaStream_t)stream>>>(
                       (const T*)src, B, reps, cols, ld_src, dst, ld_dst, accumulate));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_prep_image(const float* img, const unsigned char* flip, int N, int H, int W, float* out,
                              void* stream) {
  if (!img || !flip || !out || N < 1 || H < 1 || W < 1) return XMC_EINVAL;
  const long long total = (long long)N * H * W;
  prep_image_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, flip, N, H, W, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_prep_caption(const float* emb, const int* len, const int* idx, int N, int M, int L, int E,
                                float* emb_out, float* len_out, float* sent_out, void* stream) {
  if (!emb || !len || !idx || !emb_out || !len_out || !sent_out || N < 1 || M < 1 || L < 1 || E < 1) return XMC_EINVAL;
  prep_caption_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(emb, len, idx, M, L, E, emb_out, len_out, sent_out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}
