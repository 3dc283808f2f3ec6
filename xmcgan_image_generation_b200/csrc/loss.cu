// Loss heads of the discriminator: InfoNCE (attention_lib.contrastive_loss), the symmetric cross-entropy shared with
// word_loss, the hinge GAN loss (losses.hinge_loss) and the projection-discriminator logit (xmc_net.py:97-104).
// All reductions are warp-shuffle based; the matrices involved are [B,B] with B = per-replica batch (local negatives,
// exactly as the reference: sync_match is NotImplemented there, attention_lib.py:58-62).
#include "common.h"
#include "devutil.cuh"

namespace xmc {

// C[i][j] = scale * <A[i,:], B[j,:]>   one block per i, one warp per j (strided)
__global__ void small_gemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int n, int m, int D,
                                     float scale, float* __restrict__ C) {
  const int i = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < m; j += nw) {
    float d = 0.f;
    for (int f = lane; f < D; f += 32) d += A[(long long)i * D + f] * Bm[(long long)j * D + f];
    d = warp_sum(d);
    if (lane == 0) C[(long long)i * m + j] = d * scale;
  }
}

// out[i][f] (+)= scale * sum_j G(i,j) * X[j][f];  G(i,j) = transposed ? Gm[j][i] : Gm[i][j]
__global__ void small_gemm_nn_kernel(const float* __restrict__ Gm, int transposed, const float* __restrict__ X, int n,
                                     int m, int D, float scale, float* __restrict__ out, int accumulate) {
  const int i = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= D) return;
  float a = 0.f;
  for (int j = 0; j < m; ++j) {
    const float g = transposed ? Gm[(long long)j * n + i] : Gm[(long long)i * m + j];
    a += g * X[(long long)j * D + f];
  }
  a *= scale;
  if (accumulate) out[(long long)i * D + f] += a; else out[(long long)i * D + f] = a;
}

// loss = mean_i CE(row i, label i) + mean_j CE(col j, label j)  (tf_cross_entropy_loss_with_logits with one-hot
// identity labels, losses.py:47-51 as used at attention_lib.py:66-74 and :173-181). One block; n <= 1024.
// dlogits = weight * d loss / d logits.
__global__ void ce_sym_kernel(const float* __restrict__ logits, int n, float weight, float* __restrict__ loss_out,
                              float* __restrict__ dlogits) {
  extern __shared__ float sm[];  // row_lse[n], col_lse[n], red[32]
  float* row_lse = sm;
  float* col_lse = sm + n;
  float* red = sm + 2 * n;
  for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
    const bool is_col = t >= n;
    const int k = is_col ? t - n : t;
    float mx = -3.0e38f;
    for (int q = 0; q < n; ++q) mx = fmaxf(mx, is_col ? logits[(long long)q * n + k] : logits[(long long)k * n + q]);
    float s = 0.f;
    for (int q = 0; q < n; ++q) s += __expf((is_col ? logits[(long long)q * n + k] : logits[(long long)k * n + q]) - mx);
    sm[t] = mx + logf(s);
  }
  __syncthreads();
  float part = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x)
    part += (row_lse[k] - logits[(long long)k * n + k]) + (col_lse[k] - logits[(long long)k * n + k]);
  part = block_sum(part, red);
  if (threadIdx.x == 0) *loss_out = part / (float)n;
  if (dlogits) {
    const float wn = weight / (float)n;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int a = t / n, b = t - a * n;
      const float l = logits[t];
      const float d = (a == b) ? 2.f : 0.f;
      dlogits[t] = wn * (__expf(l - row_lse[a]) + __expf(l - col_lse[b]) - d);
    }
  }
}

// losses.tf_cross_entropy_loss_with_logits (losses.py:47-51) for arbitrary labels: out[r] = -sum_k labels[r][k] *
// log_softmax(logits[r])[k]. One warp per row, max-subtracted, warp-shuffle reductions.
__global__ void softmax_xent_kernel(const float* __restrict__ labels, const float* __restrict__ logits, long long rows,
                                    int n, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* lg = logits + r * n;
  const float* lb = labels + r * n;
  float mx = -3.0e38f;
  for (int k = lane; k < n; k += 32) mx = fmaxf(mx, lg[k]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int k = lane; k < n; k += 32) s += expf(lg[k] - mx);
  s = warp_sum(s);
  const float lse = mx + logf(s);
  float acc = 0.f;
  for (int k = lane; k < n; k += 32) acc += lb[k] * (lg[k] - lse);
  acc = warp_sum(acc);
  if (lane == 0) out[r] = -acc;
}

// get_statistics (attention_lib.py:36-43) of a square logit matrix with identity labels, both directions at once:
// out[0] = accuracy = 0.5 * (mean_i [argmax_j logits[i][j] == i] + mean_j [argmax_i logits[i][j] == j])  (first maximum
// wins, as jnp.argmax), out[1] = entropy = 0.5 * (mean row entropy + mean column entropy) with -sum p log(p + 1e-8).
__global__ void ce_stats_kernel(const float* __restrict__ logits, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float hits = 0.f, ent = 0.f;
  for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
    const bool is_col = t >= n;
    const int k = is_col ? t - n : t;
    float mx = -3.0e38f;
    int am = 0;
    for (int q = 0; q < n; ++q) {
      const float l = is_col ? logits[(long long)q * n + k] : logits[(long long)k * n + q];
      if (l > mx) { mx = l; am = q; }
    }
    float s = 0.f;
    for (int q = 0; q < n; ++q) s += expf((is_col ? logits[(long long)q * n + k] : logits[(long long)k * n + q]) - mx);
    const float inv = 1.f / s;
    float e = 0.f;
    for (int q = 0; q < n; ++q) {
      const float pr = expf((is_col ? logits[(long long)q * n + k] : logits[(long long)k * n + q]) - mx) * inv;
      e -= pr * logf(pr + 1e-8f);
    }
    hits += (am == k) ? 1.f : 0.f;
    ent += e;
  }
  hits = block_sum(hits, red);
  ent = block_sum(ent, red);
  if (threadIdx.x == 0) {
    out[0] = 0.5f * hits / (float)n;
    out[1] = 0.5f * ent / (float)n;
  }
}

// hinge (losses.py:30-35): d = mean(relu(1-real) + relu(1+fake)), g = -mean(fake). logit = [real(B); fake(B)].
__global__ void hinge_kernel(const float* __restrict__ logit, int B, float* __restrict__ d_loss,
                             float* __restrict__ g_loss, float* __restrict__ dlogit_d, float* __restrict__ dlogit_g) {
  __shared__ float red[32];
  float d = 0.f, g = 0.f;
  const float invB = 1.f / (float)B;
  for (int k = threadIdx.x; k < B; k += blockDim.x) {
    const float r = logit[k], f = logit[B + k];
    d += fmaxf(1.f - r, 0.f) + fmaxf(1.f + f, 0.f);
    g -= f;
    if (dlogit_d) {
      dlogit_d[k] = (1.f - r > 0.f) ? -invB : 0.f;
      dlogit_d[B + k] = (1.f + f > 0.f) ? invB : 0.f;
    }
    if (dlogit_g) {
      dlogit_g[k] = 0.f;
      dlogit_g[B + k] = -invB;
    }
  }
  d = block_sum(d, red);
  g = block_sum(g, red);
  if (threadIdx.x == 0) {
    *d_loss = d * invB;
    *g_loss = g * invB;
  }
}

// out[n] = <xpool[n], w1*inv_sigma + emb[n % B]> + b1      (xmc_net.py:99-104; w1 = SpectralDense(1) kernel)
__global__ void proj_logit_kernel(const float* __restrict__ xpool, const float* __restrict__ w1,
                                  const float* __restrict__ inv_sigma, const float* __restrict__ b1,
                                  const float* __restrict__ emb, int N2, int B, int C, float* __restrict__ out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N2) return;
  const float inv = inv_sigma ? *inv_sigma : 1.f;
  float a = 0.f;
  for (int c = lane; c < C; c += 32) a += xpool[(long long)n * C + c] * (w1[c] * inv + emb[(long long)(n % B) * C + c]);
  a = warp_sum(a);
  if (lane == 0) out[n] = a + b1[0];
}

// backward of proj_logit for images [n0, n0+cnt): dxpool (+)=, dw1~ +=, db1 +=, demb +=  (grads optional)
__global__ void proj_logit_bwd_kernel(const float* __restrict__ dlogit, const float* __restrict__ xpool,
                                      const float* __restrict__ w1, const float* __restrict__ inv_sigma,
                                      const float* __restrict__ emb, int n0, int cnt, int B, int C,
                                      float* __restrict__ dxpool, int accumulate_x, float* __restrict__ dw1,
                                      float* __restrict__ db1, float* __restrict__ demb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const float inv = inv_sigma ? *inv_sigma : 1.f;
  if (c < C) {
    float gw = 0.f;
    for (int k = 0; k < cnt; ++k) {
      const int n = n0 + k;
      const float dl = dlogit[n];
      const float xp = xpool[(long long)n * C + c];
      const float v = dl * (w1[c] * inv + emb[(long long)(n % B) * C + c]);
      if (accumulate_x) dxpool[(long long)n * C + c] += v; else dxpool[(long long)n * C + c] = v;
      gw += dl * xp;
      if (demb) demb[(long long)(n % B) * C + c] += dl * xp;
    }
    if (dw1) dw1[c] += gw;
  }
  if (db1 && blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < cnt; ++k) s += dlogit[n0 + k];
    db1[0] += s;
  }
}

// fp32 column sums: out[c] += sum_r x[r][c]
__global__ void colsum_f32_kernel(const float* __restrict__ x, int rows, int cols, int ld, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[(long long)r * ld + c];
  out[c] += s;
}

// tanh output head (xmc_net.py:246-247): img = (tanh(x)+1)/2 ; writes fp32 image and a bf16 copy
__global__ void axpy_f32_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_small_gemm_nt(const float* A, const float* B, int n, int m, int D, float scale, float* C,
                                 void* stream) {
  if (!A || !B || !C || n < 1 || m < 1 || D < 1) return XMC_EINVAL;
  small_gemm_nt_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(A, B, n, m, D, scale, C);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_small_gemm_nn(const float* G, int transposed, const float* X, int n, int m, int D, float scale,
                                 float* out, int accumulate, void* stream) {
  if (!G || !X || !out || n < 1 || m < 1 || D < 1) return XMC_EINVAL;
  small_gemm_nn_kernel<<<dim3(ceil_div(D, 128), n), 128, 0, (cudaStream_t)stream>>>(G, transposed, X, n, m, D, scale,
                                                                                   out, accumulate);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_ce_sym(const float* logits, int n, float weight, float* loss_out, float* dlogits, void* stream) {
  if (!logits || !loss_out || n < 1 || n > 2048) return XMC_EINVAL;
  const size_t smem = (size_t)(2 * n + 32) * sizeof(float);
  ce_sym_kernel<<<1, 512, smem, (cudaStream_t)stream>>>(logits, n, weight, loss_out, dlogits);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_softmax_xent(const float* labels, const float* logits, long long rows, int n, float* out,
                                void* stream) {
  if (!labels || !logits || !out || rows < 1 || n < 1) return XMC_EINVAL;
  softmax_xent_kernel<<<(unsigned)ceil_div_ll(rows, 8), 256, 0, (cudaStream_t)stream>>>(labels, logits, rows, n, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_ce_stats(const float* logits, int n, float* out, void* stream) {
  if (!logits || !out || n < 1) return XMC_EINVAL;
  ce_stats_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(logits, n, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_hinge(const float* logit, int B, float* d_loss, float* g_loss, float* dlogit_d, float* dlogit_g,
                         void* stream) {
  if (!logit || !d_loss || !g_loss || B < 1) return XMC_EINVAL;
  hinge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logit, B, d_loss, g_loss, dlogit_d, dlogit_g);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_proj_logit(const float* xpool, const float* w1, const float* inv_sigma, const float* b1,
                              const float* emb, int N2, int B, int C, float* out, void* stream) {
  if (!xpool || !w1 || !b1 || !emb || !out || N2 < 1 || B < 1 || C < 1) return XMC_EINVAL;
  proj_logit_kernel<<<ceil_div(N2, 8), 256, 0, (cudaStream_t)stream>>>(xpool, w1, inv_sigma, b1, emb, N2, B, C, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_proj_logit_bwd(const float* dlogit, const float* xpool, const float* w1, const float* inv_sigma,
                                  const float* emb, int n0, int cnt, int B, int C, float* dxpool, int accumulate_x,
                                  float* dw1, float* db1, float* demb, void* stream) {
  if (!dlogit || !xpool || !w1 || !emb || !dxpool || cnt < 1 || B < 1 || C < 1) return XMC_EINVAL;
  proj_logit_bwd_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(dlogit, xpool, w1, inv_sigma, emb, n0, cnt,
                                                                           B, C, dxpool, accumulate_x, dw1, db1, demb);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_colsum_f32(const float* x, int rows, int cols, int ld, float* out, void* stream) {
  if (!x || !out || rows < 1 || cols < 1) return XMC_EINVAL;
  colsum_f32_kernel<<<ceil_div(cols, 128), 128, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, out);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

extern "C" int xmc_axpy_f32(float* y, const float* x, float a, long long n, void* stream) {
  if (!x || !y || n < 1) return XMC_EINVAL;
  long long blocks = ceil_div_ll(n, 256);
  if (blocks > 4096) blocks = 4096;
  axpy_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, x, a, n);
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}
