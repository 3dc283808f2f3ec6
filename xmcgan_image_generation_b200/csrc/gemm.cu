// tcgen05 implicit-GEMM kernels (sm_100a): forward/dgrad convolution + dense + batched GEMM (K-major operands),
// and weight-gradient / A^T B GEMM (pixel-major operands, split-K).
//
// Structure of both kernels: one CTA per SM, 6 warps:
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer    (one elected lane issues tcgen05.mma, tcgen05.commit releases smem / signals epilogue)
//   warps 2..   : epilogue      (tcgen05.ld TMEM -> registers -> fused bias/mask/residual/relu -> global)
// SAME padding, image borders, ragged M/N/K edges and channel counts that are not a multiple of 64 are all handled
// by TMA out-of-bounds zero fill; nothing is ever im2col'ed in memory.
#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include "common.h"
#include "ptx.cuh"

namespace xmc {

constexpr int kStages = 4;
constexpr int kABytes = 16384;  // 128 rows x 64 bf16
constexpr int kBBytes = 32768;  // up to 256 rows x 64 bf16
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kBarBytes = 256;
constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + slack for 1024B alignment
constexpr int kThreads = 192;     // wgrad kernel: TMA, MMA, 4 epilogue warps
constexpr int kWgradStagesMax = 8; // wgrad kernel: ring depth for small stages (barrier slots)
constexpr int kWgradRing = 224 * 1024;  // wgrad kernel: bytes of the operand ring (stage size and depth per launch)
constexpr int kSmemWgrad = kWgradRing + kBarBytes + 1024;
constexpr int kEpiParts = 3;      // forward kernel: epilogue warps per TMEM lane quarter
constexpr int kThreadsFwd = 64 + 128 * kEpiParts;  // TMA warp, MMA warp, 4 * kEpiParts epilogue warps
constexpr int kTmemCols = 512;
constexpr int kBiasMax = 2048;    // forward kernel: floats of bias staged in shared memory
constexpr int kSmemFwd = kSmemBytes + kBiasMax * 4;
// resident-weights 3x3 kernel: all 9 taps of a [Cout <= 96][C <= 96] kernel + a 3-stage ring of 130-pixel halo rows
constexpr int kResBBytes = 9 * (96 * 128 + 96 * 64);  // 165888
constexpr int kResAStage = 17408;                     // 130 rows x 128 B, rounded up to the 1024 B swizzle repeat
constexpr int kResStages = 3;
constexpr int kSmemRes = kResBBytes + kResStages * kResAStage + kBarBytes + kBiasMax * 4 + 1024;

struct FwdParams {
  int tiles_w, tiles_h, tiles_n;
  int tw, th, tn;
  int n_tiles, BN;
  int KH, KW, pad_h, pad_w, C, cchunks;
  int last_mmas;  // K=16 MMA slices of the last 64-channel chunk that hold real channels
  int N, H, W, Cout;
  int batched;
  int strideH, strideW;
  int parities;  // 1, or 4: sub-pixel mode (nearest-2x-upsample fused into a 3x3 conv as four 2x2 convs)
  uint32_t stage_tx_bytes;
  int debug;   // profiling aid (XMC_FWD_DEBUG): 1 = no MMAs (fetch pipeline alone), 2 = no loads (MMA + epilogue alone)
  uint32_t idesc;
  void* out;
  int out_dtype, ldOut;
  const float* bias;
  const void* residual;  // bf16, or fp32 in the kF32 instantiation
  const void* mask;
  int ldRes, ldMask, res_shift, relu, mask_last;
  float alpha;
  int vec_ok;
  int bias_smem;  // 1: the bias vector (Cout <= kBiasMax floats) is staged in shared memory by the epilogue warps
  // fp32-activation mode with the two-part operand (act_f32 = 2): the A tensor is bf16 [.., hi(pairC) | lo(pairC)] and
  // the K chunks of a tap run hi, lo, hi again (the weights stay [hi | hi | lo], 3 * pairC per tap)
  int pairC;
  bf16* out_pair;  // optional second output of the fp32 epilogue: the result as bf16 [.., hi(Cout) | lo(Cout)]
  int ldPair;
  // fp32 epilogue reading its residual / mask from two-part bf16 tensors [.., hi(Cout) | lo(Cout)] (pitch ldRes / ldMask
  // in bf16 elements): residual = hi + lo, mask = sign of hi. `out` may then be null (pair output only).
  int res_pair, mask_pair;
};

struct WgradParams {
  int chunks_w, chunks_h, chunks_n;  // pixel-chunk grid (K dimension)
  int tw, th, tn;
  int total_chunks, chunks_per_split, ksplit;
  int m_tiles, n_tiles, BN, nslabs;
  int taps, KW, pad_h, pad_w;
  int Ca, Cb;
  int batched;
  int items_per_split, total_items;
  int subpixel;         // 1: 16 parity taps (a,dh,b,dw); B is the 2x up-sampled gradient read with stride 2
                        // 2: pool-fused: 16 taps (r,s) of the 4x4/stride-2 form; A is the 2x larger input read with stride 2
  int tap3;             // 1: items are filter rows; the three kw taps share one 66-pixel halo chunk of A (3 accumulators)
  int tg;               // taps per item that share the B chunk (one A chunk and one accumulator each); 1 = one tap
  int groups;           // items per (batch, output tile): tap groups (tg > 1), filter rows (tap3) or taps
  int a_slabs;          // 64-channel slabs of A loaded per tap (1 when Ca <= 64)
  int stages;           // depth of the shared-memory ring (192 KB / stage_bytes, at most kWgradStagesMax)
  int debug;            // profiling aid (XMC_WGRAD_DEBUG): 1 = no MMAs (fetch pipeline alone), 2 = no loads (MMA pipeline alone)
  int merge_a, merge_b; // 1: the 64-channel slabs of a tap's A chunk / of the B chunk arrive as ONE 5-D TMA box
  uint32_t stage_bytes;
  uint32_t slab_bytes;  // bytes one TMA box writes
  uint32_t idesc;
  void* out;
  int out_mode, ldOut;
  long long out_tap_stride, out_batch_stride;
  float alpha;
  int vec_ok;
  // deterministic accumulation (out_mode 0). ws == nullptr: every output element has exactly one owner item, which
  // adds its tile with a plain read-modify-write. ws != nullptr (K split across CTAs, or sub-pixel parity taps that
  // share destinations): every item stores its partial tile into ws[split][batch][source tap][Ca][Cb] and
  // wgrad_reduce_kernel adds the partials to the destination in a fixed order.
  float* ws;
  long long ws_split_stride;  // floats between consecutive K splits
  int ws_taps;                // source-tap slots per batch entry
};

__device__ __forceinline__ uint8_t* align_1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t h) { return __uint_as_float(h << 16); }

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256); addresses must be 32-byte aligned
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Lean epilogue for the commonest short-K case, compiled without any of the generic machinery (no mask, alpha == 1,
// bf16 output, 32-byte aligned pitches, Cout % 16 == 0, bias staged in shared memory): kRes adds the residual.
template <bool kRes, int kParts = kEpiParts>
__device__ __forceinline__ void fwd_epilogue_tile_lean(const FwdParams& p, uint32_t t_addr, int nt, bool row_ok,
                                                       long long pix, long long rpix, const float* bias_s, int half) {
  const int col0 = nt * p.BN;
  const int ncols = p.Cout - col0;
  const float* b_row = bias_s + col0;
  bf16* o_row = reinterpret_cast<bf16*>(p.out) + pix * p.ldOut + col0;
  const bf16* r_row = kRes ? reinterpret_cast<const bf16*>(p.residual) + rpix * p.ldRes + col0 : nullptr;
  const bool relu = p.relu != 0;
  for (int c0 = half * 16; c0 < p.BN; c0 += 16 * kParts) {
    uint32_t v[16];
    tmem_ld16(t_addr + c0, v);
    uint4 r0, r1;
    const bool live = row_ok && c0 < ncols;
    if (kRes && live) ldg256(r_row + c0, r0, r1);
    tmem_ld_wait();
    if (live) {
      float f[16];
      const float4* b4 = reinterpret_cast<const float4*>(b_row + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 bz = b4[i];
        f[4 * i] = __uint_as_float(v[4 * i]) + bz.x;
        f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bz.y;
        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bz.z;
        f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bz.w;
      }
      if (kRes) {
        const uint32_t rw_[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[2 * i] += bf16_bits_to_float(rw_[i] & 0xFFFFu);
          f[2 * i + 1] += bf16_bits_to_float(rw_[i] >> 16);
        }
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      uint4 a, b;
      a.x = pack_bf16x2(f[0], f[1]);   a.y = pack_bf16x2(f[2], f[3]);
      a.z = pack_bf16x2(f[4], f[5]);   a.w = pack_bf16x2(f[6], f[7]);
      b.x = pack_bf16x2(f[8], f[9]);   b.y = pack_bf16x2(f[10], f[11]);
      b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
      stg256(o_row + c0, a, b);
    }
  }
}

// Epilogue of one 128 x BN output tile for one warp: t_addr = TMEM address of the warp's 32 lanes in the tile's
// accumulator, (row_ok, pix, rpix) = this lane's output row, `part` = which of the kEpiParts warps sharing the lane
// quarter this is. Shared by the streaming and the resident-weights forward kernels.
// kF32: the fp32-activation instantiation (config.dtype = "float32" / the frozen ResNet branch): residual and mask are
// fp32 tensors (16 columns = 64 bytes = two 32-byte accesses each), the output is fp32.
template <bool kF32, int kParts = kEpiParts>
__device__ __forceinline__ void fwd_epilogue_tile(const FwdParams& p, uint32_t t_addr, int nt, bool row_ok,
                                                  long long pix, long long rpix, const float* bias_s, int half) {
  constexpr int kV = kF32 ? 4 : 2;  // uint4 vectors per 16-column chunk of a mask / residual row
  // One 16-column chunk = issue (TMEM load + the chunk's mask / residual vectors) ... finish (fused math, store).
  // The chunks of a warp are software-pipelined over two register sets: chunk i+1 is issued before chunk i is
  // finished, so its TMEM / global latency hides behind the math and the stores of chunk i.
  auto load16 = [&](const void* base, long long elem, uint4 (&dst)[kV]) {
    const char* ptr = reinterpret_cast<const char*>(base) + elem * (kF32 ? 4 : 2);
    if (p.vec_ok == 2) {
#pragma unroll
      for (int i = 0; i < kV; i += 2) ldg256(ptr + 16 * i, dst[i], dst[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < kV; ++i) dst[i] = __ldg(reinterpret_cast<const uint4*>(ptr) + i);
    }
  };
  // element i (0..15) of a loaded row chunk as float
  auto elem = [&](const uint4 (&src)[kV], int i) -> float {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src);
    if (kF32) return __uint_as_float(w[i]);
    return (i & 1) ? bf16_bits_to_float(w[i >> 1] >> 16) : bf16_bits_to_float(w[i >> 1] & 0xFFFFu);
  };
  auto issue = [&](int c0, uint32_t (&v)[16], uint4 (&mk)[kV], uint4 (&rs)[kV]) {
    tmem_ld16(t_addr + c0, v);  // asynchronous until tmem_ld_wait
    const int col = nt * p.BN + c0;
    if (row_ok && p.vec_ok && p.Cout - col >= 16) {
      if (kF32 && p.mask && p.mask_pair) {   // 16 bf16 hi values = 32 bytes
        ldg256(reinterpret_cast<const bf16*>(p.mask) + pix * p.ldMask + col, mk[0], mk[1]);
      } else if (p.mask) {
        load16(p.mask, pix * p.ldMask + col, mk);
      }
      if (kF32 && p.residual && p.res_pair) {  // hi and lo parts, 32 bytes each
        const bf16* rp = reinterpret_cast<const bf16*>(p.residual) + rpix * p.ldRes + col;
        ldg256(rp, rs[0], rs[1]);
        ldg256(rp + p.Cout, rs[kV - 2], rs[kV - 1]);
      } else if (p.residual) {
        load16(p.residual, rpix * p.ldRes + col, rs);
      }
    }
  };
  // residual element i: fp32 / bf16 tensor, or hi + lo of a two-part tensor (fp32 instantiation only)
  auto res_elem = [&](const uint4 (&src)[kV], int i) -> float {
    if (kF32 && p.res_pair) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(src);
      const uint32_t h = w[i >> 1], l = w[8 + (i >> 1)];
      return (i & 1) ? bf16_bits_to_float(h >> 16) + bf16_bits_to_float(l >> 16)
                     : bf16_bits_to_float(h & 0xFFFFu) + bf16_bits_to_float(l & 0xFFFFu);
    }
    return elem(src, i);
  };
  auto finish = [&](int c0, const uint32_t (&v)[16], const uint4 (&mk)[kV], const uint4 (&rs)[kV]) {
    const int col = nt * p.BN + c0;
    const bool active = row_ok && col < p.Cout;
    const int nvalid = min(16, p.Cout - col);
    if (active && p.vec_ok && nvalid == 16) {
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
      if (p.bias) {
        // bias_s: the whole bias vector staged in shared memory once per CTA (an LDG per chunk was the single
        // largest stall of the short-K epilogue); vectors too long for the staging area come from global memory
        const float4* b4 = p.bias_smem ? reinterpret_cast<const float4*>(bias_s + col)
                                       : reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bz = b4[i];
          f[4 * i] += bz.x; f[4 * i + 1] += bz.y; f[4 * i + 2] += bz.z; f[4 * i + 3] += bz.w;
        }
      }
      if (p.residual && p.mask_last) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += res_elem(rs, i);
      }
      if (p.mask) {
        if (kF32 && !p.mask_pair) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (!(elem(mk, i) > 0.f)) f[i] = 0.f;
        } else {
          const uint32_t* mw = reinterpret_cast<const uint32_t*>(mk);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // bf16 > 0  <=>  sign bit clear and magnitude non-zero
            const uint32_t lo = mw[i] & 0xFFFFu, hi = mw[i] >> 16;
            if (!(lo != 0 && lo < 0x8000u)) f[2 * i] = 0.f;
            if (!(hi != 0 && hi < 0x8000u)) f[2 * i + 1] = 0.f;
          }
        }
      }
      if (p.residual && !p.mask_last) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += res_elem(rs, i);
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      if (kF32 && p.out_pair) {
        // the result once more as the next GEMM's operand: hi = bf16(v), lo = bf16(v - hi), 16 columns = 32 bytes each
        bf16* o = p.out_pair + pix * p.ldPair + col;
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          h[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
          l[i] = pack_bf16x2(f[2 * i] - bf16_bits_to_float(h[i] & 0xFFFFu), f[2 * i + 1] - bf16_bits_to_float(h[i] >> 16));
        }
        stg256(o, make_uint4(h[0], h[1], h[2], h[3]), make_uint4(h[4], h[5], h[6], h[7]));
        stg256(o + p.Cout, make_uint4(l[0], l[1], l[2], l[3]), make_uint4(l[4], l[5], l[6], l[7]));
      }
      if (!p.out) {
        // pair output only
      } else if (p.out_dtype == 0) {
        bf16* o = reinterpret_cast<bf16*>(p.out) + pix * p.ldOut + col;
        uint4 a, b;
        a.x = pack_bf16x2(f[0], f[1]);   a.y = pack_bf16x2(f[2], f[3]);
        a.z = pack_bf16x2(f[4], f[5]);   a.w = pack_bf16x2(f[6], f[7]);
        b.x = pack_bf16x2(f[8], f[9]);   b.y = pack_bf16x2(f[10], f[11]);
        b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
        if (p.vec_ok == 2) {
          stg256(o, a, b);
        } else {
          reinterpret_cast<uint4*>(o)[0] = a;
          reinterpret_cast<uint4*>(o)[1] = b;
        }
      } else {
        float* o = reinterpret_cast<float*>(p.out) + pix * p.ldOut + col;
        if (p.vec_ok == 2) {
          const uint4* fv = reinterpret_cast<const uint4*>(f);
          stg256(o, fv[0], fv[1]);
          stg256(o + 8, fv[2], fv[3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
      }
    } else if (active) {
      // generic path (ragged N edge or unaligned pitches): scalar accesses
      auto rd = [&](const void* base, long long e) -> float {
        return kF32 ? reinterpret_cast<const float*>(base)[e] : __bfloat162float(reinterpret_cast<const bf16*>(base)[e]);
      };
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < nvalid) {
          if (p.bias) f[i] += p.bias[col + i];
          if (p.residual && p.mask_last) f[i] += rd(p.residual, rpix * p.ldRes + col + i);
          if (p.mask && !(rd(p.mask, pix * p.ldMask + col + i) > 0.f)) f[i] = 0.f;
          if (p.residual && !p.mask_last) f[i] += rd(p.residual, rpix * p.ldRes + col + i);
          if (p.relu) f[i] = fmaxf(f[i], 0.f);
          if (p.out_dtype == 0) reinterpret_cast<bf16*>(p.out)[pix * p.ldOut + col + i] = __float2bfloat16(f[i]);
          else reinterpret_cast<float*>(p.out)[pix * p.ldOut + col + i] = f[i];
        }
      }
    }
  };
  constexpr int kStep = 16 * kParts;
  uint32_t va[16], vb[16];
  uint4 mka[kV], mkb[kV], rsa[kV], rsb[kV];
  int c0 = half * 16;
  if (c0 < p.BN) issue(c0, va, mka, rsa);
  while (c0 < p.BN) {  // all conditions are warp-uniform (tcgen05.ld / wait::ld are .sync.aligned)
    tmem_ld_wait();
    if (c0 + kStep < p.BN) issue(c0 + kStep, vb, mkb, rsb);
    finish(c0, va, mka, rsa);
    c0 += kStep;
    if (c0 >= p.BN) break;
    tmem_ld_wait();
    if (c0 + kStep < p.BN) issue(c0 + kStep, va, mka, rsa);
    finish(c0, vb, mkb, rsb);
    c0 += kStep;
  }
}

// =====================================================================================================================
// Forward / dgrad / dense / batched GEMM:  D[pixels, Cout] = sum_taps A_tap[pixels, C] * B[Cout, tap*C + c]
// =====================================================================================================================
// Epilogue warps per TMEM lane quarter of an instantiation. The fp32 epilogue (kMode 3) keeps twice the residual / mask
// registers in flight; at 14 warps the register file caps a thread at 128 registers and it spilled its loop state to
// local memory (the reloads were the top stall of the ResNet branch's 1x1 layers) — 10 warps leave it 200.
template <int kMode> struct FwdParts { static constexpr int v = kMode == 3 ? 2 : kEpiParts; };

template <int kMode>  // 0: generic epilogue; 1 / 2: lean epilogue (bias [+ residual]); 3: fp32 residual / mask / output
__global__ void __launch_bounds__(64 + 128 * FwdParts<kMode>::v, 1)
gemm_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FwdParams p) {
  constexpr int kParts = FwdParts<kMode>::v;
  constexpr int kThreadsK = 64 + 128 * kParts;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + kStages * kStageBytes;
  // barrier layout: full[kStages], empty[kStages], tfull[2], tempty[2], then tmem pointer
  // (a deeper ring of smaller stages for narrow tiles was measured: no gain — unlike the weight-gradient kernel, this
  // loop is not bound by the bytes in flight)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kStages * kStageBytes + 8 * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();   // programmatic dependent launch: the prologue below may overlap the previous kernel's tail
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * kParts);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();                // ... nothing in global memory is touched before the previous kernel has completed

  float* bias_s = reinterpret_cast<float*>(smem + kStages * kStageBytes + kBarBytes);
  if (p.bias_smem && warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.Cout; i += kThreadsK - 64) bias_s[i] = p.bias[i];
    asm volatile("bar.sync 1, %0;" ::"n"(kThreadsK - 64) : "memory");  // epilogue warps only
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles * p.parities;
  const int taps = p.KH * p.KW;
  const int kiters = taps * p.cchunks;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      // tile -> (m tile = (in, ih, iw), parity, n tile) as a mixed-radix counter advanced by gridDim.x per step (digit
      // steps computed once: no divisions per tile). Parity (a,b) of sub-pixel mode: output pixel (2h+a, 2w+b) reads
      // the 2x2 input window starting at (h-(1-a), w-(1-b)) and the parity's own summed weights (rows par*Cout.. of B)
      const int radix[4] = {p.n_tiles, p.parities, p.tiles_w, p.tiles_h};
      int dig[5], stp[5];
      {
        int t = blockIdx.x, g = gridDim.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dig[i] = t % radix[i]; t /= radix[i];
          stp[i] = g % radix[i]; g /= radix[i];
        }
        dig[4] = t; stp[4] = g;
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = dig[0], par = dig[1], iw = dig[2], ih = dig[3], in = dig[4];
        {
          int carry = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int v = dig[i] + stp[i] + carry;
            carry = v >= radix[i] ? 1 : 0;
            dig[i] = v - (carry ? radix[i] : 0);
          }
          dig[4] += stp[4] + carry;
        }
        const int w0 = iw * p.tw, h0 = ih * p.th, n0 = in * p.tn;
        const int pad_h = p.pad_h - (par >> 1), pad_w = p.pad_w - (par & 1);
        // no divisions on the per-stage path (this thread's loop is on the critical path of the short-K layers)
        const int per = p.pairC >> 6;   // two-part operand: 64-channel chunks per part
        int kh = 0, kw = 0;
        for (int tap = 0; tap < taps; ++tap) {
          int ca = 0, in_part = 0;
          for (int c = 0; c < p.cchunks; ++c) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_arrive_expect_tx(full_bar(stage), p.debug == 2 ? 0u : p.stage_tx_bytes);
            const uint32_t sa = sbase + stage * kStageBytes;
            if (p.debug != 2) {
              tma_load_4d(sa, &tmA, full_bar(stage), ca, w0 * p.strideW + kw - pad_w,
                          h0 * p.strideH + kh - pad_h, n0);
              tma_load_3d(sa + kABytes, &tmB, full_bar(stage), tap * p.C + c * 64, par * p.Cout + nt * p.BN,
                          p.batched ? n0 : 0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            // next chunk's channel offset; two-part operand: chunks walk [hi | lo | hi] of the stored [hi | lo]
            ca += 64;
            if (p.pairC && ++in_part == per) {
              in_part = 0;
              ca = (ca == p.pairC) ? p.pairC : 0;   // after hi -> lo (offset pairC), after lo -> hi again (offset 0)
            }
          }
          if (++kw == p.KW) { kw = 0; ++kh; }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      int c = 0;
      for (int k = 0; k < kiters; ++k) {
        // the last 64-channel chunk of a tap may be ragged (C = 96: 64 + 32): only the K=16 slices that hold real
        // channels are issued, the TMA zero fill beyond C is never multiplied
        const int nmma = (c == p.cchunks - 1) ? p.last_mmas : 4;
        if (++c == p.cchunks) c = 0;
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = sbase + stage * kStageBytes;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sa + kABytes, 16, 1024);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // advance 16 K-elements = 32 bytes inside the 128B swizzle row: +2 in the (addr>>4) field
            if (j < nmma && !(p.debug == 1 && k > 0))
              umma_bf16(d_tmem, adesc + 2 * j, bdesc + 2 * j, p.idesc, (k > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (k == kiters - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // 4 * kEpiParts epilogue warps: warp w may only touch TMEM lanes 32*(w%4)..+31; the kEpiParts warps that share a
    // lane quarter take the 16-column chunks of the tile round-robin. A chunk is a dependent chain of ~200
    // instructions, so short-K tiles (1x1 convolutions) are bound by the epilogue's latency: more warps, not wider
    // accesses, is what shortens it (measured: grouping loads or staging through shared memory made it slower).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    // this lane's row of every tile: (rw, rh, rn) inside the tw x th x tn pixel box. Hoisted out of the tile loop,
    // and the tile decode skips the divisions of the common n_tiles == 1 / parities == 1 cases: on short-K tiles the
    // epilogue is the critical path and a dozen emulated integer divisions per tile are a visible part of it.
    const int r = q * 32 + lane;
    const int rw = r % p.tw, rh = (r / p.tw) % p.th, rn = r / (p.tw * p.th);
    const int tiles_wh = p.tiles_w * p.tiles_h;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int nt = 0, t2 = tile;
      if (p.n_tiles > 1) { nt = tile % p.n_tiles; t2 = tile / p.n_tiles; }
      int par = 0, mt = t2;
      if (p.parities > 1) { par = t2 & 3; mt = t2 >> 2; }
      const int in = mt / tiles_wh;
      const int rem = mt - in * tiles_wh;
      const int ih = rem / p.tiles_w;
      const int iw = rem - ih * p.tiles_w;
      const int w = iw * p.tw + rw, h = ih * p.th + rh, n = in * p.tn + rn;
      const bool row_ok = (rn < p.tn) && (n < p.N) && (h < p.H) && (w < p.W);
      // sub-pixel mode scatters the tile into the 2x up-sampled output at (2h+a, 2w+b)
      const long long pix = p.parities == 1
                                ? ((long long)n * p.H + h) * p.W + w
                                : ((long long)n * 2 * p.H + 2 * h + (par >> 1)) * (2 * p.W) + 2 * w + (par & 1);
      long long rpix = 0;
      if (p.residual) {
        const int Hs = p.H >> p.res_shift, Ws = p.W >> p.res_shift;
        rpix = ((long long)n * Hs + (h >> p.res_shift)) * Ws + (w >> p.res_shift);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      if (kMode == 1) fwd_epilogue_tile_lean<false, kParts>(p, t_addr, nt, row_ok, pix, rpix, bias_s, half);
      else if (kMode == 2) fwd_epilogue_tile_lean<true, kParts>(p, t_addr, nt, row_ok, pix, rpix, bias_s, half);
      else if (kMode == 3) fwd_epilogue_tile<true, kParts>(p, t_addr, nt, row_ok, pix, rpix, bias_s, half);
      else fwd_epilogue_tile<false, kParts>(p, t_addr, nt, row_ok, pix, rpix, bias_s, half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// Resident-weights 3x3 convolution for the widest feature maps (C <= 96 channels, W a multiple of 128: the 96 -> 96
// layers at 128x128 / 256x256). In the streaming kernel those layers are bound by L2 -> SM bandwidth, not by the
// tensor pipe: every 128-pixel tile re-fetches all 9 weight taps (216 KB) and the same input row once per kw tap
// (216 KB) for 2.6k cycles of MMA. Here
//   * the whole weight tensor (9 taps x [Cout][64 + 32 channels], 162 KB at Cout = 96) is loaded ONCE per CTA and stays
//     in shared memory; the ragged second channel chunk uses a 32-channel box in a SWIZZLE_64B layout so that it costs
//     half the space (the A operand keeps its 128-byte rows: the two descriptors of an MMA are independent);
//   * one 130-pixel halo row of the input per (kh, channel chunk) serves all three kw taps: the A descriptor of tap kw
//     simply starts kw rows (kw * 128 bytes) into the tile.
// L2 -> SM traffic per 128 output pixels drops from 432 KB to 50 KB and the TMA request count from 36 to 4; the
// epilogue is the streaming kernel's. Measured (112 x 128 x 128, 96 -> 96): streaming 500 us, resident weights alone
// 567 us (only 3 ring stages fit beside them), + halo rows 370 us, + two output rows per tile: see DESIGN.md.
// =====================================================================================================================
__global__ void __launch_bounds__(kThreadsFwd, 1)
conv3x3_resident_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmB2, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sA = sbase + kResBBytes;
  const uint32_t bar_base = sA + kResStages * kResAStage;
  // barriers: full[kResStages], empty[kResStages], bfull, tfull[2], tempty[2], then the tmem pointer
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kResStages + s); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * kResStages);
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kResStages + 1 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kResStages + 3 + a); };
  uint8_t* bar_ptr = smem + kResBBytes + kResStages * kResAStage;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * kResStages + 5));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kResStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(bfull_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * kEpiParts);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  float* bias_s = reinterpret_cast<float*>(bar_ptr + kBarBytes);
  if (p.bias_smem && warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.Cout; i += kThreadsFwd - 64) bias_s[i] = p.bias[i];
    asm volatile("bar.sync 1, %0;" ::"n"(kThreadsFwd - 64) : "memory");
  }

  // one tile = 128 pixels of TWO consecutive image rows (h, h+1): the four input rows h-1 .. h+2 it needs are each
  // loaded once and feed up to two accumulators (row h with kh = ir, row h+1 with kh = ir-1), which doubles the MMA
  // work per TMA box and with it the latency the 3-stage ring can hide
  const int hpairs = p.H >> 1;
  const int total_tiles = p.tiles_w * hpairs * p.N;
  const uint32_t bmain = (uint32_t)p.BN * 128u;    // bytes of one tap's [BN][64-channel] weight box
  const uint32_t btail = (uint32_t)p.BN * 64u;     // bytes of one tap's [BN][32-channel] box (SWIZZLE_64B)
  const uint32_t sBt = sbase + 9u * bmain;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      // resident weights: all taps, both channel chunks, one barrier
      mbar_arrive_expect_tx(bfull_bar, 9u * (bmain + (p.cchunks > 1 ? btail : 0u)));
      for (int tap = 0; tap < 9; ++tap) {
        tma_load_3d(sbase + tap * bmain, &tmB, bfull_bar, tap * p.C, 0, 0);
        if (p.cchunks > 1) tma_load_3d(sBt + tap * btail, &tmB2, bfull_bar, tap * p.C + 64, 0, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int iw = tile % p.tiles_w;
        const int h = 2 * ((tile / p.tiles_w) % hpairs);
        const int n = tile / (p.tiles_w * hpairs);
        const int w0 = iw * 128;
        for (int ir = 0; ir < 4; ++ir) {
          for (int c = 0; c < p.cchunks; ++c) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_arrive_expect_tx(full_bar(stage), 130u * 128u);
            // 130-pixel halo row (w0-1 .. w0+128) of input row h-1+ir; image borders are TMA zero fill = SAME padding
            tma_load_4d(sA + stage * kResAStage, &tmA, full_bar(stage), c * 64, w0 - 1, h - 1 + ir, n);
            if (++stage == kResStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    mbar_wait(bfull_bar, 0);
    tc_fence_after();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      for (int ir = 0; ir < 4; ++ir) {
        for (int c = 0; c < p.cchunks; ++c) {
          const int nmma = (c == p.cchunks - 1) ? p.last_mmas : 4;
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = sA + stage * kResAStage;
#pragma unroll
            for (int r = 0; r < 2; ++r) {       // output row h + r uses this input row as filter row kh = ir - r
              const int kh = ir - r;
              if (kh < 0 || kh > 2) continue;
              const uint32_t d_tmem = tmem_base + acc * 256 + r * 128;
              for (int kw = 0; kw < 3; ++kw) {
                const int tap = kh * 3 + kw;
                // tap kw reads rows kw .. kw+127 of the 130-row tile (row = pixel, 128 B each). The swizzle XOR follows
                // the absolute shared-memory address (what TMA wrote), so a start address kw rows into the 1024-byte
                // repeat needs no base offset (measured on B200: setting it to kw gives wrong results).
                const uint64_t adesc = make_smem_desc_ex(sa + kw * 128u, 16, 1024, 2, 0);
                const uint64_t bdesc = (c == 0) ? make_smem_desc_ex(sbase + tap * bmain, 16, 1024, 2, 0)
                                                : make_smem_desc_ex(sBt + tap * btail, 16, 512, 4, 0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < nmma)
                    umma_bf16(d_tmem, adesc + 2 * j, bdesc + 2 * j, p.idesc, (kh | c | kw | j) ? 1u : 0u);
              }
            }
            umma_commit(empty_bar(stage));
            if (ir == 3 && c == p.cchunks - 1) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == kResStages) { stage = 0; phase ^= 1; }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int iw = tile % p.tiles_w;
      const int h0 = 2 * ((tile / p.tiles_w) % hpairs);
      const int n = tile / (p.tiles_w * hpairs);
      const int w = iw * 128 + q * 32 + lane;
      const bool row_ok = w < p.W;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int r = 0; r < 2; ++r) {
        const int h = h0 + r;
        const long long pix = ((long long)n * p.H + h) * p.W + w;
        long long rpix = 0;
        if (p.residual) {
          const int Hs = p.H >> p.res_shift, Ws = p.W >> p.res_shift;
          rpix = ((long long)n * Hs + (h >> p.res_shift)) * Ws + (w >> p.res_shift);
        }
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + r * 128;
        fwd_epilogue_tile<false>(p, t_addr, 0, row_ok, pix, rpix, bias_s, half);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// Weight gradient:  D[Ca-tile(128), Cb-tile(BN)] = sum_{pixel chunks} A[pixels(+tap shift), ca]^T * B[pixels, cb]
// Both operands arrive pixel-major (64-channel slabs of [pixels][128B] rows) and are consumed as MN-major UMMA tiles.
// Persistent: work item = (K split, batch, tap, output tile); items that run concurrently share the same pixel range
// (HBM reads it once, L2 serves the rest). Two TMEM accumulators: the epilogue (vector reductions to global) of one
// item overlaps the MMAs of the next.
// =====================================================================================================================
__global__ void __launch_bounds__(kThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + kWgradRing;
  constexpr int kSM = kWgradStagesMax;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kSM + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kSM + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kSM + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kWgradRing + 8 * (2 * kSM + 4));
  const int nstages = p.stages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
  // A pixel chunk can hold fewer than 64 pixels (tiny tensors); rows the TMA never writes must read as zero
  // because they are part of the reduction dimension.
  if (p.slab_bytes < 8192) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < kWgradRing / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  // A slab slot: 64 pixels x 128 B, or the 66-pixel halo chunk of tap3 mode rounded up to the 1024 B swizzle repeat
  const uint32_t a_slab = p.tap3 ? 9216u : 8192u;
  const uint32_t acc_cols = (uint32_t)((p.BN + 31) / 32) * 32u;  // tap3: column pitch of the three kw accumulators

  // item -> (split, batch, tap, m tile, n tile)
  auto decode = [&](int item, int& split, int& bz, int& tap, int& mt, int& nt) {
    split = item / p.items_per_split;
    int rem = item - split * p.items_per_split;
    const int tiles = p.m_tiles * p.n_tiles;
    bz = rem / (p.groups * tiles);
    rem -= bz * p.groups * tiles;
    tap = rem / tiles;   // tap, filter row (tap3) or tap group (tg > 1)
    rem -= tap * tiles;
    mt = rem / p.n_tiles;
    nt = rem - mt * p.n_tiles;
  };

  // tap groups (tg > 1): member k of group g. Plain and pool-fused taps: consecutive tap indices; sub-pixel parity
  // taps a<<3 | dh<<2 | b<<1 | dw share B per (a, b): groups of 4 = (a, b) with members (dh, dw), of 2 = (a, dh, b)
  // with members dw.
  auto group_taps = [&](int g) { return min(p.tg, p.taps - g * p.tg); };
  auto member_tap = [&](int g, int k) {
    if (p.subpixel == 1) {
      if (p.tg == 4) return ((g >> 1) << 3) | ((k >> 1) << 2) | ((g & 1) << 1) | (k & 1);
      if (p.tg == 2) return ((g >> 2) << 3) | (((g >> 1) & 1) << 2) | ((g & 1) << 1) | k;
    }
    return g * p.tg + k;
  };
  // A shift (ah, aw), A stride, B start (bh, bw) and B stride of a tap
  auto tap_geom = [&](int tap, int& ah, int& aw, int& as, int& bh, int& bw, int& bs) {
    bh = 0; bw = 0; bs = 1; as = 1;
    if (p.subpixel == 2) {
      // pool-fused conv3x3 -> 2x2 mean: tap = r*4 + s of the equivalent 4x4 / stride-2 / pad-1 convolution:
      // x[2i + r - 1, 2j + s - 1] * dy_low[i, j]
      ah = (tap >> 2) - 1; aw = (tap & 3) - 1; as = 2;
    } else if (p.subpixel) {
      // tap = a<<3 | dh<<2 | b<<1 | dw : x[i+dh-(1-a), j+dw-(1-b)] * dy[2i+a, 2j+b]
      const int a = (tap >> 3) & 1, dh = (tap >> 2) & 1, b = (tap >> 1) & 1, dw = tap & 1;
      ah = dh - (1 - a); aw = dw - (1 - b); bh = a; bw = b; bs = 2;
    } else {
      const int kh = tap / p.KW, kw = tap - kh * p.KW;
      ah = kh - p.pad_h; aw = kw - p.pad_w;
    }
  };
  const uint32_t a_tap_bytes = (uint32_t)p.a_slabs * 8192u;   // tg > 1: one tap's A region inside a stage

  if (warp == 0) {
    // Producer: one elected thread (UTMALDG takes uniform operands: per-lane boxes would be serialised by the compiler
    // anyway). The per-chunk path is kept short — the tap shifts of a group are packed into two registers, the loops
    // over taps / slabs are unrolled with predicates, the chunk coordinates advance incrementally (no divisions).
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx3 = 2u * 66u * 128u + p.nslabs * p.slab_bytes;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int split, bz, tap, mt, nt;
        decode(item, split, bz, tap, mt, nt);
        const int nt_item = p.tg > 1 ? group_taps(tap) : 1;
        // A shifts of the member taps, one signed byte each (+8 bias); A stride; B start and stride
        uint32_t ahp = 0, awp = 0;
        int as = 1, bh = 0, bw = 0, bs = 1;
        for (int k2 = 0; k2 < nt_item; ++k2) {
          int ah, aw;
          tap_geom(p.tg > 1 ? member_tap(tap, k2) : tap, ah, aw, as, bh, bw, bs);
          if (p.tap3) {  // item = filter ROW kh: one 66-pixel halo chunk of A serves the three kw taps
            ah = tap - p.pad_h; aw = -p.pad_w; bh = 0; bw = 0; bs = 1; as = 1;
          }
          ahp |= (uint32_t)(ah + 8) << (8 * k2);
          awp |= (uint32_t)(aw + 8) << (8 * k2);
        }
        const uint32_t tx = p.tap3 ? tx3 : (uint32_t)(nt_item * p.a_slabs + p.nslabs) * p.slab_bytes;
        const uint32_t a_stride = p.tap3 ? a_slab : 8192u;
        const uint32_t b_off = p.tap3 ? 2 * a_slab : p.tg * a_tap_bytes;
        const bool two_slabs = p.tap3 || p.a_slabs == 2;
        const int j0 = split * p.chunks_per_split;
        const int j1 = min(p.total_chunks, j0 + p.chunks_per_split);
        const int m0 = mt * 128, n_off = nt * p.BN;
        int iw = j0 % p.chunks_w, ih = (j0 / p.chunks_w) % p.chunks_h, in = j0 / (p.chunks_w * p.chunks_h);
        for (int j = j0; j < j1; ++j) {
          const int w0 = iw * p.tw, h0 = ih * p.th, n0 = p.batched ? bz : in * p.tn;
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_arrive_expect_tx(full_bar(stage), p.debug == 2 ? 0u : tx);
          const uint32_t sa = sbase + stage * p.stage_bytes;
          const uint32_t fb = full_bar(stage);
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            if (k2 < nt_item && p.debug != 2) {
              const int cw = as * w0 + (int)((awp >> (8 * k2)) & 0xff) - 8;
              const int ch = as * h0 + (int)((ahp >> (8 * k2)) & 0xff) - 8;
              if (p.merge_a) {   // both slabs of the tap in one box (last coordinate = first slab)
                tma_load_5d(sa + k2 * a_tap_bytes, &tmA, fb, 0, cw, ch, n0, m0 >> 6);
              } else {
                tma_load_4d(sa + k2 * a_tap_bytes, &tmA, fb, m0, cw, ch, n0);
                if (two_slabs) tma_load_4d(sa + k2 * a_tap_bytes + a_stride, &tmA, fb, m0 + 64, cw, ch, n0);
              }
            }
          }
          if (p.debug == 2) {
          } else if (p.merge_b) {
            tma_load_5d(sa + b_off, &tmB, fb, 0, bs * w0 + bw, bs * h0 + bh, n0, n_off >> 6);
          } else {
            for (int s2 = 0; s2 < p.nslabs; ++s2)
              tma_load_4d(sa + b_off + s2 * 8192, &tmB, fb, n_off + s2 * 64, bs * w0 + bw, bs * h0 + bh, n0);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1; }
          if (++iw == p.chunks_w) {
            iw = 0;
            if (++ih == p.chunks_h) { ih = 0; ++in; }
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int split, bz, tap, mt, nt;
      decode(item, split, bz, tap, mt, nt);
      const int nt_item = p.tg > 1 ? group_taps(tap) : 1;
      const int j0 = split * p.chunks_per_split;
      const int j1 = min(p.total_chunks, j0 + p.chunks_per_split);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int j = j0; j < j1; ++j) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = sbase + stage * p.stage_bytes;
          // MN-major SW128: LBO = byte distance between 64-channel slabs, SBO = distance between 8-pixel groups.
          if (p.debug == 1 && j > j0) {
          } else if (p.tap3) {
            const uint64_t adesc = make_smem_desc(sa, a_slab, 1024);
            const uint64_t bdesc = make_smem_desc(sa + 2 * a_slab, 8192, 1024);
            // tap kw multiplies pixels kw .. kw+63 of the 66-pixel halo chunk (one pixel = one 128 B row; the start
            // address moves kw rows into the swizzle repeat, the XOR follows the absolute address) into its own
            // accumulator: three taps per loaded chunk
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int s = 0; s < 4; ++s)
                umma_bf16(tmem_base + kw * acc_cols, adesc + 8 * kw + 128 * s, bdesc + 128 * s, p.idesc,
                          (j > j0 || s > 0) ? 1u : 0u);
            }
          } else if (p.tg > 1) {
            // tap group: every member tap has its own A chunk and accumulator, all multiply the one B chunk
            const uint64_t bdesc = make_smem_desc(sa + p.tg * a_tap_bytes, 8192, 1024);
            for (int k2 = 0; k2 < nt_item; ++k2) {
              const uint64_t adesc = make_smem_desc(sa + k2 * a_tap_bytes, 8192, 1024);
#pragma unroll
              for (int s = 0; s < 4; ++s)
                umma_bf16(tmem_base + k2 * acc_cols, adesc + 128 * s, bdesc + 128 * s, p.idesc,
                          (j > j0 || s > 0) ? 1u : 0u);
            }
          } else {
            const uint64_t adesc = make_smem_desc(sa, 8192, 1024);
            const uint64_t bdesc = make_smem_desc(sa + a_tap_bytes, 8192, 1024);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              // 16 pixels (K) per MMA = two 8-pixel groups = 2048 bytes: +128 in the (addr>>4) field
              umma_bf16(d_tmem, adesc + 128 * s, bdesc + 128 * s, p.idesc, (j > j0 || s > 0) ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (j == j1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
      if (p.tap3 || p.tg > 1) {  // several accumulators fill TMEM: single-buffered, the barrier pair alternates phase
        acc_phase ^= 1;
      } else {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int split, bz, tap, mt, nt;
      decode(item, split, bz, tap, mt, nt);
      const int m = mt * 128 + q * 32 + lane;
      const int n_off = nt * p.BN;
      const bool row_ok = m < p.Ca;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int naccs = p.tap3 ? 3 : (p.tg > 1 ? group_taps(tap) : 1);
      for (int kw3 = 0; kw3 < naccs; ++kw3) {
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (naccs > 1 || p.tg > 1 ? kw3 * acc_cols : acc * 256);
      // tap this accumulator belongs to: filter row `tap` x kw (tap3), member kw3 of tap group `tap`, or `tap`
      const int tap_id = p.tap3 ? tap * 3 + kw3 : (p.tg > 1 ? member_tap(tap, kw3) : tap);
      // destinations: a plain tap writes its own [Ca][Cb] matrix; a sub-pixel parity tap (a,dh,b,dw) is the gradient of
      // a SUM of 3x3 taps, W_a[dh] = sum_{kh in S(a,dh)} W[kh] with S(0,0)={0}, S(0,1)={1,2}, S(1,0)={0,1}, S(1,1)={2},
      // so it is added to every (kh, kw) in S(a,dh) x S(b,dw)
      int dest[4], ndest = 1;
      dest[0] = tap_id;
      if (p.subpixel) {
        const int a = (tap_id >> 3) & 1, dh = (tap_id >> 2) & 1, b = (tap_id >> 1) & 1, dw = tap_id & 1;
        const int kh0 = (a == 0) ? (dh == 0 ? 0 : 1) : (dh == 0 ? 0 : 2), nkh = (a != dh) ? 2 : 1;
        const int kw0 = (b == 0) ? (dw == 0 ? 0 : 1) : (dw == 0 ? 0 : 2), nkw = (b != dw) ? 2 : 1;
        ndest = 0;
        for (int i = 0; i < nkh; ++i)
          for (int j2 = 0; j2 < nkw; ++j2) dest[ndest++] = (kh0 + i) * 3 + kw0 + j2;
      }
      const long long obase0 = (long long)bz * p.out_batch_stride + (long long)m * p.ldOut;
      // partial-tile slot of this item in the workspace (deterministic split-K / shared destinations)
      float* ws_row = p.ws ? p.ws + (long long)split * p.ws_split_stride +
                                 (((long long)bz * p.ws_taps + tap_id) * p.Ca + m) * p.Cb
                           : nullptr;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_addr + c0, v);
        tmem_ld_wait();
        const int col = n_off + c0;
        if (row_ok && col < p.Cb) {
          const int nvalid = min(16, p.Cb - col);
          const bool vec = p.vec_ok && nvalid == 16;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
          if (ws_row) {  // Cb % 8 == 0 and a 16-byte aligned workspace: vector stores are always legal here
            float* o = ws_row + col;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (4 * i < nvalid)
                reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            continue;
          }
          for (int di = 0; di < ndest; ++di) {
            const long long obase = obase0 + (long long)dest[di] * p.out_tap_stride;
            if (p.out_mode == 2) {
              bf16* o = reinterpret_cast<bf16*>(p.out) + obase + col;
              if (vec) {
                uint4 a, b;
                a.x = pack_bf16x2(f[0], f[1]);   a.y = pack_bf16x2(f[2], f[3]);
                a.z = pack_bf16x2(f[4], f[5]);   a.w = pack_bf16x2(f[6], f[7]);
                b.x = pack_bf16x2(f[8], f[9]);   b.y = pack_bf16x2(f[10], f[11]);
                b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
                reinterpret_cast<uint4*>(o)[0] = a;
                reinterpret_cast<uint4*>(o)[1] = b;
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nvalid) o[i] = __float2bfloat16(f[i]);
              }
            } else {
              float* o = reinterpret_cast<float*>(p.out) + obase + col;
              if (p.out_mode == 1) {
                if (vec) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (i < nvalid) o[i] = f[i];
                }
              } else {
                // out_mode 0 without a workspace: this item is the only writer of these elements in this launch
                if (vec) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float4 t = reinterpret_cast<float4*>(o)[i];
                    t.x += f[4 * i]; t.y += f[4 * i + 1]; t.z += f[4 * i + 2]; t.w += f[4 * i + 3];
                    reinterpret_cast<float4*>(o)[i] = t;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (i < nvalid) o[i] += f[i];
                }
              }
            }
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (p.tap3 || p.tg > 1) {
        acc_phase ^= 1;
      } else {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================================
// Second stage of the deterministic weight-gradient accumulation: out[b][t][ca][cb] += sum over the source taps of
// destination tap t (one, or the four parity taps of sub-pixel mode) and over the K splits, in a FIXED order, of the
// partial tiles gemm_wgrad_kernel left in the workspace. Replaces the red.global.add epilogue of round 1, whose
// summation order depended on CTA scheduling (two runs on the same inputs differed in the last bits).
// =====================================================================================================================
struct WgradReduceParams {
  const float* ws;
  float* out;
  int ksplit, nbatch, src_taps, dst_taps, Ca, Cb, ldOut, subpixel, vec_ok;
  int store;   // 1: out = sum (XmcWgradDesc.out_mode 1), 0: out += sum
  long long ws_split_stride, out_tap_stride, out_batch_stride;
};

__global__ void wgrad_reduce_kernel(const WgradReduceParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int cb4 = p.Cb >> 2;
  const long long total = (long long)p.nbatch * p.dst_taps * p.Ca * cb4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % cb4);
    long long rem = idx / cb4;
    const int m = (int)(rem % p.Ca);
    rem /= p.Ca;
    const int t = (int)(rem % p.dst_taps);
    const int bz = (int)(rem / p.dst_taps);
    int src[4], nsrc = 1;
    src[0] = t;
    if (p.subpixel == 2) {
      // pool-fused form: destination (kh, kw) = sum of the 4x4 taps (r, s) with r in {kh, kh+1}, s in {kw, kw+1}
      const int kh = t / 3, kw = t - kh * 3;
      nsrc = 0;
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) src[nsrc++] = (kh + i) * 4 + kw + j;
    } else if (p.subpixel) {
      // destination (kh, kw) collects the parity taps (a,dh) with kh in S(a,dh): S(0,0)={0}, S(0,1)={1,2},
      // S(1,0)={0,1}, S(1,1)={2} (see gemm_wgrad_kernel); source tap index = a<<3 | dh<<2 | b<<1 | dw
      const int kh = t / 3, kw = t - kh * 3;
      const int ah[2] = {kh == 0 ? 0 : (kh == 1 ? 1 : 1), kh == 0 ? 2 : (kh == 1 ? 2 : 3)};  // (a<<1|dh) pairs
      const int aw[2] = {kw == 0 ? 0 : (kw == 1 ? 1 : 1), kw == 0 ? 2 : (kw == 1 ? 2 : 3)};
      nsrc = 0;
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) src[nsrc++] = (ah[i] << 2) | aw[j];
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < nsrc; ++i) {
      const float* w = p.ws + (((long long)bz * p.src_taps + src[i]) * p.Ca + m) * p.Cb + c4 * 4;
      for (int sp = 0; sp < p.ksplit; ++sp) {
        const float4 v = *reinterpret_cast<const float4*>(w + (long long)sp * p.ws_split_stride);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    float* o = p.out + (long long)bz * p.out_batch_stride + (long long)t * p.out_tap_stride + (long long)m * p.ldOut +
               c4 * 4;
    if (p.store) {
      if (p.vec_ok) {
        *reinterpret_cast<float4*>(o) = acc;
      } else {
        o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
      }
    } else if (p.vec_ok) {
      float4 cur = *reinterpret_cast<float4*>(o);
      cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
      *reinterpret_cast<float4*>(o) = cur;
    } else {
      o[0] += acc.x; o[1] += acc.y; o[2] += acc.z; o[3] += acc.w;
    }
  }
}

// =====================================================================================================================
// Host side
// =====================================================================================================================
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
      set_cuda_error(e != cudaSuccess ? e : cudaErrorUnknown);
      return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// bf16 tensor map, SWIZZLE_128B, zero OOB fill. dims/box innermost first; strides in bytes for dims 1..rank-1.
static int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, const uint32_t* estrides = nullptr,
                     CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = get_encode_fn();
  if (!fn) return XMC_ECUDA;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = estrides ? estrides[i] : 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue);
    return XMC_ECUDA;
  }
  return XMC_OK;
}

// N-tile width: multiple of 16 in [16,256] minimising the padded width, ties -> wider tile.
static int pick_bn(int n) {
  if (n <= 256) return ceil_div(n, 16) * 16;
  int best = 256, best_pad = ceil_div(n, 256) * 256;
  for (int bn = 240; bn >= 128; bn -= 16) {
    int pad = ceil_div(n, bn) * bn;
    if (pad < best_pad) { best_pad = pad; best = bn; }
  }
  return best;
}

// N-tile width of the forward kernel from a small cost model (cycles): tensor time = waves * k-iterations * 2*BN
// (four 128xBNx16 MMAs per 64-channel k-iteration), L2->SM time = all tiles' operand bytes / the chip-wide L2
// throughput cap (~6300 B/clk), plus the exposed epilogue of the last tile. For layers with many waves this reduces to
// "fewest padded columns, widest tile"; for the small-M layers (4x4 / 8x8 feature maps: fewer tiles than SMs at
// BN=256) it trades tile width against occupied SMs.
static int pick_bn_fwd(int cout, long long m_tiles, int kiters, int sms) {
  const int cmax = ceil_div(cout, 16) * 16;
  int best = cmax < 256 ? cmax : 256;
  double best_cost = 1e300;
  for (int bn = (cmax < 256 ? cmax : 256); bn >= 32; bn -= 16) {
    const long long tiles = m_tiles * ceil_div(cout, bn);
    const long long waves = (tiles + sms - 1) / sms;
    const double mma = (double)waves * kiters * 2.0 * bn;
    const double l2 = (double)tiles * kiters * (16384.0 + 128.0 * bn) / 6300.0;
    const double cost = (mma > l2 ? mma : l2) + 1000.0 + 8.0 * bn;
    if (cost < best_cost * 0.999) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace xmc

using namespace xmc;

// The opt-in for more than 48 KB of dynamic shared memory is a per-DEVICE function attribute: one process may drive
// several GPUs (the reference's pmap mode; an XLA custom call), so the "already set" flags are kept per device and
// guarded by a mutex (launches from several host threads).
enum { kAttrFwd = 0, kAttrRes = 1, kAttrWgrad = 2, kAttrKinds = 3 };
static cudaError_t ensure_smem_attr(int kind) {
  constexpr int kMaxDev = 64;
  static std::mutex mu;
  static bool done[kMaxDev][kAttrKinds] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool cached = dev >= 0 && dev < kMaxDev;
  std::lock_guard<std::mutex> lock(mu);
  if (cached && done[dev][kind]) return cudaSuccess;
  switch (kind) {
    case kAttrFwd:
      e = cudaFuncSetAttribute(gemm_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd);
      break;
    case kAttrRes:
      e = cudaFuncSetAttribute(conv3x3_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemRes);
      break;
    default:
      e = cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemWgrad);
      break;
  }
  if (e == cudaSuccess && cached) done[dev][kind] = true;
  return e;
}

extern "C" int xmc_conv2d_fwd(const XmcConvDesc* d, const void* x, const void* wk, const float* bias,
                              const void* residual, const void* mask, void* y, void* y_pair, void* stream) {
  if (!d || !x || !wk || (!y && !y_pair)) return XMC_EINVAL;
  if (d->N < 1 || d->H < 1 || d->W < 1 || d->C < 1 || d->Cout < 1 || d->KH < 1 || d->KW < 1) return XMC_EINVAL;
  if ((d->ldA % 8) || (d->ldB % 8) || d->ldB < d->KH * d->KW * d->C) return XMC_EINVAL;
  if (d->pitchW <= 0 && d->ldA < (d->act_f32 == 2 ? d->C / 3 * 2 : d->C)) return XMC_EINVAL;
  if (d->C % 8) return XMC_EINVAL;
  if (!aligned16(x) || !aligned16(wk)) return XMC_EALIGN;
  if (d->batched && (d->strideB_batch % 8)) return XMC_EINVAL;

  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.Cout = d->Cout;
  p.KH = d->KH; p.KW = d->KW; p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  p.cchunks = ceil_div(d->C, 64);
  p.last_mmas = ceil_div(d->C - (p.cchunks - 1) * 64, 16);
  p.batched = d->batched;
  p.strideH = d->strideH > 0 ? d->strideH : 1;
  p.strideW = d->strideW > 0 ? d->strideW : 1;
  p.parities = d->subpixel ? 4 : 1;
  if (d->subpixel && (d->KH != 2 || d->KW != 2 || d->pad_h != 1 || d->pad_w != 1 || residual || d->batched))
    return XMC_EINVAL;
  if (p.strideH > 8 || p.strideW > 8) return XMC_EINVAL;
  p.tw = d->W < 128 ? d->W : 128;
  p.th = 128 / p.tw; if (p.th > d->H) p.th = d->H;
  p.tn = d->batched ? 1 : 128 / (p.tw * p.th); if (p.tn > d->N) p.tn = d->N; if (p.tn < 1) p.tn = 1;
  p.tiles_w = ceil_div(d->W, p.tw);
  p.tiles_h = ceil_div(d->H, p.th);
  p.tiles_n = ceil_div(d->N, p.tn);
  p.BN = pick_bn_fwd(d->Cout, (long long)p.tiles_w * p.tiles_h * p.tiles_n * p.parities,
                     d->KH * d->KW * p.cchunks, num_sms());
  p.n_tiles = ceil_div(d->Cout, p.BN);
  p.idesc = make_idesc_bf16(128, p.BN, 0, 0);
  p.stage_tx_bytes = (uint32_t)(64 * p.tw * p.th * p.tn * 2 + 64 * p.BN * 2);
  static const int fwd_debug = [] { const char* e = getenv("XMC_FWD_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = fwd_debug;
  p.out = y; p.out_dtype = d->out_dtype; p.ldOut = d->ldOut;
  p.bias = bias; p.residual = residual; p.mask = mask;
  // fp32 activations: residual / mask are fp32 tensors and the output must be fp32 (the A operand is the bf16
  // [hi | lo | hi] split of the fp32 input, made by xmc_split3; see XmcConvDesc.act_f32)
  const bool f32io = d->act_f32 != 0;
  if (f32io && d->out_dtype != 1) return XMC_EINVAL;
  if (d->act_f32 == 2) {   // two-part operand [hi | lo]: d->C = 3 * real channels, real channels % 64 == 0
    if (d->C % 3 || (d->C / 3) % 64) return XMC_EINVAL;
    p.pairC = d->C / 3;
  }
  if (y_pair) {
    // second output [hi | lo] of the result: needs the vector path (asserted below) and Cout % 16 == 0
    if (!f32io || (d->Cout % 16) || (d->ldPair % 16) || d->ldPair < 2 * d->Cout ||
        (reinterpret_cast<uintptr_t>(y_pair) & 31))
      return XMC_EINVAL;
    p.out_pair = reinterpret_cast<bf16*>(y_pair);
    p.ldPair = d->ldPair;
  }
  const int rm_unit = f32io ? 4 : 8;  // residual / mask elements per 16 bytes
  if ((d->res_pair || d->mask_pair) && !f32io) return XMC_EINVAL;
  p.res_pair = (residual && d->res_pair) ? 1 : 0;
  p.mask_pair = (mask && d->mask_pair) ? 1 : 0;
  p.ldRes = d->ldRes; p.ldMask = d->ldMask; p.res_shift = d->res_shift; p.relu = d->relu;
  p.mask_last = d->mask_last;
  p.alpha = d->alpha;
  bool vec = !y || (aligned16(y) && (d->out_dtype == 0 ? (d->ldOut % 8 == 0) : (d->ldOut % 4 == 0)));
  if (residual) vec = vec && aligned16(residual) && (d->ldRes % (p.res_pair ? 8 : rm_unit) == 0);
  if (mask) vec = vec && aligned16(mask) && (d->ldMask % (p.mask_pair ? 8 : rm_unit) == 0);
  // 32-byte accesses when every pointer and pitch allows it (16 bf16 / 8 fp32 columns per access)
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  bool vec32 = vec && (!y || (al32(y) && (d->out_dtype == 0 ? (d->ldOut % 16 == 0) : (d->ldOut % 8 == 0))));
  if (residual) vec32 = vec32 && al32(residual) && (d->ldRes % (p.res_pair ? 16 : 2 * rm_unit) == 0);
  if (mask) vec32 = vec32 && al32(mask) && (d->ldMask % (p.mask_pair ? 16 : 2 * rm_unit) == 0);
  p.vec_ok = vec32 ? 2 : (vec ? 1 : 0);
  // the two-part tensors are only handled by the 32-byte vector path (Cout % 16 == 0 keeps every chunk on it)
  if ((y_pair || p.res_pair || p.mask_pair) && (p.vec_ok != 2 || (d->Cout % 16))) return XMC_EALIGN;
  p.bias_smem = (bias && d->Cout <= kBiasMax && aligned16(bias)) ? 1 : 0;

  // ---- resident-weights path for the wide 3x3 layers (see conv3x3_resident_kernel) ------------------------------------
  // XMC_RESIDENT=0 switches the path off (debugging aid; process-global, read once)
  static const int resident_mode = [] { const char* e = getenv("XMC_RESIDENT"); return e ? atoi(e) : 1; }();
  const int bn_res = ceil_div(d->Cout, 16) * 16;
  const bool res_ok = resident_mode > 0 && !f32io && d->KH == 3 && d->KW == 3 && d->pad_h == 1 && d->pad_w == 1 &&
                      p.strideH == 1 && p.strideW == 1 && !d->batched && !d->subpixel && d->pitchW <= 0 &&
                      d->Hin <= 0 && d->Win <= 0 && (d->W % 128) == 0 && (d->H % 2) == 0 && d->C <= 96 &&
                      (d->C % 16) == 0 && bn_res <= 128 && 9 * (bn_res * 128 + (d->C > 64 ? bn_res * 64 : 0)) <= kResBBytes;
  if (res_ok) {
    p.BN = bn_res;
    p.n_tiles = 1;
    p.tw = 128; p.th = 1; p.tn = 1;
    p.tiles_w = d->W / 128; p.tiles_h = d->H; p.tiles_n = d->N;
    p.idesc = make_idesc_bf16(128, p.BN, 0, 0);
    CUtensorMap tmA, tmB, tmB2;
    {
      uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
      uint64_t str[3] = {(uint64_t)d->ldA * 2, (uint64_t)d->ldA * 2 * d->W, (uint64_t)d->ldA * 2 * d->W * d->H};
      uint32_t box[4] = {64, 130, 1, 1};
      int r = make_tmap(&tmA, x, 4, dims, str, box);
      if (r) return r;
    }
    {
      uint64_t dims[3] = {(uint64_t)9 * d->C, (uint64_t)d->Cout, 1};
      uint64_t str[2] = {(uint64_t)d->ldB * 2, (uint64_t)d->ldB * 2 * d->Cout};
      uint32_t box[3] = {64, (uint32_t)p.BN, 1};
      int r = make_tmap(&tmB, wk, 3, dims, str, box);
      if (r) return r;
      uint32_t box2[3] = {32, (uint32_t)p.BN, 1};
      r = make_tmap(&tmB2, wk, 3, dims, str, box2, nullptr, CU_TENSOR_MAP_SWIZZLE_64B);
      if (r) return r;
    }
    XMC_CUDA_CHECK(ensure_smem_attr(kAttrRes));
    const int total = p.tiles_w * (d->H / 2) * d->N;
    const int grid = total < grid_sms() ? total : grid_sms();
    XMC_CUDA_CHECK(launch_pdl(conv3x3_resident_kernel, dim3(grid), dim3(kThreadsFwd), kSmemRes, (cudaStream_t)stream, tmA,
                              tmB, tmB2, p));
    XMC_LAUNCH_CHECK();
    return XMC_OK;
  }

  CUtensorMap tmA, tmB;
  {
    const int Hin = d->Hin > 0 ? d->Hin : d->H * p.strideH;
    const int Win = d->Win > 0 ? d->Win : d->W * p.strideW;
    const uint64_t pW = d->pitchW > 0 ? (uint64_t)d->pitchW : (uint64_t)d->ldA;
    const uint64_t pH = d->pitchH > 0 ? (uint64_t)d->pitchH : pW * Win;
    const uint64_t pN = d->pitchN > 0 ? (uint64_t)d->pitchN : pH * Hin;
    if ((pW % 8) || (pH % 8) || (pN % 8)) return XMC_EINVAL;
    if (p.tw * p.strideW > 256 || p.th * p.strideH > 256) return XMC_EINVAL;
    uint64_t dims[4] = {(uint64_t)(p.pairC ? 2 * p.pairC : d->C), (uint64_t)Win, (uint64_t)Hin, (uint64_t)d->N};
    uint64_t str[3] = {pW * 2, pH * 2, pN * 2};
    uint32_t box[4] = {64, (uint32_t)(p.tw * p.strideW), (uint32_t)(p.th * p.strideH), (uint32_t)p.tn};
    uint32_t est[4] = {1, (uint32_t)p.strideW, (uint32_t)p.strideH, 1};
    int r = make_tmap(&tmA, x, 4, dims, str, box, est);
    if (r) return r;
  }
  {
    const int nb = d->batched ? d->N : 1;
    uint64_t dims[3] = {(uint64_t)d->KH * d->KW * d->C, (uint64_t)d->Cout * p.parities, (uint64_t)nb};
    uint64_t sb = d->batched ? (uint64_t)d->strideB_batch * 2 : (uint64_t)d->ldB * 2 * d->Cout;
    uint64_t str[2] = {(uint64_t)d->ldB * 2, sb};
    uint32_t box[3] = {64, (uint32_t)p.BN, 1};
    int r = make_tmap(&tmB, wk, 3, dims, str, box);
    if (r) return r;
  }
  XMC_CUDA_CHECK(ensure_smem_attr(kAttrFwd));
  const int total = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles * p.parities;
  const int grid = total < grid_sms() ? total : grid_sms();
  // XMC_LEAN_EPI=0 forces the generic epilogue (debugging aid; process-global, read once)
  static const int lean_mode = [] { const char* e = getenv("XMC_LEAN_EPI"); return e ? atoi(e) : 1; }();
  auto al32p = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  const bool lean = lean_mode && p.bias_smem && !mask && d->alpha == 1.f && d->out_dtype == 0 && p.vec_ok == 2 &&
                    (d->Cout % 16) == 0 && (!residual || al32p(residual));
  if (f32io)
    XMC_CUDA_CHECK(launch_pdl(gemm_fwd_kernel<3>, dim3(grid), dim3(64 + 128 * FwdParts<3>::v), kSmemFwd,
                              (cudaStream_t)stream, tmA, tmB, p));
  else if (lean && !residual)
    XMC_CUDA_CHECK(launch_pdl(gemm_fwd_kernel<1>, dim3(grid), dim3(kThreadsFwd), kSmemFwd, (cudaStream_t)stream, tmA, tmB,
                              p));
  else if (lean)
    XMC_CUDA_CHECK(launch_pdl(gemm_fwd_kernel<2>, dim3(grid), dim3(kThreadsFwd), kSmemFwd, (cudaStream_t)stream, tmA, tmB,
                              p));
  else
    XMC_CUDA_CHECK(launch_pdl(gemm_fwd_kernel<0>, dim3(grid), dim3(kThreadsFwd), kSmemFwd, (cudaStream_t)stream, tmA, tmB,
                              p));
  XMC_LAUNCH_CHECK();
  return XMC_OK;
}

// Validates the descriptor and fills the launch plan (tiling, K split, work items). Shared by the workspace query and
// the launch so that both always agree.
static int plan_wgrad(const XmcWgradDesc* d, WgradParams& p) {
  if (!d) return XMC_EINVAL;
  if (d->N < 1 || d->H < 1 || d->W < 1 || d->Ca < 1 || d->Cb < 1 || d->KH < 1 || d->KW < 1) return XMC_EINVAL;
  if ((d->ldA % 8) || (d->ldB % 8) || (d->pitchWA <= 0 && d->ldA < d->Ca) || d->ldB < d->Cb) return XMC_EINVAL;
  if ((d->Ca % 8) || (d->Cb % 8)) return XMC_EINVAL;
  memset(&p, 0, sizeof(p));
  p.tw = d->W < 64 ? d->W : 64;
  p.th = 64 / p.tw; if (p.th > d->H) p.th = d->H;
  p.tn = d->batched ? 1 : 64 / (p.tw * p.th); if (p.tn > d->N) p.tn = d->N; if (p.tn < 1) p.tn = 1;
  p.chunks_w = ceil_div(d->W, p.tw);
  p.chunks_h = ceil_div(d->H, p.th);
  p.chunks_n = d->batched ? 1 : ceil_div(d->N, p.tn);
  p.total_chunks = p.chunks_w * p.chunks_h * p.chunks_n;
  p.BN = pick_bn(d->Cb);
  p.n_tiles = ceil_div(d->Cb, p.BN);
  p.nslabs = ceil_div(p.BN, 64);
  const int m_tiles = ceil_div(d->Ca, 128);
  // fp32 destinations (out_mode 0: accumulate, 1: store) take every decomposition below; bf16 stores stay one item per
  // output element
  const bool f32_out = d->out_mode == 0 || d->out_mode == 1;
  if (d->subpixel && (d->KH != 3 || d->KW != 3 || !f32_out || d->batched || d->subpixel > 2)) return XMC_EINVAL;
  // tap3: for 3x3 kernels on rows of >= 64 pixels whose three kw accumulators fit TMEM (Cb <= 160), a work item is a
  // filter ROW: one 66-pixel halo chunk of xa and one chunk of xb feed three taps (3x less operand traffic; these
  // narrow-channel layers are L2 -> SM bandwidth bound). XMC_WGRAD_TAP3=0 switches it off (debugging aid, read once per
  // process).
  static const int tap3_mode = [] { const char* e = getenv("XMC_WGRAD_TAP3"); return e ? atoi(e) : 1; }();
  p.tap3 = (tap3_mode && d->KH == 3 && d->KW == 3 && d->pad_h == 1 && d->pad_w == 1 && !d->subpixel && !d->batched &&
            d->pitchWA <= 0 && f32_out && d->W >= 64 && p.n_tiles == 1 && 3 * (((p.BN + 31) / 32) * 32) <= 512)
               ? 1 : 0;
  const int taps = d->subpixel ? 16 : d->KH * d->KW;
  p.subpixel = d->subpixel;
  const int nbatch = d->batched ? d->N : 1;
  const int acc_cols = ((p.BN + 31) / 32) * 32;
  p.a_slabs = (d->Ca <= 64 && !p.tap3) ? 1 : 2;
  // Tap groups: the taps of one output tile all multiply the same B chunk (sub-pixel parity taps: per parity), so an
  // item can take tg of them — tg A chunks, ONE B chunk, tg accumulators per ring stage. The kernel is bound by
  // L2 -> SM operand traffic (ncu: 12-13 TB/s at 54 % tensor activity for 192 -> 384 at 32x32); per pixel and item the
  // traffic drops from tg x (A + B) to tg x A + B. Limits: tg accumulators in the 512 TMEM columns, a stage of at most
  // a third (BN <= 128) or a quarter (BN <= 192) of the 224 KB ring — the loop is latency-bound on the bytes in flight
  // (ring / ~2 us loaded L2 latency = ~50 B/clk per SM), so grouping must not cost ring depth. Measured: 3 -> 96 packed
  // window 395 -> 270 us, pool-fused 96 -> 96 316 -> 240 us, sub-pixel 192 -> 96 315 -> 250 us, 192 -> 384 at 32x32
  // 203 -> 173 us; BN = 256 stays ungrouped (1536 -> 1536 at 4x4: 118 -> 152 us with 3 stages and an epilogue that is
  // no longer overlapped). XMC_WGRAD_TG=1 switches the grouping off (debugging aid).
  static const int tg_max = [] { const char* e = getenv("XMC_WGRAD_TG"); return e ? atoi(e) : 4; }();
  p.tg = 1;
  if (!p.tap3 && f32_out && taps > 1 && p.BN <= 192) {
    const int tg_minst = p.BN <= 128 ? 3 : 4;   // ring stages a grouped item must keep
    for (int tg = 4; tg >= 2; --tg) {
      if (tg > tg_max || tg > taps || tg * acc_cols > 512) continue;
      if (d->subpixel == 1 && tg == 3) continue;
      if ((tg * p.a_slabs + p.nslabs) * 8192 > kWgradRing / tg_minst) continue;
      p.tg = tg;
      break;
    }
  }
  p.groups = p.tap3 ? 3 : ceil_div(taps, p.tg);
  p.stage_bytes = p.tap3 ? (uint32_t)(2 * 9216 + p.nslabs * 8192) : (uint32_t)((p.tg * p.a_slabs + p.nslabs) * 8192);
  p.stages = kWgradRing / (int)p.stage_bytes;
  if (p.stages > kWgradStagesMax) p.stages = kWgradStagesMax;
  const int per_item = p.tap3 ? 3 : p.tg;   // accumulators (taps) an item computes
  const int base_ctas = m_tiles * p.n_tiles * p.groups * nbatch;
  int ksplit = 1;
  if (f32_out) {
    // Split K (pixels) across CTAs with a small cost model (microseconds): tensor time = waves x chunks per CTA x
    // (four 128 x BN x 16 MMAs = 2 BN cycles per 64-pixel chunk and tap) + one epilogue per item, plus — as soon as
    // the reduction is split — the partial tiles' round trip through the workspace and the second-stage launch.
    // Layers with large weights (1536 x 1536 x 9 = 85 MB per split) therefore stay unsplit even if their last wave is
    // ragged; narrow layers (96 x 96) split ~50 ways. Keeps >= 8 chunks (512 pixels) per CTA.
    const int sms = num_sms();
    const int max_split = p.total_chunks / 8 > 0 ? p.total_chunks / 8 : 1;
    const double clk_mhz = 1600.0, hbm_bytes_per_us = 5.0e6;
    const double tile_bytes = 4.0 * (double)taps * d->Ca * d->Cb * nbatch;
    double best_cost = 1e300;
    for (int ks = 1; ks <= max_split && ks <= 1024; ++ks) {
      const int cps = ceil_div(p.total_chunks, ks);
      const int ks_eff = ceil_div(p.total_chunks, cps);
      const long long ctas = (long long)base_ctas * ks_eff;
      const long long waves = (ctas + sms - 1) / sms;
      const double epi = 1500.0 + 12.0 * p.BN * per_item;
      double cost = (double)waves * ((double)cps * 2.0 * p.BN * per_item + epi) / clk_mhz;
      if (ks_eff > 1) cost += 3.0 + (2.0 * ks_eff + 2.0) * tile_bytes / hbm_bytes_per_us;
      if (cost < best_cost * 0.999) { best_cost = cost; ksplit = ks_eff; }
      if (ctas > 16LL * sms) break;
    }
  }
  p.chunks_per_split = ceil_div(p.total_chunks, ksplit);
  p.ksplit = ceil_div(p.total_chunks, p.chunks_per_split);
  p.KW = d->KW; p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  p.taps = taps; p.m_tiles = m_tiles;
  p.items_per_split = base_ctas;
  p.total_items = base_ctas * p.ksplit;
  p.Ca = d->Ca; p.Cb = d->Cb; p.batched = d->batched;
  p.slab_bytes = (uint32_t)(64 * p.tw * p.th * p.tn * 2);
  // One box per operand instead of one per 64-channel slab: the TMA unit spends a fixed ~100 cycles per box on top of
  // ~128 B/clk (an 8 KB slab box: ~50 B/clk, which is where the narrow-slab loads of this kernel sat). A 5-D tensor map
  // whose LAST dimension walks the slabs (stride 128 B) lands them slab-major in shared memory, exactly the layout the
  // MN-major descriptors expect. Needs whole slabs in global memory (channels % 64 == 0), full 64-pixel chunks (the slab
  // pitch in shared memory is then the 8 KB of a box row block) and, for B, tiles that start on a slab boundary.
  static const int merge_mode = [] { const char* e = getenv("XMC_WGRAD_MERGE"); return e ? atoi(e) : 1; }();
  static const int debug_mode = [] { const char* e = getenv("XMC_WGRAD_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = debug_mode;
  p.merge_a = (merge_mode && !p.tap3 && p.a_slabs == 2 && d->Ca % 64 == 0 && p.slab_bytes == 8192 && d->pitchWA <= 0) ? 1 : 0;
  p.merge_b = (merge_mode && p.nslabs > 1 && d->Cb % 64 == 0 && p.slab_bytes == 8192 &&
               (p.n_tiles == 1 || p.BN % 64 == 0)) ? 1 : 0;
  p.idesc = make_idesc_bf16(128, p.BN, 1, 1);
  p.out_mode = d->out_mode; p.ldOut = d->ldOut;
  p.out_tap_stride = d->out_tap_stride; p.out_batch_stride = d->out_batch_stride;
  p.alpha = d->alpha;
  // source-tap slots of the workspace: tap3 items carry three kw accumulators each
  p.ws_taps = d->subpixel ? 16 : d->KH * d->KW;
  p.ws_split_stride = (long long)nbatch * p.ws_taps * d->Ca * d->Cb;
  return XMC_OK;
}

// A workspace is needed whenever an output element has more than one producer item: K split across CTAs, or the
// sub-pixel parity taps (each destination tap sums four of them).
static bool wgrad_needs_ws(const XmcWgradDesc* d, const WgradParams& p) {
  return (d->out_mode == 0 || d->out_mode == 1) && (p.ksplit > 1 || d->subpixel);
}

extern "C" int xmc_conv2d_wgrad_workspace_bytes(const XmcWgradDesc* d, long long* bytes) {
  if (!bytes) return XMC_EINVAL;
  WgradParams p;
  const int r = plan_wgrad(d, p);
  if (r) return r;
  *bytes = wgrad_needs_ws(d, p) ? (long long)p.ksplit * p.ws_split_stride * (long long)sizeof(float) : 0;
  return XMC_OK;
}

extern "C" int xmc_conv2d_wgrad(const XmcWgradDesc* d, const void* xa, const void* xb, void* dw, void* workspace,
                                long long workspace_bytes, void* stream) {
  if (!d || !xa || !xb || !dw) return XMC_EINVAL;
  WgradParams p;
  {
    const int r = plan_wgrad(d, p);
    if (r) return r;
  }
  if (!aligned16(xa) || !aligned16(xb)) return XMC_EALIGN;
  p.out = dw;
  const int unit = d->out_mode == 2 ? 8 : 4;
  p.vec_ok = (aligned16(dw) && d->ldOut % unit == 0 && d->out_tap_stride % unit == 0 &&
              d->out_batch_stride % unit == 0) ? 1 : 0;
  const bool use_ws = wgrad_needs_ws(d, p);
  if (use_ws) {
    const long long need = (long long)p.ksplit * p.ws_split_stride * (long long)sizeof(float);
    if (!workspace || workspace_bytes < need) return XMC_EINVAL;
    if (!aligned16(workspace)) return XMC_EALIGN;
    p.ws = reinterpret_cast<float*>(workspace);
  }

  CUtensorMap tmA, tmB;
  if (d->pitchWA > 0) {
    // re-pitched view of xa (packed-window convolutions): extents (Ca, W, HinA, N), caller-given pitches
    if ((d->pitchWA % 8) || (d->pitchHA % 8) || (d->pitchNA % 8) || d->HinA < d->H || d->subpixel || d->batched)
      return XMC_EINVAL;
    uint64_t dims[4] = {(uint64_t)d->Ca, (uint64_t)d->W, (uint64_t)d->HinA, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)d->pitchWA * 2, (uint64_t)d->pitchHA * 2, (uint64_t)d->pitchNA * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    int r = make_tmap(&tmA, xa, 4, dims, str, box);
    if (r) return r;
  } else {
    // pool-fused mode (subpixel == 2): xa is the [N,2H,2W,Ca] input, traversed with stride 2 (one 4x4 tap per item)
    const int as = d->subpixel == 2 ? 2 : 1;
    if (p.tw * as > 256 || p.th * as > 256) return XMC_EINVAL;
    uint64_t dims[4] = {(uint64_t)d->Ca, (uint64_t)d->W * as, (uint64_t)d->H * as, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)d->ldA * 2, (uint64_t)d->ldA * 2 * d->W * as,
                       (uint64_t)d->ldA * 2 * d->W * as * d->H * as};
    uint32_t box[4] = {64, (uint32_t)((p.tap3 ? p.tw + 2 : p.tw) * as), (uint32_t)(p.th * as), (uint32_t)p.tn};
    uint32_t est[4] = {1, (uint32_t)as, (uint32_t)as, 1};
    int r;
    if (p.merge_a) {   // (64 channels, W, H, N, slab): the slab dimension last, 128 bytes apart
      uint64_t dims5[5] = {64, dims[1], dims[2], dims[3], (uint64_t)(d->Ca / 64)};
      uint64_t str5[4] = {str[0], str[1], str[2], 128};
      uint32_t box5[5] = {64, box[1], box[2], box[3], 2};
      uint32_t est5[5] = {1, est[1], est[2], 1, 1};
      r = make_tmap(&tmA, xa, 5, dims5, str5, box5, est5);
    } else {
      r = make_tmap(&tmA, xa, 4, dims, str, box, est);
    }
    if (r) return r;
  }
  {
    // sub-pixel mode (subpixel == 1): xb is the [N,2H,2W,Cb] gradient, traversed with stride 2 (one parity per tap)
    const int bs = d->subpixel == 1 ? 2 : 1;
    if (p.tw * bs > 256 || p.th * bs > 256) return XMC_EINVAL;
    uint64_t dims[4] = {(uint64_t)d->Cb, (uint64_t)d->W * bs, (uint64_t)d->H * bs, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)d->ldB * 2, (uint64_t)d->ldB * 2 * d->W * bs,
                       (uint64_t)d->ldB * 2 * d->W * bs * d->H * bs};
    uint32_t box[4] = {64, (uint32_t)(p.tw * bs), (uint32_t)(p.th * bs), (uint32_t)p.tn};
    uint32_t est[4] = {1, (uint32_t)bs, (uint32_t)bs, 1};
    int r;
    if (p.merge_b) {
      uint64_t dims5[5] = {64, dims[1], dims[2], dims[3], (uint64_t)(d->Cb / 64)};
      uint64_t str5[4] = {str[0], str[1], str[2], 128};
      uint32_t box5[5] = {64, box[1], box[2], box[3], (uint32_t)p.nslabs};
      uint32_t est5[5] = {1, est[1], est[2], 1, 1};
      r = make_tmap(&tmB, xb, 5, dims5, str5, box5, est5);
    } else {
      r = make_tmap(&tmB, xb, 4, dims, str, box, est);
    }
    if (r) return r;
  }
  XMC_CUDA_CHECK(ensure_smem_attr(kAttrWgrad));
  const int grid = p.total_items < grid_sms() ? p.total_items : grid_sms();
  XMC_CUDA_CHECK(launch_pdl(gemm_wgrad_kernel, dim3(grid), dim3(kThreads), kSmemWgrad, (cudaStream_t)stream, tmA, tmB, p));
  XMC_LAUNCH_CHECK();
  if (use_ws) {
    WgradReduceParams r;
    r.ws = p.ws; r.out = reinterpret_cast<float*>(dw);
    r.ksplit = p.ksplit; r.nbatch = d->batched ? d->N : 1; r.src_taps = p.ws_taps;
    r.dst_taps = d->KH * d->KW; r.Ca = d->Ca; r.Cb = d->Cb; r.ldOut = d->ldOut; r.subpixel = p.subpixel;
    r.vec_ok = p.vec_ok;
    r.store = d->out_mode == 1 ? 1 : 0;
    r.ws_split_stride = p.ws_split_stride; r.out_tap_stride = d->out_tap_stride;
    r.out_batch_stride = d->out_batch_stride;
    const long long total = (long long)r.nbatch * r.dst_taps * r.Ca * (r.Cb / 4);
    long long blocks = ceil_div_ll(total, 256);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    XMC_CUDA_CHECK(launch_pdl(wgrad_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, r));
    XMC_LAUNCH_CHECK();
  }
  return XMC_OK;
}
