"""Thin tensor-level wrappers over the libxmc.so C ABI. torch tensors are only device-memory containers here; every
arithmetic operation below is a hand-written sm_100a kernel. Each wrapper counts its kernel launches in LAUNCHES."""
import ctypes

import torch

from . import _lib
from ._lib import BnDesc, ConvDesc, WgradDesc, ptr, stream

import contextlib

BF16 = torch.bfloat16
F32 = torch.float32
LAUNCHES = [0]
# Activation dtype of the program being issued: bf16 (the reference's config.dtype default, coco_xmc.py:45) or fp32
# (config.dtype = "float32"; the frozen ResNet branch). Host-side bookkeeping only: every wrapper below derives the
# kernel's activation type from the dtype of the tensors it is given, ACT only decides what `empty()` allocates.
ACT = [BF16]


@contextlib.contextmanager
def act_dtype(dtype):
  prev = ACT[0]
  ACT[0] = dtype
  try:
    yield
  finally:
    ACT[0] = prev


def _f32(t):
  return int(t.dtype == F32)


def _call(name, *args, launches=1):
  fn = getattr(_lib.lib(), name)
  _lib.check(fn(*args))
  LAUNCHES[0] += launches


# ---- SM reservation for a concurrent collective ------------------------------------------------------------------------
# The tensor-core kernels are persistent (one thread block per SM, ~200 KB of shared memory): an NCCL kernel launched on
# another stream cannot co-reside and either waits for a kernel boundary or, once resident, pushes the tail of every
# GEMM grid into a second wave — measured on 8 GPUs, the "overlapped" all-reduces cost about their full duration. While
# a gradient all-reduce is in flight the GEMM launches therefore leave `n` SMs free (xmc_set_sm_limit); the window
# closes after `tflop` of executed GEMM work has been issued (about the all-reduce's duration) or at release_sms().
_RESERVE = [0.0]


def reserve_sms(n, tflop):
  if n <= 0:
    return release_sms()
  _lib.check(_lib.lib().xmc_set_sm_limit(max(_lib.lib().xmc_num_sms() - n, 1)))
  _RESERVE[0] = tflop * 1e12


def release_sms():
  if _RESERVE[0] != 0.0:
    _lib.check(_lib.lib().xmc_set_sm_limit(0))
    _RESERVE[0] = 0.0


def _spend(flops):
  if _RESERVE[0] > 0.0:
    _RESERVE[0] -= flops
    if _RESERVE[0] <= 0.0:
      _RESERVE[0] = 1.0   # force release_sms to act
      release_sms()


def _pix_ld(t):
  """Pitch (elements) between consecutive pixels/rows of a tensor whose last dim is contiguous."""
  assert t.stride(-1) == 1
  return t.stride(-2)


def _check_dense_rows(t):
  """All leading dims must collapse into one row index with pitch stride(-2)."""
  ld = t.stride(-2)
  exp = ld
  for d in range(t.dim() - 2, -1, -1):
    assert t.stride(d) == exp or t.shape[d] == 1, f"non-collapsible layout {t.shape} {t.stride()}"
    exp *= t.shape[d]


def empty(shape, dtype=None):
  return torch.empty(shape, device="cuda", dtype=ACT[0] if dtype is None else dtype)


def zeros(shape, dtype=F32):
  return torch.zeros(shape, device="cuda", dtype=dtype)


# ------------------------------------------------------------------------------------------------------------- GEMMs
def split3(x2d, weights=False, pair=False):
  """fp32 [rows, C] (pitched) -> bf16 [rows, 3C]: [hi | lo | hi] for an activation, [hi | hi | lo] for the B operand
  (a "weight" that is itself an activation, as in the word-loss GEMMs); pair=True: the two-part form [rows, 2C] =
  [hi | lo] — see XmcConvDesc.act_f32."""
  rows, C = x2d.shape
  parts = 2 if pair else 3
  out = empty((rows, parts * C), BF16)
  _call("xmc_split3", ptr(x2d), rows, C, x2d.stride(0), 2 if pair else int(weights), ptr(out), parts * C, stream())
  return out


def pair_ok(C):
  """The forward GEMM can read the two-part operand [hi | lo] when the channel count is a multiple of 64."""
  return C % 64 == 0


def _split_nhwc(x, c=None, pair=False):
  """[N,H,W,>=C] fp32 view -> bf16 [N,H,W,3C] split copy, or [N,H,W,2C] with pair=True (rows must collapse)."""
  _check_dense_rows(x)
  C = x.shape[-1] if c is None else c
  rows = x.numel() // x.shape[-1]
  flat = torch.as_strided(x, (rows, C), (x.stride(-2), 1))
  return split3(flat, pair=pair).view(*x.shape[:-1], (2 if pair else 3) * C)


def conv_fwd(x, wk, kh, cout, *, bias=None, residual=None, res_shift=0, mask=None, relu=False, out=None,
             out_dtype=None, alpha=1.0, ldb=None, batched=False, stride_b=0, c=None, stride=1, pad=None,
             mask_last=False, view=None, subpixel=False, pre_split=False, alg_scale=1.0, x_pair=None,
             want_pair=False, want_f32=True, residual_pair=None, mask_pair=None):
  """x: [N,H,W,>=C] bf16 view (channel-contiguous); wk: bf16 tensor whose data pointer is the [cout][kh*kh*C] K-major
  matrix (row pitch ldb). Returns y [N,H/stride,W/stride,cout] (or writes into the `out` view).
  stride=2 reads the input at (h*2+kh-pad, w*2+kw-pad) (XLA SAME: pad low = 0 for a 3x3 or 1x1 kernel on an even
  input). alg_scale: bookkeeping only (bench.py): algorithmic FLOPs of this launch relative to the dense count, e.g.
  0.25 for a stride-1 convolution over a zero-inserted gradient (the transpose of a stride-2 convolution).
  fp32-activation mode, chains of convolutions: want_pair=True also returns the result as the bf16 two-part operand
  [.., hi | lo] (written by the epilogue), x_pair= passes such an operand in place of splitting x here; x may then be
  None (shape taken from x_pair). residual_pair / mask_pair: residual / mask given as such two-part tensors (value =
  hi + lo, mask = [hi > 0]); want_f32=False (with want_pair) skips the fp32 output, so a chain of convolutions never
  materialises fp32 activations. `view`: dict overriding the input view (Hout, Wout, KH, KW, pitches, strides) for the ResNet stem."""
  pairC = 0
  if x_pair is not None:   # fp32 mode, operand already in the two-part form (a previous launch's want_pair output)
    pairC = x_pair.shape[-1] // 2
    x, pre_split, c = x_pair, True, 3 * pairC
  elif x is not None and x.dtype == F32 and view is None and c is None and pair_ok(x.shape[-1]):
    pairC = x.shape[-1]
    x, pre_split, c = _split_nhwc(x, pair=True), True, 3 * pairC
  f32 = x.dtype == F32 or pre_split   # pre_split: x already is a bf16 split operand, I/O tensors are fp32
  if f32 and not pre_split:
    # fp32-activation mode, channel counts that are not a multiple of 64: A = [hi | lo | hi] split of x (3C channels)
    x = _split_nhwc(x, c)
    c = None
  if f32 and wk.dtype == F32:
    # wk is either already a split weight copy of the arena (bf16, ld given by the caller) or, as here, an fp32
    # activation used as the B operand (word-loss GEMMs): split as [hi | hi | lo]
    Kb = wk.shape[-1]
    ldb_in = ldb if ldb is not None else Kb
    rows_b = wk.numel() // wk.shape[-1]
    wk = split3(torch.as_strided(wk, (rows_b, Kb), (ldb_in, 1)), weights=True)
    ldb, stride_b = 3 * Kb, 3 * stride_b
  if out_dtype is None:
    out_dtype = F32 if f32 else BF16
  N, H, W = x.shape[0], x.shape[1], x.shape[2]
  C = x.shape[3] if c is None else c
  _check_dense_rows(x)
  d = ConvDesc()
  d.act_f32 = 2 if pairC else int(f32)
  d.N, d.H, d.W, d.C, d.ldA = N, H // stride, W // stride, C, _pix_ld(x)
  d.KH = d.KW = kh
  d.pad_h = d.pad_w = (kh // 2 if stride == 1 else 0) if pad is None else pad
  if stride != 1:
    d.strideH = d.strideW = stride
    d.Hin, d.Win = H, W
  if view is not None:
    d.H, d.W, d.KH, d.KW = view["Hout"], view["Wout"], view["KH"], view["KW"]
    d.pad_h = d.pad_w = 0
    d.strideH, d.strideW, d.Hin, d.Win = view["strideH"], view["strideW"], view["Hin"], view["Win"]
    d.pitchW, d.pitchH, d.pitchN = view["pitchW"], view["pitchH"], view["pitchN"]
  H, W = d.H, d.W
  d.mask_last = 1 if mask_last else 0
  if subpixel:  # four 2x2 convs writing the 2x up-sampled output (see XmcConvDesc.subpixel)
    d.subpixel = 1
    H, W = 2 * d.H, 2 * d.W
  d.Cout = cout
  d.ldB = ldb if ldb is not None else d.KH * d.KW * C
  d.batched = 1 if batched else 0
  d.strideB_batch = stride_b
  if not want_f32:
    assert want_pair and f32 and out is None
    d.out_dtype, d.ldOut = 1, cout
  else:
    if out is None:
      out = empty((N, H, W, cout), out_dtype)
    else:
      _check_dense_rows(out)
    d.out_dtype = 0 if out.dtype == BF16 else 1
    d.ldOut = _pix_ld(out)
  if residual_pair is not None:
    assert residual is None and f32
    residual, d.res_pair = residual_pair, 1
  if mask_pair is not None:
    assert mask is None and f32
    mask, d.mask_pair = mask_pair, 1
  d.alpha = alpha
  d.relu = 1 if relu else 0
  d.res_shift = res_shift
  d.ldRes = _pix_ld(residual) if residual is not None else 0
  d.ldMask = _pix_ld(mask) if mask is not None else 0
  out_pair = None
  if want_pair:
    assert f32 and cout % 16 == 0
    out_pair = empty((N, H, W, 2 * cout), BF16)
    d.ldPair = 2 * cout
  _call("xmc_conv2d_fwd", ctypes.byref(d), ptr(x), ptr(wk), ptr(bias), ptr(residual), ptr(mask), ptr(out),
        ptr(out_pair), stream())
  _spend(2.0 * N * d.H * d.W * (4 if subpixel else d.KH * d.KW) * C * cout)
  return (out, out_pair) if want_pair else out


def wgrad(xa, xb, kh, out, *, out_mode=0, batched=False, ld_out=None, tap_stride=None, batch_stride=0, alpha=1.0,
          ca=None, cb=None, subpixel=False, view_a=None):
  """out[b][tap][ca][cb] (+)= sum_pixels xa[p+shift][ca] * xb[p][cb]. xa, xb: [N,H,W,C*] bf16 views.
  view_a: dict re-pitching xa (packed-window form, see XmcWgradDesc.HinA): KH, KW, Hin, pitchW, pitchH, pitchN."""
  if xa.dtype == F32 or xb.dtype == F32:
    # fp32-activation mode: dw = a_hi^T b_hi + a_lo^T b_hi + a_hi^T b_lo, three passes of the bf16 kernel over pitched
    # views of the [hi | lo] split copies, accumulated in the fp32 output (16 mantissa bits per operand)
    assert xa.dtype == F32 and xb.dtype == F32 and view_a is None and ca is None and cb is None
    Ca, Cb = xa.shape[-1], xb.shape[-1]
    a3, b3 = _split_nhwc(xa, pair=True), _split_nhwc(xb, pair=True)   # [hi | lo]: only the two views are read
    a_hi, a_lo, b_hi, b_lo = a3[..., :Ca], a3[..., Ca:2 * Ca], b3[..., :Cb], b3[..., Cb:2 * Cb]
    assert out.dtype == F32
    first = 1 if out_mode in (1, 2) else 0   # store modes: the first pass stores, the others accumulate
    for i, (pa, pb) in enumerate(((a_hi, b_hi), (a_lo, b_hi), (a_hi, b_lo))):
      _wgrad_bf16(pa, pb, kh, out, out_mode=first if i == 0 else 0, batched=batched, ld_out=ld_out,
                  tap_stride=tap_stride, batch_stride=batch_stride, alpha=alpha, subpixel=subpixel)
    return out
  return _wgrad_bf16(xa, xb, kh, out, out_mode=out_mode, batched=batched, ld_out=ld_out, tap_stride=tap_stride,
                     batch_stride=batch_stride, alpha=alpha, ca=ca, cb=cb, subpixel=subpixel, view_a=view_a)


def _wgrad_bf16(xa, xb, kh, out, *, out_mode=0, batched=False, ld_out=None, tap_stride=None, batch_stride=0, alpha=1.0,
                ca=None, cb=None, subpixel=False, view_a=None):
  """One launch of the bf16 tensor-core weight-gradient kernel (+ its deterministic second stage)."""
  # pixel grid of the reduction: xa's (the low-resolution input in sub-pixel mode), xb's for a re-pitched xa view and
  # in pool-fused mode (subpixel == 2: xa is the 2x larger input, xb the low-resolution gradient)
  N, H, W = (xa if (view_a is None and subpixel != 2) else xb).shape[:3]
  _check_dense_rows(xa)
  _check_dense_rows(xb)
  d = WgradDesc()
  d.N, d.H, d.W = N, H, W
  d.Ca = xa.shape[3] if ca is None else ca
  d.Cb = xb.shape[3] if cb is None else cb
  d.ldA, d.ldB = _pix_ld(xa), _pix_ld(xb)
  d.KH = d.KW = kh
  d.pad_h = d.pad_w = kh // 2
  if view_a is not None:
    d.KH, d.KW, d.pad_h, d.pad_w = view_a["KH"], view_a["KW"], 0, 0
    d.HinA, d.pitchWA, d.pitchHA, d.pitchNA = view_a["Hin"], view_a["pitchW"], view_a["pitchH"], view_a["pitchN"]
  d.batched = 1 if batched else 0
  d.out_mode = out_mode
  d.ldOut = d.Cb if ld_out is None else ld_out
  d.out_tap_stride = d.Ca * d.Cb if tap_stride is None else tap_stride
  d.out_batch_stride = batch_stride
  d.alpha = alpha
  d.subpixel = int(subpixel)
  # deterministic accumulation: partial tiles of a split reduction go through a caller-owned workspace and are added in
  # a fixed order by a second kernel (counted as a launch); the caching allocator hands the same block back each step
  need = ctypes.c_longlong(0)
  _lib.check(_lib.lib().xmc_conv2d_wgrad_workspace_bytes(ctypes.byref(d), ctypes.byref(need)))
  ws = torch.empty(need.value, device="cuda", dtype=torch.uint8) if need.value else None
  _call("xmc_conv2d_wgrad", ctypes.byref(d), ptr(xa), ptr(xb), ptr(out), ptr(ws), need.value, stream(),
        launches=2 if need.value else 1)
  _spend(2.0 * N * H * W * (16 if subpixel else d.KH * d.KW) * d.Ca * d.Cb)
  return out


# ------------------------------------------------------------------ 3-channel image convolutions (packed-window form)
def c3_pad(x3):
  """[N,H,W,3] bf16 -> zero-bordered 8-channel copy [N,H+2,W+2,8]: per output pixel and kh the 3 kw taps x 8 channels
  are one contiguous 24-element run, which a tensor map with a 16-byte pixel pitch hands to the GEMM kernels."""
  N, H, W, _ = x3.shape
  out = empty((N, H + 2, W + 2, 8))
  _call("xmc_pad_c3_to_c8", ptr(x3), N, H, W, ptr(out), stream())
  return out


def c3_pack_weights(w, ldw, cout):
  """bf16 [cout][ldw] with k = tap*3+c (a forward or dgrad arena matrix) -> bf16 [cout][72] with k = kh*24+kw*8+c."""
  out = empty((cout, 72))
  _call("xmc_pack_c3_weights", ptr(w), ldw, cout, ptr(out), stream())
  return out


def _c3_view(xpad):
  S = xpad.shape[1] - 2
  WP = xpad.shape[2]
  return dict(Hout=S, Wout=WP - 2, KH=3, KW=1, strideH=1, strideW=1, Hin=S + 2, Win=WP - 2, pitchW=8, pitchH=WP * 8,
              pitchN=(S + 2) * WP * 8, real_taps=9, real_c=3)


def c3_conv(xpad, wpacked, cout, **kw):
  """conv3x3 (3 image channels -> cout) on the tcgen05 GEMM kernel: K = 3 kh-taps x 24 (two K=16 MMA slices each)."""
  return conv_fwd(xpad, wpacked, 3, cout, ldb=72, c=24, view=_c3_view(xpad), **kw)


def c3_wgrad(xpad, y, flip, s_tap, s_c3, s_c, out):
  """out[tap_o*s_tap + c3*s_c3 + c*s_c] += sum_p x3[p + d(tap)][c3] * y[p][c] (the contract of xmc_wgrad_c3) on the
  tcgen05 wgrad kernel: [3 kh][24][C] partial result, then a 27*C-element scatter-add."""
  C = y.shape[3]
  v = _c3_view(xpad)
  tmp = zeros((3, 24, C))
  LAUNCHES[0] += 1
  wgrad(xpad, y, 3, tmp, out_mode=0, ld_out=C, tap_stride=24 * C, ca=24,
        view_a=dict(KH=3, KW=1, Hin=v["Hin"], pitchW=8, pitchH=v["pitchH"], pitchN=v["pitchN"]))
  _call("xmc_unpack_c3_wgrad", ptr(tmp), C, int(flip), s_tap, s_c3, s_c, ptr(out), stream())


def wgrad_c3(x3, y, kh, flip, s_tap, s_c3, s_c, out):
  """out[tap_o*s_tap + c3*s_c3 + c*s_c] += sum_p x3[p + d(tap)][c3] * y[p][c] on the CUDA-core kernel (any activation
  dtype; fp32 FMAs): x3 [N,H,W,3], y [N,H,W,C], kh in {1, 3}. Deterministic (block partials added in order)."""
  N, H, W, C = y.shape
  assert x3.dtype == y.dtype
  part = empty(N * ((H + 7) // 8) * ((W + 63) // 64) * kh * kh * 3 * C, F32)
  _call("xmc_wgrad_c3", ptr(x3), ptr(y), _f32(y), N, H, W, C, kh, kh, int(flip), s_tap, s_c3, s_c, ptr(out), ptr(part),
        stream(), launches=2)


# --------------------------------------------------------------------------------------------------------- batch norm
def _bn_desc(N, H, W, C, Hc, ldG, goff, boff, relu, upsample, f32=0):
  d = BnDesc()
  d.act_f32 = f32
  d.N, d.H, d.W, d.C, d.Hc = N, H, W, C, Hc
  d.ldG, d.goff, d.boff = ldG, goff, boff
  d.relu, d.upsample = int(relu), int(upsample)
  return d


_SMS = [0]


def partial_rows(units, per_sm=4):
  """Thread blocks (= partial rows) of a two-stage deterministic reduction over `units` work units: a few per SM."""
  if not _SMS[0]:
    _SMS[0] = _lib.lib().xmc_num_sms()
  return int(max(1, min(per_sm * _SMS[0], units)))


def bn_stats(x):
  C = x.shape[-1]
  P = x.numel() // C
  sums = empty(2 * C, F32)
  rows = partial_rows(P // 64)
  part = empty(rows * 2 * C, F32)
  _call("xmc_bn_stats", ptr(x), _f32(x), P, C, C, ptr(sums), ptr(part), rows, stream(), launches=2)
  return sums, P


def bn_finalize(sums, P, C, ra_mean, ra_var, new_ra_mean, new_ra_var, eps=1e-5, momentum=0.9):
  mr = empty(2 * C, F32)
  _call("xmc_bn_finalize", ptr(sums), P, C, eps, momentum, ptr(ra_mean), ptr(ra_var), ptr(new_ra_mean),
        ptr(new_ra_var), ptr(mr), stream())
  return mr


def bn_eval_stats(ra_mean, ra_var, C, eps=1e-5):
  mr = empty(2 * C, F32)
  _call("xmc_bn_eval_stats", ptr(ra_mean), ptr(ra_var), C, eps, ptr(mr), stream())
  return mr


def bn_apply(x, mr, gb, Hc, goff, boff, relu, upsample):
  N, H, W, C = x.shape
  assert gb.dtype == x.dtype
  d = _bn_desc(N, H, W, C, Hc, gb.stride(0), goff, boff, relu, upsample, _f32(x))
  y = empty((N, 2 * H, 2 * W, C) if upsample else (N, H, W, C), x.dtype)
  _call("xmc_bn_apply", ctypes.byref(d), ptr(x), ptr(mr), ptr(gb), ptr(y), stream())
  return y


def bn_bwd(dy, x, mr, gb, dgb, Hc, goff, boff, relu, upsample, group=None):
  """Returns dx (bf16, shape of x); writes d(gamma), d(beta) into the fp32 matrix dgb (same layout as gb).
  group: (process_group, size) of a cross-replica BatchNorm — the two per-channel sums of the backward are then
  all-reduced over the group (the transpose of the forward's pmean), see parallel.bn_group."""
  N, H, W, C = x.shape
  assert gb.stride(0) == dgb.stride(0) and gb.dtype == x.dtype == dy.dtype
  d = _bn_desc(N, H, W, C, Hc, gb.stride(0), goff, boff, relu, upsample, _f32(x))
  sums = empty(2 * C, F32)
  rows = partial_rows(N * Hc * Hc * 8, per_sm=8)   # the library clamps to the blocks it can fill
  part = empty(rows * 2 * C, F32)
  _call("xmc_bn_bwd_reduce", ctypes.byref(d), ptr(dy), ptr(x), ptr(mr), ptr(gb), ptr(dgb), ptr(sums), ptr(part), rows,
        stream(), launches=2)
  if group is not None:
    from . import parallel
    parallel.all_reduce_sum_(sums, group=group[0])
    d.replicas = group[1]
  dx = empty((N, H, W, C), x.dtype)
  _call("xmc_bn_bwd_apply", ctypes.byref(d), ptr(dy), ptr(x), ptr(mr), ptr(gb), ptr(sums), ptr(dx), stream())
  return dx


# ------------------------------------------------------------------------------------------------- pooling and misc
def pool2(a, b=None, low=None, scale=0.25, want_relu=False):
  N, H2, W2, C = a.shape
  out = empty((N, H2 // 2, W2 // 2, C), a.dtype)
  out_relu = empty((N, H2 // 2, W2 // 2, C), a.dtype) if want_relu else None
  _call("xmc_pool2", ptr(a), ptr(b), ptr(low), _f32(a), N, H2 // 2, W2 // 2, C, scale, ptr(out), ptr(out_relu),
        stream())
  return (out, out_relu) if want_relu else out


def relu(x):
  """y = max(x, 0) as a stand-alone pass (where no producing kernel can fold it in)."""
  y = torch.empty_like(x)
  _call("xmc_relu_or_add", ptr(x), None, _f32(x), x.numel(), ptr(y), stream())
  return y


def unpool2(dout, scale=0.25):
  N, H, W, C = dout.shape
  g = empty((N, 2 * H, 2 * W, C), dout.dtype)
  _call("xmc_unpool2", ptr(dout), _f32(dout), N, H, W, C, scale, ptr(g), stream())
  return g


def colsum(x, out, c=None):
  """out[c] += sum over all leading dims of x[..., c] (x bf16)."""
  C = x.shape[-1] if c is None else c
  _check_dense_rows(x)
  P = x.numel() // x.shape[-1]
  # small tensors: one block adds straight into `out`; larger ones use a machine-sized grid of partial rows
  rows = 1 if P * C <= 65536 else partial_rows((P + 63) // 64)
  part = empty(rows * C, F32) if rows > 1 else None
  _call("xmc_colsum", ptr(x), _f32(x), P, C, _pix_ld(x), ptr(out), ptr(part), rows, stream(),
        launches=2 if rows > 1 else 1)


def relu_sumhw(x):
  N, H, W, C = x.shape
  out = empty((N, C), F32)
  _call("xmc_relu_sumhw", ptr(x), _f32(x), N, H * W, C, ptr(out), stream())
  return out


def relu_sumhw_bwd(x, dout):
  N, H, W, C = x.shape
  dx = empty((N, H, W, C), x.dtype)
  _call("xmc_relu_sumhw_bwd", ptr(x), _f32(x), ptr(dout), N, H * W, C, ptr(dx), stream())
  return dx


def cast_to_bf16(src, dst=None):
  """2-D (rows, cols) fp32 -> the activation dtype (bf16 cast; in fp32 mode the tensor itself, or a device-to-device
  copy into `dst`), pitched views allowed."""
  rows, cols = src.shape
  if (dst.dtype if dst is not None else ACT[0]) == F32:
    if dst is None:
      return src
    dst.copy_(src)   # plumbing: device-to-device copy, no arithmetic
    return dst
  if dst is None:
    dst = empty((rows, cols), BF16)
  _call("xmc_cast_f32_to_bf16", ptr(src), rows, cols, src.stride(0), ptr(dst), dst.stride(0), stream())
  return dst


def cast_to_f32(src, dst=None, accumulate=False):
  rows, cols = src.shape
  if dst is None:
    dst = empty((rows, cols), F32)
  _call("xmc_cast_bf16_to_f32", ptr(src), rows, cols, src.stride(0), ptr(dst), dst.stride(0), int(accumulate),
        stream())
  return dst


def bcast_rows(src, reps, dst):
  """dst[b*reps + r, :] = src[b, :]; dst is a [B*reps, cols] pitched bf16 view."""
  B, cols = src.shape
  assert src.dtype == dst.dtype
  _call("xmc_bcast_rows", ptr(src), _f32(src), B, reps, cols, src.stride(0), ptr(dst), dst.stride(0), stream())


def sum_rows(src, B, reps, dst, accumulate=False):
  cols = src.shape[-1]
  _call("xmc_sum_rows", ptr(src), _f32(src), B, reps, cols, src.stride(0), ptr(dst), dst.stride(0), int(accumulate),
        stream())


def axpy(y, x, a=1.0):
  _call("xmc_axpy_f32", ptr(y), ptr(x), a, y.numel(), stream())


def colsum_f32(x, out):
  rows, cols = x.shape
  _call("xmc_colsum_f32", ptr(x), rows, cols, x.stride(0), ptr(out), stream())


# -------------------------------------------------------------------------------------------- attention / word loss
def l2norm_rows(x, out_dtype=F32, want_out=True, eps=1e-12):
  """x: [rows, D] (fp32 or bf16, pitched). Returns (xhat or None, invnorm)."""
  rows, D = x.shape
  y = empty((rows, D), out_dtype) if want_out else None
  inv = empty(rows, F32)
  _call("xmc_l2norm_rows", ptr(x), int(x.dtype == F32), rows, D, x.stride(0), ptr(y), int(out_dtype == F32),
        D, ptr(inv), eps, stream())
  return y, inv


def l2norm_rows_bwd(dxhat, xhat, inv, out=None, accumulate=False):
  rows, D = xhat.shape
  if out is None:
    out = empty((rows, D), dxhat.dtype)
  _call("xmc_l2norm_rows_bwd", ptr(dxhat), int(dxhat.dtype == F32), dxhat.stride(0), ptr(xhat),
        int(xhat.dtype == F32), xhat.stride(0), ptr(inv), rows, D, ptr(out), int(out.dtype == F32), out.stride(0),
        int(accumulate), stream())
  return out


def attention_g_fwd(q, what, max_len, gamma, ctx_out):
  """q: [B,R,D] bf16; what: [B,L,D] fp32 normalised words; ctx_out: [B*R, >=D] bf16 pitched view."""
  B, R, D = q.shape
  L = what.shape[1]
  attn = empty((B * R, L), F32)
  assert q.dtype == ctx_out.dtype
  _call("xmc_attention_g_fwd", ptr(q), _f32(q), q.stride(1), ptr(what), ptr(max_len), B, R, L, D, float(gamma),
        ptr(ctx_out), ctx_out.stride(0), ptr(attn), stream())
  return attn


def attention_g_bwd(dctx, q, what, attn, gamma):
  B, R, D = q.shape
  L = what.shape[1]
  dq = empty((B, R, D), q.dtype)
  assert dctx.dtype == q.dtype
  _call("xmc_attention_g_bwd", ptr(dctx), _f32(q), dctx.stride(0), ptr(q), q.stride(1), ptr(what), ptr(attn), B, R, L, D,
        float(gamma), ptr(dq), D, stream())
  return dq


def transpose_bf16(src, ld_dst):
  rows, cols = src.shape
  dst = empty((cols, ld_dst), src.dtype)
  _call("xmc_transpose_bf16", ptr(src), _f32(src), rows, cols, src.stride(0), ptr(dst), ld_dst, stream())
  return dst


# ----------------------------------------------------------------------------------------------------------- losses
def small_gemm_nt(A, B, scale):
  n, D = A.shape
  m = B.shape[0]
  C = empty((n, m), F32)
  _call("xmc_small_gemm_nt", ptr(A), ptr(B), n, m, D, scale, ptr(C), stream())
  return C


def small_gemm_nn(G, transposed, X, scale, out=None, accumulate=False):
  n = G.shape[1] if transposed else G.shape[0]
  m = G.shape[0] if transposed else G.shape[1]
  D = X.shape[1]
  if out is None:
    out = empty((n, D), F32)
  _call("xmc_small_gemm_nn", ptr(G), int(transposed), ptr(X), n, m, D, scale, ptr(out), int(accumulate), stream())
  return out


def ce_sym(logits, loss_slot, weight=1.0, want_grad=True):
  n = logits.shape[0]
  dl = empty((n, n), F32) if want_grad else None
  _call("xmc_ce_sym", ptr(logits), n, weight, ptr(loss_slot), ptr(dl), stream())
  return dl


def ce_stats(logits, out):
  """out[0] = accuracy, out[1] = entropy of get_statistics (attention_lib.py:36-43), both directions averaged."""
  _call("xmc_ce_stats", ptr(logits), logits.shape[0], ptr(out), stream())


def hinge(logit, B, d_slot, g_slot):
  dd = empty(2 * B, F32)
  dg = empty(2 * B, F32)
  _call("xmc_hinge", ptr(logit), B, ptr(d_slot), ptr(g_slot), ptr(dd), ptr(dg), stream())
  return dd, dg
