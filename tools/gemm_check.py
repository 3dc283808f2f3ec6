"""GPU debugging aid (not a test): runs the tcgen05 GEMM entry points on a list of shapes and prints errors against
torch fp32 on the same bf16-rounded inputs. Usage: python tools/gemm_check.py [group ...]"""
import ctypes
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from xmcgan_image_generation_b200 import _lib

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def conv_fwd(x, wk, KH, KW, bias=None, residual=None, mask=None, relu=0, res_shift=0, out_dtype=1, alpha=1.0,
             batched=False):
  """x [N,H,W,C] bf16; wk [Cout, KH*KW*C] bf16 (or [N,Cout,K] if batched)."""
  L = _lib.lib()
  N, H, W, C = x.shape
  Cout = wk.shape[-2]
  d = _lib.ConvDesc()
  d.N, d.H, d.W, d.C, d.ldA = N, H, W, C, x.stride(2)
  d.KH, d.KW, d.pad_h, d.pad_w = KH, KW, KH // 2, KW // 2
  d.Cout, d.ldB = Cout, wk.stride(-2)
  d.batched = 1 if batched else 0
  d.strideB_batch = wk.stride(0) if batched else 0
  d.out_dtype = out_dtype
  y = torch.full((N, H, W, Cout), float("nan"), device=dev, dtype=torch.float32 if out_dtype else torch.bfloat16)
  d.ldOut = Cout
  d.alpha = alpha
  d.relu = relu
  d.res_shift = res_shift
  d.ldRes = residual.stride(2) if residual is not None else 0
  d.ldMask = mask.stride(2) if mask is not None else 0
  _lib.check(L.xmc_conv2d_fwd(ctypes.byref(d), _lib.ptr(x), _lib.ptr(wk), _lib.ptr(bias), _lib.ptr(residual),
                              _lib.ptr(mask), _lib.ptr(y), None, _lib.stream()))
  torch.cuda.synchronize()
  return y


def wgrad(xa, xb, KH, KW, out_mode=0, batched=False, alpha=1.0):
  L = _lib.lib()
  N, H, W, Ca = xa.shape
  Cb = xb.shape[-1]
  d = _lib.WgradDesc()
  d.N, d.H, d.W = N, H, W
  d.Ca, d.ldA, d.Cb, d.ldB = Ca, xa.stride(2), Cb, xb.stride(2)
  d.KH, d.KW, d.pad_h, d.pad_w = KH, KW, KH // 2, KW // 2
  d.batched = 1 if batched else 0
  d.out_mode = out_mode
  d.ldOut = Cb
  d.out_tap_stride = Ca * Cb
  d.out_batch_stride = KH * KW * Ca * Cb
  d.alpha = alpha
  nb = N if batched else 1
  dt = torch.bfloat16 if out_mode == 2 else torch.float32
  out = torch.zeros((nb, KH * KW, Ca, Cb), device=dev, dtype=dt)
  need = ctypes.c_longlong(0)
  _lib.check(L.xmc_conv2d_wgrad_workspace_bytes(ctypes.byref(d), ctypes.byref(need)))
  ws = torch.empty(max(need.value, 16), device=dev, dtype=torch.uint8)
  _lib.check(L.xmc_conv2d_wgrad(ctypes.byref(d), _lib.ptr(xa), _lib.ptr(xb), _lib.ptr(out), _lib.ptr(ws), need.value,
                                _lib.stream()))
  torch.cuda.synchronize()
  return out


def report(name, got, ref):
  got = got.float()
  ref = ref.float()
  err = (got - ref).abs().max().item()
  scale = ref.abs().max().item() + 1e-6
  nan = torch.isnan(got).sum().item()
  ok = (err / scale < 2e-2) and nan == 0
  print(f"{'PASS' if ok else 'FAIL'} {name}: max_abs_err={err:.4g} ref_max={scale:.4g} rel={err/scale:.3g} nan={nan}",
        flush=True)
  if not ok and got.dim() >= 2:
    g2 = got.reshape(-1, got.shape[-1])
    r2 = ref.reshape(-1, ref.shape[-1])
    bad = ((g2 - r2).abs() > 2e-2 * scale) | torch.isnan(g2)
    rows = bad.any(1).nonzero().flatten()
    cols = bad.any(0).nonzero().flatten()
    print(f"   bad rows: {rows.numel()}/{g2.shape[0]} first {rows[:12].tolist()}  bad cols: {cols.numel()}/{g2.shape[1]}"
          f" first {cols[:12].tolist()}")
    print("   got[0,:8]", g2[0, :8].tolist())
    print("   ref[0,:8]", r2[0, :8].tolist())
  return ok


def ref_conv(x, wk, KH, KW):
  N, H, W, C = x.shape
  Cout = wk.shape[0]
  w4 = wk.float().reshape(Cout, KH, KW, C).permute(0, 3, 1, 2).contiguous()
  y = F.conv2d(x.float().permute(0, 3, 1, 2), w4, padding=(KH // 2, KW // 2))
  return y.permute(0, 2, 3, 1).contiguous()


def rnd(*shape, scale=1.0):
  return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def group_gemm():
  torch.manual_seed(0)
  for (M, K, Cout) in [(128, 64, 64), (256, 128, 128), (384, 256, 256), (1000, 96, 96), (300, 768, 952), (64, 8, 24),
                       (56, 256, 10752)]:
    x = rnd(1, 1, M, K)
    wk = rnd(Cout, K, scale=0.1)
    y = conv_fwd(x, wk, 1, 1)
    report(f"gemm M={M} K={K} N={Cout}", y[0, 0], x[0, 0].float() @ wk.float().t())
  # dense layout: pixels along N (H=W=1)
  x = rnd(56, 1, 1, 768)
  wk = rnd(128, 768, scale=0.1)
  y = conv_fwd(x, wk, 1, 1)
  report("dense N=56 K=768 Cout=128", y[:, 0, 0], x[:, 0, 0].float() @ wk.float().t())


def group_conv():
  torch.manual_seed(1)
  for (N, H, W, C, Cout) in [(2, 16, 16, 64, 64), (1, 4, 128, 64, 96), (2, 64, 64, 96, 192), (4, 8, 8, 128, 256),
                             (16, 4, 4, 192, 64), (3, 32, 32, 96, 96), (5, 8, 8, 64, 32), (2, 4, 4, 16, 16),
                             (1, 128, 128, 96, 96)]:
    x = rnd(N, H, W, C)
    wk = rnd(Cout, 9 * C, scale=0.05)
    y = conv_fwd(x, wk, 3, 3)
    report(f"conv3x3 N={N} H={H} W={W} C={C} Cout={Cout}", y, ref_conv(x, wk, 3, 3))


def group_epilogue():
  torch.manual_seed(2)
  N, H, W, C, Cout = 2, 16, 16, 64, 96
  x = rnd(N, H, W, C)
  wk = rnd(Cout, 9 * C, scale=0.05)
  bias = torch.randn(Cout, device=dev)
  ref = ref_conv(x, wk, 3, 3)
  y = conv_fwd(x, wk, 3, 3, bias=bias)
  report("epi bias", y, ref + bias)
  y = conv_fwd(x, wk, 3, 3, bias=bias, relu=1, out_dtype=0)
  report("epi bias+relu bf16", y, torch.relu(ref + bias))
  mask = rnd(N, H, W, Cout)
  res = rnd(N, H, W, Cout)
  y = conv_fwd(x, wk, 3, 3, mask=mask, residual=res, alpha=0.5)
  report("epi mask+res+alpha", y, 0.5 * ref * (mask.float() > 0) + res.float())
  res2 = rnd(N, H // 2, W // 2, Cout)
  y = conv_fwd(x, wk, 3, 3, bias=bias, residual=res2, res_shift=1, out_dtype=0)
  up = res2.float().repeat_interleave(2, 1).repeat_interleave(2, 2)
  report("epi bias+upsampled residual bf16", y, ref + bias + up)
  # channel-sliced input view (ldA > C)
  xb = rnd(N, H, W, 128)
  xs = xb[..., 64:128]
  y = conv_fwd(xs, wk, 3, 3)
  report("sliced input ldA=128", y, ref_conv(xs.contiguous(), wk, 3, 3))


def group_batched():
  torch.manual_seed(3)
  Bn, M, K, Cout = 5, 256, 768, 952
  x = rnd(Bn, 1, M, K, scale=0.1)
  wk = rnd(Bn, Cout, K, scale=0.1)
  y = conv_fwd(x, wk, 1, 1, batched=True)
  ref = torch.einsum("bmk,bnk->bmn", x[:, 0].float(), wk.float())
  report("batched gemm", y[:, 0], ref)


def ref_wgrad(xa, xb, KH, KW):
  N, H, W, Ca = xa.shape
  Cb = xb.shape[-1]
  ph, pw = KH // 2, KW // 2
  xp = F.pad(xa.float(), (0, 0, pw, pw, ph, ph))
  out = torch.zeros(KH * KW, Ca, Cb, device=dev)
  for kh in range(KH):
    for kw in range(KW):
      sl = xp[:, kh:kh + H, kw:kw + W, :]
      out[kh * KW + kw] = torch.einsum("nhwa,nhwb->ab", sl, xb.float())
  return out


def group_wgrad():
  torch.manual_seed(4)
  for (N, H, W, Ca, Cb, KH) in [(2, 16, 16, 64, 64, 3), (1, 8, 8, 128, 128, 1), (2, 64, 64, 96, 192, 3),
                                (4, 8, 8, 256, 128, 3), (16, 4, 4, 192, 64, 3), (2, 128, 128, 96, 96, 3),
                                (56, 1, 1, 256, 1536, 1), (8, 1, 1, 768, 128, 1), (3, 32, 32, 1024, 4224, 1)]:
    xa = rnd(N, H, W, Ca)
    xb = rnd(N, H, W, Cb, scale=0.1)
    out = wgrad(xa, xb, KH, KH)
    report(f"wgrad N={N} H={H} W={W} Ca={Ca} Cb={Cb} k={KH}", out[0], ref_wgrad(xa, xb, KH, KH))


def group_wgrad_batched():
  torch.manual_seed(5)
  Bn, R, M, Nf = 4, 256, 952, 768
  xa = rnd(Bn, 1, R, M, scale=0.1)   # alpha [i][r][jw]
  xb = rnd(Bn, 1, R, Nf)             # regions [i][r][f]
  out = wgrad(xa, xb, 1, 1, out_mode=1, batched=True)
  ref = torch.einsum("brm,brn->bmn", xa[:, 0].float(), xb[:, 0].float())
  report("wgrad batched store fp32", out[:, 0], ref)
  out = wgrad(xa, xb, 1, 1, out_mode=2, batched=True)
  report("wgrad batched store bf16", out[:, 0], ref)


GROUPS = dict(gemm=group_gemm, conv=group_conv, epilogue=group_epilogue, batched=group_batched, wgrad=group_wgrad,
              wgrad_batched=group_wgrad_batched)

if __name__ == "__main__":
  names = sys.argv[1:] or list(GROUPS)
  for n in names:
    print(f"== {n}", flush=True)
    GROUPS[n]()
