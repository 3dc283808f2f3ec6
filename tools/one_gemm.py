"""GPU debugging aid: runs one conv_fwd (or, with the `wgrad` flag, one weight-gradient) shape in a loop (for ncu).
python tools/one_gemm.py N H W C k Cout [res] [relu] [wgrad] [pair]
pair: the fp32-activation chain form of the ResNet trunk (two-part bf16 operand in, two-part result out, two-part
residual, split weights, no fp32 tensor)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from xmcgan_image_generation_b200 import ops

N, H, W, C, k, Cout = (int(a) for a in sys.argv[1:7])
flags = sys.argv[7:]
x = (torch.randn(N, H, W, C, device="cuda") * 0.5).to(torch.bfloat16)
wk = (torch.randn(Cout, k * k * C, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(Cout, device="cuda")
res = (torch.randn(N, H, W, Cout, device="cuda")).to(torch.bfloat16) if "res" in flags else None
if "wgrad" in flags:
  dy = (torch.randn(N, H, W, Cout, device="cuda") * 0.1).to(torch.bfloat16)
  dw = torch.zeros(k * k * C * Cout, device="cuda")
  run = lambda: ops.wgrad(x, dy, k, dw, out_mode=0, ld_out=Cout, tap_stride=C * Cout)
  for _ in range(3):
    run()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(10):
    run()
  e.record()
  torch.cuda.synchronize()
  ms = s.elapsed_time(e) / 10
  print(f"wgrad {ms*1e3:.1f} us  {2.0 * N * H * W * k * k * C * Cout / ms / 1e9:.1f} TFLOP/s")
  sys.exit(0)
if "pair" in flags:
  xp = (torch.randn(N, H, W, 2 * C, device="cuda") * 0.5).to(torch.bfloat16)
  wk3 = (torch.randn(Cout, k * k * 3 * C, device="cuda") * 0.05).to(torch.bfloat16)
  resp = torch.randn(N, H, W, 2 * Cout, device="cuda").to(torch.bfloat16) if "res" in flags else None
  run = lambda: ops.conv_fwd(None, wk3, k, Cout, bias=bias, relu="relu" in flags, ldb=k * k * 3 * C, x_pair=xp,
                             want_pair=True, want_f32=False, residual_pair=resp)
  for _ in range(3):
    run()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(10):
    run()
  e.record()
  torch.cuda.synchronize()
  ms = s.elapsed_time(e) / 10
  by = (xp.numel() + N * H * W * 2 * Cout + (resp.numel() if resp is not None else 0)) * 2
  print(f"pair {ms*1e3:.1f} us  {2.0 * N * H * W * k * k * C * Cout / ms / 1e9:.1f} TFLOP/s (algorithmic)  "
        f"{by/ms/1e6:.1f} GB/s (algorithmic bytes)")
  sys.exit(0)
for _ in range(3):
  y = ops.conv_fwd(x, wk, k, Cout, bias=bias, residual=res, relu="relu" in flags)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
  y = ops.conv_fwd(x, wk, k, Cout, bias=bias, residual=res, relu="relu" in flags)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
fl = 2.0 * N * H * W * k * k * C * Cout
by = (x.numel() + y.numel() + (res.numel() if res is not None else 0)) * 2
print(f"{ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s  {by/ms/1e6:.1f} GB/s (algorithmic bytes)")
