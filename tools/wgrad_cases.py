"""GPU debugging aid: single weight-gradient launches of the shape classes of a train_step (for ncu / timing).
python tools/wgrad_cases.py [case ...]   cases: c3 sp2 sp1 plain tap3 small"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from xmcgan_image_generation_b200 import ops

cases = sys.argv[1:] or ["c3", "sp2", "sp1", "plain", "tap3", "small"]
bf = torch.bfloat16


def rnd(*shape):
  return (torch.randn(*shape, device="cuda") * 0.1).to(bf)


def run(name, fn, flops):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(10):
    fn()
  e.record()
  torch.cuda.synchronize()
  ms = s.elapsed_time(e) / 10
  print(f"{name}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s")


for c in cases:
  if c == "c3":     # D block 0 conv1: 3 -> 96 at 128x128, packed-window form
    x3, y = rnd(112, 128, 128, 3), rnd(112, 128, 128, 96)
    xpad = ops.c3_pad(x3)
    out = torch.zeros(27 * 96, device="cuda")
    run(c, lambda: ops.c3_wgrad(xpad, y, False, 3 * 96, 96, 1, out), 2.0 * 112 * 128 * 128 * 27 * 96)
  elif c == "sp2":  # D block 0 conv2 + pool: x 128x128x96, dy 64x64x96
    x, dy = rnd(112, 128, 128, 96), rnd(112, 64, 64, 96)
    out = torch.zeros(9 * 96 * 96, device="cuda")
    run(c, lambda: ops.wgrad(x, dy, 3, out, out_mode=0, ld_out=96, tap_stride=96 * 96, alpha=0.25, subpixel=2),
        2.0 * 112 * 64 * 64 * 16 * 96 * 96)
  elif c == "sp1":  # G block 4 conv1 (sub-pixel): x 64x64x192, dy 128x128x96
    x, dy = rnd(56, 64, 64, 192), rnd(56, 128, 128, 96)
    out = torch.zeros(9 * 192 * 96, device="cuda")
    run(c, lambda: ops.wgrad(x, dy, 3, out, out_mode=0, ld_out=96, tap_stride=192 * 96, subpixel=1),
        2.0 * 56 * 64 * 64 * 16 * 192 * 96)
  elif c == "plain":
    x, dy = rnd(112, 32, 32, 192), rnd(112, 32, 32, 384)
    out = torch.zeros(9 * 192 * 384, device="cuda")
    run(c, lambda: ops.wgrad(x, dy, 3, out, out_mode=0, ld_out=384, tap_stride=192 * 384),
        2.0 * 112 * 32 * 32 * 9 * 192 * 384)
  elif c == "tap3":
    x, dy = rnd(56, 128, 128, 96), rnd(56, 128, 128, 96)
    out = torch.zeros(9 * 96 * 96, device="cuda")
    run(c, lambda: ops.wgrad(x, dy, 3, out, out_mode=0, ld_out=96, tap_stride=96 * 96),
        2.0 * 56 * 128 * 128 * 9 * 96 * 96)
  elif c in ("wide", "wide2", "wide3"):
    H, ca, cb = {"wide": (16, 384, 768), "wide2": (8, 768, 1536), "wide3": (8, 1536, 1536)}[c]
    x, dy = rnd(112, H, H, ca), rnd(112, H, H, cb)
    out = torch.zeros(9 * ca * cb, device="cuda")
    run(c, lambda: ops.wgrad(x, dy, 3, out, out_mode=0, ld_out=cb, tap_stride=ca * cb), 2.0 * 112 * H * H * 9 * ca * cb)
  elif c == "small":
    x, dy = rnd(112, 4, 4, 1536), rnd(112, 4, 4, 1536)
    out = torch.zeros(9 * 1536 * 1536, device="cuda")
    run(c, lambda: ops.wgrad(x, dy, 3, out, out_mode=0, ld_out=1536, tap_stride=1536 * 1536),
        2.0 * 112 * 4 * 4 * 9 * 1536 * 1536)
