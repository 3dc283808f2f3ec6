"""GPU debugging aid: the BatchNorm kernels (and colsum) at the generator's layer shapes, time vs HBM time of the bytes
they must move. python tools/bn_bench.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from xmcgan_image_generation_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 56
PEAK = 6543.4e9


def timeit(fn, n=5):
  """Sum of the CUPTI kernel durations of one call (us), per kernel name."""
  from torch.profiler import ProfilerActivity, profile
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
      fn()
    torch.cuda.synchronize()
  per = {}
  for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and "xmc::" in e.name:
      k = e.name.split("xmc::")[1].split("<")[0].split("(")[0]
      per[k] = per.get(k, 0.0) + (e.time_range.end - e.time_range.start) / n
  return sum(per.values()), per


# (H, C, Hc, upsample): the generator's conditional BatchNorms at gf = 96
layers = [(4, 1536, 1, True), (8, 1536, 1, False), (8, 1536, 1, True), (16, 768, 1, False), (16, 768, 16, True),
          (32, 384, 16, False), (32, 384, 16, True), (64, 192, 16, False), (64, 192, 16, True), (128, 96, 16, False),
          (128, 96, 16, False)]
tot = {}
print("| layer | kernel | us | HBM-time us | frac |")
print("|---|---|---:|---:|---:|")
for H, C, Hc, up in layers:
  x = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
  gb = (torch.randn(B * Hc * Hc, 2 * C, device="cuda") * 0.1).to(torch.bfloat16)
  dgb = torch.zeros(B * Hc * Hc, 2 * C, device="cuda")
  Ho = 2 * H if up else H
  dy = torch.randn(B, Ho, Ho, C, device="cuda").to(torch.bfloat16)
  sums, P = ops.bn_stats(x)
  mr = ops.bn_finalize(sums, P, C, None, None, None, None) if hasattr(ops, "bn_finalize") else None
  n = x.numel()
  cases = [("bn_stats", lambda: ops.bn_stats(x), 2 * n),
           ("bn_apply", lambda: ops.bn_apply(x, mr, gb, Hc, 0, C, True, up), n * (2 + (8 if up else 2))),
           ("bn_bwd", lambda: ops.bn_bwd(dy, x, mr, gb, dgb, Hc, 0, C, True, up),
            n * ((8 if up else 2) + 2) + n * ((8 if up else 2) + 2 + 2)),
           ("colsum", lambda: ops.colsum(dy, torch.zeros(C, device="cuda")), dy.numel() * 2)]
  for name, fn, nbytes in cases:
    us, per = timeit(fn)
    ideal = nbytes / PEAK * 1e6
    a = tot.setdefault(name, [0.0, 0.0])
    a[0] += us
    a[1] += ideal
    detail = " ".join(f"{k.replace('_kernel','')}={v:.1f}" for k, v in per.items())
    print(f"| {H}x{H}x{C} Hc={Hc} up={int(up)} | {name} | {us:.1f} | {ideal:.1f} | {ideal/us:.2f} | {detail} |")
for k, v in tot.items():
  print(f"| all | {k} | {v[0]:.1f} | {v[1]:.1f} | {v[1]/v[0]:.2f} |")
