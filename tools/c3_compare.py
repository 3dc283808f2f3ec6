"""GPU debugging aid: the 3 -> C image-side 3x3 convolution, CUDA-core kernel vs packed-window tensor-core form."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from xmcgan_image_generation_b200 import _lib, ops

N, S, C = int(sys.argv[1]) if len(sys.argv) > 1 else 112, 128, 96
img = (torch.randn(N, S, S, 3, device="cuda") * 0.5).to(torch.bfloat16)
wk = (torch.randn(C, 27, device="cuda") * 0.1).to(torch.bfloat16)
bias = torch.randn(C, device="cuda")


def timeit(fn):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(10):
    fn()
  e.record()
  torch.cuda.synchronize()
  return s.elapsed_time(e) / 10 * 1e3


y0 = ops.empty((N, S, S, C))
cuda_core = lambda: ops._call("xmc_conv_c3_in", img.data_ptr(), 0, wk.data_ptr(), 27, bias.data_ptr(), N, S, S, C, 3, 3, 1,
                              y0.data_ptr(), _lib.stream())
xpad = ops.c3_pad(img)
wp = ops.c3_pack_weights(wk, 27, C)
y1 = [None]


def tensor():
  y1[0] = ops.c3_conv(xpad, wp, C, bias=bias, relu=True)


print(f"cuda-core {timeit(cuda_core):.1f} us   tensor-core (packed window) {timeit(tensor):.1f} us   "
      f"pad {timeit(lambda: ops.c3_pad(img)):.1f} us")
print("max |diff|", (y0.float() - y1[0].float()).abs().max().item(), "of", y0.float().abs().max().item())
