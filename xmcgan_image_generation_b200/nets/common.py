"""Counterparts of the resampling helpers of xmcgan/nets/common.py on CUDA bf16 NHWC tensors. The residual blocks
themselves (GenBlock, GenSpatialBlock, DiscBlock, DiscOptimizedBlock; common.py:58-186) are executed by engine.py as
fused kernel programs and are reachable through nets.xmc_net.Generator / Discriminator."""
import torch

from .. import ops


def _bf16(x):
  return torch.as_tensor(x).to("cuda", torch.bfloat16).contiguous()


def dsample(x):
  """common.dsample / tensorflow_style_avg_pooling (common.py:23-45,54-55): 2x2 stride-2 mean."""
  x = _bf16(x)
  if x.shape[1] % 2 or x.shape[2] % 2:
    raise ValueError("dsample needs even spatial sizes")
  if x.shape[3] % 8:
    out = ops.empty((x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]), x.dtype)
    ops._call("xmc_pool2_small", x.data_ptr(), ops._f32(x), x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3],
              0.25, out.data_ptr(), ops._lib.stream())
    return out
  return ops.pool2(x, scale=0.25)


def upsample(x, factor=2):
  """common.upsample (common.py:48-51): nearest, out[i] = in[i // 2]."""
  if factor != 2:
    raise NotImplementedError("only factor=2 is used by the reference networks")
  x = _bf16(x)
  if x.shape[3] % 8:
    raise ValueError("channel count must be a multiple of 8")
  return ops.unpool2(x, scale=1.0)
