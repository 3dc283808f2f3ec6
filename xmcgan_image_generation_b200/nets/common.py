"""Counterparts of xmcgan/nets/common.py on CUDA NHWC tensors: the resampling helpers (common.py:23-55) and the
residual blocks as module-level callables with the reference's constructor arguments and Flax call surface
(DiscBlock :58-79, DiscOptimizedBlock :117-133, GenBlock :136-160, GenSpatialBlock :163-186), written as the reference
writes them (relu, upsample / dsample, conv_fn / dense_fn / norm_fn sub-modules with Flax auto-naming). Forward values;
training runs the same mathematics as fused forward + backward kernel programs in engine.py (sub-pixel convolutions,
shortcut before the upsample, pooling after the residual add — exact rearrangements), which these blocks are tested
against."""
import torch

from .. import ops
from ..libml import layers


def _bf16(x):
  """Plumbing: accepts CUDA bf16 / fp32 tensors as they are, anything else is moved to the device as bf16."""
  x = torch.as_tensor(x)
  if x.is_cuda and x.dtype in (torch.bfloat16, torch.float32):
    return x.contiguous()
  return x.to("cuda", torch.bfloat16).contiguous()


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


class DiscBlock(layers.Module):
  """common.DiscBlock (common.py:58-79)."""

  def __init__(self, filters, downsample, conv_fn, activation_fn=layers.relu, dtype=None, name=None):
    self.filters, self.downsample, self.conv_fn, self.activation_fn, self.name = (filters, downsample, conv_fn,
                                                                                   activation_fn, name)

  def forward(self, scope, x):
    needs_projection = self.downsample or x.shape[-1] != self.filters
    x0 = x
    x = self.activation_fn(x)
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3)), x)
    x = self.activation_fn(x)
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3)), x)
    if needs_projection:
      x0 = scope.call(self.conv_fn(self.filters, kernel_size=(1, 1)), x0)
    if self.downsample:
      x = dsample(x)
      x0 = dsample(x0)
    return layers.add(x0, x)


class DiscOptimizedBlock(layers.Module):
  """common.DiscOptimizedBlock (common.py:117-133); the 3-channel image is zero-padded to 8 channels for the
  tensor-core convolution of this module-level form (the padded channels meet zero weights rows: exact)."""

  def __init__(self, filters, conv_fn, activation_fn=layers.relu, dtype=None, name=None):
    self.filters, self.conv_fn, self.activation_fn, self.name = filters, conv_fn, activation_fn, name

  def forward(self, scope, x):
    x0 = x
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3)), x)
    x = self.activation_fn(x)
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3)), x)
    x = dsample(x)
    x0 = dsample(x0)
    x0 = scope.call(self.conv_fn(self.filters, kernel_size=(1, 1)), x0)
    return layers.add(x, x0)


class GenBlock(layers.Module):
  """common.GenBlock (common.py:136-160)."""

  def __init__(self, filters, conv_fn, dense_fn, norm_fn, activation_fn=layers.relu, dtype=None, name=None):
    self.filters, self.conv_fn, self.dense_fn, self.norm_fn = filters, conv_fn, dense_fn, norm_fn
    self.activation_fn, self.name = activation_fn, name

  def forward(self, scope, x, cond):
    x0 = x
    x = scope.call(layers.ConditionalBatchNorm(norm_fn=self.norm_fn, dense_fn=self.dense_fn), x, cond)
    x = self.activation_fn(x)
    x = upsample(x)
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3), use_bias=True), x)
    x = scope.call(layers.ConditionalBatchNorm(norm_fn=self.norm_fn, dense_fn=self.dense_fn), x, cond)
    x = self.activation_fn(x)
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3), use_bias=True), x)
    x0 = upsample(x0)
    x0 = scope.call(self.conv_fn(self.filters, kernel_size=(1, 1), use_bias=True), x0)
    return layers.add(x, x0)


class GenSpatialBlock(layers.Module):
  """common.GenSpatialBlock (common.py:163-186)."""

  def __init__(self, filters, conv_fn, dense_fn, norm_fn, activation_fn=layers.relu, dtype=None, name=None):
    self.filters, self.conv_fn, self.dense_fn, self.norm_fn = filters, conv_fn, dense_fn, norm_fn
    self.activation_fn, self.name = activation_fn, name

  def forward(self, scope, x, cond0, cond1):
    x0 = x
    x = scope.call(layers.LocalConditionalBatchNorm(norm_fn=self.norm_fn, conv_fn=self.conv_fn), x, cond0)
    x = self.activation_fn(x)
    x = upsample(x)
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3), use_bias=True), x)
    x = scope.call(layers.LocalConditionalBatchNorm(norm_fn=self.norm_fn, conv_fn=self.conv_fn), x, cond1)
    x = self.activation_fn(x)
    x = scope.call(self.conv_fn(self.filters, kernel_size=(3, 3), use_bias=True), x)
    x0 = upsample(x0)
    x0 = scope.call(self.conv_fn(self.filters, kernel_size=(1, 1), use_bias=True), x0)
    return layers.add(x, x0)
