"""A numpy-backed, apply-only stand-in for the slice of `flax.linen` / `jax` that the reference's network code uses
(xmcgan/nets/xmc_net.py, xmcgan/nets/common.py, xmcgan/libml/layers.py, xmcgan/utils/resnet_v1.py), so that THOSE FILES can be executed from
/root/reference without JAX / Flax (neither is installable here). Test infrastructure, used only by
tests/golden/make_reference_golden.py.

What is the reference's own code when a network runs on this stand-in: the whole model wiring — block order, which
tensor feeds which layer, conditional / local-conditional BatchNorm modulation, the spectral-norm power iteration and
sigma of layers.SpectralConv / SpectralDense, attention_for_g and its mask, the projection logit, the contrastive
statistics, up / down-sampling calls, every reshape / tile / concatenate — and Flax's auto-naming of sub-modules and
parameters (a `self.param` looks its value up under the path the call order produces: a naming mismatch is a KeyError).

What is restated here (from flax 0.3.3 / jax semantics, the same statements oracle/xmc_oracle.py makes): the
primitives nn.Conv (NHWC x HWIO, XLA's SAME padding, any stride), nn.max_pool, nn.Dense, nn.BatchNorm (mean / mean-of-squares statistics, running
averages momentum * old + (1 - momentum) * new, rsqrt(var + eps)), lax.conv_general_dilated / dot_general,
jax.image.resize(nearest, integer factor), lax.reduce_window(add) 2x2 / stride 2, stop_gradient = identity (forward).
All arithmetic is numpy float32."""
import collections
import types

import numpy as np

_STACK = []   # modules whose __call__ is executing, innermost last
DTYPE = [np.float32]   # arithmetic type of the primitives; the finite-difference gradient run switches it to float64
# jax.lax.stop_gradient under finite differences: the base run RECORDS every stopped value, the perturbed runs REPLAY
# them, so that the difference quotient treats them as the constants jax.grad sees (d stop_gradient(u(w)) / dw = 0)
SG = {"mode": None, "tape": [], "pos": 0}


def stop_gradient(x):
  if SG["mode"] == "record":
    SG["tape"].append(np.array(x, copy=True))
    return x
  if SG["mode"] == "replay":
    v = SG["tape"][SG["pos"]]
    SG["pos"] += 1
    assert v.shape == np.shape(x)
    return v
  return x


def _same_pads(n, k, s):
  """XLA 'SAME': out = ceil(n / s), total = max((out - 1) * s + k - n, 0), low = total // 2."""
  total = max((-(-n // s) - 1) * s + k - n, 0)
  return total // 2, total - total // 2


def conv2d_same(x, kernel, strides=(1, 1)):
  """NHWC x HWIO -> NHWC, SAME padding as XLA computes it, any stride."""
  kh, kw, cin, cout = kernel.shape
  x = np.asarray(x, DTYPE[0])
  xp = np.pad(x, ((0, 0), _same_pads(x.shape[1], kh, strides[0]), _same_pads(x.shape[2], kw, strides[1]), (0, 0)))
  win = np.lib.stride_tricks.sliding_window_view(xp, (kh, kw), axis=(1, 2))[:, ::strides[0], ::strides[1]]
  return np.einsum("nhwcij,ijco->nhwo", win, np.asarray(kernel, DTYPE[0]), optimize=True).astype(DTYPE[0])


def max_pool(x, window_shape, strides=None, padding="VALID"):
  """flax.linen.max_pool on NHWC: lax.reduce_window(max) with -inf padding."""
  strides = strides or (1, 1)
  assert padding == "SAME"
  xp = np.pad(np.asarray(x, DTYPE[0]), ((0, 0), _same_pads(x.shape[1], window_shape[0], strides[0]),
                                          _same_pads(x.shape[2], window_shape[1], strides[1]), (0, 0)),
              constant_values=-np.inf)
  win = np.lib.stride_tricks.sliding_window_view(xp, tuple(window_shape), axis=(1, 2))[:, ::strides[0], ::strides[1]]
  return win.max(axis=(-2, -1))


class _Ctx:
  def __init__(self, variables, mutable):
    self.variables, self.mutable, self.mutated = variables, set(mutable or ()), {}

  def get(self, col, path):
    for src in (self.mutated, self.variables):
      node = src.get(col)
      for p in path:
        node = node.get(p) if isinstance(node, dict) else None
      if node is not None:
        return node
    raise KeyError(f"no variable {col}/{'/'.join(path)}")

  def put(self, col, path, value):
    if col not in self.mutable:
      raise ValueError(f"collection {col} is not mutable")
    node = self.mutated.setdefault(col, {})
    for p in path[:-1]:
      node = node.setdefault(p, {})
    node[path[-1]] = value


class _Variable:
  def __init__(self, ctx, col, path):
    self._ctx, self._col, self._path = ctx, col, path

  @property
  def value(self):
    return self._ctx.get(self._col, self._path)

  @value.setter
  def value(self, v):
    self._ctx.put(self._col, self._path, np.asarray(v, DTYPE[0]))


def compact(fn):
  return fn


class Module:
  """flax.linen.Module, apply-only: dataclass-style fields, `setup`, compact `__call__`, auto-named sub-modules."""
  _fields = ()

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    fields = []
    for klass in reversed(cls.__mro__):
      for name in klass.__dict__.get("__annotations__", {}):
        if name not in fields:
          fields.append(name)
    cls._fields = tuple(fields)
    if "__call__" in cls.__dict__:
      inner = cls.__dict__["__call__"]

      def call(self, *args, __inner=inner, **kwargs):
        if not self._setup_done:
          self._setup_done = True
          self.setup()
        _STACK.append(self)
        try:
          return __inner(self, *args, **kwargs)
        finally:
          _STACK.pop()

      cls.__call__ = call

  def __init__(self, *args, **kwargs):
    name = kwargs.pop("name", None)
    kwargs.pop("parent", None)
    for f, v in zip(self._fields, args):
      object.__setattr__(self, f, v)
    for f in self._fields[len(args):]:
      if f in kwargs:
        object.__setattr__(self, f, kwargs.pop(f))
      elif hasattr(type(self), f):
        object.__setattr__(self, f, getattr(type(self), f))   # the class-level default, stored on the instance
      else:
        raise TypeError(f"{type(self).__name__}: missing field {f}")
    if kwargs:
      raise TypeError(f"{type(self).__name__}: unexpected fields {sorted(kwargs)}")
    self._setup_done, self._counters = False, collections.Counter()
    if _STACK:   # constructed inside a parent's compact method: bound and named now, in construction order
      parent = _STACK[-1]
      if name is None:
        name = f"{type(self).__name__}_{parent._counters[type(self).__name__]}"
        parent._counters[type(self).__name__] += 1
      self._path, self._ctx = parent._path + (name,), parent._ctx
    else:
      self._path, self._ctx = (), None
    self.name = name

  def setup(self):
    pass

  def param(self, name, init_fn, *init_args):
    return np.asarray(self._ctx.get("params", self._path + (name,)), DTYPE[0])

  def variable(self, col, name, init_fn=None, *init_args):
    return _Variable(self._ctx, col, self._path + (name,))

  def make_rng(self, name):
    raise RuntimeError("apply-only stand-in: variables must be given")

  def apply(self, variables, *args, mutable=False, rngs=None, **kwargs):
    self._ctx = _Ctx(variables, mutable if mutable else ())
    self._path = ()
    out = self(*args, **kwargs)
    return (out, self._ctx.mutated) if mutable else out


class Conv(Module):
  features: int
  kernel_size: tuple
  strides: tuple = None
  padding: str = "SAME"
  input_dilation: tuple = None
  kernel_dilation: tuple = None
  feature_group_count: int = 1
  use_bias: bool = True
  dtype: type = np.float32
  precision: object = None
  kernel_init: object = None
  bias_init: object = None

  def __call__(self, inputs):
    assert self.padding == "SAME" and self.feature_group_count == 1
    kernel = self.param("kernel", None, tuple(self.kernel_size) + (inputs.shape[-1], self.features))
    y = conv2d_same(inputs, kernel, tuple(self.strides) if self.strides else (1, 1))
    return y + self.param("bias", None, (self.features,)) if self.use_bias else y


class Dense(Module):
  features: int
  use_bias: bool = True
  dtype: type = np.float32
  precision: object = None
  kernel_init: object = None
  bias_init: object = None

  def __call__(self, inputs):
    y = np.matmul(np.asarray(inputs, DTYPE[0]), self.param("kernel", None, (inputs.shape[-1], self.features)))
    return y + self.param("bias", None, (self.features,)) if self.use_bias else y


class BatchNorm(Module):
  """flax 0.3.3 linen.BatchNorm over the last axis (normalization.py): statistics in float32."""
  use_running_average: bool = None
  axis: int = -1
  momentum: float = 0.99
  epsilon: float = 1e-5
  dtype: type = np.float32
  use_bias: bool = True
  use_scale: bool = True
  bias_init: object = None
  scale_init: object = None
  axis_name: str = None
  axis_index_groups: object = None

  def __call__(self, x, use_running_average=None):
    use_ra = self.use_running_average if use_running_average is None else use_running_average
    assert self.axis_name is None, "cross-replica statistics are not part of the stand-in"
    x = np.asarray(x, DTYPE[0])
    red = tuple(range(x.ndim - 1))
    ra_mean, ra_var = self.variable("batch_stats", "mean"), self.variable("batch_stats", "var")
    if use_ra:
      mean, var = ra_mean.value, ra_var.value
    else:
      mean = np.mean(x, axis=red)
      var = np.mean(np.square(x), axis=red) - np.square(mean)
      m = DTYPE[0](self.momentum)
      new_mean, new_var = m * ra_mean.value + (1 - m) * mean, m * ra_var.value + (1 - m) * var
      ra_mean.value, ra_var.value = new_mean, new_var
    mul = 1.0 / np.sqrt(var + DTYPE[0](self.epsilon))
    if self.use_scale:
      mul = mul * self.param("scale", None)
    y = (x - mean) * mul
    if self.use_bias:
      y = y + self.param("bias", None)
    return y.astype(DTYPE[0])


def install(jax, sys_modules):
  """Adds the lax / image pieces to the jax stand-in of make_reference_golden and registers flax.linen & friends."""
  lax = jax.lax
  lax.stop_gradient = stop_gradient
  lax.add = np.add
  lax.Precision = types.SimpleNamespace(HIGHEST=None, DEFAULT=None)
  lax.ConvDimensionNumbers = collections.namedtuple("ConvDimensionNumbers", "lhs_spec rhs_spec out_spec")

  def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None, rhs_dilation=None,
                           dimension_numbers=None, feature_group_count=1, precision=None):
    assert tuple(window_strides) == (1, 1) and padding == "SAME" and not lhs_dilation and not rhs_dilation
    assert feature_group_count == 1
    assert tuple(dimension_numbers.lhs_spec) == (0, 3, 1, 2) and tuple(dimension_numbers.rhs_spec) == (3, 2, 0, 1)
    return conv2d_same(lhs, rhs)

  def dot_general(lhs, rhs, dimension_numbers, precision=None):
    ((lc, rc), (lb, rb)) = dimension_numbers
    assert tuple(lc) == (lhs.ndim - 1,) and tuple(rc) == (0,) and not lb and not rb
    return np.matmul(lhs, rhs)

  def reduce_window(x, init, op, window, strides, padding):
    assert op is np.add and tuple(window) == (1, 2, 2, 1) and tuple(strides) == (1, 2, 2, 1)
    n, h, w, c = x.shape
    assert h % 2 == 0 and w % 2 == 0   # SAME == VALID on even maps
    return x.reshape(n, h // 2, 2, w // 2, 2, c).sum(axis=(2, 4)) + init

  lax.conv_general_dilated, lax.dot_general, lax.reduce_window = conv_general_dilated, dot_general, reduce_window

  def resize(x, shape, method):
    assert method == "nearest"
    fh, fw = shape[1] // x.shape[1], shape[2] // x.shape[2]
    assert fh * x.shape[1] == shape[1] and fw * x.shape[2] == shape[2]
    return np.repeat(np.repeat(x, fh, axis=1), fw, axis=2)

  image = types.ModuleType("jax.image")
  image.resize = resize
  jax.image = image
  init = types.ModuleType("jax.nn.initializers")
  init.glorot_normal = lambda *a, **k: None
  jax.nn.initializers = init
  linen = types.ModuleType("flax.linen")
  linen.Module, linen.compact, linen.Conv, linen.Dense, linen.BatchNorm = Module, compact, Conv, Dense, BatchNorm
  linen.relu = lambda x: np.maximum(x, 0)
  linit = types.ModuleType("flax.linen.initializers")
  linit.lecun_normal = linit.normal = lambda *a, **k: None
  linit.zeros = None
  linen.initializers = linit
  linen.max_pool = max_pool
  flax = types.ModuleType("flax")
  flax.linen = linen
  mlc = types.ModuleType("ml_collections")
  mlc.ConfigDict = dict
  sys_modules.update({"flax": flax, "flax.linen": linen, "flax.linen.initializers": linit, "jax.image": image,
                      "jax.nn.initializers": init, "ml_collections": mlc})
  return linen
