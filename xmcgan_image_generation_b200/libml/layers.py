"""Module-level counterparts of xmcgan/libml/layers.py (and of the flax.linen modules the reference composes them
with) on CUDA tensors, with the Flax call surface the reference's callers use:

    conv_fn = functools.partial(layers.SpectralConv, train=True, dtype=torch.bfloat16)
    variables = conv_fn(96, kernel_size=(3, 3)).init(rng, x)          # {"params": ..., "spectral_norm_stats": ...}
    y, new_vars = conv_fn(96, kernel_size=(3, 3)).apply(variables, x, mutable=["spectral_norm_stats"])

    SpectralDense (layers.py:49-113)   SpectralConv (:125-241)   ConditionalBatchNorm (:244-258)
    LocalConditionalBatchNorm (:261-273)   + Dense / Conv / BatchNorm (flax.linen, as configured at xmc_net.py:176-201)

Variables are nested dicts of CUDA fp32 tensors with Flax's auto-naming (`<ClassName>_<index>` in call order), i.e. the
sub-trees of SURVEY.md Appendix A. Forward values only: every call is a program of libxmc.so kernels (spectral-norm power
iteration, weight preparation, tcgen05 implicit-GEMM convolution, BatchNorm statistics / modulation); training uses the
fused forward + backward programs of engine.py, which these modules are tested against. There is no CPU path."""
import functools
import math

import torch

from .. import engine as _engine
from .. import ops
from ..nets import xmc_net as _xn

BF16, F32 = ops.BF16, ops.F32


def _act(dtype):
  if dtype in (None, torch.bfloat16, "bfloat16"):
    return BF16
  if dtype in (torch.float32, "float32"):
    return F32
  raise ValueError(f"dtype {dtype!r} is not supported (bfloat16 or float32)")


class _Scope:
  """Variable scope of one module call: Flax auto-naming, parameter creation (init) / lookup (apply), mutable
  collections."""

  def __init__(self, variables, init, gen, new_vars, path=()):
    self.variables, self.init, self.gen, self.new_vars, self.path = variables, init, gen, new_vars, path
    self.counts = {}

  def child(self, module):
    name = getattr(module, "name", None)
    if name is None:
      cls = type(module).__name__
      idx = self.counts.get(cls, 0)
      self.counts[cls] = idx + 1
      name = f"{cls}_{idx}"
    return _Scope(self.variables, self.init, self.gen, self.new_vars, self.path + (name,))

  def _node(self, root, create):
    node = root
    for k in self.path:
      if k not in node:
        if not create:
          raise KeyError(f"missing variables for {'/'.join(self.path)}")
        node[k] = {}
      node = node[k]
    return node

  def variable(self, collection, name, shape, init_fn):
    if self.init:
      node = self._node(self.variables.setdefault(collection, {}), True)
      if name not in node:
        node[name] = init_fn(shape, self.gen).to("cuda", torch.float32).contiguous()
      return node[name]
    if collection not in self.variables:
      raise KeyError(f"collection {collection!r} is missing")
    t = self._node(self.variables[collection], False)[name]
    t = torch.as_tensor(t).to("cuda", torch.float32).contiguous()
    if tuple(t.shape) != tuple(shape):
      raise ValueError(f"{'/'.join(self.path)}/{name}: shape {tuple(t.shape)} != {tuple(shape)}")
    return t

  def param(self, name, shape, init_fn):
    return self.variable("params", name, shape, init_fn)

  def put(self, collection, name, value):
    self._node(self.new_vars.setdefault(collection, {}), True)[name] = value

  def call(self, module, *args):
    return module.forward(self.child(module), *args)


def _glorot(shape, gen):
  rf = int(math.prod(shape[:-2])) if len(shape) > 2 else 1
  std = math.sqrt(2.0 / ((shape[-2] + shape[-1]) * rf))
  return torch.randn(shape, generator=gen) * std


def _zeros(shape, gen):
  return torch.zeros(shape)


def _ones(shape, gen):
  return torch.ones(shape)


def _normal001(shape, gen):   # flax.linen.initializers.normal() (stddev 1e-2), layers.py:86-91
  return torch.randn(shape, generator=gen) * 0.01


def _merge(base, upd):
  out = dict(base)
  for k, v in upd.items():
    out[k] = _merge(base.get(k, {}), v) if isinstance(v, dict) and isinstance(base.get(k, {}), dict) else v
  return out


class Module:
  """Minimal flax.linen.Module stand-in: subclasses implement forward(scope, *inputs)."""
  name = None

  def init(self, rng, *inputs):
    gen = torch.Generator().manual_seed(_xn._seed_of(rng))
    variables = {}
    self.forward(_Scope(variables, True, gen, {}), *inputs)
    return variables

  def apply(self, variables, *inputs, mutable=False, rngs=None):
    new_vars = {}
    out = self.forward(_Scope(variables, False, None, new_vars), *inputs)
    if mutable is False:
      return out
    cols = list(mutable) if not isinstance(mutable, str) else [mutable]
    return out, {c: _merge(variables.get(c, {}), new_vars.get(c, {})) for c in cols if c in variables or c in new_vars}


def _to_act(x, dtype):
  """jnp.asarray(inputs, dtype): plumbing cast of the caller's tensor to the module's compute dtype."""
  return torch.as_tensor(x).to(device="cuda", dtype=dtype).contiguous()


def _kernel_matrices(kernel, taps, cin, cout, act, spectral=None):
  """fp32 HWIO kernel [taps*cin, cout] -> K-major operand copy for xmc_conv2d_fwd (bf16, or the [hi|hi|lo] split of
  fp32 mode), optionally spectrally normalised: spectral = (u0 [1,cout], eps) -> also returns the new u0."""
  S = 3 if act == F32 else 1
  flat = kernel.reshape(-1).contiguous()
  if flat.numel() % 4:   # the multi-tensor kernels read the buffer in 16-byte vectors
    flat = torch.cat([flat, flat.new_zeros(4 - flat.numel() % 4)])
  sn, u_new, scalars, n_sn = -1, None, None, 0
  if spectral is not None:
    u0, eps = spectral
    tab = _engine._SnTable()
    sn = tab.add(("k",), 0, taps * cin, cout)
    tab.upload()
    u_new = torch.empty(tab.u_layout.total, device="cuda")
    u_flat = torch.zeros(tab.u_layout.total, device="cuda")
    u_flat[:cout] = u0.reshape(-1)
    ops._call("xmc_sn_forward", tab.dev.data_ptr(), tab.n, float(eps), flat.data_ptr(), u_flat.data_ptr(),
              u_new.data_ptr(), tab.t_ws.data_ptr(), tab.s_ws.data_ptr(), tab._s, tab.scalars.data_ptr(), tab._rb, tab._ct,
              ops.stream(), launches=3)
    scalars, n_sn = tab.scalars, tab.n
    u_new = u_new[:cout].view(1, cout)
  ld = _engine._r8(S * taps * cin)
  arena = torch.zeros(_engine._r8(cout) * ld, device="cuda", dtype=BF16)
  prep = _engine._PrepTable()
  prep.add(0, taps, cin, cout, 0, ld, -1, 0, sn, split=S == 3)
  prep.upload()
  ops._call("xmc_prep_weights", prep.dev.data_ptr(), prep.n, prep.tiles, flat.data_ptr(),
            scalars.data_ptr() if scalars is not None else None, n_sn, arena.data_ptr(), None, None, ops.stream())
  return arena, ld, u_new


class _ConvBase(Module):
  spectral = False

  def _apply_kernel(self, scope, x, kh, cin, features, use_bias, dtype, train, eps, kernel_shape):
    kernel = scope.param("kernel", kernel_shape, _glorot)
    spectral = None
    if self.spectral:
      u0 = scope.variable("spectral_norm_stats", "u0", (1, features), _normal001)
      spectral = (u0, eps)
    bias = scope.param("bias", (features,), _zeros) if use_bias else None
    if x is None:   # init without running the kernels
      return None
    arena, ld, u_new = _kernel_matrices(kernel, kh * kh, cin, features, dtype, spectral)
    if self.spectral and train:   # u0 advances only in train mode (layers.py:98-99,215-216)
      scope.put("spectral_norm_stats", "u0", u_new)
    with ops.act_dtype(dtype):
      if cin == 3:   # the image-side convolutions (K = 27 or 3): the direct CUDA-core kernel of smallconv.cu
        N, H, W, _ = x.shape
        y = ops.empty((N, H, W, features))
        ops._call("xmc_conv_c3_in", x.data_ptr(), ops._f32(x), arena.data_ptr(), ld, bias.data_ptr() if bias is not None
                  else None, N, H, W, features, kh, kh, 0, y.data_ptr(), ops.stream())
        return y
      return ops.conv_fwd(x, arena, kh, features, bias=bias, ldb=ld)


class Conv(_ConvBase):
  """flax.linen.Conv as the reference configures it (xmc_net.py:188-191): NHWC / HWIO, stride 1, 'SAME', bias."""

  def __init__(self, features, kernel_size, strides=None, padding="SAME", use_bias=True, dtype=None, train=None,
               eps=1e-10, name=None, **unused):
    ks = (kernel_size,) * 2 if isinstance(kernel_size, int) else tuple(kernel_size)
    if ks not in ((1, 1), (3, 3)) or (strides not in (None, (1, 1))) or padding != "SAME":
      raise NotImplementedError("built for the reference's uses: 1x1 / 3x3, stride 1, 'SAME'")
    self.features, self.kh, self.use_bias, self.dtype = features, ks[0], use_bias, _act(dtype)
    self.train, self.eps, self.name = train, eps, name

  def forward(self, scope, inputs):
    x = _to_act(inputs, self.dtype)
    single = x.dim() == 3
    if single:
      x = x[None]
    if x.dim() != 4 or (x.shape[-1] % 8 and x.shape[-1] != 3):
      raise ValueError("inputs must be [N,H,W,C] with C a multiple of 8 (or 3 image channels)")
    cin = x.shape[-1]
    y = self._apply_kernel(scope, x, self.kh, cin, self.features, self.use_bias, self.dtype, self.train, self.eps,
                           (self.kh, self.kh, cin, self.features))
    return y[0] if single else y


class SpectralConv(Conv):
  """layers.SpectralConv (layers.py:125-241): one power-iteration step, W / (sigma + eps), then the convolution."""
  spectral = True

  def __init__(self, features, train, kernel_size, **kw):
    super().__init__(features, kernel_size, train=train, **kw)


class Dense(_ConvBase):
  """flax.linen.Dense over the last axis."""

  def __init__(self, features, use_bias=True, dtype=None, train=None, eps=1e-10, name=None, **unused):
    self.features, self.use_bias, self.dtype, self.train, self.eps, self.name = (features, use_bias, _act(dtype), train,
                                                                                 eps, name)

  def forward(self, scope, inputs):
    x = _to_act(inputs, self.dtype)
    cin = x.shape[-1]
    if cin % 8:
      raise ValueError("the input feature count must be a multiple of 8")
    rows = x.reshape(-1, cin)
    y = self._apply_kernel(scope, _engine.as4(rows), 1, cin, self.features, self.use_bias, self.dtype, self.train,
                           self.eps, (cin, self.features))
    return y.view(*x.shape[:-1], self.features)


class SpectralDense(Dense):
  """layers.SpectralDense (layers.py:49-113)."""
  spectral = True

  def __init__(self, features, train, **kw):
    super().__init__(features, train=train, **kw)


class BatchNorm(Module):
  """flax.linen.BatchNorm as configured at xmc_net.py:192-201: fp32 statistics over (N,H,W), biased variance
  E[x^2]-E[x]^2, running average with `momentum`; scale / bias off (the conditional variants supply them)."""

  def __init__(self, use_running_average, momentum=0.9, epsilon=1e-5, dtype=None, use_bias=False, use_scale=False,
               axis_name=None, axis_index_groups=None, name=None):
    if use_bias or use_scale:
      raise NotImplementedError("the reference uses BatchNorm without scale / bias (layers.py:256,271)")
    if axis_index_groups is not None:
      raise NotImplementedError("cross-replica statistics run through the fused engine (config.batch_norm_group_size)")
    self.eval, self.momentum, self.eps, self.dtype, self.name = use_running_average, momentum, epsilon, _act(dtype), name

  def stats(self, scope, x):
    C = x.shape[-1]
    mean = scope.variable("batch_stats", "mean", (C,), _zeros)
    var = scope.variable("batch_stats", "var", (C,), _ones)
    if self.eval:
      return ops.bn_eval_stats(mean, var, C, self.eps)
    sums, P = ops.bn_stats(x)
    nm, nv = torch.empty_like(mean), torch.empty_like(var)
    mr = ops.bn_finalize(sums, P, C, mean, var, nm, nv, self.eps, self.momentum)
    scope.put("batch_stats", "mean", nm)
    scope.put("batch_stats", "var", nv)
    return mr

  def forward(self, scope, x, gb=None, Hc=1):
    """gb: optional [rows, 2C] (gamma | beta) modulation matrix in the activation dtype (conditional variants)."""
    x = _to_act(x, self.dtype)
    N, H, W, C = x.shape
    mr = self.stats(scope, x)
    if gb is None:
      gb = torch.zeros(N, 2 * C, device="cuda", dtype=self.dtype)
    return ops.bn_apply(x, mr, gb, Hc, 0, C, False, False)


class ConditionalBatchNorm(Module):
  """layers.ConditionalBatchNorm (layers.py:244-258): gamma = Dense(C)(emb), beta = Dense(C)(emb),
  BN_noaffine(x) * (gamma + 1) + beta."""

  def __init__(self, norm_fn, dense_fn, name=None):
    self.norm_fn, self.dense_fn, self.name = norm_fn, dense_fn, name

  def forward(self, scope, x, emb):
    filters = x.shape[-1]
    gamma = scope.call(self.dense_fn(filters), emb)
    beta = scope.call(self.dense_fn(filters), emb)
    norm = self.norm_fn(use_bias=False, use_scale=False)
    gb = torch.cat([gamma.reshape(-1, filters), beta.reshape(-1, filters)], dim=1).to(norm.dtype)   # plumbing
    return norm.forward(scope.child(norm), x, gb, 1)


class LocalConditionalBatchNorm(Module):
  """layers.LocalConditionalBatchNorm (layers.py:261-273): gamma / beta = Conv1x1(C)(emb) per pixel of `emb`, whose
  spatial size is x's or a power-of-two fraction of it (a 1x1 convolution commutes with nearest up-sampling, so the
  reference's up-sampled condition and the condition at its native 16x16 give the same result)."""

  def __init__(self, norm_fn, conv_fn, name=None):
    self.norm_fn, self.conv_fn, self.name = norm_fn, conv_fn, name

  def forward(self, scope, x, emb):
    filters = x.shape[-1]
    gamma = scope.call(self.conv_fn(filters, kernel_size=(1, 1)), emb)
    beta = scope.call(self.conv_fn(filters, kernel_size=(1, 1)), emb)
    norm = self.norm_fn(use_bias=False, use_scale=False)
    Hc = gamma.shape[1]
    if gamma.shape[1] != gamma.shape[2]:
      raise ValueError("the spatial condition must be square")
    gb = torch.cat([gamma.reshape(-1, filters), beta.reshape(-1, filters)], dim=1).to(norm.dtype)   # plumbing
    return norm.forward(scope.child(norm), x, gb, Hc)


def _dev_act(x):
  """Plumbing: CUDA bf16 / fp32 tensors pass as they are, anything else (host tensors, other dtypes) is moved to the
  device as bf16 — the kernels below only ever see device pointers."""
  x = torch.as_tensor(x)
  if x.is_cuda and x.dtype in (BF16, F32):
    return x.contiguous()
  return x.to("cuda", BF16 if x.dtype != F32 or not x.is_cuda else F32).contiguous()


def relu(x):
  """flax nn.relu as a stand-alone kernel (the fused engine folds it into the producing kernel instead)."""
  x = _dev_act(x)
  if x.numel() % 8:
    raise ValueError("element count must be a multiple of 8")
  y = torch.empty_like(x)
  ops._call("xmc_relu_or_add", x.data_ptr(), None, ops._f32(x), x.numel(), y.data_ptr(), ops.stream())
  return y


def add(a, b):
  a, b = _dev_act(a), _dev_act(b)
  if a.shape != b.shape or a.dtype != b.dtype or a.numel() % 8:
    raise ValueError("add needs two tensors of the same shape / dtype with a multiple of 8 elements")
  y = torch.empty_like(a)
  ops._call("xmc_relu_or_add", a.data_ptr(), b.data_ptr(), ops._f32(a), a.numel(), y.data_ptr(), ops.stream())
  return y
