"""Module-level API of SURVEY.md 8(b): libml.layers.{SpectralConv, SpectralDense, ConditionalBatchNorm,
LocalConditionalBatchNorm}, nets.common.{GenBlock, GenSpatialBlock, DiscBlock, DiscOptimizedBlock},
attention_lib.attention and losses.tf_cross_entropy_loss_with_logits against the oracle's restatement of the same
reference lines, through the Flax call surface (init / apply / mutable collections, Flax auto-naming)."""
import functools

import pytest
import torch

from oracle import xmc_oracle as orc
from tests import helpers

gpu = pytest.mark.gpu


def _cpu(tree):
  return {k: _cpu(v) for k, v in tree.items()} if isinstance(tree, dict) else tree.detach().float().cpu()


def _randomize(tree, gen, scale=0.1):
  """Non-trivial biases / statistics (the initialisers give zeros), same values for both sides."""
  for k, v in tree.items():
    if isinstance(v, dict):
      _randomize(v, gen, scale)
    elif k in ("bias", "mean"):
      v.copy_(torch.randn(v.shape, generator=gen) * scale)
    elif k == "var":
      v.copy_(1.0 + 0.3 * torch.rand(v.shape, generator=gen))


def _fns(spectral, train, dtype):
  from xmcgan_image_generation_b200.libml import layers
  conv = functools.partial(layers.SpectralConv, train=train, dtype=dtype) if spectral else \
      functools.partial(layers.Conv, dtype=dtype)
  dense = functools.partial(layers.SpectralDense, train=train, dtype=dtype) if spectral else \
      functools.partial(layers.Dense, dtype=dtype)
  norm = functools.partial(layers.BatchNorm, use_running_average=not train, momentum=0.9, dtype=dtype)
  return conv, dense, norm


def _q(t):
  return t.to(torch.bfloat16).float()


@gpu
@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 6e-3), (torch.float32, 2e-5)])
def test_spectral_conv_and_dense_modules(dtype, tol):
  """layers.SpectralConv (layers.py:125-241) / SpectralDense (:49-113): output and the advanced u0 vs the oracle;
  u0 stays put in eval mode; variable names / shapes as Flax creates them."""
  from xmcgan_image_generation_b200.libml import layers
  pol = orc.FP32 if dtype == torch.float32 else orc.Policy("bfloat16")
  g = torch.Generator().manual_seed(0)
  x = torch.randn(2, 16, 16, 32, generator=g)
  x = x if dtype == torch.float32 else _q(x)
  for train in (True, False):
    m = layers.SpectralConv(48, train, (3, 3), dtype=dtype)
    v = m.init(1, x)
    assert set(v) == {"params", "spectral_norm_stats"}
    assert v["params"]["kernel"].shape == (3, 3, 32, 48) and v["spectral_norm_stats"]["u0"].shape == (1, 48)
    _randomize(v["params"], g)
    y, new = m.apply(v, x, mutable=["spectral_norm_stats"])
    want, st = orc.spectral_conv(x, _cpu(v["params"]), _cpu(v["spectral_norm_stats"]), train, pol)
    assert helpers.rel(y, want) < tol
    assert helpers.rel(new["spectral_norm_stats"]["u0"], st["u0"]) < 1e-5
    assert torch.equal(new["spectral_norm_stats"]["u0"].cpu(), v["spectral_norm_stats"]["u0"].cpu()) == (not train)
  e = torch.randn(5, 64, generator=g)
  e = e if dtype == torch.float32 else _q(e)
  m = layers.SpectralDense(24, True, dtype=dtype)
  v = m.init(2, e)
  assert v["params"]["kernel"].shape == (64, 24)
  y = m.apply(v, e)
  want, _ = orc.spectral_dense(e, _cpu(v["params"]), _cpu(v["spectral_norm_stats"]), True, pol)
  assert y.shape == (5, 24) and helpers.rel(y, want) < tol


@gpu
@pytest.mark.parametrize("spectral", [False, True])
def test_conditional_batch_norm_modules(spectral):
  """layers.ConditionalBatchNorm (layers.py:244-258) and LocalConditionalBatchNorm (:261-273) in train mode: output
  1e-2 (bf16), new running statistics 1e-5, Flax auto-names (Dense_0 / Dense_1 / BatchNorm_0 ...)."""
  from xmcgan_image_generation_b200.libml import layers
  conv, dense, norm = _fns(spectral, True, torch.bfloat16)
  pol = orc.Policy("bfloat16")
  g = torch.Generator().manual_seed(3)
  x = _q(torch.randn(3, 8, 8, 32, generator=g) * 1.5 + 0.3)
  emb = _q(torch.randn(3, 16, generator=g))
  m = layers.ConditionalBatchNorm(norm_fn=norm, dense_fn=dense)
  v = m.init(4, x, emb)
  pre = "SpectralDense" if spectral else "Dense"
  assert set(v["params"]) == {pre + "_0", pre + "_1"} and set(v["batch_stats"]) == {"BatchNorm_0"}
  _randomize(v["params"], g)
  _randomize(v["batch_stats"], g)
  y, new = m.apply(v, x, emb, mutable=["batch_stats"])
  sc = orc._Scope(_cpu(v))
  want = orc.conditional_batch_norm(sc, x, emb, spectral, True, pol)
  assert helpers.rel(y, want) < 1e-2
  assert helpers.rel(new["batch_stats"]["BatchNorm_0"]["mean"], sc.updates["batch_stats"]["BatchNorm_0"]["mean"]) < 1e-5
  # local variant: the condition at x's resolution (as the reference feeds it) and at a quarter of it
  for hc in (8, 2):
    cond = _q(torch.randn(3, hc, hc, 24 if not spectral else 24, generator=g))
    m = layers.LocalConditionalBatchNorm(norm_fn=norm, conv_fn=conv)
    v = m.init(5, x, cond)
    _randomize(v["params"], g)
    y = m.apply(v, x, cond)
    sc = orc._Scope(_cpu(v))
    up = cond
    while up.shape[1] < 8:
      up = orc.upsample(up)
    want = orc.local_conditional_batch_norm(sc, x, up, spectral, True, pol)
    assert helpers.rel(y, want) < 1e-2, hc


@gpu
@pytest.mark.parametrize("spectral", [True, False])
def test_residual_block_modules(spectral):
  """nets.common.{DiscBlock, DiscOptimizedBlock, GenBlock, GenSpatialBlock} (common.py:58-186) with the reference's
  constructor arguments, against the oracle's restatement of the same lines on the same variables: 2e-2 rel-L2 (bf16
  activations through up to two BatchNorms), mutated collections present with Flax names."""
  from xmcgan_image_generation_b200.nets import common
  conv, dense, norm = _fns(spectral, True, torch.bfloat16)
  pol = orc.Policy("bfloat16")
  g = torch.Generator().manual_seed(7)
  pre = "SpectralConv" if spectral else "Conv"
  # DiscOptimizedBlock on 3-channel images
  img = _q(torch.rand(2, 32, 32, 3, generator=g))
  m = common.DiscOptimizedBlock(16, conv_fn=conv)
  v = m.init(1, img)
  assert set(v["params"]) == {pre + "_0", pre + "_1", pre + "_2"}
  _randomize(v["params"], g)
  y = m.apply(v, img)
  want = orc.disc_optimized_block(orc._Scope(_cpu(v)), img, spectral, True, pol)
  assert y.shape == (2, 16, 16, 16) and helpers.rel(y, want) < 1e-2
  # DiscBlock: down-sampling with projection, and the final block (no down-sampling, same width: no projection)
  for filters, down, cin in ((32, True, 16), (16, False, 16)):
    x = _q(torch.randn(2, 16, 16, cin, generator=g))
    m = common.DiscBlock(filters, down, conv_fn=conv)
    v = m.init(2, x)
    assert len(v["params"]) == (3 if (down or cin != filters) else 2)
    _randomize(v["params"], g)
    y, new = m.apply(v, x, mutable=["spectral_norm_stats"])
    want = orc.disc_block(orc._Scope(_cpu(v)), x, filters, down, spectral, True, pol)
    assert helpers.rel(y, want) < 1e-2, (filters, down)
    assert (set(new.get("spectral_norm_stats", {})) == set(v["params"])) == spectral or not spectral
  # GenBlock / GenSpatialBlock
  x = _q(torch.randn(3, 4, 4, 32, generator=g))
  cond = _q(torch.randn(3, 16, generator=g))
  m = common.GenBlock(16, conv_fn=conv, dense_fn=dense, norm_fn=norm)
  v = m.init(3, x, cond)
  assert {"ConditionalBatchNorm_0", "ConditionalBatchNorm_1", pre + "_0", pre + "_1", pre + "_2"} == set(v["params"])
  _randomize(v["params"], g)
  y, new = m.apply(v, x, cond, mutable=["batch_stats"])
  sc = orc._Scope(_cpu(v))
  want = orc.gen_block(sc, x, cond, spectral, True, pol)
  assert y.shape == (3, 8, 8, 16) and helpers.rel(y, want) < 2e-2
  assert set(new["batch_stats"]) == {"ConditionalBatchNorm_0", "ConditionalBatchNorm_1"}
  c0 = _q(torch.randn(3, 4, 4, 24, generator=g))
  c1 = orc.upsample(c0)
  m = common.GenSpatialBlock(16, conv_fn=conv, dense_fn=dense, norm_fn=norm)
  v = m.init(4, x, c0, c1)
  _randomize(v["params"], g)
  y = m.apply(v, x, c0, c1)
  want = orc.gen_spatial_block(orc._Scope(_cpu(v)), x, c0, c1, spectral, True, pol)
  assert y.shape == (3, 8, 8, 16) and helpers.rel(y, want) < 2e-2


@gpu
def test_attention_and_generic_cross_entropy():
  """attention_lib.attention (attention_lib.py:105-127; softmax over REGIONS, context from the normalised regions):
  5e-3 (bf16 operands); with a word-padding mask the padded words get uniform attention (fp32 absorption of the
  scores by the -1e9 shift), as in the reference. losses.tf_cross_entropy_loss_with_logits
  (losses.py:47-51) with soft labels: 1e-6."""
  from xmcgan_image_generation_b200.libml import attention_lib, losses
  g = torch.Generator().manual_seed(11)
  B, R, L, D = 3, 256, 17, 64
  region = _q(torch.randn(B, R, D, generator=g))
  words = torch.randn(B, L, D, generator=g) * 0.5
  want = orc.attention(region, words, 5.0)
  got = attention_lib.attention(region, words, 5.0)
  assert got.shape == (B, L, D) and helpers.rel(got, want) < 5e-3
  max_len = torch.tensor([[3.0], [17.0], [9.0]])
  mask = (torch.arange(L)[None, :] >= max_len).float()[:, None, :].repeat(1, R, 1)
  assert helpers.rel(attention_lib.attention(region, words, 5.0, mask), orc.attention(region, words, 5.0, mask)) < 5e-3
  labels = torch.softmax(torch.randn(7, 13, generator=g), -1)
  logits = torch.randn(7, 13, generator=g) * 3
  got = losses.tf_cross_entropy_loss_with_logits(labels, logits)
  want = orc.tf_cross_entropy_loss_with_logits(labels, logits)
  assert got.shape == (7,) and helpers.rel(got, want) < 1e-6
  eye = torch.eye(5)
  lg = torch.randn(5, 5, generator=g)
  sym = losses.tf_cross_entropy_loss_with_logits(eye, lg).mean() + losses.tf_cross_entropy_loss_with_logits(eye, lg.t()).mean()
  assert abs(sym.item() - losses.symmetric_identity_cross_entropy(lg).item()) < 1e-5
