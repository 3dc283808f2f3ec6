"""Pins the CPU oracle (oracle/xmc_oracle.py) with analytic known answers and with the committed
golden fixture. The reference ships no numeric goldens for this path, so these KATs are what anchors the oracle."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import xmc_oracle as orc
from tests import helpers


def test_hinge_known_answer():
  d, g = orc.hinge_loss(torch.zeros(5, 1), torch.zeros(5, 1))
  assert d.item() == 2.0 and g.item() == 0.0
  d, g = orc.hinge_loss(torch.full((4, 1), 3.0), torch.full((4, 1), -2.0))
  assert d.item() == 0.0 and g.item() == 2.0


def test_contrastive_orthonormal_and_constant():
  B = 6
  x = torch.eye(B, 16)
  loss, acc, _ = orc.contrastive_loss(x, x)
  assert abs(loss.item() - 2 * math.log(1 + (B - 1) * math.exp(-10.0))) < 1e-6
  assert acc.item() == 1.0
  c = torch.ones(B, 8)
  loss, _, _ = orc.contrastive_loss(c, c)
  assert abs(loss.item() - 2 * math.log(B)) < 1e-5


def test_contrastive_sync_match_raises():
  with pytest.raises(NotImplementedError):
    orc.contrastive_loss(torch.ones(2, 4), torch.ones(2, 4), sync_match=True)


def test_attention_for_g_single_word():
  torch.manual_seed(0)
  q, w = torch.randn(2, 5, 8), torch.randn(2, 4, 8)
  mask = torch.ones(2, 5, 4)
  mask[:, :, 0] = 0.0
  ctx, attn = orc.attention_for_g(q, w, 15.0, mask)
  want = orc.l2_normalize(w)[:, :1].expand(2, 5, 8)
  assert torch.allclose(ctx, want, atol=1e-6)
  assert torch.allclose(attn[..., 0], torch.ones(2, 5))


def test_word_loss_ignores_padded_words():
  torch.manual_seed(1)
  B, R, L, D = 3, 6, 5, 8
  img, words = torch.randn(B, R, D), torch.randn(B, L, D)
  max_len = torch.tensor([[2.0], [5.0], [3.0]])
  a, _, _ = orc.word_loss(img, words, max_len)
  w2 = words.clone()
  w2[0, 2:] = torch.randn(3, D) * 7
  w2[2, 3:] = torch.randn(2, D) * 7
  b, _, _ = orc.word_loss(img, w2, max_len)
  assert abs(a.item() - b.item()) < 1e-4 * max(1.0, abs(a.item()))


def test_spectral_norm_rank_one():
  a, b = torch.randn(12), torch.randn(7)
  k = torch.outer(a, b)
  u0 = (b / b.norm())[None]
  kn, u1 = orc.spectral_normalize(k, u0)
  sigma = a.norm() * b.norm()
  assert torch.allclose(kn, k / (sigma + 1e-10), rtol=1e-4, atol=1e-6)
  assert torch.allclose(u1.abs(), u0.abs(), atol=1e-5)


def test_spectral_norm_backward_formula():
  """Appendix B: dW = dWt/s' - <dWt,W>/s'^2 * v0^T u1 equals autograd through sigma with u, v stop-gradient."""
  torch.manual_seed(2)
  w = torch.randn(3, 3, 4, 5, requires_grad=True)
  u0 = torch.randn(1, 5) * 0.01
  g = torch.randn(3, 3, 4, 5)
  kn, u1 = orc.spectral_normalize(w, u0)
  (kn * g).sum().backward()
  w2 = w.detach().reshape(-1, 5)
  v0 = orc._l2_normalize_sn(u0 @ w2.t(), 1e-10)
  s = (v0 @ w2 @ u1.t())[0, 0] + 1e-10
  want = g.reshape(-1, 5) / s - (g.reshape(-1, 5) * w2).sum() / s ** 2 * (v0.t() @ u1)
  assert torch.allclose(w.grad.reshape(-1, 5), want, rtol=1e-4, atol=1e-6)


def test_batch_norm_train_statistics():
  x = torch.randn(4, 6, 6, 3) * 3 + 2
  y, new = orc.batch_norm(x, {"mean": torch.zeros(3), "var": torch.ones(3)}, True)
  assert y.mean(dim=(0, 1, 2)).abs().max() < 1e-5
  assert (y.var(dim=(0, 1, 2), unbiased=False) - 1).abs().max() < 1e-3
  assert torch.allclose(new["mean"], 0.1 * x.mean(dim=(0, 1, 2)), atol=1e-6)


def test_conv1x1_commutes_with_resampling():
  torch.manual_seed(3)
  x, k, b = torch.randn(2, 4, 4, 5), torch.randn(1, 1, 5, 7), torch.randn(7)
  assert torch.equal(orc.upsample(orc.conv2d(x, k, b)), orc.conv2d(orc.upsample(x), k, b))
  assert torch.allclose(orc.dsample(orc.conv2d(x, k, b)), orc.conv2d(orc.dsample(x), k, b), atol=1e-5)


def test_upsample_dsample_index_semantics():
  x = torch.arange(2 * 2 * 2 * 1, dtype=torch.float32).reshape(2, 2, 2, 1)
  up = orc.upsample(x)
  for i in range(4):
    for j in range(4):
      assert torch.equal(up[:, i, j], x[:, i // 2, j // 2])
  assert torch.equal(orc.dsample(up), x)


def test_adam_first_step_and_split():
  p = {"w": torch.tensor([1.0, -2.0])}
  g = {"w": torch.tensor([0.5, -0.25])}
  newp, st = orc.adam_apply(p, orc.adam_init(p), g, lr=0.1, beta1=0.5, beta2=0.999)
  assert torch.allclose(newp["w"], p["w"] - 0.1 * torch.sign(g["w"]), atol=1e-5)
  assert st["step"] == 1
  parts = orc.split_input_dict({"a": torch.arange(8).reshape(4, 2)}, 2)
  assert torch.equal(parts[0]["a"], torch.arange(4).reshape(2, 2))
  assert torch.equal(parts[1]["a"], torch.arange(4, 8).reshape(2, 2))


def test_image_size_other_than_128_256_raises():
  cfg = helpers.small_config(image_size=64)
  with pytest.raises(ValueError):
    orc._channel_dims_g(cfg.image_size)
  with pytest.raises(ValueError):
    orc._channel_dims_d(cfg.image_size)


def test_oracle_grads_match_finite_differences():
  """train_d's d_loss gradient (autograd over the restatement) against central differences in float64-free fp32 on a
  handful of coordinates of a tiny model."""
  cfg = helpers.small_config(gf_dim=8, df_dim=8)
  _, _, g_vars, d_vars = helpers.cpu_variables(cfg, E=16, seed=3)
  state = orc.make_state(g_vars, d_vars)
  batch = helpers.make_batch(2, cfg, E=16, L=5, seed=1, min_len=2)
  r = orc.d_losses_and_grads(state, batch, cfg, orc.FP32, want_g=False)

  def loss_with(path, idx, delta):
    node = state["d_params"]
    for k in path[:-1]:
      node = node[k]
    t = node[path[-1]]
    old = t.view(-1)[idx].item()
    t.view(-1)[idx] = old + delta
    out = orc.d_losses_and_grads(state, batch, cfg, orc.FP32, want_g=False)["d_loss"].item()
    t.view(-1)[idx] = old
    return out

  for path, idx in ((("SpectralDense_0", "bias"), 0), (("SpectralDense_1", "kernel"), 5),
                    (("DiscBlock_4", "SpectralConv_1", "bias"), 3)):
    eps = 1e-2
    fd = (loss_with(path, idx, eps) - loss_with(path, idx, -eps)) / (2 * eps)
    node = r["d_grad"]
    for k in path:
      node = node[k]
    an = node.reshape(-1)[idx].item()
    assert abs(fd - an) < 5e-2 * max(1e-2, abs(an)) + 2e-3, (path, fd, an)


def test_golden_fixture():
  """tests/golden/oracle_tiny.npz was produced by tests/golden/make_golden.py from this oracle; it guards the
  restatement against accidental drift (upstream has no goldens for the path: parity unpinned by the reference)."""
  path = os.path.join(os.path.dirname(__file__), "golden", "oracle_tiny.npz")
  gold = np.load(path)
  from tests.golden import make_golden
  now = make_golden.compute()
  for k in gold.files:
    assert np.allclose(now[k], gold[k], rtol=2e-4, atol=1e-5), k


def test_jax_image_resize_bilinear_weights_known_answer():
  """get_pretrained_embs resizes with jax.image.resize(..., "bilinear") (pretrained_model_utils.py:118-121). Restated
  from jax's compute_weight_mat: sample_f = (o + 0.5) / scale - 0.5, triangle kernel of width max(1/scale, 1) (the
  anti-aliasing when down-sampling), taps outside the image dropped, weights renormalised. The oracle uses torch's
  interpolate (antialias only when down-sampling): both must agree for 128 -> 224 and for 256 -> 224."""
  import numpy as np

  def weight_mat(S, T):
    inv = S / T
    ks = max(inv, 1.0)
    M = np.zeros((T, S))
    for o in range(T):
      c = (o + 0.5) * inv - 0.5
      j = np.arange(S)
      w = np.maximum(0.0, 1.0 - np.abs(j - c) / ks)
      M[o] = w / w.sum()
    return M

  for S in (128, 256):
    x = torch.rand(2, S, S, 3, dtype=torch.float64)
    M = torch.from_numpy(weight_mat(S, 224))
    want = torch.einsum("ij,njkc,lk->nilc", M, x, M)
    got = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2), size=(224, 224), mode="bilinear", align_corners=False,
                                          antialias=S > 224).permute(0, 2, 3, 1)
    assert (got - want).abs().max() < 1e-6
  # 2-tap rows when up-sampling, up to 3 when 256 -> 224; border rows keep their full weight on the edge pixel
  assert (weight_mat(128, 224) > 0).sum(1).max() == 2 and (weight_mat(256, 224) > 0).sum(1).max() == 3
  assert abs(weight_mat(128, 224)[0, 0] - 1.0) < 1e-12


def test_preprocess_batch_known_answers():
  """COCODataset.preprocess (coco_dataset.py:127-167) restated: flip is an exact index reversal along W, values are
  clipped to [0,1], the sentence embedding sums ALL word slots (padded ones included) and divides by the caption's
  length, and the chosen caption's tensors are plain gathers."""
  img = torch.arange(2 * 2 * 3 * 3, dtype=torch.float32).reshape(2, 2, 3, 3) / 20.0 - 0.2
  emb = torch.ones(2, 3, 4, 5)
  emb[1, 2] = 2.0
  lens = torch.tensor([[4, 2, 3], [1, 4, 2]])
  out = orc.preprocess_batch({"image": img, "caption/embedding": emb, "caption/max_len": lens},
                             flip=torch.tensor([True, False]), sentence_idx=torch.tensor([1, 2]), z=torch.zeros(2, 8))
  assert torch.equal(out["image"][0], img[0].flip(1).clamp(0, 1)) and torch.equal(out["image"][1], img[1].clamp(0, 1))
  assert out["image"].min() >= 0 and out["image"].max() <= 1
  assert torch.equal(out["max_len"], torch.tensor([[2.0], [2.0]]))
  assert torch.equal(out["embedding"][1], emb[1, 2])
  assert torch.allclose(out["sentence_embedding"][0], torch.full((5,), 4 * 1.0 / 2))   # 4 slots summed / len 2
  assert torch.allclose(out["sentence_embedding"][1], torch.full((5,), 4 * 2.0 / 2))


def test_generator_spectral_norm_state_advances_only_in_train_g_d():
  """xmc_gan.py:225 (train_d discards the generator's new collections) vs :159,181 (train_g_d keeps them), with
  config.g_spectral_norm=True: u0 of every generator layer is one power-iteration step ahead after a train_step, not
  two, and SpectralConv/SpectralDense replace every Conv/Dense name in the generator tree."""
  cfg = helpers.small_config(gf_dim=8, df_dim=8, g_spectral_norm=True)
  _, _, g_vars, d_vars = helpers.cpu_variables(cfg, E=16, seed=3)
  names = {k for path, _ in orc.tree_leaves(g_vars["params"]) for k in path.split("/")}
  assert not any(n.startswith(("Conv_", "Dense_")) for n in names)
  state = orc.make_state(g_vars, d_vars)
  batch = helpers.make_batch(4, cfg, E=16, L=5, seed=9, min_len=2)
  new_state, _ = orc.train_step(state, batch, cfg, orc.FP32)
  p = state["g_params"]["SpectralDense_1"]["kernel"]
  u0 = state["generator_state"]["spectral_norm_stats"]["SpectralDense_1"]["u0"]
  _, u1 = orc.spectral_normalize(p, u0)
  got = new_state["generator_state"]["spectral_norm_stats"]["SpectralDense_1"]["u0"]
  assert torch.allclose(got, u1, atol=1e-6)
  _, u2 = orc.spectral_normalize(p, u1)
  assert not torch.allclose(got, u2, atol=1e-4)
