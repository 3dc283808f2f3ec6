"""The oracle — and the product's host-side helpers — against outputs of THE REFERENCE'S OWN CODE
(tests/golden/reference_libml.npz, made by tests/golden/make_reference_golden.py: xmcgan/libml/losses.py,
xmcgan/libml/attention_lib.py, utils/image_utils.make_grid, utils/device_utils.get_device_groups and
train_utils.split_input_dict executed from /root/reference on a numpy stand-in for the few `jax` entry points they
use). This pins the loss / attention layer of oracle/xmc_oracle.py to the reference itself rather than to a reading of
it: float32 throughout, 2e-6 relative (summation order). The fixture travels; /root/reference is not read here."""
import os

import numpy as np
import pytest
import torch

from oracle import xmc_oracle as orc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_libml.npz"))
T = lambda k: torch.from_numpy(G[k])


def close(got, want, tol=2e-6):
  got = np.asarray([float(v) for v in got]) if isinstance(got, (tuple, list)) else np.asarray(got, np.float64)
  want = np.asarray(want, np.float64)
  assert got.shape == want.shape, (got.shape, want.shape)
  assert np.max(np.abs(got - want)) <= tol * max(1.0, np.max(np.abs(want))), np.max(np.abs(got - want))


def test_word_loss_matches_the_reference_code():
  # (loss, accuracy, entropy): attention_lib.py:130-191 incl. the region-axis softmax, the -1e9 masks, un-normalised
  # words in the cosine, logsumexp over words, both cross-entropy directions
  close(orc.word_loss(T("region"), T("words"), T("max_len")), G["word_loss"], 5e-6)


def test_contrastive_loss_matches_the_reference_code():
  close(orc.contrastive_loss(T("img"), T("cond")), G["contrastive_loss"])
  close(orc.contrastive_loss(T("img"), T("cond"), l2_norm=False, temperature=0.5), G["contrastive_loss_t05_nonorm"])


def test_attention_functions_match_the_reference_code():
  close(orc.attention(T("region"), T("words"), 5.0, T("mask")), G["attention_ctx"])
  close(orc.attention(T("region"), T("words"), 5.0), G["attention_ctx_nomask"])
  ctx, attn = orc.attention_for_g(T("region"), T("words"), 15.0, T("mask"))
  close(ctx, G["attention_for_g_ctx"])
  close(attn, G["attention_for_g_attn"])
  # padded words carry exactly zero weight in the reference too (exp(-1e9) underflows)
  assert (G["attention_for_g_attn"][0, :, 3:] == 0).all() and (attn[0, :, 3:] == 0).all()
  close(orc.l2_normalize(T("region"), -1), G["l2_normalize"])
  close(orc.cosine_similarity(T("words"), T("attention_ctx")), G["cosine_similarity"])


def test_losses_match_the_reference_code():
  labels = torch.eye(4)
  close(orc.tf_cross_entropy_loss_with_logits(labels, T("logits")), G["tf_cross_entropy"])
  # the integer-label form (losses.py:38-44) is the one-hot form on the diagonal labels
  close(orc.tf_cross_entropy_loss_with_logits(labels, T("logits"))[:, None], G["cross_entropy_int"])
  close(orc.get_statistics(T("logits"), labels), G["get_statistics"])
  d, g = orc.hinge_loss(T("real_logit"), T("fake_logit"))
  close([d, g], G["hinge_loss"])
  # hinge_loss_d / hinge_loss_g (losses.py:20-27) are the two halves of hinge_loss
  close([d], G["hinge_loss_d"], 1e-6)
  close([g], G["hinge_loss_g"])


def test_host_helpers_match_the_reference_code():
  from xmcgan_image_generation_b200 import parallel, train_utils
  samples = T("grid_samples")
  assert np.array_equal(train_utils.make_grid(samples, 9).numpy(), G["make_grid_9"])
  assert np.array_equal(train_utils.make_grid(samples, 64).numpy(), G["make_grid_64"])
  assert parallel.get_device_groups(16, 4, device_count=8) == G["device_groups_8_16_4"].tolist()
  assert parallel.get_device_groups(8, 4, device_count=8) == G["device_groups_8_8_4"].tolist()
  a, b = torch.arange(24, dtype=torch.float32).reshape(8, 3), torch.arange(8)
  for split in (train_utils.split_input_dict, orc.split_input_dict):
    parts = split({"a": a, "b": b}, 2)
    assert np.array_equal(np.asarray(parts[0]["a"]), G["split_a0"]) and np.array_equal(np.asarray(parts[1]["a"]), G["split_a1"])
    assert np.array_equal(np.asarray(parts[1]["b"]), G["split_b1"])


def test_fixture_is_what_the_reference_produces_today():
  """In the build container (where /root/reference exists) the committed fixture is regenerated and compared, so it
  cannot drift from the reference's code; elsewhere the check is skipped."""
  if not os.path.isdir("/root/reference/xmcgan"):
    pytest.skip("/root/reference is not present on this machine")
  import subprocess
  import sys
  code = ("import numpy as np; from tests.golden import make_reference_golden as m; d = m.compute(); "
          f"g = np.load({m_path!r}); "
          "assert sorted(d) == sorted(g.files); "
          "assert all(np.array_equal(np.asarray(d[k]), g[k]) for k in g.files); "
          "d = m.compute_nets(); d.update(m.compute_resnet()); d.update(m.compute_grads()); g = np.load(m.OUT_NETS); assert sorted(d) == sorted(g.files); "
          "assert all(np.allclose(np.asarray(d[k]), g[k], rtol=1e-4 if k.startswith('grads/fd') else 1e-6, atol=1e-7) for k in g.files); print('same')")
  out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  assert out.returncode == 0 and "same" in out.stdout, out.stderr[-2000:]


m_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_libml.npz")


# ----------------------------------------------------------------------------------------------------------------------
# The networks: oracle.generator_apply / discriminator_apply against the reference's own xmc_net.py / common.py /
# layers.py executed on the numpy stand-in for flax.linen (tests/golden/flax_stand_in.py)
# ----------------------------------------------------------------------------------------------------------------------
N = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_nets.npz"))


def _net_inputs():
  from tests.golden import make_reference_golden as m
  cfg, g_vars, d_vars, batch = m.net_inputs()
  for name, tree in (("g_vars", g_vars), ("d_vars", d_vars), ("batch", batch)):
    leaves = m.flatten(tree)
    got = [sum(float(v.astype(np.float64).sum()) for v in leaves.values()),
           sum(float(np.abs(v.astype(np.float64)).sum()) for v in leaves.values())]
    assert np.allclose(got, N["checksum/" + name], rtol=1e-9), f"{name}: the seeded inputs changed, regenerate the fixture"
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  return cfg, to_t(g_vars), to_t(d_vars), to_t(batch), m


def test_generator_matches_the_reference_network_code():
  """Image in train mode (batch statistics) and in inference mode (running averages), and every new running average,
  against the reference's Generator / GenBlock / GenSpatialBlock / (Local)ConditionalBatchNorm / attention_for_g code
  run on the stand-in. That the reference's code finds every parameter under the product's tree names is part of the
  check (the fixture could not have been generated otherwise)."""
  cfg, g_vars, _, batch, m = _net_inputs()
  img, new = orc.generator_apply(g_vars, (batch, batch["z"]), cfg, True, orc.FP32)
  close(img, N["g_train/image"], 2e-5)
  got = m.flatten({k: v for k, v in new.items()})
  want = {k[len("g_train/new/"):]: N[k] for k in N.files if k.startswith("g_train/new/")}
  assert sorted(got) == sorted(want)
  for k in want:
    close(got[k], want[k], 2e-5)
  img_e, _ = orc.generator_apply(g_vars, (batch, batch["z"]), cfg, False, orc.FP32)
  close(img_e, N["g_eval/image"], 2e-5)
  assert np.abs(N["g_eval/image"] - N["g_train/image"]).max() > 1e-3     # the two modes really differ


def test_spectrally_normalised_generator_matches_the_reference_network_code():
  """config.g_spectral_norm = True: every generator layer through the reference's layers.SpectralConv /
  SpectralDense; image, new running averages and new power-iteration vectors."""
  from tests.golden import make_reference_golden as m
  cfg, g_vars, _, batch = m.net_inputs(g_spectral_norm=True)
  leaves = m.flatten(g_vars)
  got = [sum(float(v.astype(np.float64).sum()) for v in leaves.values()),
         sum(float(np.abs(v.astype(np.float64)).sum()) for v in leaves.values())]
  assert np.allclose(got, N["checksum/g_vars_sn"], rtol=1e-9)
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  g_vars, batch = to_t(g_vars), to_t(batch)
  img, new = orc.generator_apply(g_vars, (batch, batch["z"]), cfg, True, orc.FP32)
  close(img[:, ::4, ::4], N["g_sn_train/image_s4"], 5e-5)
  got = m.flatten({k: v for k, v in new.items()})
  want = {k[len("g_sn_train/new/"):]: N[k] for k in N.files if k.startswith("g_sn_train/new/")}
  assert sorted(got) == sorted(want) and any(k.startswith("spectral_norm_stats/") for k in want)
  for k in want:
    close(got[k], want[k], 5e-5)


def test_discriminator_matches_the_reference_network_code():
  """Logits, all 15 entries of the statistics dictionary (word / sentence / image contrastive losses, accuracies,
  entropies) and every advanced power-iteration vector, on [real; fake] with the fake half taken from the fixture."""
  cfg, _, d_vars, batch, m = _net_inputs()
  both = torch.cat([batch["image"], torch.from_numpy(N["g_train/image"])], 0)
  (logit, stats), new = orc.discriminator_apply(d_vars, (both, batch), cfg, True, orc.FP32)
  close(logit, N["d_train/logit"], 5e-5)
  want_stats = {k[len("d_train/stats/"):]: float(N[k]) for k in N.files if k.startswith("d_train/stats/")}
  assert len(want_stats) == 15
  for k, v in want_stats.items():
    assert k in stats, k
    assert abs(float(stats[k]) - v) <= 5e-5 * max(1.0, abs(v)), (k, float(stats[k]), v)
  got = m.flatten({k: v for k, v in new.items()})
  want = {k[len("d_train/new/"):]: N[k] for k in N.files if k.startswith("d_train/new/")}
  assert sorted(got) == sorted(want)
  for k in want:
    close(got[k], want[k], 2e-5)
  (logit_e, _), _ = orc.discriminator_apply(d_vars, (both, batch), cfg, False, orc.FP32)
  close(logit_e, N["d_eval/logit"], 5e-5)


def test_resnet50_matches_the_reference_network_code():
  """oracle.resnet50_apply (the frozen feature branch, inference-mode BatchNorm) against the reference's
  utils/resnet_v1.py — ResNet / ResNetStage / BottleneckResNetBlock with their explicit layer names, the stride on the
  3x3 convolution, no ReLU after init_bn, max_pool 3x3 / 2 SAME, global mean, head — run on the stand-in with the same
  synthetic weights and 224 x 224 images: pooled features and logits."""
  from tests.golden import make_reference_golden as m
  variables, images = m.resnet_inputs()
  leaves = m.flatten(variables)
  got = [sum(float(v.astype(np.float64).sum()) for v in leaves.values()),
         sum(float(np.abs(v.astype(np.float64)).sum()) for v in leaves.values()), float(images.astype(np.float64).sum())]
  assert np.allclose(got, N["resnet/checksum"], rtol=1e-9), "the seeded ResNet inputs changed, regenerate the fixture"
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  pool, logits = orc.resnet50_apply(to_t(variables), torch.from_numpy(images), orc.FP32)
  assert tuple(pool.shape) == tuple(N["resnet/pool_shape"]) == (2, 7, 7, 2048)
  close(pool[..., ::16], N["resnet/pool_c16"], 5e-5)
  close(pool.mean(dim=(1, 2)), N["resnet/pool_mean"], 5e-5)
  close(logits, N["resnet/logits"], 5e-5)


def test_gradients_match_finite_differences_of_the_reference_loss_code():
  """The two pull-backs of xmc_gan.train_g_d (d_loss wrt the discriminator's parameters, g_loss wrt the generator's,
  xmc_gan.py:162-167; train_d's gradient is the first one) as the oracle's autograd computes them, against central
  differences (float64, eps 3e-8) of the reference's OWN `loss_fn` — taken out of xmcgan/xmc_gan.py and executed with
  the reference's networks on the stand-in, `jax.lax.stop_gradient` honoured by replaying the base run's stopped values
  (the power-iteration vectors u0 / v0 of every spectrally normalised layer) in the perturbed runs. 19 directional
  derivatives: three dense random directions per network and one per selected leaf (first / deep / shortcut
  convolutions, both dense heads, the word projections, conditional-BatchNorm gamma / beta layers, the output conv).
  This pins WHICH parameters each loss differentiates, where stop_gradient cuts, and the BatchNorm-statistics and
  attention terms of the backward pass — everything autograd derives from the forward pinned above."""
  from tests.golden import make_reference_golden as m
  cfg, g_np, d_np, batch_np = m.net_inputs()
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  state = orc.make_state(to_t(g_np), to_t(d_np))
  r = orc.d_losses_and_grads(state, to_t(batch_np), cfg, orc.FP32, want_g=True)
  close([r["d_loss"], r["g_loss"]], N["grads/base"], 2e-6)
  key = lambda p: p if isinstance(p, str) else "/".join(p)
  grads = {"d": {key(p): g for p, g in orc.tree_leaves(r["d_grad"])},
           "g": {key(p): g for p, g in orc.tree_leaves(r["g_grad"])}}
  assert int(N["grads/stopped_values"]) == 40      # u0 and v0 of the discriminator's 20 spectrally normalised layers
  worst = 0.0
  for name, net, direction in m.grad_directions(g_np["params"], d_np["params"]):
    got = sum(float((grads[net][k].double().numpy() * v).sum()) for k, v in direction.items())
    want = float(N["grads/fd/" + name])
    err = abs(got - want) / max(abs(want), 0.05)
    worst = max(worst, err)
    assert err < 1e-3, (name, got, want)
  print(f"[oracle autograd vs finite differences of the reference's loss_fn] worst relative difference {worst:.1e}")


def _stats_close(stats, prefix, tol=5e-5):
  want = {k[len(prefix):]: float(N[k]) for k in N.files if k.startswith(prefix)}
  assert len(want) == 15
  for k, v in want.items():
    assert abs(float(stats[k]) - v) <= tol * max(1.0, abs(v)), (prefix, k, float(stats[k]), v)


def test_256px_configuration_matches_the_reference_network_code():
  """image_size = 256 (six blocks in both networks, xmc_net.py:81-86,202-205): generated image, discriminator logits
  and statistics."""
  from tests.golden import make_reference_golden as m
  cfg, g_np, d_np, batch_np = m.net_inputs(image_size=256)
  assert np.allclose(m.checksum(dict(g=g_np, d=d_np, b=batch_np)), N["checksum/vars_256"], rtol=1e-9)
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  g_vars, d_vars, batch = to_t(g_np), to_t(d_np), to_t(batch_np)
  img, _ = orc.generator_apply(g_vars, (batch, batch["z"]), cfg, True, orc.FP32)
  assert tuple(img.shape) == (2, 256, 256, 3)
  close(img[:, ::8, ::8], N["px256/image_s8"], 5e-5)
  close(img.mean(dim=(1, 2)), N["px256/image_mean"], 5e-5)
  (logit, stats), _ = orc.discriminator_apply(d_vars, (torch.cat([batch["image"], img.detach()]), batch), cfg, True,
                                              orc.FP32)
  close(logit, N["px256/logit"], 1e-4)
  _stats_close(stats, "px256/stats/", 1e-4)


@pytest.mark.parametrize("switch", ["word_contrastive", "sentence_contrastive", "image_contrastive"])
def test_loss_switches_match_the_reference_network_code(switch):
  """coco_xmc.py's contrastive-loss switches off one at a time: the discriminator's statistics dictionary (zeros where
  the reference leaves its initial 0) and logits; without word_contrastive the word projection is not even created."""
  from tests.golden import make_reference_golden as m
  cfg, _, d_np, batch_np = m.net_inputs(**{switch: False})
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  d_vars, batch = to_t(d_np), to_t(batch_np)
  both = torch.cat([batch["image"], torch.from_numpy(N["g_train/image"])], 0)
  (logit, stats), _ = orc.discriminator_apply(d_vars, (both, batch), cfg, True, orc.FP32)
  close(logit, N[f"no_{switch}/logit"], 5e-5)
  _stats_close(stats, f"no_{switch}/stats/")
  zeroed = {"word_contrastive": "real_word_loss", "sentence_contrastive": "fake_sentence_loss",
            "image_contrastive": "image_contrastive_loss"}[switch]
  assert float(N[f"no_{switch}/stats/{zeroed}"]) == 0.0 and float(stats[zeroed]) == 0.0
