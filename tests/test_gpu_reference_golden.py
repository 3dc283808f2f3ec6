"""The CUDA networks (fp32 mode) directly against the fixture produced by the REFERENCE'S OWN network code
(tests/golden/reference_nets.npz: xmcgan/nets/xmc_net.py + common.py + libml/layers.py executed on the numpy stand-in
for flax.linen, see tests/golden/make_reference_golden.py). Same seeded weights and batch as the fixture."""
import functools
import os

import numpy as np
import pytest
import torch

from tests import helpers

gpu = pytest.mark.gpu
N = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_nets.npz"))


@gpu
def test_cuda_networks_match_the_reference_network_code():
  """config.dtype = "float32" (3 x bf16 split GEMMs, fp32 activations): generated image (train and inference mode),
  new running averages, discriminator logits, the five contrastive losses and the advanced power-iteration vectors
  within 2e-4 of what the reference's Generator / Discriminator code computes (measured on B200: image 1.2e-5, inference
  image 1.4e-5, running averages 6e-6, logits 6e-6, losses 7e-6, u0 2e-7)."""
  from tests.golden import make_reference_golden as m
  from xmcgan_image_generation_b200.nets import xmc_net
  cfg, g_np, d_np, batch_np = m.net_inputs(dtype="float32")
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  g_vars, d_vars, batch = to_t(g_np), to_t(d_np), to_t(batch_np)
  flat = lambda lay, tree: xmc_net.FlatTree(lay, xmc_net.as_flat(lay, tree))
  g_eng, d_eng = xmc_net.get_engine(cfg, "g", m.NET_E), xmc_net.get_engine(cfg, "d", m.NET_E)
  g_params, g_stats = flat(g_eng.layout, g_vars["params"]), flat(g_eng.stats_layout, g_vars["batch_stats"])
  d_params, d_u = flat(d_eng.layout, d_vars["params"]), flat(d_eng.u_layout, d_vars["spectral_norm_stats"])

  gen = functools.partial(xmc_net.Generator, config=cfg)
  img, new = gen(train=True).apply({"params": g_params, "batch_stats": g_stats}, (batch, batch["z"]),
                                   mutable=["batch_stats"])
  e_img = helpers.rel(img, torch.from_numpy(N["g_train/image"]))
  got = m.flatten({"batch_stats": {k: v for k, v in new["batch_stats"].to_cpu_tree().items()}})
  e_stats = max(helpers.rel(torch.as_tensor(got[k[len("g_train/new/"):]]), torch.from_numpy(N[k]))
                for k in N.files if k.startswith("g_train/new/"))
  img_e = gen(train=False).apply({"params": g_params, "batch_stats": g_stats}, (batch, batch["z"]), mutable=False)
  e_eval = helpers.rel(img_e, torch.from_numpy(N["g_eval/image"]))

  both = torch.cat([batch["image"], torch.from_numpy(N["g_train/image"])], 0)
  disc = functools.partial(xmc_net.Discriminator, config=cfg)
  (logit, stat), new_d = disc(train=True).apply({"params": d_params, "spectral_norm_stats": d_u}, (both, batch),
                                                mutable=["spectral_norm_stats"])
  e_logit = helpers.rel(logit, torch.from_numpy(N["d_train/logit"]))
  e_loss = max(abs(float(stat[k]) - float(N["d_train/stats/" + k])) / max(1.0, abs(float(N["d_train/stats/" + k])))
               for k in ("real_word_loss", "fake_word_loss", "real_sentence_loss", "fake_sentence_loss",
                         "image_contrastive_loss"))
  got_u = m.flatten({"spectral_norm_stats": new_d["spectral_norm_stats"].to_cpu_tree()})
  e_u = max(helpers.rel(torch.as_tensor(got_u[k[len("d_train/new/"):]]), torch.from_numpy(N[k]))
            for k in N.files if k.startswith("d_train/new/"))
  print(f"\n[vs reference-code fixture, fp32 mode] image {e_img:.2e} eval image {e_eval:.2e} batch_stats {e_stats:.2e} "
        f"logit {e_logit:.2e} losses {e_loss:.2e} u0 {e_u:.2e}")
  assert e_img < 2e-4 and e_eval < 2e-4 and e_stats < 2e-4
  assert e_logit < 2e-4 and e_loss < 2e-4 and e_u < 2e-4


@gpu
def test_cuda_gradients_match_finite_differences_of_the_reference_loss_code():
  """The CUDA path's two pull-backs (fp32 mode) against the 19 directional derivatives obtained by central differences
  of the reference's own `loss_fn` (tests/golden/reference_nets.npz, `grads/fd/*`; see
  tests/test_reference_golden.py). A random direction v ~ N(0, 1) gives <g, v> with standard deviation |g|, so the
  difference is judged against |g| over the leaves of the direction: 2e-2 (measured on B200: 4.9e-3 worst)."""
  from tests.golden import make_reference_golden as m
  from xmcgan_image_generation_b200 import ops
  from xmcgan_image_generation_b200.nets import xmc_net
  cfg, g_np, d_np, batch_np = m.net_inputs(dtype="float32")
  to_t = lambda t: {k: to_t(v) if isinstance(v, dict) else torch.from_numpy(v) for k, v in t.items()}
  g_vars, d_vars, batch = to_t(g_np), to_t(d_np), to_t(batch_np)
  flat = lambda lay, tree: xmc_net.FlatTree(lay, xmc_net.as_flat(lay, tree))
  g_eng, d_eng = xmc_net.get_engine(cfg, "g", m.NET_E), xmc_net.get_engine(cfg, "d", m.NET_E)
  g_params, g_stats = flat(g_eng.layout, g_vars["params"]), flat(g_eng.stats_layout, g_vars["batch_stats"])
  d_params, d_u = flat(d_eng.layout, d_vars["params"]), flat(d_eng.u_layout, d_vars["spectral_norm_stats"])
  B, S = m.NET_B, cfg.image_size
  dev = xmc_net.batch_to_device(batch)
  g_eng.prep_weights(g_params.buf, None, None)
  u_new = torch.empty_like(d_u.buf)
  d_eng.prep_weights(d_params.buf, d_u.buf, u_new)
  with ops.act_dtype(g_eng.act):
    all_images = ops.empty((2 * B, S, S, 3))
    ops.cast_to_bf16(dev["image"].reshape(-1, 3), all_images[:B].view(-1, 3))
  img, gctx = g_eng.forward(g_params.buf, g_stats.buf, dev, dev["z"], train=True, fake_bf16=all_images[B:])
  losses = torch.zeros(16, device="cuda")
  _, dctx = d_eng.forward(d_params.buf, all_images, dev, losses, need_g=True)
  d_grads = torch.zeros_like(d_params.buf)
  d_eng.backward_d(dctx, d_params.buf, d_grads)
  d_eng.sn_backward(d_params.buf, d_grads, u_new)
  d_fake = d_eng.backward_g(dctx, d_params.buf)
  g_grads = torch.zeros_like(g_params.buf)
  g_eng.backward(gctx, d_fake, g_params.buf, g_grads)
  torch.cuda.synchronize()
  grads = {"d": m.flatten_any(xmc_net.FlatTree(d_eng.layout, d_grads).to_cpu_tree()),
           "g": m.flatten_any(xmc_net.FlatTree(g_eng.layout, g_grads).to_cpu_tree())}
  worst = 0.0
  for name, net, direction in m.grad_directions(g_np["params"], d_np["params"]):
    got = sum(float((grads[net][k].double().numpy() * v).sum()) for k, v in direction.items())
    norm = float(np.sqrt(sum(float((grads[net][k].double() ** 2).sum()) for k in direction)))
    want = float(N["grads/fd/" + name])
    err = abs(got - want) / max(norm, 1e-6)
    worst = max(worst, err)
    assert err < 2e-2, (name, got, want, norm)
  print(f"\n[CUDA fp32-mode gradients vs finite differences of the reference's loss_fn] worst |diff| / |g| = {worst:.1e}")
