"""Oracle parity at BASELINE width: gf = df = 96, E = 768 (coco_xmc.py defaults), B = 8 per sub-batch — the shapes of
BASELINE.json configs 1/2 (C = 1536 at 4x4 / 8x8, 24 channel chunks, the [4224][1024] LocalConditionalBatchNorm GEMM,
the cost-model tile widths and split-K counts the small-width tests never reach) — and the 256 px variant at full
width. Tolerances are the small-config ones of tests/test_gpu_parity.py or tighter; every test prints its measured
distances so that the log of a GPU run documents the margins."""
import pytest
import torch

from oracle import xmc_oracle as orc
from tests import helpers
from tests.test_gpu_parity import _build, _mods, grad_tree_report

gpu = pytest.mark.gpu


def _full_config(**kw):
  from xmcgan_image_generation_b200.configs import coco_xmc
  cfg = coco_xmc.get_config()
  cfg.update(dict(batch_size=16, pretrained_image_contrastive=False))
  cfg.update(kw)
  return cfg


def _forward_and_pullbacks(cfg, B, E, seed):
  """CUDA: one forward of G and D, both pull-backs (xmc_gan.py:127-167). Returns what the oracle is compared with."""
  _, engine, ops, _, _, xmc_net = _mods()
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, E=E, seed=seed)
  batch = helpers.make_batch(B, cfg, E=E, seed=seed + 1)
  dev = xmc_net.batch_to_device(batch)
  g_eng, d_eng = xmc_net.get_engine(cfg, "g", E), xmc_net.get_engine(cfg, "d", E)
  S = cfg.image_size
  g_eng.prep_weights(g_params.buf, None, None)
  u_new = torch.empty_like(d_u.buf)
  d_eng.prep_weights(d_params.buf, d_u.buf, u_new)
  with ops.act_dtype(g_eng.act):
    all_images = ops.empty((2 * B, S, S, 3))
    ops.cast_to_bf16(dev["image"].reshape(-1, 3), all_images[:B].view(-1, 3))
  new_stats = torch.empty_like(g_stats.buf)
  img, gctx = g_eng.forward(g_params.buf, g_stats.buf, dev, dev["z"], train=True, new_stats=new_stats,
                            fake_bf16=all_images[B:])
  losses = torch.zeros(16, device="cuda")
  logit, dctx = d_eng.forward(d_params.buf, all_images, dev, losses, need_g=True)
  d_grads = torch.zeros_like(d_params.buf)
  d_eng.backward_d(dctx, d_params.buf, d_grads)
  d_eng.sn_backward(d_params.buf, d_grads, u_new)
  d_fake = d_eng.backward_g(dctx, d_params.buf)
  g_grads = torch.zeros_like(g_params.buf)
  g_eng.backward(gctx, d_fake, g_params.buf, g_grads)
  torch.cuda.synchronize()
  got = dict(img=img.cpu(), logit=logit.cpu(), losses=losses.cpu(),
             d_grad=xmc_net.FlatTree(d_eng.layout, d_grads).to_cpu_tree(),
             g_grad=xmc_net.FlatTree(g_eng.layout, g_grads).to_cpu_tree(),
             u0=xmc_net.FlatTree(d_eng.u_layout, u_new).to_cpu_tree(),
             stats=xmc_net.FlatTree(g_eng.stats_layout, new_stats).to_cpu_tree())
  return got, batch, orc.make_state(g_vars, d_vars)


def _check_forward_and_pullbacks(cfg, B, E, seed, grad_tol, grad_cos, deep_tol=None, deep_cos=None, loss_tol=2e-3):
  """fp32 configurations are compared with the fp32 oracle, bf16 ones with the bf16 policy that also rounds cotangents.
  deep_*: separate bars for the generator leaves below the 16x16 stage (see the bf16 test's docstring)."""
  got, batch, ostate = _forward_and_pullbacks(cfg, B, E, seed)
  fp32 = cfg.dtype == "float32"
  r = orc.d_losses_and_grads(ostate, batch, cfg, orc.FP32 if fp32 else orc.Policy("bfloat16", round_grads=True),
                             want_g=True)
  S = engine_slots()
  l = got["losses"]
  e_img = helpers.rel(got["img"], r["fake"])
  e_logit = helpers.rel(got["logit"].reshape(-1), r["logit"].reshape(-1))
  print(f"\n[width {cfg.gf_dim} px {cfg.image_size} B {B} {cfg.dtype}] image rel-L2 {e_img:.3e}  logits rel-L2 {e_logit:.3e}")
  assert e_img < (1e-3 if fp32 else 2e-2)
  assert e_logit < (1e-3 if fp32 else 3e-2)
  names = dict(real_word="real_word_loss", fake_word="fake_word_loss", real_sent="real_sentence_loss",
               fake_sent="fake_sentence_loss", image="image_contrastive_loss")
  for slot, key in names.items():
    a, b = l[S[slot]].item(), r["result"][key].item()
    print(f"  {key:24s} cuda {a:.5f} oracle {b:.5f} rel {abs(a - b) / abs(b):.2e}")
    assert abs(a - b) < (2e-3 if fp32 else 5e-3) * abs(b), key
  d_scale = (l[S["hinge_d"]].abs() + l[S["real_word"]].abs() + l[S["real_sent"]].abs()).item()
  g_scale = sum(l[S[k]].abs().item() for k in ("hinge_g", "fake_word", "fake_sent", "image"))
  d_loss = (l[S["hinge_d"]] + l[S["real_word"]] + l[S["real_sent"]]).item()
  g_loss = sum(l[S[k]].item() for k in ("hinge_g", "fake_word", "fake_sent", "image"))
  print(f"  d_loss cuda {d_loss:.5f} oracle {r['d_loss'].item():.5f}   g_loss cuda {g_loss:.5f} oracle {r['g_loss'].item():.5f}")
  assert abs(d_loss - r["d_loss"].item()) < loss_tol * d_scale
  assert abs(g_loss - r["g_loss"].item()) < loss_tol * g_scale
  for (p, a), (_, b) in zip(orc.tree_leaves(got["u0"]),
                            orc.tree_leaves(r["new_discriminator_state"]["spectral_norm_stats"])):
    assert helpers.rel(a, b) < 1e-4, p
  for (p, a), (_, b) in zip(orc.tree_leaves(got["stats"]), orc.tree_leaves(r["new_generator_state"]["batch_stats"])):
    assert helpers.rel(a, b) < 1e-2, p
  for name in ("d_grad", "g_grad"):
    deep = (deep_tol, deep_cos) if name == "g_grad" else (None, None)
    worst, bad = grad_tree_report(got[name], r[name], grad_tol, grad_cos, *deep)
    print(f"  {name}: worst leaf rel-L2 {worst[0]:.3e} ({worst[1]}), lowest cosine {worst[2]:.7f} ({worst[3]})")
    assert not bad, bad


def engine_slots():
  from xmcgan_image_generation_b200 import engine
  return engine.LOSS_SLOTS


@gpu
def test_forward_and_both_pullbacks_at_baseline_width():
  """128 px, gf = df = 96, E = 768, B = 8, bf16 (the reference default): generated image 2e-2 rel-L2, logits 3e-2, the
  five contrastive losses 5e-3, d_loss / g_loss 2e-3 of their term sizes, new u0 1e-4, new batch statistics 1e-2.
  Gradients per leaf vs the oracle whose bf16 policy also rounds cotangents: discriminator 3e-2 / cosine 0.999
  (measured 1.6e-2 / 0.9999). The generator's gradients come down through the whole discriminator and then the
  generator; their distance grows with depth from 3 % (128x128 end) to 12 % (4x4 end: Dense_1, GenBlock_0), exactly
  like the distance between the ORACLE's own bf16 and fp32 policies on the same leaves (0.10 rel-L2 / cosine 0.995 at
  Dense_1; rounding cotangents as well moves the oracle by only 5e-3): bf16 conditioning, not a formula error. They
  get 1.6e-1 / 0.985 here as a sanity bound; the sharp check of every backward formula at this width is the fp32-mode
  test below (cosine >= 0.9999 on every leaf of both networks)."""
  _check_forward_and_pullbacks(_full_config(), 8, 768, 21, 3e-2, 0.999, 1.6e-1, 0.985)


@gpu
def test_forward_and_both_pullbacks_at_256px_full_width():
  """BASELINE config 4's network (image_size = 256, gf = df = 96: one more block in G and D) at B = 2. With two images
  the 4x4 BatchNorm statistics come from 32 elements per channel; d_loss / g_loss move by +-0.015 (1e-3) with the
  summation order of those statistics alone (measured across builds), hence 5e-3 of the term sizes here."""
  _check_forward_and_pullbacks(_full_config(image_size=256, batch_size=4), 2, 768, 31, 5e-2, 0.999, 2e-1, 0.98,
                               loss_tol=5e-3)


@gpu
def test_fp32_mode_forward_and_both_pullbacks_at_baseline_width():
  """config.dtype = "float32" (train_utils.py:148-151) at gf = df = 96, E = 768, B = 8 vs the fp32 oracle: fp32
  activations, every GEMM as three bf16 tensor-core passes over [hi | lo] splits of both operands (16 mantissa bits
  per operand, fp32 accumulation — SURVEY.md 8c(4)). Bars of SURVEY.md 8c(4): losses rel 2e-3, gradient cosine
  >= 0.9999 on EVERY leaf of both networks (rel-L2 1.5e-2); image / logits 1e-3."""
  _check_forward_and_pullbacks(_full_config(dtype="float32"), 8, 768, 21, 1.5e-2, 0.9999)


@gpu
def test_train_step_at_baseline_width_matches_oracle():
  """One full train_step (train_d + train_g_d, Adam x3, EMA) at gf = df = 96, E = 768, B = 8 per sub-batch, from a
  mid-training optimiser state (helpers.warm_adam). train_g_d's metrics are taken after train_d's update of 88 M
  discriminator parameters, which turns a 1.6 % gradient difference (bf16) into a 0.6 % difference of hinge_g, so the
  two halves are checked separately and each sharply:
    (1) train_d: the discriminator UPDATE per leaf rel-L2 5e-2 vs the oracle's;
    (2) train_g_d from the SAME updated discriminator (the oracle continues from the CUDA path's parameters / moments
        / u0): all metrics within 2e-3 of the largest, generator and discriminator updates per leaf (5e-2; 2e-1 bf16
        sanity bound on the generator, see above), EMA 1e-5, step counters exact."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  cfg = _full_config()
  B, E = 8, 768
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, E=E, seed=41)
  batch = helpers.make_batch(2 * B, cfg, E=E, seed=42)
  ostate = orc.make_state(g_vars, d_vars)
  state = train_utils.TrainState(0, train_utils.Optimizer(g_params, cfg.g_lr, cfg.beta1, cfg.beta2),
                                 train_utils.Optimizer(d_params, cfg.d_lr, cfg.beta1, cfg.beta2),
                                 {"batch_stats": g_stats}, {"spectral_norm_stats": d_u}, g_params.clone())
  helpers.warm_adam(state, ostate)
  pol = orc.Policy("bfloat16", round_grads=True)
  b0, b1 = train_utils.split_input_dict(xmc_net.batch_to_device(batch), 2)
  ob0, ob1 = orc.split_input_dict(batch, 2)
  # ---- (1) train_d ----------------------------------------------------------------------------------------------
  d_old = d_params.to_cpu_tree()
  state = xmc_gan.train_d(None, state, b0, None, None, cfg)
  ostate1, _ = orc.train_d(ostate, ob0, cfg, pol)
  d_mid = state.d_optimizer.target.to_cpu_tree()
  helpers.check_updates("train_d: d", d_mid, d_old, ostate1["d_params"], d_vars["params"], cfg.d_lr)
  # ---- (2) train_g_d from the same discriminator -------------------------------------------------------------------
  lay = state.d_optimizer.target.layout
  ostate1["d_params"] = d_mid
  ostate1["d_opt"] = {"step": state.d_optimizer.step, "m": xmc_net.FlatTree(lay, state.d_optimizer.m).to_cpu_tree(),
                      "v": xmc_net.FlatTree(lay, state.d_optimizer.v).to_cpu_tree()}
  ostate1["discriminator_state"] = {"spectral_norm_stats": state.discriminator_state["spectral_norm_stats"].to_cpu_tree()}
  g_old = state.g_optimizer.target.to_cpu_tree()
  state, metrics = xmc_gan.train_g_d(None, state, b1, None, None, cfg, {})
  got = metrics.compute()
  ostate2, want, _ = orc.train_g_d(ostate1, ob1, cfg, pol)
  scale = max(abs(v) for v in want.values())
  print("\n[train_step @ baseline width]", {k: (round(got[k], 5), round(want[k], 5)) for k in want})
  for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g"):
    assert abs(got[k] - want[k]) < 2e-3 * scale, (k, got[k], want[k])
  assert (state.step, state.d_optimizer.step, state.g_optimizer.step) == (1, 102, 101)
  helpers.check_updates("train_g_d: g", state.g_optimizer.target.to_cpu_tree(), g_old, ostate2["g_params"],
                        g_vars["params"], cfg.g_lr)
  helpers.check_updates("train_g_d: d", state.d_optimizer.target.to_cpu_tree(), d_mid, ostate2["d_params"], d_mid,
                        cfg.d_lr)
  worst = max((helpers.rel(a, b), p) for (p, a), (_, b) in
              zip(orc.tree_leaves(state.ema_params.to_cpu_tree()), orc.tree_leaves(ostate2["ema_params"])))
  assert worst[0] < 1e-5, worst
