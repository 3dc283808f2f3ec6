"""Parity of the B200 CUDA path (through the C ABI / the reference-shaped Python API) against the CPU oracle.
Tolerances (stated per test): the kernels use bf16 operands / bf16 activation storage with fp32 accumulation, the
oracle's bf16 policy rounds at the same storage points; index / mask / split semantics are exact."""
import pytest
import torch

from oracle import xmc_oracle as orc
from tests import helpers

gpu = pytest.mark.gpu
# per-leaf gradient bars vs the oracle with rounded cotangents: (rel-L2, cosine). The 256 px variant runs at width 8
# with batch 2 (its BatchNorm statistics come from very few elements per channel at 4x4).
GRAD_TOL = (3e-2, 0.999)
GRAD_TOL_256 = (6e-2, 0.998)
# generator leaves in bf16 mode (see _is_deep_generator_leaf and tests/test_gpu_baseline_width.py): conditioned at up to
# ~20x the bf16 epsilon at the 4x4 end (the oracle's bf16 and fp32 policies differ by 0.10 there); a sanity bound only —
# the SHARP gradient check is the float32 parametrisation of the same test (cosine >= 0.9999 on every leaf)
GRAD_TOL_DEEP = (2e-1, 0.98)


def _mods():
  from xmcgan_image_generation_b200 import _lib, engine, ops, train_utils, xmc_gan
  from xmcgan_image_generation_b200.nets import xmc_net
  return _lib, engine, ops, train_utils, xmc_gan, xmc_net


def _q(t):
  return t.to(torch.bfloat16).float()


@gpu
def test_native_library_is_loaded():
  _lib, *_ = _mods()
  L = _lib.lib()
  assert L.xmc_num_sms() > 0
  assert _lib._LIB is not None


@gpu
@pytest.mark.parametrize("N,H,W,C,Cout,k", [(2, 16, 16, 64, 64, 3), (3, 32, 32, 96, 192, 3), (16, 4, 4, 192, 64, 3),
                                             (1, 128, 128, 96, 96, 3), (5, 8, 8, 64, 32, 1), (2, 4, 4, 16, 16, 3),
                                             # resident-weights / halo-row kernel and tap3 wgrad: two W tiles with a
                                             # 16-channel ragged chunk, a single 64-channel chunk, non-square maps
                                             (1, 4, 256, 80, 48, 3), (2, 2, 128, 64, 96, 3), (1, 6, 128, 96, 32, 3)])
def test_conv_forward_and_wgrad_match_oracle(N, H, W, C, Cout, k):
  """tcgen05 implicit-GEMM conv vs the oracle's conv2d on bf16-rounded operands: fp32 results within 1e-4 rel."""
  _, _, ops, *_ = _mods()
  torch.manual_seed(N * 100 + C)
  x = _q(torch.randn(N, H, W, C))
  kern = _q(torch.randn(k, k, C, Cout) * 0.05)
  bias = torch.randn(Cout)
  want = orc.conv2d(x, kern, bias)
  wk = kern.permute(3, 0, 1, 2).reshape(Cout, k * k * C).contiguous().cuda().to(torch.bfloat16)
  got = ops.conv_fwd(x.cuda().to(torch.bfloat16), wk, k, Cout, bias=bias.cuda(), out_dtype=torch.float32)
  assert helpers.rel(got, want) < 1e-4
  # weight gradient for a given output cotangent
  dy = _q(torch.randn(N, H, W, Cout) * 0.1)
  kp = kern.clone().requires_grad_(True)
  (orc.conv2d(x, kp, None) * dy).sum().backward()
  out = torch.zeros(k * k * C * Cout, device="cuda")
  ops.wgrad(x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16), k, out, out_mode=0, ld_out=Cout,
            tap_stride=C * Cout)
  assert helpers.rel(out.view(k, k, C, Cout), kp.grad) < 1e-4


@gpu
def test_wgrad_is_the_adjoint_of_conv_at_full_size():
  """Size-independent property at BASELINE's largest layer (56x128x128, 96->96): <conv(x,W), dy> == <W, wgrad(x,dy)>."""
  _, _, ops, *_ = _mods()
  torch.manual_seed(0)
  N, S, C = 56, 128, 96
  x = (torch.randn(N, S, S, C, device="cuda") * 0.5).to(torch.bfloat16)
  dy = (torch.randn(N, S, S, C, device="cuda") * 0.1).to(torch.bfloat16)
  w = (torch.randn(C, 9 * C, device="cuda") * 0.05).to(torch.bfloat16)
  y = ops.conv_fwd(x, w, 3, C, out_dtype=torch.float32)
  lhs = (y.double() * dy.double()).sum().item()
  dw = torch.zeros(9 * C * C, device="cuda")
  ops.wgrad(x, dy, 3, dw, out_mode=0, ld_out=C, tap_stride=C * C)
  w_hwio = w.float().view(C, 9, C).permute(1, 2, 0).reshape(-1)  # [co][tap][ci] -> [tap][ci][co]
  rhs = (dw.double() * w_hwio.double()).sum().item()
  assert abs(lhs - rhs) < 2e-3 * abs(lhs)


@gpu
def test_attention_for_g_matches_oracle_and_single_word_kat():
  from xmcgan_image_generation_b200.libml import attention_lib
  torch.manual_seed(1)
  B, R, L, D = 3, 256, 17, 64
  q = _q(torch.randn(B, R, D))
  w = torch.randn(B, L, D) * 0.5
  max_len = torch.tensor([[3.0], [17.0], [9.0]])
  mask = (torch.arange(L)[None, :] >= max_len).float()[:, None, :].repeat(1, R, 1)
  want, want_attn = orc.attention_for_g(q, w, 15.0, mask)
  got, attn = attention_lib.attention_for_g(q, w, 15.0, mask)
  assert helpers.rel(got, want) < 5e-3  # bf16 output rounding
  assert helpers.rel(attn, want_attn) < 1e-4
  assert (attn.cpu()[0, :, 3:] == 0).all()  # padded words get exactly zero weight
  one = torch.ones(B, R, L)
  one[:, :, 0] = 0
  got1, _ = attention_lib.attention_for_g(q, w, 15.0, one)
  assert helpers.rel(got1, orc.l2_normalize(w)[:, :1].expand(B, R, D)) < 5e-3


@gpu
@pytest.mark.parametrize("B", [3, 8])
def test_word_loss_and_contrastive_match_oracle(B):
  """B=3 exercises the padded (B*L not a multiple of 8) layout. bf16 region/word operands: 3e-3 rel on the loss."""
  from xmcgan_image_generation_b200.libml import attention_lib
  torch.manual_seed(2)
  R, L, D = 256, 17, 64
  img = _q(torch.randn(B, R, D))
  words = torch.randn(B, L, D) * 0.5
  max_len = torch.randint(1, L + 1, (B, 1)).float()
  want, _, _ = orc.word_loss(img, words, max_len)
  got, _, _ = attention_lib.word_loss(img, words, max_len)
  assert abs(got.item() - want.item()) < 3e-3 * abs(want.item())
  a, b = torch.randn(B, 96), torch.randn(B, 96)
  want, _, _ = orc.contrastive_loss(a, b)
  got, _, _ = attention_lib.contrastive_loss(a, b)
  assert abs(got.item() - want.item()) < 1e-4 * abs(want.item())
  eye = torch.eye(B, 96)
  got, _, _ = attention_lib.contrastive_loss(eye, eye)
  import math
  assert abs(got.item() - 2 * math.log(1 + (B - 1) * math.exp(-10.0))) < 1e-5


@gpu
def test_hinge_and_resampling_known_answers():
  from xmcgan_image_generation_b200.libml import losses
  from xmcgan_image_generation_b200.nets import common
  d, g = losses.hinge_loss(torch.zeros(5, 1), torch.zeros(5, 1))
  assert d.item() == 2.0 and g.item() == 0.0
  r, f = torch.randn(7, 1), torch.randn(7, 1)
  wd, wg = orc.hinge_loss(r, f)
  d, g = losses.hinge_loss(r, f)
  assert abs(d.item() - wd.item()) < 1e-6 and abs(g.item() - wg.item()) < 1e-6
  x = _q(torch.randn(2, 8, 8, 16))
  assert torch.equal(common.upsample(x).float().cpu(), orc.upsample(x))            # index op: exact
  assert helpers.rel(common.dsample(x), orc.dsample(x)) < 4e-3
  img = _q(torch.rand(2, 8, 8, 3))
  assert helpers.rel(common.dsample(img), orc.dsample(img)) < 4e-3


@gpu
@pytest.mark.parametrize("Hc,H,up", [(1, 4, True), (1, 8, False), (16, 16, True), (16, 64, False), (16, 32, True)])
def test_conditional_batch_norm_forward_backward_single_op(Hc, H, up):
  """(Local)ConditionalBatchNorm + relu (+ nearest upsample) and its full backward (dx, dgamma, dbeta) on identical
  bf16 inputs vs oracle autograd: 1e-2 rel-L2 (outputs are bf16). No cross-layer error amplification here, so this is
  the sharp check of the backward formulas."""
  _, _, ops, *_ = _mods()
  torch.manual_seed(H * 10 + Hc)
  N, C = 3, 32
  x = _q(torch.randn(N, H, H, C) * 2 + 0.5).requires_grad_(True)
  rows = N * Hc * Hc
  gb = _q(torch.randn(rows, 2 * C + 16) * 0.5).requires_grad_(True)   # gamma at col 8, beta at col 8+C
  goff, boff = 8, 8 + C
  s = H // Hc
  idx = torch.arange(H) // s
  g4 = gb[:, goff:goff + C].reshape(N, Hc, Hc, C)[:, idx][:, :, idx]
  b4 = gb[:, boff:boff + C].reshape(N, Hc, Hc, C)[:, idx][:, :, idx]
  xh, _ = orc.batch_norm(x, {"mean": torch.zeros(C), "var": torch.ones(C)}, True)
  y = torch.relu(xh * (g4 + 1.0) + b4)
  if up:
    y = orc.upsample(y)
  dy = _q(torch.randn_like(y) * 0.1)
  (y * dy).sum().backward()
  xd, gbd = x.detach().cuda().to(torch.bfloat16), gb.detach().cuda().to(torch.bfloat16)
  sums, P = ops.bn_stats(xd)
  mr = ops.bn_finalize(sums, P, C, torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), None, None)
  got = ops.bn_apply(xd, mr, gbd, Hc, goff, boff, True, up)
  assert helpers.rel(got, y) < 1e-2
  dgb = torch.zeros(rows, 2 * C + 16, device="cuda")
  dx = ops.bn_bwd(dy.cuda().to(torch.bfloat16), xd, mr, gbd, dgb, Hc, goff, boff, True, up)
  assert helpers.rel(dx, x.grad) < 1e-2
  assert helpers.rel(dgb[:, goff:goff + C], gb.grad[:, goff:goff + C]) < 1e-2
  assert helpers.rel(dgb[:, boff:boff + C], gb.grad[:, boff:boff + C]) < 1e-2


@gpu
def test_attention_and_loss_heads_backward_single_op():
  """Backward of attention_for_g, word_loss and contrastive_loss on identical inputs vs oracle autograd."""
  _, engine, ops, *_ = _mods()
  torch.manual_seed(5)
  B, R, L, D = 4, 256, 17, 64
  max_len = torch.tensor([[5.0], [17.0], [1.0], [9.0]])
  words = torch.randn(B, L, D) * 0.5
  # attention_for_g: d(ctx)/dq
  q = _q(torch.randn(B, R, D)).requires_grad_(True)
  mask = (torch.arange(L)[None, :] >= max_len).float()[:, None, :].repeat(1, R, 1)
  ctx, _ = orc.attention_for_g(q, words, 15.0, mask)
  dctx = _q(torch.randn(B, R, D) * 0.1)
  (ctx * dctx).sum().backward()
  what, _ = ops.l2norm_rows(words.cuda().reshape(B * L, D))
  qd = q.detach().cuda().to(torch.bfloat16)
  out = ops.empty((B * R, D))
  attn = ops.attention_g_fwd(qd, what.view(B, L, D), max_len.cuda().reshape(B), 15.0, out)
  dq = ops.attention_g_bwd(dctx.cuda().to(torch.bfloat16).view(B * R, D), qd, what.view(B, L, D), attn, 15.0)
  assert helpers.rel(dq, q.grad) < 1e-2
  # word_loss: d(loss)/d(image regions)
  img = _q(torch.randn(B, R, D)).requires_grad_(True)
  loss, _, _ = orc.word_loss(img, words, max_len)
  loss.backward()
  ws = engine.WordShared(words.cuda(), max_len.cuda())
  slot = ops.empty(1, torch.float32)
  wl = engine.WordLoss(img.detach().cuda().to(torch.bfloat16), ws, slot)
  dR = wl.bwd()
  assert abs(slot.item() - loss.item()) < 3e-3 * abs(loss.item())
  assert helpers.rel(dR.view(B, R, D), img.grad) < 3e-2   # bf16 alpha / dS / dctx operands
  # contrastive_loss: both cotangents
  a = torch.randn(B, 96).requires_grad_(True)
  b = torch.randn(B, 96).requires_grad_(True)
  loss, _, _ = orc.contrastive_loss(a, b)
  loss.backward()
  c = engine.Contrastive(a.detach().cuda(), b.detach().cuda(), slot)
  da, db = torch.zeros(B, 96, device="cuda"), torch.zeros(B, 96, device="cuda")
  c.bwd_a(da)
  c.bwd_b(db)
  assert helpers.rel(da, a.grad) < 1e-4 and helpers.rel(db, b.grad) < 1e-4


def _build(config, E=64, seed=1):
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  g_eng, d_eng, g_vars, d_vars = helpers.cpu_variables(config, E, seed)
  to_flat = lambda lay, tree: xmc_net.FlatTree(lay, xmc_net.as_flat(lay, tree))
  g_params = to_flat(xmc_net.get_engine(config, "g", E).layout, g_vars["params"])
  g_stats = to_flat(xmc_net.get_engine(config, "g", E).stats_layout, g_vars["batch_stats"])
  d_params = to_flat(xmc_net.get_engine(config, "d", E).layout, d_vars["params"])
  d_u = to_flat(xmc_net.get_engine(config, "d", E).u_layout, d_vars["spectral_norm_stats"])
  return g_vars, d_vars, g_params, g_stats, d_params, d_u


def _g_u0(config, g_vars, E=64):
  """The generator's spectral_norm_stats collection on the device (g_spectral_norm=True only)."""
  *_, xmc_net = _mods()
  lay = xmc_net.get_engine(config, "g", E).u_layout
  return xmc_net.FlatTree(lay, xmc_net.as_flat(lay, g_vars["spectral_norm_stats"]))


@gpu
def test_generator_and_discriminator_apply_match_oracle():
  """Through the Flax-shaped module API. Image: 2e-2 rel-L2 (bf16 activations through 11 BN layers); logits 3e-2;
  contrastive losses 5e-3; new batch_stats / u0 1e-2 / 1e-4."""
  *_, xmc_net = _mods()
  import functools
  cfg = helpers.small_config()
  B = 4
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg)
  batch = helpers.make_batch(B, cfg)
  gen = functools.partial(xmc_net.Generator, config=cfg)
  img, new = gen(train=True).apply({"params": g_params, "batch_stats": g_stats}, (batch, batch["z"]),
                                   mutable=["batch_stats"])
  want, upd = orc.generator_apply(g_vars, (batch, batch["z"]), cfg, True, orc.Policy("bfloat16"))
  assert img.shape == (B, 128, 128, 3) and img.min() >= 0 and img.max() <= 1
  assert helpers.rel(img, want) < 2e-2
  for (p, a), (_, b) in zip(orc.tree_leaves(new["batch_stats"].to_cpu_tree()), orc.tree_leaves(upd["batch_stats"])):
    assert helpers.rel(a, b) < 1e-2, p
  # eval mode uses the running statistics and leaves them untouched
  img_eval = gen(train=False).apply({"params": g_params, "batch_stats": g_stats}, (batch, batch["z"]), mutable=False)
  want_eval, _ = orc.generator_apply(g_vars, (batch, batch["z"]), cfg, False, orc.Policy("bfloat16"))
  assert helpers.rel(img_eval, want_eval) < 2e-2

  disc = functools.partial(xmc_net.Discriminator, config=cfg)
  all_images = torch.cat([batch["image"], want.detach()])
  (logit, stat), new_d = disc(train=True).apply({"params": d_params, "spectral_norm_stats": d_u},
                                                (all_images, batch), mutable=["spectral_norm_stats"])
  (wlogit, wstat), wupd = orc.discriminator_apply(d_vars, (all_images, batch), cfg, True, orc.Policy("bfloat16"))
  assert logit.shape == (2 * B, 1)
  assert helpers.rel(logit, wlogit) < 3e-2
  for k in ("real_word_loss", "fake_word_loss", "real_sentence_loss", "fake_sentence_loss", "image_contrastive_loss"):
    assert abs(stat[k].item() - wstat[k].item()) < 5e-3 * abs(wstat[k].item()), k
  for (p, a), (_, b) in zip(orc.tree_leaves(new_d["spectral_norm_stats"].to_cpu_tree()),
                            orc.tree_leaves(wupd["spectral_norm_stats"])):
    assert helpers.rel(a, b) < 1e-4, p


def _is_deep_generator_leaf(path):
  """Every generator leaf: in bf16 mode the generator's gradients pass through the whole discriminator and then down
  the generator; the distance to the oracle grows steadily with depth (3 % at the 128x128 end, 10-15 % at the 4x4
  end) — and so does the distance between the ORACLE's own bf16 and fp32 policies (tests/test_gpu_baseline_width.py).
  The discriminator's own gradients stay within 3e-2."""
  return True


def grad_tree_report(got_tree, ref_tree, tol, min_cos, deep_tol=None, deep_cos=None):
  """Per-leaf comparison of a gradient tree with the oracle's: rel-L2 <= tol and cosine >= min_cos for every leaf
  whose true gradient is not numerically zero; the others (biases in front of a BatchNorm) are compared in absolute
  terms against the largest leaf. deep_tol / deep_cos: separate bars for the generator's leaves in bf16 mode.
  Returns ((worst rel, its leaf, lowest cosine, its leaf), [violations])."""
  ref = orc.tree_leaves(ref_tree)
  scale = max(r.norm().item() for _, r in ref)
  bad, worst_e, worst_c = [], (-1.0, ""), (2.0, "")
  for (path, g), (_, r) in zip(orc.tree_leaves(got_tree), ref):
    g = g.float().cpu()
    if r.norm().item() > 1e-4 * scale:
      e = helpers.rel(g, r)
      cos = torch.nn.functional.cosine_similarity(g.reshape(-1), r.reshape(-1), dim=0).item()
      worst_e = max(worst_e, (e, path))
      worst_c = min(worst_c, (cos, path))
      t, c = (deep_tol, deep_cos) if (deep_tol is not None and _is_deep_generator_leaf(path)) else (tol, min_cos)
      if not (e <= t and cos >= c):
        bad.append((path, round(e, 4), round(cos, 6)))
    elif (g - r).norm().item() > 1e-3 * scale:
      bad.append((path, "abs", (g - r).norm().item(), scale))
  return (worst_e[0], worst_e[1], worst_c[0], worst_c[1]), bad


@gpu
@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
@pytest.mark.parametrize("variant", ["default", "no_sn", "no_word", "ragged", "px256", "g_sn"])
def test_both_pullbacks_match_oracle(variant, dtype):
  """d(d_loss)/d(params_d) and d(g_loss)/d(params_g) from ONE forward (xmc_gan.py:162-167) vs oracle autograd under
  Policy("bfloat16", round_grads=True) — the bf16 policy that also rounds activation COTANGENTS at the storage points,
  which is what the CUDA path (and the reference's bf16 graph) does. Per leaf: rel-L2 <= GRAD_TOL, cosine >= GRAD_COS;
  losses 2e-3 of their term sizes."""
  _, engine, ops, _, _, xmc_net = _mods()
  kw = {"no_sn": dict(d_spectral_norm=False), "no_word": dict(word_contrastive=False),
        "px256": dict(image_size=256, gf_dim=8, df_dim=8), "g_sn": dict(g_spectral_norm=True)}.get(variant, {})
  cfg = helpers.small_config(dtype=dtype, **kw)
  fp32 = dtype == "float32"
  B = {"ragged": 3, "px256": 2}.get(variant, 4)
  # 256 px: one more block in G and D, batch 2, width 8 — the configuration most sensitive to bf16 perturbations
  # (the oracle's own bf16-vs-fp32 gradients differ by > 10 % there); the sharp backward checks are the single-op tests
  tol, min_cos = GRAD_TOL_256 if variant == "px256" else GRAD_TOL
  if fp32:   # fp32 mode vs the fp32 oracle: SURVEY.md 8c(4) bars on every leaf
    tol, min_cos = 1.5e-2, 0.9999
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=4)
  batch = helpers.make_batch(B, cfg, seed=2, min_len=1 if variant == "ragged" else 3)
  dev = xmc_net.batch_to_device(batch)
  g_eng, d_eng = xmc_net.get_engine(cfg, "g", 64), xmc_net.get_engine(cfg, "d", 64)
  S = cfg.image_size
  g_u = _g_u0(cfg, g_vars).buf if g_eng.sn else None
  g_u_new = torch.empty_like(g_u) if g_eng.sn else None
  g_eng.prep_weights(g_params.buf, g_u, g_u_new)
  u_new = torch.empty_like(d_u.buf)
  d_eng.prep_weights(d_params.buf, d_u.buf if d_eng.sn else None, u_new if d_eng.sn else None)
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
  g_eng.sn_backward(g_params.buf, g_grads, g_u_new)
  torch.cuda.synchronize()
  state = orc.make_state(g_vars, d_vars if d_eng.sn else {"params": d_vars["params"]})
  r = orc.d_losses_and_grads(state, batch, cfg, orc.FP32 if fp32 else orc.Policy("bfloat16", round_grads=True),
                             want_g=True)
  l = losses.cpu()
  # the totals are sums of terms of mixed sign (hinge_g = -mean(fake logit)): tolerance relative to the term sizes
  d_scale = (l[0].abs() + l[2].abs() + l[4].abs()).item()
  g_scale = (l[1].abs() + l[3].abs() + l[5].abs() + l[6].abs()).item()
  assert abs((l[0] + l[2] + l[4]).item() - r["d_loss"].item()) < 2e-3 * d_scale
  assert abs((l[1] + l[3] + l[5] + l[6]).item() - r["g_loss"].item()) < 2e-3 * g_scale
  for name, lay, flat in (("d_grad", d_eng.layout, d_grads), ("g_grad", g_eng.layout, g_grads)):
    deep = GRAD_TOL_DEEP if (name == "g_grad" and not fp32) else (None, None)
    worst, bad = grad_tree_report(xmc_net.FlatTree(lay, flat).to_cpu_tree(), r[name], tol, min_cos, *deep)
    print(f"\n[{variant} {dtype}] {name}: worst leaf rel-L2 {worst[0]:.3e} ({worst[1]}), lowest cosine {worst[2]:.5f} ({worst[3]})")
    assert not bad, bad


@gpu
def test_train_step_matches_oracle_for_two_steps():
  """Public API: train_utils.train_step == oracle train_step (bf16 policy): metrics 5e-3 rel, parameters 5e-3,
  EMA 1e-4, batch_stats 1e-2, u0 3e-2; step counters exact (D's Adam advances twice per step). u0 is one
  power-iteration step on weights that went through 3 Adam updates: a relative weight difference d shows up in u0
  amplified by s1/(s1-s2) of that kernel, and the split-K reductions of the weight gradients are atomics (run-to-run
  order), so the worst layer sits around 1e-2 here; the single-forward test above holds u0 to 1e-4."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  cfg = helpers.small_config()
  B = 4
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=6)
  batch = helpers.make_batch(2 * B, cfg, seed=7)
  ostate = orc.make_state(g_vars, d_vars)
  state = train_utils.TrainState(0, train_utils.Optimizer(g_params, cfg.g_lr, cfg.beta1, cfg.beta2),
                                 train_utils.Optimizer(d_params, cfg.d_lr, cfg.beta1, cfg.beta2),
                                 {"batch_stats": g_stats}, {"spectral_norm_stats": d_u}, g_params.clone())
  for _ in range(2):
    state, metrics = train_utils.train_step(None, state, batch, xmc_gan, None, None, cfg, {})
    got = metrics.compute()
    ostate, want = orc.train_step(ostate, batch, cfg, orc.Policy("bfloat16"))
    scale = max(abs(v) for v in want.values())
    for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g"):
      assert abs(got[k] - want[k]) < 5e-3 * scale, (k, got[k], want[k])
    assert got["c_loss_g_pretrained"] == 0.0
  assert (state.step, state.d_optimizer.step, state.g_optimizer.step) == (2, 4, 2)
  pairs = [(state.g_optimizer.target, ostate["g_params"], 5e-3), (state.d_optimizer.target, ostate["d_params"], 5e-3),
           (state.ema_params, ostate["ema_params"], 1e-4),
           (state.generator_state["batch_stats"], ostate["generator_state"]["batch_stats"], 1e-2),
           (state.discriminator_state["spectral_norm_stats"], ostate["discriminator_state"]["spectral_norm_stats"],
            3e-2)]
  for got_t, want_t, tol in pairs:
    for (p, a), (_, b) in zip(orc.tree_leaves(got_t.to_cpu_tree()), orc.tree_leaves(want_t)):
      assert helpers.rel(a, b) < tol, (p, helpers.rel(a, b))


@gpu
def test_train_step_runs_at_baseline_width_and_decreases_nothing_to_nan():
  """Full-width networks (gf=df=96, E=768) at a small batch: losses finite, EMA moves, D sees 2B images."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  from xmcgan_image_generation_b200.configs import coco_xmc
  cfg = coco_xmc.get_config()
  cfg.update(dict(batch_size=8, pretrained_image_contrastive=False))
  batch = helpers.make_batch(16, cfg, E=768)
  gen, disc, state = train_utils.create_train_state(cfg, 42, batch)
  ema0 = state.ema_params.buf.clone()
  state, metrics = train_utils.train_step(None, state, batch, xmc_gan, gen, disc, cfg, {})
  m = metrics.compute()
  assert all(torch.isfinite(torch.tensor(v)) for v in m.values())
  assert not torch.equal(ema0, state.ema_params.buf)
  assert torch.isfinite(state.g_optimizer.target.buf).all() and torch.isfinite(state.d_optimizer.target.buf).all()


@gpu
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_frozen_resnet50_branch_matches_oracle(dtype):
  """calculate_contrastive_loss_on_pretrained (xmc_gan.py:74-90) on the frozen ResNet-50.
  float32 (default; the reference's precision for this network): logits 1e-3 rel-L2, loss 1e-4, and the end-to-end
  input gradient through all 50 layers, max-pool, stem and resize vs the fp32 oracle: rel-L2 2e-2, cosine 0.9995.
  bfloat16 (opt-in): logits 2e-2, loss 2e-3; its input gradient is chaotic under bf16 perturbations (the oracle's own
  bf16-vs-fp32 gradients differ by 20-40 %), so it is only bounded by 1.5x that distance — the reason the default is
  float32."""
  _, engine, ops, _, xmc_gan, _ = _mods()
  torch.manual_seed(3)
  variables = orc.resnet50_random_variables(1)
  model = engine.ResNetEngine(dtype=dtype)
  model.load(variables)
  fp32 = dtype == "float32"
  B, S = 3, 128
  real = torch.rand(B, S, S, 3)
  fake = torch.rand(B, S, S, 3)
  both = torch.cat([real, fake]).cuda()
  logits, rctx = model.forward(both)
  pol = orc.FP32 if fp32 else orc.Policy("bfloat16")
  _, want = orc.get_pretrained_embs(variables, torch.cat([real, fake]), pol)
  assert logits.shape == (2 * B, 1000)
  e_logits = helpers.rel(logits, want)
  slot = ops.empty(1, torch.float32)
  c = engine.Contrastive(logits[:B], logits[B:], slot)
  grads = {}
  for name, p in (("own", pol), ("fp32", orc.FP32)):
    fk = fake.clone().requires_grad_(True)
    loss = orc.calculate_contrastive_loss_on_pretrained(variables, real, fk, p)
    loss.backward()
    grads[name] = fk.grad.reshape(-1)
    if name == "own":
      e_loss = abs(slot.item() - loss.item()) / abs(loss.item())
  dl = ops.empty((B, 1000), torch.float32)
  c.bwd_b(dl, accumulate=False)
  d_fake = torch.zeros(B, S, S, 3, device="cuda")
  model.backward(rctx, dl, B, d_fake)
  g = d_fake.cpu().reshape(-1)
  e_grad = helpers.rel(g, grads["own"])
  cos = torch.nn.functional.cosine_similarity(g, grads["own"], dim=0).item()
  print(f"\n[resnet {dtype}] logits rel-L2 {e_logits:.3e} loss rel {e_loss:.3e} input-gradient rel-L2 {e_grad:.3e} cos {cos:.6f}")
  if fp32:
    assert e_logits < 1e-3 and e_loss < 1e-4
    assert e_grad < 2e-2 and cos > 0.9995
  else:
    assert e_logits < 2e-2 and e_loss < 2e-3
    noise = helpers.rel(grads["own"], grads["fp32"])
    assert e_grad < max(0.1, 1.5 * noise), (e_grad, noise)
    assert cos > 0.9


@gpu
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_resnet_pieces_forward_and_backward_sharp(dtype):
  """Sharp checks of every piece of the ResNet branch on identical inputs: stem (resize + 7x7/2 conv + folded BN
  + 3x3/2 max-pool) forward and its (linear) input gradient; bottleneck blocks with stride 1 / stride 2 + projection:
  forward and input gradient (single block, so relu-mask disagreement is negligible).
  bfloat16: 1e-2 / 2e-2 / 1e-2 / 3e-2. float32 (vs the fp32 oracle): forward 1e-4 (measured 2e-6), block input gradients 1e-3 (measured 1.4e-4: relu masks
  of pre-activations within the forward's agreement of zero); the
  stem's input gradient 1e-2: the max-pool routes each window's gradient to its arg-max, and the few windows whose two
  largest elements differ by less than the forward's 4e-6 agreement route it to a different pixel than the oracle
  (measured 5e-3; bf16: 3e-3)."""
  _, engine, ops, *_ = _mods()
  torch.manual_seed(5)
  variables = orc.resnet50_random_variables(4)
  model = engine.ResNetEngine(dtype=dtype)
  model.load(variables)
  fp32 = dtype == "float32"
  pol = orc.FP32 if fp32 else orc.Policy("bfloat16")
  adt = torch.float32 if fp32 else torch.bfloat16
  q = (lambda t: t) if fp32 else _q
  t_fwd, t_sbwd, t_bbwd = (1e-4, 1e-2, 1e-3) if fp32 else (1e-2, 2e-2, 3e-2)
  n, S = 2, 128
  img = torch.rand(n, S, S, 3).requires_grad_(True)
  x224 = torch.nn.functional.interpolate(img.permute(0, 3, 1, 2), size=(224, 224), mode="bilinear",
                                         align_corners=False).permute(0, 2, 3, 1)
  stem_o, pool_o = orc.resnet_stem(variables, x224, pol)
  stem, pooled = model.stem_forward(img.detach().cuda())
  print(f"\n[resnet pieces {dtype}] stem {helpers.rel(stem, stem_o):.2e} pooled {helpers.rel(pooled, pool_o):.2e}")
  assert helpers.rel(stem, stem_o) < t_fwd and helpers.rel(pooled, pool_o) < t_fwd
  dpool = q(torch.randn_like(pool_o) * 0.1)
  (pool_o * dpool).sum().backward()
  d_img = torch.zeros(n, S, S, 3, device="cuda")
  model.stem_backward(dpool.cuda().to(adt), stem, pooled, S, d_img)
  print(f"  stem input gradient {helpers.rel(d_img, img.grad):.2e}")
  assert helpers.rel(d_img, img.grad) < t_sbwd
  for idx in (1, 3, 7):  # stage1/block2 (identity), stage2/block1 (stride 2 + projection), stage3/block1
    spec = model.blocks[idx]
    pre, cin, f, stride, proj = spec
    Hin = {0: 56, 1: 56, 2: 28, 3: 14}[int(pre[0][-1]) - 1 if stride == 1 else int(pre[0][-1]) - 2]
    x = q(torch.relu(torch.randn(n, Hin, Hin, cin))).requires_grad_(True)
    p, s_ = variables["params"][pre[0]][pre[1]], variables["batch_stats"][pre[0]][pre[1]]
    out_o = orc.bottleneck_block(x, p, s_, stride, pol)
    out, sv = model.block_forward(x.detach().cuda().to(adt), spec)
    dout = q(torch.randn_like(out_o) * 0.1) * (out_o.detach() > 0)
    (out_o * dout).sum().backward()
    dx = model.block_backward(dout.cuda().to(adt), sv["x"], sv["r1"], sv["r2"], spec, mask_input=False)
    print(f"  {pre}: forward {helpers.rel(out, out_o):.2e} input gradient {helpers.rel(dx, x.grad):.2e}")
    assert helpers.rel(out, out_o) < t_fwd, pre
    assert helpers.rel(dx, x.grad) < t_bbwd, (pre, helpers.rel(dx, x.grad))


def _pretrained_setup(seed=8):
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  cfg = helpers.small_config(pretrained_image_contrastive=True)
  B = 3
  variables = orc.resnet50_random_variables(2)
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=seed)
  batch = helpers.make_batch(2 * B, cfg, seed=seed + 1)
  ostate = orc.make_state(g_vars, d_vars)
  state = train_utils.TrainState(0, train_utils.Optimizer(g_params, cfg.g_lr, cfg.beta1, cfg.beta2),
                                 train_utils.Optimizer(d_params, cfg.d_lr, cfg.beta1, cfg.beta2),
                                 {"batch_stats": g_stats}, {"spectral_norm_stats": d_u}, g_params.clone())
  additional = xmc_gan.create_additional_data(cfg, variables=variables)
  pol = orc.Policy("bfloat16", round_grads=True)
  # the frozen branch runs in fp32 in the reference (and by default here) whatever config.dtype says
  pre = lambda real, fake: orc.calculate_contrastive_loss_on_pretrained(variables, real, fake, orc.FP32)
  return cfg, batch, state, ostate, additional, pol, pre


@gpu
def test_train_step_with_pretrained_image_contrastive():
  """The reference's default configuration (pretrained_image_contrastive=True) through train_step vs the oracle, from
  a mid-training optimiser state (helpers.warm_adam): all five metrics within 5e-3 of the largest, generator
  parameter updates per leaf rel-L2 5e-2 (2e-1 below the 16x16 stage)."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  cfg, batch, state, ostate, additional, pol, pre = _pretrained_setup()
  helpers.warm_adam(state, ostate)
  g_old = state.g_optimizer.target.to_cpu_tree()
  o_old = ostate["g_params"]
  state, metrics = train_utils.train_step(None, state, batch, xmc_gan, None, None, cfg, additional)
  got = metrics.compute()
  ostate, want = orc.train_step(ostate, batch, cfg, pol, pretrained_fn=pre)
  scale = max(abs(v) for v in want.values())
  print("\n[pretrained, warm Adam]", {k: (round(got[k], 5), round(want[k], 5)) for k in want})
  for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g", "c_loss_g_pretrained"):
    assert abs(got[k] - want[k]) < 5e-3 * scale, (k, got[k], want[k])
  assert got["c_loss_g_pretrained"] > 0
  helpers.check_updates("g", state.g_optimizer.target.to_cpu_tree(), g_old, ostate["g_params"], o_old, cfg.g_lr)


@gpu
def test_cold_start_train_step_with_pretrained_branch_is_sharp_once_the_sign_step_is_shared():
  """Cold start (Adam t = 1) of the same configuration. train_g_d's metrics are taken AFTER train_d's first Adam
  update, which is -lr * sign(g): hinge_g jumps from -8.4 to +7.4 here, and an element whose gradient sits at
  rounding-noise level moves by +-lr on a coin flip of the implementation's rounding (the fp32 and the bf16 policy of
  the ORACLE differ by 0.17 in g_loss for that reason; tools/bias_bisect.py, profiles/r02_bias_bisect.md). So the two
  halves are checked separately and each sharply: (1) train_d's update: the updated discriminator parameters agree in
  SIGN of the step on all but a small share of the elements; (2) train_g_d from the SAME updated discriminator (the
  oracle continues from the CUDA path's parameters): all five metrics within 5e-3."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  cfg, batch, state, ostate, additional, pol, pre = _pretrained_setup()
  d_old = state.d_optimizer.target.to_cpu_tree()
  b0, b1 = train_utils.split_input_dict(xmc_net.batch_to_device(batch), 2)
  ob0, ob1 = orc.split_input_dict(batch, 2)
  state = xmc_gan.train_d(None, state, b0, None, None, cfg)
  ostate1, _ = orc.train_d(ostate, ob0, cfg, pol)
  d_new = state.d_optimizer.target.to_cpu_tree()
  same = total = 0
  for (p, a), (_, a0), (_, b), (_, b0_) in zip(orc.tree_leaves(d_new), orc.tree_leaves(d_old),
                                               orc.tree_leaves(ostate1["d_params"]), orc.tree_leaves(ostate["d_params"])):
    same += int((torch.sign(a - a0) == torch.sign(b - b0_)).sum())
    total += a.numel()
  print(f"\n[cold start] train_d sign-step agreement {same / total:.4%}")
  assert same / total > 0.97
  # the oracle continues from the CUDA path's discriminator (parameters, moments, u0): same starting point for train_g_d
  ostate1["d_params"] = d_new
  ostate1["d_opt"] = {"step": 1,
                      "m": xmc_net.FlatTree(state.d_optimizer.target.layout, state.d_optimizer.m).to_cpu_tree(),
                      "v": xmc_net.FlatTree(state.d_optimizer.target.layout, state.d_optimizer.v).to_cpu_tree()}
  ostate1["discriminator_state"] = {"spectral_norm_stats": state.discriminator_state["spectral_norm_stats"].to_cpu_tree()}
  state, metrics = xmc_gan.train_g_d(None, state, b1, None, None, cfg, additional)
  got = metrics.compute()
  _, want, _ = orc.train_g_d(ostate1, ob1, cfg, pol, pre)
  scale = max(abs(v) for v in want.values())
  print("  ", {k: (round(got[k], 5), round(want[k], 5)) for k in want})
  for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g", "c_loss_g_pretrained"):
    assert abs(got[k] - want[k]) < 5e-3 * scale, (k, got[k], want[k])


@gpu
@pytest.mark.parametrize("N,H,C,Cout", [(2, 8, 64, 64), (3, 16, 96, 192), (1, 64, 192, 96), (5, 4, 128, 64)])
def test_subpixel_conv_equals_upsample_then_conv(N, H, C, Cout):
  """conv3x3(upsample2x(x)) computed as four 2x2 convs with pre-summed weights (forward), its 4x4/stride-2 input
  gradient and its weight gradient, against the oracle's upsample -> conv2d and autograd. Forward 5e-3 (the summed
  weights are rounded to bf16 once instead of per tap), gradients 1e-2."""
  _, _, ops, *_ = _mods()
  from xmcgan_image_generation_b200 import _lib
  torch.manual_seed(H + C)
  x = _q(torch.randn(N, H, H, C)).requires_grad_(True)
  kern = (torch.randn(3, 3, C, Cout) * 0.05).requires_grad_(True)
  bias = torch.randn(Cout)
  want = orc.conv2d(orc.upsample(x), kern, bias)
  dy = _q(torch.randn(N, 2 * H, 2 * H, Cout) * 0.1)
  (want * dy).sum().backward()
  wf = ops.empty((4 * Cout, 4 * C))
  vd = ops.empty((C, 16 * Cout))
  kd = kern.detach().cuda().contiguous()
  ops._call("xmc_subpixel_prep", kd.data_ptr(), None, C, Cout, 0, wf.data_ptr(), vd.data_ptr(), _lib.stream())
  xd = x.detach().cuda().to(torch.bfloat16)
  got = ops.conv_fwd(xd, wf, 2, Cout, bias=bias.cuda(), ldb=4 * C, pad=1, subpixel=True, out_dtype=torch.float32)
  assert got.shape == (N, 2 * H, 2 * H, Cout)
  assert helpers.rel(got, want) < 5e-3
  dyd = dy.cuda().to(torch.bfloat16)
  dx = ops.conv_fwd(dyd, vd, 4, C, ldb=16 * Cout, stride=2, pad=1, out_dtype=torch.float32)
  assert helpers.rel(dx, x.grad) < 1e-2
  dw = torch.zeros(9 * C * Cout, device="cuda")
  ops.wgrad(xd, dyd, 3, dw, out_mode=0, ld_out=Cout, tap_stride=C * Cout, subpixel=True)
  assert helpers.rel(dw.view(3, 3, C, Cout), kern.grad) < 1e-2


@gpu
def test_train_step_with_generator_spectral_norm():
  """config.g_spectral_norm=True (xmc_net.py:176-191): every generator conv / dense is spectrally normalised, its u0
  collection advances in train_g_d only (train_d discards the generator's new state, xmc_gan.py:225). One train_step vs
  the oracle: metrics 5e-3, parameters 5e-3, generator u0 1e-3."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  cfg = helpers.small_config(g_spectral_norm=True)
  B = 3
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=11)
  g_u = _g_u0(cfg, g_vars)
  batch = helpers.make_batch(2 * B, cfg, seed=12)
  ostate = orc.make_state(g_vars, d_vars)
  state = train_utils.TrainState(0, train_utils.Optimizer(g_params, cfg.g_lr, cfg.beta1, cfg.beta2),
                                 train_utils.Optimizer(d_params, cfg.d_lr, cfg.beta1, cfg.beta2),
                                 {"batch_stats": g_stats, "spectral_norm_stats": g_u}, {"spectral_norm_stats": d_u},
                                 g_params.clone())
  state, metrics = train_utils.train_step(None, state, batch, xmc_gan, None, None, cfg, {})
  got = metrics.compute()
  ostate, want = orc.train_step(ostate, batch, cfg, orc.Policy("bfloat16"))
  scale = max(abs(v) for v in want.values())
  for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g"):
    assert abs(got[k] - want[k]) < 5e-3 * scale, (k, got[k], want[k])
  pairs = [(state.g_optimizer.target, ostate["g_params"], 5e-3),
           (state.generator_state["spectral_norm_stats"], ostate["generator_state"]["spectral_norm_stats"], 1e-3),
           (state.generator_state["batch_stats"], ostate["generator_state"]["batch_stats"], 1e-2)]
  for got_t, want_t, tol in pairs:
    for (p, a), (_, b) in zip(orc.tree_leaves(got_t.to_cpu_tree()), orc.tree_leaves(want_t)):
      assert helpers.rel(a, b) < tol, (p, helpers.rel(a, b))
  # the module API carries the collection as well
  import functools
  gen = functools.partial(xmc_net.Generator, config=cfg)
  v = gen(train=False).init(0, (batch, batch["z"]))
  assert set(v) == {"params", "batch_stats", "spectral_norm_stats"}
  img, new = gen(train=True).apply(v, (batch, batch["z"]), mutable=["batch_stats", "spectral_norm_stats"])
  assert img.shape[0] == 2 * B and not torch.equal(new["spectral_norm_stats"].buf, v["spectral_norm_stats"].buf)


@gpu
@pytest.mark.parametrize("B", [3, 8])
def test_accuracy_and_entropy_side_statistics(B):
  """get_statistics (attention_lib.py:36-43) through contrastive_loss / word_loss: accuracy is an index op (argmax ==
  label) and must be exact; entropy 1e-4 (InfoNCE, fp32 logits) / 2e-2 (word_loss, bf16 region operands)."""
  from xmcgan_image_generation_b200.libml import attention_lib
  torch.manual_seed(20 + B)
  a, b = torch.randn(B, 96), torch.randn(B, 96)
  a[: B // 2] = b[: B // 2] + 0.1 * torch.randn(B // 2, 96)   # some matched pairs, some not
  _, wacc, went = orc.contrastive_loss(a, b)
  _, acc, ent = attention_lib.contrastive_loss(a, b)
  assert acc.item() == wacc.item()
  assert abs(ent.item() - went.item()) < 1e-4 * max(1.0, abs(went.item()))
  R, L, D = 256, 17, 64
  img = _q(torch.randn(B, R, D))
  words = torch.randn(B, L, D) * 0.5
  max_len = torch.randint(1, L + 1, (B, 1)).float()
  _, wacc, went = orc.word_loss(img, words, max_len)
  _, acc, ent = attention_lib.word_loss(img, words, max_len)
  assert abs(acc.item() - wacc.item()) <= 0.5 / B + 1e-6      # at most one near-tie flipped by bf16 operands
  assert abs(ent.item() - went.item()) < 2e-2 * max(1.0, abs(went.item()))


@gpu
def test_generate_batch_and_checkpoint_round_trip(tmp_path):
  """train_utils.generate_batch (train_utils.py:245-309): inference-mode samples with the current and the EMA
  parameters vs the oracle's generator_apply(train=False) on the same z (2e-2 rel-L2, bf16 activations), grid layout
  exact; then the state survives a ckpt-N.flax round trip bit for bit and keeps training."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  import functools
  from xmcgan_image_generation_b200 import checkpoint as ck
  cfg = helpers.small_config(show_num=4)
  B = 3
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=13)
  batch = helpers.make_batch(2 * B, cfg, seed=14)
  state = train_utils.TrainState(0, train_utils.Optimizer(g_params, cfg.g_lr, cfg.beta1, cfg.beta2),
                                 train_utils.Optimizer(d_params, cfg.d_lr, cfg.beta1, cfg.beta2),
                                 {"batch_stats": g_stats}, {"spectral_norm_stats": d_u}, g_params.clone())
  state, _ = train_utils.train_step(None, state, batch, xmc_gan, None, None, cfg, {})   # EMA != params afterwards
  gen = functools.partial(xmc_net.Generator, config=cfg)
  out = train_utils.generate_batch(0, state, batch, gen, cfg, z=batch["z"])
  assert set(out) == {"generated_image_batch", "ema_generated_image_batch", "ori_image_batch"}
  assert out["generated_image_batch"].shape == (1, 2 * 128, 2 * 128, 3)
  pol = orc.Policy("bfloat16")
  stats = state.generator_state["batch_stats"].to_cpu_tree()
  for key, params in (("generated_image_batch", state.g_optimizer.target), ("ema_generated_image_batch", state.ema_params)):
    want, _ = orc.generator_apply({"params": params.to_cpu_tree(), "batch_stats": stats}, (batch, batch["z"]), cfg,
                                  False, pol)
    assert helpers.rel(out[key][0], train_utils.make_grid(want.detach(), 4)) < 2e-2, key
  assert torch.equal(out["ori_image_batch"][0].cpu(), train_utils.make_grid(batch["image"], 4))
  assert not torch.equal(out["generated_image_batch"], out["ema_generated_image_batch"])
  # checkpoint round trip through the reference's file format, then one more step on both copies
  path = ck.save_checkpoint(str(tmp_path), state)
  g2, d2, g_params2, g_stats2, d_params2, d_u2 = _build(cfg, seed=99)
  other = train_utils.TrainState(0, train_utils.Optimizer(g_params2, cfg.g_lr, cfg.beta1, cfg.beta2),
                                 train_utils.Optimizer(d_params2, cfg.d_lr, cfg.beta1, cfg.beta2),
                                 {"batch_stats": g_stats2}, {"spectral_norm_stats": d_u2}, g_params2.clone())
  ck.restore_checkpoint(other, path)
  assert (other.step, other.d_optimizer.step, other.g_optimizer.step) == (1, 2, 1)
  assert torch.equal(other.d_optimizer.target.buf, state.d_optimizer.target.buf)
  assert torch.equal(other.ema_params.buf, state.ema_params.buf)
  state, m1 = train_utils.train_step(None, state, batch, xmc_gan, None, None, cfg, {})
  other, m2 = train_utils.train_step(None, other, batch, xmc_gan, None, None, cfg, {})
  assert m1.compute() == m2.compute()   # deterministic reductions: the restored copy continues bit for bit


@gpu
@pytest.mark.parametrize("N,S,C", [(2, 16, 16), (3, 32, 96), (1, 128, 96)])
def test_packed_window_image_convs_match_oracle(N, S, C):
  """The 3x3 convolutions with 3 image channels on the input side, run on the tcgen05 GEMM kernels over a
  zero-bordered 8-channel copy of the image (K = 3 kh-taps x [3 kw x 8 channels]): forward (+bias, relu) 1e-3 rel-L2
  (bf16 output), weight gradient 1e-4 (fp32 out) and the flipped / transposed indexing used for the generator's output
  conv, all vs the oracle's conv2d / autograd on identical bf16 operands. The padded copy is an index op: exact."""
  _, _, ops, *_ = _mods()
  torch.manual_seed(S + C)
  x = _q(torch.rand(N, S, S, 3))
  kern = _q(torch.randn(3, 3, 3, C) * 0.2).requires_grad_(True)
  bias = torch.randn(C)
  want = torch.relu(orc.conv2d(x, kern, bias))
  dy = _q(torch.randn(N, S, S, C) * 0.1)
  orc.conv2d(x, kern, None).mul(dy).sum().backward()
  xd = x.cuda().to(torch.bfloat16)
  xpad = ops.c3_pad(xd)
  ref_pad = torch.zeros(N, S + 2, S + 2, 8)
  ref_pad[:, 1:-1, 1:-1, :3] = x
  assert torch.equal(xpad.float().cpu(), ref_pad)
  wk = kern.detach().permute(3, 0, 1, 2).reshape(C, 27)
  wk32 = torch.zeros(C, 32)
  wk32[:, :27] = wk
  wp = ops.c3_pack_weights(wk32.cuda().to(torch.bfloat16), 32, C)
  got = ops.c3_conv(xpad, wp, C, bias=bias.cuda(), relu=True)
  assert got.shape == (N, S, S, C)
  assert helpers.rel(got, want) < 4e-3
  dw = torch.zeros(27 * C, device="cuda")
  ops.c3_wgrad(xpad, dy.cuda().to(torch.bfloat16), 0, 3 * C, C, 1, dw)
  assert helpers.rel(dw.view(3, 3, 3, C), kern.grad) < 1e-4
  # generator output conv (C -> 3): dW[tap][ci][c3] = sum_p h[p + d(tap)][ci] * dpre[p][c3] in the flipped indexing
  h = _q(torch.randn(N, S, S, C) * 0.5)
  k2 = _q(torch.randn(3, 3, C, 3) * 0.1).requires_grad_(True)
  orc.conv2d(h, k2, None).mul(x).sum().backward()      # x plays d(pre-tanh)
  dw2 = torch.zeros(27 * C, device="cuda")
  ops.c3_wgrad(xpad, h.cuda().to(torch.bfloat16), 1, C * 3, 1, 3, dw2)
  assert helpers.rel(dw2.view(3, 3, C, 3), k2.grad) < 1e-4


@gpu
@pytest.mark.parametrize("S", [128, 256])
def test_bilinear_resize_to_224_matches_jax_semantics(S):
  """jax.image.resize(..., "bilinear") of get_pretrained_embs (pretrained_model_utils.py:118-121): 128 -> 224 is plain
  half-pixel bilinear; 256 -> 224 down-samples with the triangle kernel widened by 256/224 and renormalised (== torch
  antialias=True, which the oracle uses). Forward 4e-3 (bf16 output), transpose (backward) 1e-5 vs autograd."""
  _, _, ops, *_ = _mods()
  from xmcgan_image_generation_b200 import _lib
  torch.manual_seed(S)
  n, T, TP, PAD = 2, 224, 229, 2
  img = torch.rand(n, S, S, 3, requires_grad=True)
  want = torch.nn.functional.interpolate(img.permute(0, 3, 1, 2), size=(T, T), mode="bilinear", align_corners=False,
                                         antialias=S > T).permute(0, 2, 3, 1)
  out = ops.empty((n, TP, TP, 8))
  ops._call("xmc_resize_bilinear_pad", img.detach().cuda().data_ptr(), n, S, T, TP, PAD, 0, out.data_ptr(), _lib.stream())
  got = out.float().cpu()
  assert helpers.rel(got[:, PAD:PAD + T, PAD:PAD + T, :3], want) < 4e-3
  assert got[:, :PAD].abs().max() == 0 and got[..., 3:].abs().max() == 0      # zero border, zero pad channels
  d = torch.randn(n, T, T, 3)
  (want * d).sum().backward()
  dimg = torch.zeros(n, S, S, 3, device="cuda")
  ops._call("xmc_resize_bilinear_bwd", d.cuda().data_ptr(), n, S, T, dimg.data_ptr(), _lib.stream())
  assert helpers.rel(dimg, img.grad) < 1e-5


@gpu
def test_input_contract_producer_matches_oracle():
  """libml.coco_dataset.preprocess (COCODataset.preprocess, coco_dataset.py:127-167) vs the oracle restatement on the
  same random draws: flipped / clipped image, chosen caption embedding and max_len are index / clamp ops -> exact;
  sentence_embedding = sum over 17 word slots / len in the same summation order -> exact as well. With
  return_text=True the shortest caption is chosen (an argsort index op), and the result drives a train_step."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  from xmcgan_image_generation_b200.libml import coco_dataset
  torch.manual_seed(31)
  N, S, M, L, E = 6, 128, 5, 17, 64
  feats = {"image": torch.rand(N, S, S, 3) * 1.2 - 0.1,                 # some values outside [0,1]: clip matters
           "caption/embedding": torch.randn(N, M, L, E) * 0.5,
           "caption/max_len": torch.randint(3, L + 1, (N, M))}
  flip = torch.tensor([1, 0, 1, 1, 0, 0], dtype=torch.bool)
  idx = torch.tensor([0, 4, 2, 1, 3, 3])
  z = torch.randn(N, 8)
  want = orc.preprocess_batch(feats, flip, idx, z)
  got = coco_dataset.preprocess(feats, flip=flip, sentence_idx=idx, z=z)
  for k in ("image", "embedding", "max_len", "sentence_embedding", "z"):
    assert got[k].shape == want[k].shape, k
    assert torch.equal(got[k].cpu(), want[k]), k
  short = coco_dataset.preprocess(feats, rng=3, z_dim=8, return_text=True)
  want_idx = feats["caption/max_len"].argmin(dim=1)
  assert torch.equal(short["max_len"].cpu()[:, 0], feats["caption/max_len"].float().gather(1, want_idx[:, None])[:, 0])
  rnd = coco_dataset.preprocess(feats, rng=5, z_dim=8)
  assert rnd["z"].shape == (N, 8) and rnd["image"].min() >= 0 and rnd["image"].max() <= 1
  cfg = helpers.small_config(batch_size=N // 2)
  gen, disc, state = train_utils.create_train_state(cfg, 1, rnd)
  state, m = train_utils.train_step(None, state, rnd, xmc_gan, gen, disc, cfg, {})
  assert all(torch.isfinite(torch.tensor(v)) for v in m.compute().values())


@gpu
def test_graphed_train_step_equals_eager_train_step():
  """train_utils.GraphedTrainStep (the whole train_step replayed from one CUDA graph) against the eager train_step on
  an identically initialised state and the same 5 batches. Every reduction of the CUDA path has a fixed summation
  order (deterministic split-K, two-stage statistics), so the two are compared BIT FOR BIT: metrics of every step,
  parameters, Adam moments, EMA, batch statistics and u0 after 5 steps, host and device step counters."""
  _, engine, ops, train_utils, xmc_gan, xmc_net = _mods()
  cfg = helpers.small_config()
  B = 3

  def fresh():
    g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=17)
    return train_utils.TrainState(0, train_utils.Optimizer(g_params, cfg.g_lr, cfg.beta1, cfg.beta2),
                                  train_utils.Optimizer(d_params, cfg.d_lr, cfg.beta1, cfg.beta2),
                                  {"batch_stats": g_stats}, {"spectral_norm_stats": d_u}, g_params.clone())

  batches = [helpers.make_batch(2 * B, cfg, seed=40 + i) for i in range(5)]

  def run(step_of):
    st = fresh()
    fn = step_of(st)
    metrics = []
    for b in batches:
      st, m = fn(st, b)
      metrics.append(m.compute())
    return st, metrics

  eager, em = run(lambda st: (lambda s, b: train_utils.train_step(None, s, b, xmc_gan, None, None, cfg, {})))

  def graphed(st):
    step = train_utils.GraphedTrainStep(st, batches[0], xmc_gan, None, None, cfg, {}, warmup=0)
    return lambda s, b: step(b)

  st, gm = run(graphed)
  assert (st.step, st.d_optimizer.step, st.g_optimizer.step) == (5, 10, 5)
  assert int(st.d_optimizer.step_dev.item()) == 10 and int(st.g_optimizer.step_dev.item()) == 5
  for i, (a, b) in enumerate(zip(gm, em)):
    assert a == b, (i, a, b)
  for name, x, y in (("g", st.g_optimizer.target.buf, eager.g_optimizer.target.buf),
                     ("d", st.d_optimizer.target.buf, eager.d_optimizer.target.buf),
                     ("g.m", st.g_optimizer.m, eager.g_optimizer.m), ("d.v", st.d_optimizer.v, eager.d_optimizer.v),
                     ("ema", st.ema_params.buf, eager.ema_params.buf)):
    assert torch.equal(x, y), name
  # leaf by leaf (the alignment padding between leaves of a flat buffer is not state)
  for got_t, want_t in ((st.generator_state["batch_stats"], eager.generator_state["batch_stats"]),
                        (st.discriminator_state["spectral_norm_stats"], eager.discriminator_state["spectral_norm_stats"])):
    for (path, a), (_, b) in zip(orc.tree_leaves(got_t.to_cpu_tree()), orc.tree_leaves(want_t.to_cpu_tree())):
      assert torch.equal(a, b), path
