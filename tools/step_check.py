"""GPU debugging aid: runs the B200 engines against the CPU oracle on a small configuration and prints per-stage and
per-parameter errors. Usage: python tools/step_check.py [stage ...]   stages: g d grads step"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import xmc_oracle as orc
from xmcgan_image_generation_b200 import engine, ops, train_utils, xmc_gan
from xmcgan_image_generation_b200.configs import coco_xmc
from xmcgan_image_generation_b200.nets import xmc_net


def small_config(**kw):
  c = coco_xmc.get_config()
  c.update(dict(gf_dim=16, df_dim=16, z_dim=8, batch_size=8, pretrained_image_contrastive=False))
  c.update(kw)
  return c


def make_batch(n, config, E=64, L=17, seed=0):
  g = torch.Generator().manual_seed(seed)
  S = config.image_size
  emb = torch.randn(n, L, E, generator=g) * 0.5
  max_len = torch.randint(3, L + 1, (n, 1), generator=g).float()
  return {
      "image": torch.rand(n, S, S, 3, generator=g),
      "embedding": emb,
      "max_len": max_len,
      "sentence_embedding": emb.sum(1) / max_len,
      "z": torch.randn(n, config.z_dim, generator=g),
  }


def randomize_biases(flat_tree, seed, scale=0.1):
  g = torch.Generator().manual_seed(seed)
  for path, (off, shape) in flat_tree.layout.entries.items():
    if path[-1] == "bias":
      n = 1
      for s in shape:
        n *= s
      flat_tree.buf[off:off + n] = (torch.randn(n, generator=g) * scale).cuda()


def rel(a, b):
  a = a.detach().float().cpu()
  b = b.detach().float().cpu()
  return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cmp(name, got, ref, tol):
  r = rel(got, ref)
  mx = (got.detach().float().cpu() - ref.detach().float().cpu()).abs().max().item()
  nan = int(torch.isnan(got.detach().float()).sum().item())
  print(f"{'PASS' if (r < tol and nan == 0) else 'FAIL'} {name}: rel_l2={r:.3e} max_abs={mx:.3e} nan={nan} "
        f"ref_norm={ref.detach().float().norm().item():.3e}", flush=True)


def tree_cmp(name, got_tree, ref_tree, tol, top=int(os.environ.get("TOP", "12"))):
  rows = []
  for (path, g), (_, r) in zip(orc.tree_leaves(got_tree), orc.tree_leaves(ref_tree)):
    rows.append((rel(g, r), path, r.float().norm().item(), g.detach().float().cpu().norm().item()))
  rows.sort(reverse=True)
  bad = [x for x in rows if not (x[0] < tol)]
  print(f"{'PASS' if not bad else 'FAIL'} {name}: {len(rows)} leaves, {len(bad)} above tol {tol}; worst:", flush=True)
  for r_, path, rn, gn in rows[:top]:
    print(f"     {r_:.3e}  {path}  ref_norm={rn:.3e} got_norm={gn:.3e}")


def build(config, E=64, seed=1):
  g_eng = xmc_net.get_engine(config, "g", E)
  d_eng = xmc_net.get_engine(config, "d", E)
  gp, gs = g_eng.init_params(seed)
  dp, du = d_eng.init_params(seed + 10)
  g_params, g_stats = xmc_net.FlatTree(g_eng.layout, gp), xmc_net.FlatTree(g_eng.stats_layout, gs)
  d_params, d_u = xmc_net.FlatTree(d_eng.layout, dp), xmc_net.FlatTree(d_eng.u_layout, du)
  randomize_biases(g_params, 5)
  randomize_biases(d_params, 6)
  # non-trivial running stats / larger u0 so state updates are visible
  g_stats.buf.add_(torch.rand_like(g_stats.buf) * 0.1)
  return g_eng, d_eng, g_params, g_stats, d_params, d_u


def oracle_state(g_params, g_stats, d_params, d_u):
  return orc.make_state({"params": g_params.to_cpu_tree(), "batch_stats": g_stats.to_cpu_tree()},
                        {"params": d_params.to_cpu_tree(), "spectral_norm_stats": d_u.to_cpu_tree()})


def stage_g(config, B):
  g_eng, d_eng, g_params, g_stats, d_params, d_u = build(config)
  batch = make_batch(B, config)
  dev = xmc_net.batch_to_device(batch)
  g_eng.prep_weights(g_params.buf)
  new_stats = torch.empty_like(g_stats.buf)
  img, ctx = g_eng.forward(g_params.buf, g_stats.buf, dev, dev["z"], train=True, new_stats=new_stats)
  torch.cuda.synchronize()
  for pol in ("bfloat16", "float32"):
    ref, upd = orc.generator_apply({"params": g_params.to_cpu_tree(), "batch_stats": g_stats.to_cpu_tree()},
                                   (batch, batch["z"]), config, True, orc.Policy(pol))
    cmp(f"G image vs oracle[{pol}]", img, ref, 2e-2 if pol == "bfloat16" else 5e-2)
  tree_cmp("G new batch_stats", xmc_net.FlatTree(g_eng.stats_layout, new_stats).to_cpu_tree(), upd["batch_stats"], 2e-2)


def stage_d(config, B):
  g_eng, d_eng, g_params, g_stats, d_params, d_u = build(config)
  batch = make_batch(B, config)
  dev = xmc_net.batch_to_device(batch)
  S = config.image_size
  gen = torch.Generator().manual_seed(3)
  fake = torch.rand(B, S, S, 3, generator=gen)
  all_images = torch.cat([batch["image"], fake])
  images = ops.cast_to_bf16(all_images.cuda().reshape(-1, 3)).view(2 * B, S, S, 3)
  u_new = torch.empty_like(d_u.buf)
  d_eng.prep_weights(d_params.buf, d_u.buf, u_new)
  losses = torch.zeros(16, device="cuda")
  logit, ctx = d_eng.forward(d_params.buf, images, dev, losses, need_g=True)
  torch.cuda.synchronize()
  for pol in ("bfloat16", "float32"):
    (rlogit, rstat), upd = orc.discriminator_apply(
        {"params": d_params.to_cpu_tree(), "spectral_norm_stats": d_u.to_cpu_tree()}, (all_images, batch), config, True,
        orc.Policy(pol))
    cmp(f"D logit vs oracle[{pol}]", logit, rlogit.reshape(-1), 3e-2)
    Sl = engine.LOSS_SLOTS
    for k, slot in (("real_word_loss", "real_word"), ("fake_word_loss", "fake_word"),
                    ("real_sentence_loss", "real_sent"), ("fake_sentence_loss", "fake_sent"),
                    ("image_contrastive_loss", "image")):
      cmp(f"D {k} vs oracle[{pol}]", losses[Sl[slot]], rstat[k], 2e-2)
  tree_cmp("D new u0", xmc_net.FlatTree(d_eng.u_layout, u_new).to_cpu_tree(), upd["spectral_norm_stats"], 1e-3)


def stage_grads(config, B):
  g_eng, d_eng, g_params, g_stats, d_params, d_u = build(config)
  batch = make_batch(B, config)
  dev = xmc_net.batch_to_device(batch)
  S = config.image_size
  ostate = oracle_state(g_params, g_stats, d_params, d_u)
  g_eng.prep_weights(g_params.buf)
  u_new = torch.empty_like(d_u.buf)
  d_eng.prep_weights(d_params.buf, d_u.buf, u_new)
  all_images = ops.empty((2 * B, S, S, 3))
  ops.cast_to_bf16(dev["image"].reshape(-1, 3), all_images[:B].view(-1, 3))
  img, gctx = g_eng.forward(g_params.buf, g_stats.buf, dev, dev["z"], train=True, new_stats=None,
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
  for pol in ("bfloat16", "float32"):
    r = orc.d_losses_and_grads(ostate, batch, config, orc.Policy(pol), want_g=True)
    tol = 5e-2 if pol == "bfloat16" else 1e-1
    cmp(f"fake image [{pol}]", img, r["fake"], 3e-2)
    cmp(f"d_loss [{pol}]", losses[0] + losses[2] + losses[4], r["d_loss"], 2e-2)
    cmp(f"g_loss [{pol}]", losses[1] + losses[3] + losses[5] + losses[6], r["g_loss"], 2e-2)
    tree_cmp(f"d_grad [{pol}]", xmc_net.FlatTree(d_eng.layout, d_grads).to_cpu_tree(), r["d_grad"], tol)
    tree_cmp(f"g_grad [{pol}]", xmc_net.FlatTree(g_eng.layout, g_grads).to_cpu_tree(), r["g_grad"], tol)


def stage_step(config, B):
  """Full train_step through the public API vs the oracle's train_step (bf16 policy)."""
  g_eng, d_eng, g_params, g_stats, d_params, d_u = build(config)
  batch = make_batch(2 * B, config, seed=7)
  ostate = oracle_state(g_params, g_stats, d_params, d_u)
  g_opt = train_utils.Optimizer(g_params, config.g_lr, config.beta1, config.beta2)
  d_opt = train_utils.Optimizer(d_params, config.d_lr, config.beta1, config.beta2)
  state = train_utils.TrainState(0, g_opt, d_opt, {"batch_stats": g_stats}, {"spectral_norm_stats": d_u},
                                 g_params.clone())
  t0 = time.time()
  for it in range(2):
    state, metrics = train_utils.train_step(None, state, batch, xmc_gan, None, None, config, {})
    m = metrics.compute()
    ostate, om = orc.train_step(ostate, batch, config, orc.Policy("bfloat16"))
    for k in m:
      print(f"  step {it} {k}: got {m[k]:.5f} oracle {om[k]:.5f}")
  print(f"2 steps wall {time.time()-t0:.1f}s (includes oracle)")
  tree_cmp("g_params after 2 steps", state.g_optimizer.target.to_cpu_tree(), ostate["g_params"], 1e-2)
  tree_cmp("d_params after 2 steps", state.d_optimizer.target.to_cpu_tree(), ostate["d_params"], 1e-2)
  tree_cmp("ema after 2 steps", state.ema_params.to_cpu_tree(), ostate["ema_params"], 1e-3)
  tree_cmp("batch_stats after 2 steps", state.generator_state["batch_stats"].to_cpu_tree(),
           ostate["generator_state"]["batch_stats"], 2e-2)
  tree_cmp("u0 after 2 steps", state.discriminator_state["spectral_norm_stats"].to_cpu_tree(),
           ostate["discriminator_state"]["spectral_norm_stats"], 2e-3)
  print("step counters", state.step, state.d_optimizer.step, state.g_optimizer.step, "oracle", ostate["step"],
        ostate["d_opt"]["step"], ostate["g_opt"]["step"])


STAGES = dict(g=stage_g, d=stage_d, grads=stage_grads, step=stage_step)

if __name__ == "__main__":
  torch.manual_seed(0)
  names = [a for a in sys.argv[1:] if "=" not in a] or list(STAGES)
  kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:] if "=" in a)}
  B = kw.pop("B", 4)
  cfg = small_config(**kw)
  for n in names:
    print(f"== {n} {kw} B={B}", flush=True)
    STAGES[n](cfg, B)
