"""Shared builders for the parity tests: small configs, synthetic COCO-shaped batches, oracle <-> product state."""
import torch

from xmcgan_image_generation_b200.configs import coco_xmc


def small_config(**kw):
  c = coco_xmc.get_config()
  c.update(dict(gf_dim=16, df_dim=16, z_dim=8, batch_size=8, pretrained_image_contrastive=False))
  c.update(kw)
  return c


def make_batch(n, config, E=64, L=17, seed=0, min_len=3):
  g = torch.Generator().manual_seed(seed)
  S = config.image_size
  emb = torch.randn(n, L, E, generator=g) * 0.5
  max_len = torch.randint(min_len, L + 1, (n, 1), generator=g).float()
  return {"image": torch.rand(n, S, S, 3, generator=g), "embedding": emb, "max_len": max_len,
          "sentence_embedding": emb.sum(1) / max_len, "z": torch.randn(n, config.z_dim, generator=g)}


def cpu_variables(config, E=64, seed=1, bias_scale=0.1):
  """Random G/D variable trees on the CPU (same layouts/names as the product, no GPU needed)."""
  from xmcgan_image_generation_b200 import engine
  g = engine.GeneratorEngine(config, E)
  d = engine.DiscriminatorEngine(config, E)
  bufs = [engine.init_flat(g.layout, seed, engine._kind), engine.init_flat(g.stats_layout, seed + 1, engine._kind),
          engine.init_flat(d.layout, seed + 2, engine._kind), engine.init_flat(d.u_layout, seed + 3, engine._kind)]
  gen = torch.Generator().manual_seed(seed + 4)
  for lay, buf in ((g.layout, bufs[0]), (d.layout, bufs[2])):
    for path, (off, shape) in lay.entries.items():
      if path[-1] == "bias":
        n = int(torch.tensor(shape).prod())
        buf[off:off + n] = torch.randn(n, generator=gen) * bias_scale
  g_vars = {"params": g.layout.tree(bufs[0]), "batch_stats": g.stats_layout.tree(bufs[1])}
  if g.sn:
    g_vars["spectral_norm_stats"] = g.u_layout.tree(engine.init_flat(g.u_layout, seed + 5, engine._kind))
  d_vars = {"params": d.layout.tree(bufs[2]), "spectral_norm_stats": d.u_layout.tree(bufs[3])}
  return g, d, g_vars, d_vars


def rel(a, b):
  a = a.detach().float().cpu()
  b = b.detach().float().cpu()
  return ((a - b).norm() / (b.norm() + 1e-12)).item()


def warm_adam(state, ostate, t=100, v0=1e-2):
  """Puts the product TrainState and the oracle state into the same mid-training optimiser state: step t, first
  moments 0, second moments v0 everywhere. Adam's very first update is -lr * g / (|g| + eps) = -lr * sign(g): any
  element whose gradient sits at rounding-noise level moves by +-lr on a coin flip, which makes quantities measured
  after it (train_g_d's metrics follow train_d's update) ill-conditioned. With v0 = (1e-1)^2 the update is ~1.5 lr g,
  linear in g wherever |g| << 3 and sign-like only where the sign is robust, so the comparison measures the gradient."""
  from oracle import xmc_oracle as orc
  for opt, key in ((state.g_optimizer, "g_opt"), (state.d_optimizer, "d_opt")):
    opt.step = t
    opt.m.zero_()
    opt.v.fill_(v0)
    o = ostate[key]
    o["step"] = t
    o["v"] = orc.tree_map(lambda x: torch.full_like(x, v0), o["v"])


def check_updates(name, new_tree, old_tree, want_new, want_old, lr, tol=5e-2, deep_tol=2e-1):
  """Per-leaf comparison of parameter UPDATES (new - old) with the oracle's: rel-L2 <= tol (deep_tol for the generator
  leaves below the 16x16 stage, see tests/test_gpu_parity.GRAD_TOL_DEEP). Leaves whose oracle update is below 1e-3 lr
  per element — a numerically zero gradient: biases in front of a BatchNorm — are skipped. Prints the worst leaf."""
  from oracle import xmc_oracle as orc
  from tests.test_gpu_parity import _is_deep_generator_leaf
  rows = []
  for (p, a), (_, a0), (_, b), (_, b0) in zip(orc.tree_leaves(new_tree), orc.tree_leaves(old_tree),
                                              orc.tree_leaves(want_new), orc.tree_leaves(want_old)):
    da, db = (a.float().cpu() - a0.float().cpu()).reshape(-1), (b - b0).reshape(-1)
    if db.norm().item() / db.numel() ** 0.5 < 1e-3 * lr:
      continue
    rows.append((rel(da, db), p))
  worst = max(rows)
  print(f"  {name} update: worst leaf rel-L2 {worst[0]:.3e} ({worst[1]}) over {len(rows)} leaves")
  bad = [(e, p) for e, p in rows if e > (deep_tol if _is_deep_generator_leaf(p) else tol)]
  assert not bad, bad
