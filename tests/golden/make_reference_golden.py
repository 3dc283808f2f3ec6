"""Generates tests/golden/reference_libml.npz by EXECUTING THE REFERENCE'S OWN SOURCE for the loss / attention layer of
the path (xmcgan/libml/losses.py, xmcgan/libml/attention_lib.py) and three host helpers (utils/image_utils.make_grid,
utils/device_utils.get_device_groups, train_utils.split_input_dict) on fixed-seed inputs.

JAX is not installable in this environment, so the reference modules are imported from /root/reference with a small
numpy-backed stand-in for the handful of `jax` entry points they use (jnp.* = numpy.*, jax.nn.softmax / log_softmax /
relu / one_hot, jax.lax.rsqrt, jax.scipy.special.logsumexp, jax.vmap over the leading axis, jax.tree_map over a dict,
jax.device_count). Those functions are pure array arithmetic in float32, so numpy reproduces XLA-CPU's results up to
summation order (~1e-7); everything else — the formulas, axes, masks, constants — is the reference's code, line for
line. This is what pins `oracle/xmc_oracle.py` for these functions (tests/test_reference_golden.py); the networks,
their backward and the optimiser need Flax and stay pinned by known answers only.

Run in the build container only (reads /root/reference):  python -m tests.golden.make_reference_golden"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_libml.npz")


def _jax_stand_in(device_count=8):
  jnp = types.ModuleType("jax.numpy")
  jnp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("_")})
  jnp.ndarray = np.ndarray

  def softmax(x, axis=-1):
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return e / np.sum(e, axis=axis, keepdims=True)

  def log_softmax(x, axis=-1):
    s = x - np.max(x, axis=axis, keepdims=True)
    return s - np.log(np.sum(np.exp(s), axis=axis, keepdims=True))

  nn = types.ModuleType("jax.nn")
  nn.softmax, nn.log_softmax = softmax, log_softmax
  nn.relu = lambda x: np.maximum(x, 0)
  nn.one_hot = lambda idx, n: (np.asarray(idx)[..., None] == np.arange(n)).astype(np.float32)
  lax = types.ModuleType("jax.lax")
  lax.rsqrt = lambda x: (1.0 / np.sqrt(x)).astype(np.asarray(x).dtype)
  special = types.ModuleType("jax.scipy.special")

  def logsumexp(a, axis=None, keepdims=False):
    m = np.max(a, axis=axis, keepdims=True)
    r = m + np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True))
    return r if keepdims else np.squeeze(r, axis=axis)

  special.logsumexp = logsumexp
  jscipy = types.ModuleType("jax.scipy")
  jscipy.special = special

  def vmap(fn):
    return lambda *args: np.stack([fn(*[a[i] for a in args]) for i in range(len(args[0]))])

  jax = types.ModuleType("jax")
  jax.numpy, jax.nn, jax.lax, jax.scipy, jax.vmap = jnp, nn, lax, jscipy, vmap
  jax.tree_map = lambda fn, d: {k: fn(v) for k, v in d.items()}
  jax.device_count = lambda: device_count
  return {"jax": jax, "jax.numpy": jnp, "jax.nn": nn, "jax.lax": lax, "jax.scipy": jscipy,
          "jax.scipy.special": special}


def _load(name, path):
  spec = importlib.util.spec_from_file_location(name, path)
  mod = importlib.util.module_from_spec(spec)
  sys.modules[name] = mod
  spec.loader.exec_module(mod)
  return mod


def load_reference(device_count=8):
  """(losses, attention_lib, image_utils, device_utils, split_input_dict) of the reference on the stand-in."""
  shim = _jax_stand_in(device_count)
  sys.modules.update(shim)
  absl = types.ModuleType("absl")
  absl.logging = types.SimpleNamespace(info=lambda *a, **k: None)
  sys.modules.setdefault("absl", absl)
  for pkg in ("xmcgan", "xmcgan.libml", "xmcgan.utils"):
    sys.modules.setdefault(pkg, types.ModuleType(pkg))
  losses = _load("xmcgan.libml.losses", f"{REF}/xmcgan/libml/losses.py")
  sys.modules["xmcgan.libml"].losses = losses
  attention_lib = _load("xmcgan.libml.attention_lib", f"{REF}/xmcgan/libml/attention_lib.py")
  image_utils = _load("xmcgan.utils.image_utils", f"{REF}/xmcgan/utils/image_utils.py")
  device_utils = _load("xmcgan.utils.device_utils", f"{REF}/xmcgan/utils/device_utils.py")
  # train_utils.py imports flax / clu / tensorflow at module level: take the one pure function out of its source
  tree = ast.parse(open(f"{REF}/xmcgan/train_utils.py").read())
  fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "split_input_dict")
  scope = {"jax": shim["jax"], "jnp": shim["jax.numpy"], "Dict": dict}
  exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_utils.py", "exec"), scope)
  return losses, attention_lib, image_utils, device_utils, scope["split_input_dict"]


def load_reference_fid():
  """calculate_fid / _calculate_frechet_distance / calculate_inception_score of xmcgan/utils/tf_inception_utils.py:
  the module imports TensorFlow at the top, the three functions are pure numpy / scipy — they are taken out of its
  source. scipy.linalg.sqrtm lost its `disp` argument in newer releases; the stand-in restores the old return value."""
  import warnings
  from scipy import linalg as sl
  tree = ast.parse(open(f"{REF}/xmcgan/utils/tf_inception_utils.py").read())
  names = ("_calculate_frechet_distance", "calculate_fid", "calculate_inception_score")
  body = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in names) or
          (isinstance(n, ast.ClassDef) and n.name.endswith("Error"))]
  linalg = types.SimpleNamespace(sqrtm=lambda a, disp=True: (sl.sqrtm(a), 0.0) if disp is False else sl.sqrtm(a))
  scope = {"np": np, "linalg": linalg, "warnings": warnings}
  exec(compile(ast.Module(body=body, type_ignores=[]), "tf_inception_utils.py", "exec"), scope)
  return scope["calculate_fid"], scope["calculate_inception_score"]


def compute():
  losses, A, image_utils, device_utils, split_input_dict = load_reference()
  rng = np.random.default_rng(20211017)
  f32 = np.float32
  B, R, L, D = 4, 16, 6, 8
  out = {}
  region = rng.normal(size=(B, R, D)).astype(f32)
  words = (rng.normal(size=(B, L, D)) * 0.7).astype(f32)
  max_len = np.array([[3.0], [6.0], [2.0], [5.0]], f32)
  out.update(region=region, words=words, max_len=max_len)
  # word_loss: (loss, accuracy, entropy), default gammas
  out["word_loss"] = np.array([float(v) for v in A.word_loss(region, words, max_len)], np.float64)
  # contrastive_loss on sentence / image features
  img, cond = rng.normal(size=(B, D)).astype(f32), rng.normal(size=(B, D)).astype(f32)
  out.update(img=img, cond=cond)
  out["contrastive_loss"] = np.array([float(v) for v in A.contrastive_loss(img, cond)], np.float64)
  out["contrastive_loss_t05_nonorm"] = np.array(
      [float(v) for v in A.contrastive_loss(img, cond, l2_norm=False, temperature=0.5)], np.float64)
  # attention (per-word region context) and attention_for_g (per-region word context) with padding masks
  mask = (np.arange(L, dtype=f32)[None, :] >= max_len).astype(f32)[:, None, :].repeat(R, 1)     # [B, R, L]
  out["mask"] = mask
  out["attention_ctx"] = A.attention(region, words, 5.0, mask).astype(f32)
  out["attention_ctx_nomask"] = A.attention(region, words, 5.0).astype(f32)
  ctx, attn = A.attention_for_g(region, words, 15.0, mask)
  out["attention_for_g_ctx"], out["attention_for_g_attn"] = ctx.astype(f32), attn.astype(f32)
  out["l2_normalize"] = A.l2_normalize(region, -1).astype(f32)
  out["cosine_similarity"] = A.cosine_similarity(words, out["attention_ctx"]).astype(f32)
  logits = (rng.normal(size=(B, B)) * 3).astype(f32)
  labels = np.eye(B, dtype=f32)
  out["logits"] = logits
  out["get_statistics"] = np.array([float(v) for v in A.get_statistics(logits, labels)], np.float64)
  out["tf_cross_entropy"] = losses.tf_cross_entropy_loss_with_logits(labels=labels, logits=logits).astype(f32)
  out["cross_entropy_int"] = losses.cross_entropy_loss_with_logits(labels=np.arange(B), logits=logits).astype(f32)
  real_logit, fake_logit = rng.normal(size=(B, 1)).astype(f32), rng.normal(size=(B, 1)).astype(f32)
  out.update(real_logit=real_logit, fake_logit=fake_logit)
  out["hinge_loss"] = np.array([float(v) for v in losses.hinge_loss(real_logit, fake_logit)], np.float64)
  out["hinge_loss_d"] = np.array([float(losses.hinge_loss_d(real_logit, fake_logit))], np.float64)
  out["hinge_loss_g"] = np.array([float(losses.hinge_loss_g(fake_logit))], np.float64)
  # host helpers
  samples = rng.random(size=(10, 4, 6, 3)).astype(f32)
  out["grid_samples"] = samples
  out["make_grid_9"] = image_utils.make_grid(samples, 9)
  out["make_grid_64"] = image_utils.make_grid(samples, 64)     # cut to the batch: 3 x 3 of 10
  out["device_groups_8_16_4"] = np.array(device_utils.get_device_groups(16, 4))     # 8 devices, groups of 4
  out["device_groups_8_8_4"] = np.array(device_utils.get_device_groups(8, 4))       # groups of 2
  # FID / Inception score statistics (tf_inception_utils.py:123-224)
  calculate_fid, calculate_inception_score = load_reference_fid()
  pool1 = rng.normal(size=(300, 12))
  pool2 = rng.normal(size=(260, 12)) * 1.5 + 0.3
  preds = rng.random(size=(205, 7)) ** 3 + 1e-3
  preds = preds / preds.sum(1, keepdims=True)
  out.update(fid_pool1=pool1, fid_pool2=pool2, is_preds=preds)
  out["fid"] = np.array([calculate_fid(pool1, pool2), calculate_fid(pool1, pool1)], np.float64)
  out["inception_score_10"] = np.array(calculate_inception_score(preds, 10), np.float64)
  out["inception_score_1"] = np.array(calculate_inception_score(preds, 1), np.float64)
  parts = split_input_dict({"a": np.arange(24, dtype=f32).reshape(8, 3), "b": np.arange(8)}, 2)
  out["split_a0"], out["split_a1"], out["split_b1"] = parts[0]["a"], parts[1]["a"], parts[1]["b"]
  return out


# ---------------------------------------------------------------------------------------------------------------------
# The networks: xmcgan/nets/xmc_net.py + nets/common.py + libml/layers.py executed on tests/golden/flax_stand_in.py
# ---------------------------------------------------------------------------------------------------------------------
OUT_NETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_nets.npz")
NET_CFG = dict(gf_dim=8, df_dim=8, z_dim=8, batch_size=4)   # the narrowest widths the product's layouts accept
NET_E, NET_L, NET_B = 8, 5, 2


def load_reference_nets():
  """xmc_net of the reference on the numpy stand-ins for jax and flax.linen."""
  from tests.golden import flax_stand_in
  losses, attention_lib, image_utils, device_utils, _ = load_reference()
  flax_stand_in.install(sys.modules["jax"], sys.modules)
  sys.modules.setdefault("xmcgan.nets", types.ModuleType("xmcgan.nets"))
  sys.modules["xmcgan.libml"].attention_lib = attention_lib
  sys.modules["xmcgan.utils"].device_utils = device_utils
  layers = _load("xmcgan.libml.layers", f"{REF}/xmcgan/libml/layers.py")
  sys.modules["xmcgan.libml"].layers = layers
  common = _load("xmcgan.nets.common", f"{REF}/xmcgan/nets/common.py")
  sys.modules["xmcgan.nets"].common = common
  return _load("xmcgan.nets.xmc_net", f"{REF}/xmcgan/nets/xmc_net.py")


def flatten(tree, prefix=""):
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(flatten(v, f"{prefix}{k}/"))
    else:
      out[prefix + k] = np.asarray(v, np.float32)
  return out


def unflatten(flat):
  tree = {}
  for key, v in flat.items():
    node = tree
    parts = key.split("/")
    for p in parts[:-1]:
      node = node.setdefault(p, {})
    node[parts[-1]] = v
  return tree


def net_inputs(**overrides):
  """Config, variables (the product's own initialiser: same Flax tree names as the reference expects) and a batch."""
  import torch
  from tests import helpers
  cfg = helpers.small_config(**dict(NET_CFG, **overrides))
  _, _, g_vars, d_vars = helpers.cpu_variables(cfg, E=NET_E, seed=21)
  to_np = lambda t: {k: to_np(v) if isinstance(v, dict) else v.detach().numpy().astype(np.float32) for k, v in t.items()}
  g_vars, d_vars = to_np(g_vars), to_np(d_vars)
  # running statistics that are not the initial (0, 1): the eval-mode pass must really use them
  rng = np.random.default_rng(5)
  for k, v in flatten(g_vars["batch_stats"]).items():
    v += (rng.normal(size=v.shape) * 0.1).astype(np.float32) if k.endswith("mean") else \
        (rng.random(size=v.shape) * 0.5).astype(np.float32)
  batch = {k: v.numpy().astype(np.float32) for k, v in helpers.make_batch(NET_B, cfg, E=NET_E, L=NET_L, seed=4).items()}
  return cfg, g_vars, d_vars, batch


def compute_nets():
  xmc_net = load_reference_nets()
  cfg, g_vars, d_vars, batch = net_inputs()
  # the weights are rebuilt from their seed by net_inputs() (they would make the fixture several MB): the fixture keeps
  # one checksum pair per tree so that a change of the initialiser shows up as such, not as a parity failure
  out = {"cfg/" + k: np.array(v) for k, v in NET_CFG.items()}
  for name, tree in (("g_vars", g_vars), ("d_vars", d_vars), ("batch", batch)):
    leaves = flatten(tree)
    out["checksum/" + name] = np.array([sum(float(v.astype(np.float64).sum()) for v in leaves.values()),
                                        sum(float(np.abs(v.astype(np.float64)).sum()) for v in leaves.values())])
  z = batch["z"]
  # Generator: train mode (batch statistics, new running averages) and inference mode (running averages)
  img, g_new = xmc_net.Generator(config=cfg, train=True).apply(g_vars, (batch, z), mutable=["batch_stats"])
  out["g_train/image"] = img
  out.update({"g_train/new/" + k: v for k, v in flatten(g_new).items()})
  out["g_eval/image"] = xmc_net.Generator(config=cfg, train=False).apply(g_vars, (batch, z), mutable=False)
  # Discriminator on [real; fake] (train mode: the power-iteration vectors advance)
  both = np.concatenate([batch["image"], img], axis=0)
  (logit, stats), d_new = xmc_net.Discriminator(config=cfg, train=True).apply(
      d_vars, (both, batch), mutable=["spectral_norm_stats"])
  out["d_train/logit"] = logit
  out.update({"d_train/stats/" + k: np.array(float(v), np.float64) for k, v in stats.items()})
  out.update({"d_train/new/" + k: v for k, v in flatten(d_new).items()})
  (logit_e, _) = xmc_net.Discriminator(config=cfg, train=False).apply(d_vars, (both, batch), mutable=False)
  out["d_eval/logit"] = logit_e
  # variant: spectrally normalised generator (every conv / dense of G through layers.SpectralConv / SpectralDense, its
  # own spectral_norm_stats collection); the image is kept at every 4th pixel to keep the fixture small
  cfg2, g2, _, batch2 = net_inputs(g_spectral_norm=True)
  leaves = flatten(g2)
  out["checksum/g_vars_sn"] = np.array([sum(float(v.astype(np.float64).sum()) for v in leaves.values()),
                                        sum(float(np.abs(v.astype(np.float64)).sum()) for v in leaves.values())])
  img2, new2 = xmc_net.Generator(config=cfg2, train=True).apply(g2, (batch2, batch2["z"]),
                                                               mutable=["batch_stats", "spectral_norm_stats"])
  out["g_sn_train/image_s4"] = img2[:, ::4, ::4]
  out.update({"g_sn_train/new/" + k: v for k, v in flatten(new2).items()})
  # variant: the 256 px configuration (six blocks in both networks, xmc_net.py:81-86,202-205); image every 8th pixel
  cfg3, g3, d3, batch3 = net_inputs(image_size=256)
  out["checksum/vars_256"] = checksum(dict(g=g3, d=d3, b=batch3))
  img3, _ = xmc_net.Generator(config=cfg3, train=True).apply(g3, (batch3, batch3["z"]), mutable=["batch_stats"])
  (logit3, stats3), _ = xmc_net.Discriminator(config=cfg3, train=True).apply(
      d3, (np.concatenate([batch3["image"], img3]), batch3), mutable=["spectral_norm_stats"])
  out["px256/image_s8"], out["px256/image_mean"], out["px256/logit"] = img3[:, ::8, ::8], img3.mean(axis=(1, 2)), logit3
  out.update({"px256/stats/" + k: np.array(float(v), np.float64) for k, v in stats3.items()})
  # variants: the contrastive-loss switches of coco_xmc.py off one at a time (discriminator on the 128 px images above)
  for switch in ("word_contrastive", "sentence_contrastive", "image_contrastive"):
    cfg4, _, d4, batch4 = net_inputs(**{switch: False})
    (logit4, stats4), _ = xmc_net.Discriminator(config=cfg4, train=True).apply(
        d4, (both, batch4), mutable=["spectral_norm_stats"])
    out[f"no_{switch}/logit"] = logit4
    out.update({f"no_{switch}/stats/" + k: np.array(float(v), np.float64) for k, v in stats4.items()})
  return out


def checksum(tree):
  leaves = flatten(tree)
  return np.array([sum(float(v.astype(np.float64).sum()) for v in leaves.values()),
                   sum(float(np.abs(v.astype(np.float64)).sum()) for v in leaves.values())])


# ---------------------------------------------------------------------------------------------------------------------
# Gradients: central differences of the reference's own loss_fn (xmc_gan.train_g_d / train_d), stop_gradient honoured
# ---------------------------------------------------------------------------------------------------------------------
GRAD_EPS = 3e-8   # the loss is piecewise smooth (ReLU kinks, B = 2): at 1e-5 kink crossings cost 2 %, below 1e-7 the quotient is stable to 1e-5
D_LEAVES = ("DiscOptimizedBlock_0/SpectralConv_0/kernel", "DiscBlock_3/SpectralConv_1/kernel",
            "DiscBlock_1/SpectralConv_2/bias", "SpectralDense_0/kernel", "SpectralDense_1/kernel", "SpectralConv_0/kernel")
G_LEAVES = ("Dense_0/kernel", "Dense_1/bias", "GenBlock_0/ConditionalBatchNorm_0/Dense_0/kernel",
            "GenBlock_1/Conv_2/kernel", "GenSpatialBlock_1/LocalConditionalBatchNorm_1/Conv_1/kernel", "Conv_0/kernel",
            "Conv_1/kernel")


def grad_directions(g_params, d_params):
  """[(name, which network, {leaf path: direction})]: three dense random directions per network and one direction per
  selected leaf, from a fixed seed over the leaves in sorted order (the test rebuilds exactly these)."""
  rng = np.random.default_rng(77)
  out = []
  for net, params, leaves in (("d", d_params, D_LEAVES), ("g", g_params, G_LEAVES)):
    flat = flatten(params)
    for i in range(3):
      out.append((f"{net}/dense{i}", net, {k: rng.normal(size=flat[k].shape) for k in sorted(flat)}))
    for leaf in leaves:
      out.append((f"{net}/{leaf}", net, {leaf: rng.normal(size=flat[leaf].shape)}))
  return out


def reference_loss_fn(xmc_net, losses, cfg, state, batch):
  """train_g_d's nested loss_fn(params_d, params_g) -> ((d_loss, g_loss), aux) and calculate_contrastive_loss, taken
  out of xmcgan/xmc_gan.py by their AST nodes and executed with the closure variables of train_g_d supplied as globals
  (train_d's loss_fn returns the same d_loss: same code up to the unused generator statistics). Float64 throughout:
  the one `jnp.float32` cast of the logits in that code is mapped to float64 so that differences of 1e-5 survive."""
  import functools
  tree = ast.parse(open(f"{REF}/xmcgan/xmc_gan.py").read())
  top = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
  nested = next(n for n in ast.walk(top["train_g_d"]) if isinstance(n, ast.FunctionDef) and n.name == "loss_fn")
  jnp64 = types.SimpleNamespace(**vars(sys.modules["jax.numpy"]))
  jnp64.float32 = np.float64
  scope = {"jnp": jnp64, "jax": sys.modules["jax"], "losses": losses, "state": state, "batch": batch, "rng": None,
           "config": cfg, "dtype": np.float64, "additional_data": {},
           "generator": functools.partial(xmc_net.Generator, config=cfg, dtype=np.float64),
           "discriminator": functools.partial(xmc_net.Discriminator, config=cfg, dtype=np.float64)}
  exec(compile(ast.Module(body=[top["calculate_contrastive_loss"], nested], type_ignores=[]), "xmc_gan.py", "exec"), scope)
  return scope["loss_fn"]


def compute_grads():
  """Directional derivatives of d_loss wrt the discriminator's parameters and of g_loss wrt the generator's (the two
  pull-backs of xmc_gan.py:162-167) by central differences of the reference's loss_fn in float64."""
  from tests.golden import flax_stand_in as F
  xmc_net = load_reference_nets()
  losses = sys.modules["xmcgan.libml.losses"]
  cfg, g_vars, d_vars, batch = net_inputs()
  to64 = lambda t: {k: to64(v) if isinstance(v, dict) else np.asarray(v, np.float64) for k, v in t.items()}
  g_vars, d_vars, batch = to64(g_vars), to64(d_vars), to64(batch)
  state = types.SimpleNamespace(generator_state={"batch_stats": g_vars["batch_stats"]},
                                discriminator_state={"spectral_norm_stats": d_vars["spectral_norm_stats"]})
  loss_fn = reference_loss_fn(xmc_net, losses, cfg, state, batch)
  F.DTYPE[0] = np.float64
  try:
    F.SG.update(mode="record", tape=[], pos=0)
    (d0, g0), _ = loss_fn(d_vars["params"], g_vars["params"])
    out = {"grads/base": np.array([float(d0), float(g0)]), "grads/eps": np.array(GRAD_EPS),
           "grads/stopped_values": np.array(len(F.SG["tape"]))}

    def shifted(params, direction, sign):
      flat = flatten_any(params)
      for k, v in direction.items():
        flat[k] = flat[k] + sign * GRAD_EPS * v
      return unflatten(flat)

    def losses_at(pd, pg):
      F.SG.update(mode="replay", pos=0)
      (d, g), _ = loss_fn(pd, pg)
      assert F.SG["pos"] == len(F.SG["tape"])
      return float(d), float(g)

    for name, net, direction in grad_directions(g_vars["params"], d_vars["params"]):
      if net == "d":
        plus = losses_at(shifted(d_vars["params"], direction, +1), g_vars["params"])[0]
        minus = losses_at(shifted(d_vars["params"], direction, -1), g_vars["params"])[0]
      else:
        plus = losses_at(d_vars["params"], shifted(g_vars["params"], direction, +1))[1]
        minus = losses_at(d_vars["params"], shifted(g_vars["params"], direction, -1))[1]
      out["grads/fd/" + name] = np.array((plus - minus) / (2 * GRAD_EPS))
  finally:
    F.DTYPE[0] = np.float32
    F.SG.update(mode=None, tape=[], pos=0)
  return out


def flatten_any(tree, prefix=""):
  """flatten() without the float32 cast."""
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(flatten_any(v, f"{prefix}{k}/"))
    else:
      out[prefix + k] = v
  return out


def resnet_inputs():
  """Synthetic frozen weights (oracle.resnet50_random_variables: the reference's checkpoint is not shipped) and two
  224 x 224 images (the bilinear resize in front of the network is jax.image.resize, not reference code)."""
  from oracle import xmc_oracle as orc
  variables = orc.resnet50_random_variables(seed=3)
  to_np = lambda t: {k: to_np(v) if isinstance(v, dict) else v.numpy().astype(np.float32) for k, v in t.items()}
  images = np.random.default_rng(9).random(size=(2, 224, 224, 3)).astype(np.float32)
  return to_np(variables), images


def compute_resnet():
  """utils/resnet_v1.py (ResNet50 = ResNet + ResNetStage + BottleneckResNetBlock) in inference mode on the stand-in."""
  load_reference_nets()
  resnet_v1 = _load("xmcgan.utils.resnet_v1", f"{REF}/xmcgan/utils/resnet_v1.py")
  variables, images = resnet_inputs()
  pool, logits = resnet_v1.ResNet50(num_classes=1000).apply(variables, images, mutable=False, train=False)
  leaves = flatten(variables)
  return {"resnet/checksum": np.array([sum(float(v.astype(np.float64).sum()) for v in leaves.values()),
                                       sum(float(np.abs(v.astype(np.float64)).sum()) for v in leaves.values()),
                                       float(images.astype(np.float64).sum())]),
          "resnet/pool_shape": np.array(pool.shape), "resnet/pool_c16": pool[..., ::16].astype(np.float32),
          "resnet/pool_mean": pool.mean(axis=(1, 2)).astype(np.float32), "resnet/logits": logits.astype(np.float32)}


if __name__ == "__main__":
  np.savez_compressed(OUT, **compute())
  print("wrote", OUT)
  nets = compute_nets()
  nets.update(compute_resnet())
  nets.update(compute_grads())
  np.savez_compressed(OUT_NETS, **nets)
  print("wrote", OUT_NETS, os.path.getsize(OUT_NETS), "bytes")
