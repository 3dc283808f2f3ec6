"""Counterparts of xmcgan/xmc_gan.py: train_d (:194-256), train_g_d (:93-191), calculate_contrastive_loss (:58-71),
create_additional_data (:43-55), TrainMetrics (:33-40).

State handling: the returned TrainState reuses (and overwrites) the parameter / moment / statistics buffers of the
state passed in — the argument is donated, as `state = p_train_step(state, ...)` in the reference loop
(train_utils.py:424) discards the old state anyway."""
import torch

from . import engine as _engine
from . import ops
from . import parallel
from .nets import xmc_net

_S = _engine.LOSS_SLOTS


class TrainMetrics:
  """Cross-replica mean of the five loss scalars (clu.metrics.Average of xmc_gan.py:33-40, gathered at :185-190)."""
  names = ("d_loss", "g_loss", "c_loss_d", "c_loss_g", "c_loss_g_pretrained")

  def __init__(self, values):
    self._values = values  # device fp32 [5], already averaged over replicas

  @classmethod
  def gather_from_model_output(cls, losses):
    l = losses
    d_c = l[_S["real_word"]] + l[_S["real_sent"]]
    g_c = l[_S["fake_word"]] + l[_S["fake_sent"]] + l[_S["image"]]
    vals = torch.stack([l[_S["hinge_d"]] + d_c, l[_S["hinge_g"]] + g_c + l[_S["pretrained"]], d_c, g_c,
                        l[_S["pretrained"]]])
    if parallel.world_size() > 1:
      parallel.all_reduce_sum_(vals)
      vals = vals / parallel.world_size()
    return cls(vals)

  def compute(self):
    host = self._values.detach().float().cpu().tolist()  # device -> host read of the step's result
    return dict(zip(self.names, host))


def calculate_contrastive_loss(result_dict):
  """xmc_gan.calculate_contrastive_loss (xmc_gan.py:58-71)."""
  real_loss = result_dict["real_word_loss"] + result_dict["real_sentence_loss"]
  fake_loss = result_dict["fake_word_loss"] + result_dict["fake_sentence_loss"]
  return real_loss, fake_loss + result_dict["image_contrastive_loss"]


def create_additional_data(config, variables=None, checkpoint_path="data/resnet_pretrained.npy",
                           image_model_dtype="float32"):
  """xmc_gan.create_additional_data (xmc_gan.py:43-55) + pretrained_model_utils.get_pretrained_model (:65-99): returns
  {"image_model", "image_model_state"} for the frozen ResNet-50. `variables` ({"params","batch_stats"} with the
  names of resnet_v1.py) takes precedence; otherwise the reference's .npy checkpoint is loaded from `checkpoint_path`
  (not shipped with the reference, README.md:60-63). image_model_dtype: "float32" = the reference's precision for this
  network (it is built without a dtype, pretrained_model_utils.py:87-91); "bfloat16" is the faster stated deviation."""
  if not config.pretrained_image_contrastive:
    return {}
  if variables is None:
    import numpy as np
    data = np.load(checkpoint_path, allow_pickle=True).item()  # same format as pretrained_model_utils.py:95-98
    variables = {"params": data["params"], "batch_stats": data["batch_stats"]}
  model = _engine.ResNetEngine(dtype=image_model_dtype)
  model.load(variables)
  return {"image_model": model, "image_model_state": variables}


def calculate_contrastive_loss_on_pretrained(model, state, real_images, fake_images):
  """xmc_gan.calculate_contrastive_loss_on_pretrained (xmc_gan.py:74-90), forward value only."""
  b = real_images.shape[0]
  both = torch.cat([xmc_net._to_dev(real_images), xmc_net._to_dev(fake_images)])  # plumbing: one batched forward
  logits, _ = model.forward(both)
  slot = ops.empty(1, ops.F32)
  _engine.Contrastive(logits[:b], logits[b:], slot)
  return slot[0]


class _Workspace:
  """Persistent per-state device buffers (gradients, double-buffered statistics, loss slots)."""

  def __init__(self, state, g_eng, d_eng):
    self.g_grads = torch.zeros_like(state.g_optimizer.target.buf)
    self.d_grads = torch.zeros_like(state.d_optimizer.target.buf)
    self.g_stats_alt = torch.empty_like(state.generator_state["batch_stats"].buf)
    self.u0_alt = torch.empty_like(state.discriminator_state["spectral_norm_stats"].buf) if d_eng.sn else None
    self.g_u0_alt = torch.empty_like(state.generator_state["spectral_norm_stats"].buf) if g_eng.sn else None
    self.graph_mode = False  # True: state buffers keep their addresses (copies instead of pointer swaps)
    self.pending_d = None    # handle of train_d's in-flight D-gradient all-reduce (train_d_deferred)


def _workspace(state, g_eng, d_eng):
  ws = getattr(state, "_ws", None)
  if ws is None:
    ws = _Workspace(state, g_eng, d_eng)
    object.__setattr__(state, "_ws", ws)
  return ws


def _engines(config, batch):
  e = int(batch["embedding"].shape[-1])
  return xmc_net.get_engine(config, "g", e), xmc_net.get_engine(config, "d", e)


def _adam(opt, grads, ema=None, decay=0.0, handles=None):
  """One Adam step (+ EMA) on a flat parameter buffer. handles: [(lo, hi, work)] from _all_reduce_sliced — the step
  is then applied slice by slice, each as soon as its part of the gradient sum has arrived, so that the update of one
  slice runs while NCCL reduces the next."""
  opt.step += 1
  t = opt.step
  # graph mode (train_utils.GraphedTrainStep): the step count lives in a device int so that a replayed launch computes
  # this step's bias corrections itself
  step_dev = opt.step_dev
  n = opt.target.buf.numel()
  for lo, hi, work in (handles or [(0, n, None)]):
    if work is not None:
      work.wait()
    last = hi == n
    ops._call("xmc_adam", opt.target.buf.data_ptr() + 4 * lo, grads.data_ptr() + 4 * lo, opt.m.data_ptr() + 4 * lo,
              opt.v.data_ptr() + 4 * lo, hi - lo, opt.learning_rate, opt.beta1, opt.beta2, opt.eps,
              1.0 - opt.beta1 ** t, 1.0 - opt.beta2 ** t, 1.0 / parallel.world_size(),
              ema.data_ptr() + 4 * lo if ema is not None else None, decay,
              step_dev.data_ptr() if step_dev is not None else None, int(last), ops._lib.stream(),
              launches=2 if (step_dev is not None and last) else 1)


def _all_reduce_sliced(flat, slices):
  """Sum all-reduce of a flat gradient buffer as `slices` consecutive in-flight all-reduces (NCCL runs them in order on
  its stream). Returns [(lo, hi, work)] for _adam, or None on one replica."""
  if parallel.world_size() == 1:
    return None
  n = flat.numel()
  per = ((n + slices - 1) // slices + 1023) // 1024 * 1024
  out = []
  for lo in range(0, n, per):
    hi = min(lo + per, n)
    out.append((lo, hi, parallel.all_reduce_sum_(flat[lo:hi], async_op=True)))
  return out


def _with_z(rng, batch, config):
  """The reference samples z = jax.random.normal(rng, (B, z_dim)) when the batch carries none (xmc_gan.py:131-135,
  227-231). Here: a torch CUDA generator seeded from `rng` (JAX's threefry stream cannot be reproduced bit for bit)."""
  if "z" in batch:
    return batch
  n = batch["image"].shape[0]
  g = torch.Generator(device="cuda").manual_seed(xmc_net._seed_of(rng))
  return dict(batch, z=torch.randn(n, config.z_dim, device="cuda", generator=g))


def _forward_both(state, batch, config, ws, g_eng, d_eng, losses, keep_g_state, need_g):
  """Generator forward, then discriminator forward on [real; fake]. Returns (gctx, dctx, state): `state` is the
  argument, or — when a deferred train_d update was pending (train_d_deferred) — the state with that update applied
  (it is completed between the two forwards: the generator forward does not read the discriminator)."""
  B = batch["z"].shape[0]
  S = config.image_size
  g_params = state.g_optimizer.target.buf
  # the generator's bf16 weights (and, with g_spectral_norm, its power-iteration step) are shared by train_d and
  # train_g_d of one train_step: same parameters, same u0 (train_d discards the generator's new state, xmc_gan.py:225)
  g_u0 = state.generator_state["spectral_norm_stats"].buf if g_eng.sn else None
  if g_eng.prepped_for != g_eng.prep_key(g_params, ws.g_u0_alt):
    g_eng.prep_weights(g_params, g_u0, ws.g_u0_alt)
  all_images = ops.empty((2 * B, S, S, 3))
  ops.cast_to_bf16(batch["image"].reshape(B * S * S, 3), all_images[:B].view(B * S * S, 3))
  fake, gctx = g_eng.forward(g_params, state.generator_state["batch_stats"].buf, batch, batch["z"], train=True,
                             new_stats=ws.g_stats_alt if keep_g_state else None, fake_bf16=all_images[B:])
  state = _finish_pending_d(state, ws, d_eng)
  d_params = state.d_optimizer.target.buf
  u0 = state.discriminator_state["spectral_norm_stats"].buf if d_eng.sn else None
  d_eng.prep_weights(d_params, u0, ws.u0_alt)
  _, dctx = d_eng.forward(d_params, all_images, batch, losses, need_g=need_g)
  gctx["fake"] = fake
  return gctx, dctx, state


def _all_reduce_async(flat):
  """Sum all-reduce of a flat gradient buffer left in flight on NCCL's stream; the tensor-core GEMMs issued while it
  runs leave parallel.RESERVED_SMS SMs to its thread blocks (ops.reserve_sms) for about its duration."""
  handle = parallel.all_reduce_sum_(flat, async_op=True)
  if handle is not None and parallel.RESERVED_SMS > 0:
    ops.reserve_sms(parallel.RESERVED_SMS, parallel.reserve_tflop(flat.numel() * 4))
  return handle


def _finish_pending_d(state, ws, d_eng):
  """Completes a train_d whose gradient all-reduce was left in flight (train_d_deferred): wait, Adam on D, keep the new
  u0. No-op when nothing is pending."""
  if ws.pending_d is None:
    return state
  handle, ws.pending_d = ws.pending_d, None
  if handle is not True:
    ops.release_sms()
    handle.wait()
  _adam(state.d_optimizer, ws.d_grads)
  new_state = state.replace(discriminator_state=_swap_d_state(state, ws, d_eng))
  object.__setattr__(new_state, "_ws", ws)
  return new_state


def _swap_d_state(state, ws, d_eng):
  if not d_eng.sn:
    return state.discriminator_state
  old = state.discriminator_state["spectral_norm_stats"]
  if ws.graph_mode:  # fixed buffer addresses: copy the new u0 back instead of swapping the two buffers
    old.buf.copy_(ws.u0_alt)
    return state.discriminator_state
  new = xmc_net.FlatTree(d_eng.u_layout, ws.u0_alt)
  ws.u0_alt = old.buf
  return {"spectral_norm_stats": new}


def train_d(rng, state, batch, generator, discriminator, config):
  """xmc_gan.train_d (xmc_gan.py:194-256): d_loss = hinge_d + real_word + real_sentence; gradient wrt params_d only;
  pmean; Adam on D; the generator's new batch statistics are discarded, the discriminator's new u0 is kept."""
  with ops.act_dtype(_engine.act_dtype_of(config)):
    return _train_d(rng, state, batch, generator, discriminator, config)


def train_d_deferred(rng, state, batch, generator, discriminator, config):
  """train_d whose gradient all-reduce is left in flight: the wait, Adam on D and the u0 hand-over happen inside the next
  train_d / train_g_d call on the returned state, after its generator forward (which does not read the discriminator)
  — the all-reduce hides behind that forward instead of stalling the stream. Used by train_utils.train_step; the
  returned state must go straight into the next train_d / train_g_d (its D parameters are the OLD ones until then)."""
  with ops.act_dtype(_engine.act_dtype_of(config)):
    return _train_d(rng, state, batch, generator, discriminator, config, defer=True)


def _train_d(rng, state, batch, generator, discriminator, config, defer=False):
  batch = _with_z(rng, xmc_net.batch_to_device(batch), config)
  g_eng, d_eng = _engines(config, batch)
  ws = _workspace(state, g_eng, d_eng)
  losses = torch.zeros(16, device="cuda")
  gctx, dctx, state = _forward_both(state, batch, config, ws, g_eng, d_eng, losses, keep_g_state=False, need_g=False)
  del gctx
  d_params = state.d_optimizer.target.buf
  ws.d_grads.zero_()
  ops.LAUNCHES[0] += 1
  d_eng.backward_d(dctx, d_params, ws.d_grads)
  d_eng.sn_backward(d_params, ws.d_grads, ws.u0_alt)
  if defer:
    ws.pending_d = _all_reduce_async(ws.d_grads) or True
    object.__setattr__(state, "_ws", ws)
    return state
  parallel.all_reduce_sum_(ws.d_grads)
  _adam(state.d_optimizer, ws.d_grads)
  new_d_state = _swap_d_state(state, ws, d_eng)
  new_state = state.replace(discriminator_state=new_d_state)
  object.__setattr__(new_state, "_ws", ws)
  return new_state


def train_g_d(rng, state, batch, generator, discriminator, config, additional_data):
  """xmc_gan.train_g_d (xmc_gan.py:93-191): one forward, two pull-backs at the old parameters (:162-167), pmean of
  both gradients (:170-171), Adam on D and G (:172-173), polyak EMA (:174-177), metrics (:185-190)."""
  with ops.act_dtype(_engine.act_dtype_of(config)):
    return _train_g_d(rng, state, batch, generator, discriminator, config, additional_data)


def _train_g_d(rng, state, batch, generator, discriminator, config, additional_data):
  batch = _with_z(rng, xmc_net.batch_to_device(batch), config)
  g_eng, d_eng = _engines(config, batch)
  ws = _workspace(state, g_eng, d_eng)
  losses = torch.zeros(16, device="cuda")
  gctx, dctx, state = _forward_both(state, batch, config, ws, g_eng, d_eng, losses, keep_g_state=True, need_g=True)
  g_params = state.g_optimizer.target.buf
  d_params = state.d_optimizer.target.buf
  # pull-back #1: d_loss -> params_d
  ws.d_grads.zero_()
  ops.LAUNCHES[0] += 1
  d_eng.backward_d(dctx, d_params, ws.d_grads)
  d_eng.sn_backward(d_params, ws.d_grads, ws.u0_alt)
  h_d = _all_reduce_async(ws.d_grads)
  # pull-back #2: g_loss -> fake images -> params_g
  d_fake = d_eng.backward_g(dctx, d_params)
  del dctx
  if config.pretrained_image_contrastive:
    # frozen ResNet-50 image-image InfoNCE between real and generated images (xmc_gan.py:148-152); its gradient
    # reaches the generator only through the fake images
    model = additional_data["image_model"]
    B = batch["z"].shape[0]
    S = config.image_size
    both = ops.empty((2 * B, S, S, 3), ops.F32)
    both[:B].copy_(batch["image"])     # device-to-device copies (plumbing)
    both[B:].copy_(gctx["fake"])
    logits, rctx = model.forward(both)
    c = _engine.Contrastive(logits[:B], logits[B:], losses[_S["pretrained"]:])
    dl_fake = ops.empty((B, logits.shape[1]), ops.F32)
    c.bwd_b(dl_fake, accumulate=False)
    model.backward(rctx, dl_fake, B, d_fake)
    del rctx
  ws.g_grads.zero_()
  ops.LAUNCHES[0] += 1
  g_eng.backward(gctx, d_fake, g_params, ws.g_grads)
  g_eng.sn_backward(g_params, ws.g_grads, ws.g_u0_alt)
  del gctx
  # the generator's all-reduce is the one nothing hides (its largest leaves are the last ones the backward produces); D's
  # Adam (whose all-reduce finished long ago, behind the generator backward) runs under it. parallel.G_SLICES > 1 issues
  # it in slices with G's Adam + EMA of slice i under the all-reduce of slice i+1 (measured: no gain, see parallel.py)
  h_g = _all_reduce_sliced(ws.g_grads, parallel.G_SLICES)
  ops.release_sms()
  if h_d is not None:
    h_d.wait()
  _adam(state.d_optimizer, ws.d_grads)
  _adam(state.g_optimizer, ws.g_grads, ema=state.ema_params.buf, decay=config.polyak_decay, handles=h_g)
  g_eng.prepped_for = None  # xmc_adam rewrote the parameters through raw pointers
  old_stats = state.generator_state["batch_stats"]
  if ws.graph_mode:
    old_stats.buf.copy_(ws.g_stats_alt)
    if g_eng.sn:
      state.generator_state["spectral_norm_stats"].buf.copy_(ws.g_u0_alt)
    new_g_state = state.generator_state
  else:
    new_g_state = {"batch_stats": xmc_net.FlatTree(g_eng.stats_layout, ws.g_stats_alt)}
    ws.g_stats_alt = old_stats.buf
    if g_eng.sn:
      old_u0 = state.generator_state["spectral_norm_stats"]
      new_g_state["spectral_norm_stats"] = xmc_net.FlatTree(g_eng.u_layout, ws.g_u0_alt)
      ws.g_u0_alt = old_u0.buf
  new_d_state = _swap_d_state(state, ws, d_eng)
  new_state = state.replace(step=state.step + 1, generator_state=new_g_state, discriminator_state=new_d_state)
  object.__setattr__(new_state, "_ws", ws)
  metrics = TrainMetrics.gather_from_model_output(losses)
  return new_state, metrics
