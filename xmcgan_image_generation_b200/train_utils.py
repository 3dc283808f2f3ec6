"""Counterparts of the hot-path symbols of xmcgan/train_utils.py: TrainState (:42-50), split_input_dict (:69-88),
train_step (:91-130), create_train_state (:133-193), and the sampling path generate_batch (:245-309, SURVEY.md §8f
rank 1). The outer loop (train/test, summaries) is out of scope; the checkpoint format is in
checkpoint.py."""
import dataclasses
import functools
import math
from typing import Any, Optional

import torch

from . import parallel
from .nets import xmc_net


class Optimizer:
  """flax.optim.Optimizer(Adam) stand-in: `target` (the parameters), Adam moments and step count."""

  def __init__(self, target, learning_rate, beta1, beta2, eps=1e-8):
    self.target = target  # FlatTree
    self.m = torch.zeros_like(target.buf)
    self.v = torch.zeros_like(target.buf)
    self.step = 0
    self.step_dev = None  # device copy of `step` (int32[1]) once a GraphedTrainStep drives this optimiser
    self.learning_rate, self.beta1, self.beta2, self.eps = learning_rate, beta1, beta2, eps

  def set_step(self, step):
    """Sets the Adam step count on the host AND, when present, in the device counter the graph-replayed xmc_adam reads
    (a checkpoint restore that only set `.step` would leave the bias corrections of later steps stale)."""
    self.step = int(step)
    if self.step_dev is not None:
      self.step_dev.fill_(self.step)


@dataclasses.dataclass
class TrainState:
  """train_utils.TrainState (train_utils.py:42-50)."""
  step: int
  g_optimizer: Optimizer
  d_optimizer: Optimizer
  generator_state: Optional[Any]
  discriminator_state: Optional[Any]
  ema_params: Any

  def replace(self, **kw):
    return dataclasses.replace(self, **kw)


def split_input_dict(input_dict, splits, axis=0):
  """train_utils.split_input_dict (train_utils.py:69-88): equal splits of every leaf along `axis` (views)."""
  output = [dict() for _ in range(splits)]
  for key, value in input_dict.items():
    n = value.shape[axis]
    if n % splits:
      raise ValueError(f"cannot split axis of size {n} into {splits} equal parts")  # jnp.split raises as well
    for i, part in enumerate(torch.split(torch.as_tensor(value), n // splits, dim=axis)):
      output[i][key] = part
  return output


def train_step(rng, state, batch, gan_model, generator, discriminator, config, additional_data):
  """train_utils.train_step (train_utils.py:91-130): d_step_per_g_step-1 discriminator steps, then one joint step."""
  batch = xmc_net.batch_to_device(batch)
  batches = split_input_dict(batch, config.d_step_per_g_step)
  # rngs = jax.random.split(rng, d_step_per_g_step) (train_utils.py:121): one derived seed per sub-step; they are only
  # consumed when the batch carries no "z" (xmc_gan.py:131-135)
  seed = xmc_net._seed_of(rng)
  rngs = [(seed * 1000003 + i + 1) & 0x7FFFFFFF for i in range(config.d_step_per_g_step)]
  # with several replicas every train_d leaves its gradient all-reduce in flight; the next call completes it behind its
  # generator forward (xmc_gan.train_d_deferred). A foreign gan_model without that entry point runs the plain train_d.
  train_d = gan_model.train_d
  if parallel.world_size() > 1 and hasattr(gan_model, "train_d_deferred"):
    train_d = gan_model.train_d_deferred
  for i in range(config.d_step_per_g_step - 1):
    state = train_d(rngs[i], state, batches[i], generator, discriminator, config)
  return gan_model.train_g_d(rngs[-1], state, batches[-1], generator, discriminator, config, additional_data)


def make_grid(samples, show_num=64):
  """image_utils.make_grid (xmcgan/utils/image_utils.py:23-38): the first h*w images tiled into one [h*H, w*W, C]
  image. A pure index permutation (views / one strided copy, no arithmetic): exact."""
  batch_size, height, width, c = samples.shape
  show_num = min(show_num, batch_size)
  h_num = int(math.sqrt(show_num))
  w_num = int(show_num / h_num)
  grid = samples[:h_num * w_num].reshape(h_num, w_num, height, width, c).transpose(1, 2)
  return grid.reshape(height * h_num, width * w_num, c)


def generate_batch(rng, state, batch, generator, config, collect_all=False, z=None):
  """train_utils.generate_batch (train_utils.py:245-309): samples with the current and with the EMA generator
  parameters in inference mode (running BatchNorm statistics; with g_spectral_norm the power iteration still runs but
  u0 is not advanced), tiled into grids next to the original images. As in the reference z is ALWAYS drawn fresh from
  `rng` (train_utils.py:269-270; batch["z"] is ignored) — a torch CUDA generator, since JAX's threefry stream cannot
  be reproduced; the extra keyword `z` pins it explicitly (parity tests). collect_all gathers over all replicas
  (jax.lax.all_gather over "batch")."""
  batch = xmc_net.batch_to_device(batch)
  n = batch["image"].shape[0]
  if z is not None:
    z = xmc_net._to_dev(z)
  else:
    g = torch.Generator(device="cuda").manual_seed(xmc_net._seed_of(rng))
    z = torch.randn(n, config.z_dim, device="cuda", generator=g)
  g_variables = dict(state.generator_state, params=state.g_optimizer.target)
  ema_g_variables = dict(state.generator_state, params=state.ema_params)
  images = {"generated_image_batch": generator(train=False).apply(g_variables, (batch, z), mutable=False),
            "ema_generated_image_batch": generator(train=False).apply(ema_g_variables, (batch, z), mutable=False),
            "ori_image_batch": batch["image"]}
  out = {}
  for key, img in images.items():
    img = img.float()
    if collect_all and parallel.world_size() > 1:
      parts = [torch.empty_like(img) for _ in range(parallel.world_size())]
      torch.distributed.all_gather(parts, img.contiguous())
      img = torch.cat(parts)
    out[key] = make_grid(img, config.show_num)[None]  # the summary writer wants a 4-D tensor
  return out


class GraphedTrainStep:
  """train_step captured once into a CUDA graph and replayed: the ~730 kernel launches of a step become one graph
  launch (no per-launch host work, no inter-kernel launch gaps). Same semantics as train_step on the same state:

      step = GraphedTrainStep(state, batch, xmc_gan, generator, discriminator, config, additional_data)
      state, metrics = step(batch)        # every call = one train_step on `batch` (host or device tensors)

  What makes the step replayable: the input batch lives in fixed device buffers (each call copies into them), the Adam
  step counts live on the device (xmc_adam's step_dev), and the state buffers keep their addresses (new batch
  statistics / u0 are copied back instead of swapping buffers). Construction runs `warmup` real train_steps eagerly
  (kernel attributes, allocator pool, NCCL communicators) and one more while capturing — the capture does not execute,
  so construction advances the state by exactly `warmup` steps."""

  def __init__(self, state, batch, gan_model, generator, discriminator, config, additional_data, warmup=2):
    from . import xmc_gan as _xg
    self.config = config
    self.static_batch = {k: v.clone() for k, v in xmc_net.batch_to_device(batch).items()}
    g_eng, d_eng = _xg._engines(config, self.static_batch)
    ws = _xg._workspace(state, g_eng, d_eng)
    ws.graph_mode = True
    for opt in (state.g_optimizer, state.d_optimizer):
      opt.step_dev = torch.tensor([opt.step], device="cuda", dtype=torch.int32)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(warmup):
        state, _ = train_step(None, state, self.static_batch, gan_model, generator, discriminator, config,
                              additional_data)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    steps_before = (state.step, state.g_optimizer.step, state.d_optimizer.step)
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph):
      self.state, self.metrics = train_step(None, state, self.static_batch, gan_model, generator, discriminator, config,
                                            additional_data)
    # capturing ran the Python side of one step (host counters advanced) without executing it on the device
    self.state.step, self.state.g_optimizer.step, self.state.d_optimizer.step = steps_before
    self._d_steps = config.d_step_per_g_step
    self._host_steps = (self.state.g_optimizer.step, self.state.d_optimizer.step)

  def __call__(self, batch):
    """One train_step on `batch`. The graph is bound to the buffers of the state it was built on: restoring a
    checkpoint INTO that state (checkpoint.from_state_dict copies in place and calls Optimizer.set_step) is fine, a
    different TrainState object needs a new GraphedTrainStep."""
    st = self.state
    if (st.g_optimizer.step, st.d_optimizer.step) != self._host_steps:
      # someone moved the host counters (Optimizer.set_step keeps the device counters in line): resynchronise
      for opt in (st.g_optimizer, st.d_optimizer):
        opt.set_step(opt.step)
    for k, dst in self.static_batch.items():
      dst.copy_(torch.as_tensor(batch[k]).reshape(dst.shape), non_blocking=True)
    self.graph.replay()
    st.step += 1
    st.g_optimizer.step += 1
    st.d_optimizer.step += self._d_steps
    self._host_steps = (st.g_optimizer.step, st.d_optimizer.step)
    return st, self.metrics


def create_train_state(config, rng, init_batch):
  """train_utils.create_train_state (train_utils.py:133-193)."""
  from . import engine as _engine
  dtype = _engine.act_dtype_of(config)   # bf16 or fp32 activations (train_utils.py:148-151); anything else raises
  if config.architecture == "xmc_net":
    generator_cls, discriminator_cls = xmc_net.Generator, xmc_net.Discriminator
  else:
    raise ValueError(f"Architecture {config.architecture} is not supported.")
  generator = functools.partial(generator_cls, config=config, dtype=dtype)
  discriminator = functools.partial(discriminator_cls, config=config, dtype=dtype)
  seed = xmc_net._seed_of(rng)
  g_vars = dict(generator(train=False).init(seed * 3 + 1, (init_batch, None)))
  g_params = g_vars.pop("params")
  d_vars = dict(discriminator(train=False).init(seed * 3 + 2, (None, init_batch)))
  d_params = d_vars.pop("params")
  g_opt = Optimizer(g_params, config.g_lr, config.beta1, config.beta2)
  d_opt = Optimizer(d_params, config.d_lr, config.beta1, config.beta2)
  return generator, discriminator, TrainState(step=0, g_optimizer=g_opt, d_optimizer=d_opt, generator_state=g_vars,
                                              discriminator_state=d_vars, ema_params=g_params.clone())
