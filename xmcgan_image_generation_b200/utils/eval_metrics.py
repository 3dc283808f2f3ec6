"""EvalMetric (reference: xmcgan/utils/eval_metrics.py:29-216): FID / Inception score of the current and of the EMA
generator over `config.eval_num` images, averaged over `config.eval_avg_num` draws of z.

What runs on the GPU here is the generator (inference mode, the same kernels as generate_batch); the Inception-v3
network is an injected callable `inception_fn(images) -> (pool [n, 2048], preds [n, 1000])` on [n, H, W, 3] images in
[0, 1] — the reference builds it from downloaded Keras weights (inception_utils.inception_model), which this
environment does not have, and its layers (1x7 / 7x1 convolutions, average pooling, branch concatenation) are not
among the hot-path kernels (open). The statistics are utils/inception_utils.py."""
import numpy as np
import torch

from . import inception_utils
from .. import parallel
from ..nets import xmc_net


def _fold_in(rng, i):
  """A new integer seed from (rng, i) — the role of jax.random.fold_in (eval_metrics.py:139,196); JAX's threefry
  stream itself cannot be reproduced here."""
  return (xmc_net._seed_of(rng) * 1000003 + i + 1) & 0x7FFFFFFF


class EvalMetric:
  """ds: iterator of evaluation batches (the input contract of train_step); config: eval_num, eval_batch_size,
  eval_avg_num, z_dim; num_splits: splits of the Inception score; inception_fn: see the module docstring."""

  def __init__(self, ds, config, num_splits=1, inception_fn=None):
    if inception_fn is None:
      raise NotImplementedError("EvalMetric needs an Inception-v3 feature function: the network is not part of this "
                                "package (DESIGN.md §8, f4)")
    self.ds, self.config = ds, config
    self.eval_num, self.eval_batch_size, self.avg_num = config.eval_num, config.eval_batch_size, config.eval_avg_num
    self.num_splits = num_splits
    self._inception = inception_fn
    # the real images' pool features are computed once (eval_metrics.py:69-88)
    self._pool = self._get_real_pool_for_evaluation()

  def _n_iter(self):
    return self.eval_num // self.eval_batch_size + 1

  @staticmethod
  def _gather(t):
    """All replicas' rows (the reference all_gathers inside its pmap, eval_metrics.py:66-68)."""
    t = torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).float().contiguous()
    if parallel.world_size() == 1:
      return t.cpu().numpy()
    t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(parallel.world_size())]
    torch.distributed.all_gather(parts, t)
    return torch.cat(parts).cpu().numpy()

  def _get_real_pool_for_evaluation(self):
    pools = []
    for _ in range(self._n_iter()):
      pool, _ = self._inception(next(self.ds)["image"])
      pools.append(self._gather(pool))
    return np.concatenate(pools)[:self.eval_num]

  def _get_generated_images(self, rng, state, batch, generator):
    """Images of the current and of the EMA generator for one batch, same z, running BatchNorm statistics
    (eval_metrics.py:90-124)."""
    batch = xmc_net.batch_to_device(batch)
    n = batch["image"].shape[0]
    g = torch.Generator(device="cuda").manual_seed(xmc_net._seed_of(rng))
    z = torch.randn(n, self.config.z_dim, device="cuda", generator=g)
    cur = dict(state.generator_state, params=state.g_optimizer.target)
    ema = dict(state.generator_state, params=state.ema_params)
    gen = generator(train=False)
    return gen.apply(cur, (batch, z), mutable=False).float(), gen.apply(ema, (batch, z), mutable=False).float()

  def _get_generated_pool_for_evaluation(self, generator_fn, state, rng):
    out = [[], [], [], []]
    for step in range(self._n_iter()):
      batch = next(self.ds)
      image, ema_image = self._get_generated_images(_fold_in(rng, step), state, batch, generator_fn)
      for k, img in ((0, image), (2, ema_image)):
        pool, preds = self._inception(img)
        out[k].append(self._gather(pool))
        out[k + 1].append(self._gather(preds))
    return tuple(np.concatenate(v)[:self.eval_num] for v in out)

  def calculate_inception_fid(self, generator_fn, state, rng):
    """(fid, fid_std, inception_score, inception_score_std, ema_fid, ema_fid_std, ema_inception_score,
    ema_inception_score_std) over eval_avg_num draws (eval_metrics.py:173-216)."""
    fid, inc, ema_fid, ema_inc = [], [], [], []
    for i in range(self.avg_num):
      pool, preds, ema_pool, ema_preds = self._get_generated_pool_for_evaluation(generator_fn, state, _fold_in(rng, i))
      inc.append(inception_utils.calculate_inception_score(preds, num_splits=self.num_splits)[0])
      ema_inc.append(inception_utils.calculate_inception_score(ema_preds, num_splits=self.num_splits)[0])
      fid.append(inception_utils.calculate_fid(pool, self._pool))
      ema_fid.append(inception_utils.calculate_fid(ema_pool, self._pool))
    return (np.mean(fid), np.std(fid), np.mean(inc), np.std(inc), np.mean(ema_fid), np.std(ema_fid),
            np.mean(ema_inc), np.std(ema_inc))
