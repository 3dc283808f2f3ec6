"""FID / Inception-score statistics (reference xmcgan/utils/tf_inception_utils.py:123-224) against
closed forms, and the EvalMetric flow with a stub feature network."""
import numpy as np
import pytest
import torch

from xmcgan_image_generation_b200.utils import inception_utils

gpu = pytest.mark.gpu


def test_frechet_distance_closed_forms():
  rng = np.random.default_rng(0)
  d = 16
  # commuting (diagonal) covariances: d^2 = |dmu|^2 + sum (sqrt(a) - sqrt(b))^2
  a, b = rng.uniform(0.5, 2.0, d), rng.uniform(0.5, 2.0, d)
  mu1, mu2 = rng.normal(size=d), rng.normal(size=d)
  want = np.sum((mu1 - mu2) ** 2) + np.sum((np.sqrt(a) - np.sqrt(b)) ** 2)
  got = inception_utils._calculate_frechet_distance(mu1, np.diag(a), mu2, np.diag(b))
  assert abs(got - want) < 1e-9 * max(1.0, want)
  # identical Gaussians: zero; a rotated full covariance against itself as well
  q, _ = np.linalg.qr(rng.normal(size=(d, d)))
  s = q @ np.diag(a) @ q.T
  assert abs(inception_utils._calculate_frechet_distance(mu1, s, mu1, s)) < 1e-8
  with pytest.raises(inception_utils.ShapeNotMatchError):
    inception_utils._calculate_frechet_distance(mu1, s, mu2[:-1], s)


def test_fid_of_samples_matches_the_population_value():
  rng = np.random.default_rng(1)
  d, n = 8, 200000
  x = rng.normal(size=(n, d))
  y = rng.normal(size=(n, d)) * 2.0 + 0.5       # N(0.5, 4 I) vs N(0, I): d^2 = d * 0.25 + d * (2 - 1)^2
  want = d * 0.25 + d * 1.0
  got = inception_utils.calculate_fid(x, y)
  assert abs(got - want) < 0.05 * want
  assert abs(inception_utils.calculate_fid(x, x)) < 1e-6


def test_inception_score_known_answers():
  n, c = 1000, 10
  uniform = np.full((n, c), 1.0 / c)
  m, s = inception_utils.calculate_inception_score(uniform, num_splits=10)
  assert abs(m - 1.0) < 1e-12 and s < 1e-12      # p(y|x) == p(y): score 1
  # confident and evenly spread over the classes: score -> number of classes
  eps = 1e-9
  sharp = np.full((n, c), eps)
  sharp[np.arange(n), np.arange(n) % c] = 1.0 - (c - 1) * eps
  m, _ = inception_utils.calculate_inception_score(sharp, num_splits=10)
  assert abs(m - c) < 1e-3
  # rows beyond num_splits * (n // num_splits) are ignored (tf_inception_utils.py:217-219)
  extra = np.concatenate([uniform, sharp[:7]])
  assert inception_utils.calculate_inception_score(extra, num_splits=10)[0] == pytest.approx(1.0, abs=1e-12)


def test_eval_metric_needs_a_feature_network():
  from xmcgan_image_generation_b200.utils import eval_metrics
  with pytest.raises(NotImplementedError):
    eval_metrics.EvalMetric(iter(()), None)


@gpu
def test_eval_metric_flow_with_a_stub_feature_network():
  """EvalMetric.calculate_inception_fid end to end on the GPU generator (current and EMA parameters, inference mode)
  with a stub in place of Inception-v3: per-channel image statistics as 'pool' features, a softmax over 6 pooled
  values as 'preds'. The EMA generator equals the current one at initialisation, so both FIDs agree; the scores are
  finite and the Inception score lies in [1, classes]."""
  from tests import helpers
  from xmcgan_image_generation_b200 import train_utils
  from xmcgan_image_generation_b200.nets import xmc_net
  from xmcgan_image_generation_b200.utils import eval_metrics
  cfg = helpers.small_config()
  cfg.eval_num, cfg.eval_batch_size, cfg.eval_avg_num = 12, 4, 2
  batch = helpers.make_batch(4, cfg, seed=3)

  def ds():
    while True:
      yield batch

  def stub(images):
    x = torch.as_tensor(np.asarray(images.cpu()) if torch.is_tensor(images) else images).float()
    pool = torch.cat([x.mean((1, 2)), x.std((1, 2)), x[:, ::8, ::8].reshape(x.shape[0], -1)[:, :10]], 1)
    preds = torch.softmax(pool[:, :6] * 5.0, -1)
    return pool.numpy(), preds.numpy()

  gen, disc, state = train_utils.create_train_state(cfg, 1, batch)
  em = eval_metrics.EvalMetric(ds(), cfg, num_splits=1, inception_fn=stub)
  out = em.calculate_inception_fid(gen, state, rng=7)
  fid, fid_std, inc, inc_std, ema_fid, ema_fid_std, ema_inc, ema_inc_std = out
  assert all(np.isfinite(v) for v in out)
  assert abs(fid - ema_fid) < 1e-6 * max(1.0, abs(fid)) and 1.0 - 1e-9 <= inc <= 6.0 and inc == pytest.approx(ema_inc)
