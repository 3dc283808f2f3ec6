"""FID / Inception-score statistics (reference: xmcgan/utils/inception_utils.py:152-181, which
forwards to xmcgan/utils/tf_inception_utils.py:123-224). As in the reference these are HOST computations on the
gathered Inception features (numpy / scipy, float64): they run once per evaluation over a few 2048 x 2048 matrices, not
on the step path. The Inception-v3 network that produces the features is NOT part of this package (its weights are a
download): eval_metrics.EvalMetric takes it as a callable."""
import warnings

import numpy as np
from scipy import linalg


class ShapeNotMatchError(ValueError):
  pass


class ImaginaryComponentError(ValueError):
  pass


def _calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
  """d^2 = |mu1 - mu2|^2 + Tr(S1 + S2 - 2 (S1 S2)^(1/2)) between N(mu1, S1) and N(mu2, S2)
  (tf_inception_utils.py:123-184): matrix square root of the product by scipy.linalg.sqrtm; a non-finite root is
  retried with eps on both diagonals; a root whose diagonal has an imaginary part above 1e-3 is an error, a smaller one
  is dropped."""
  mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
  sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
  if mu1.shape != mu2.shape:
    raise ShapeNotMatchError("mean vectors of different length")
  if sigma1.shape != sigma2.shape:
    raise ShapeNotMatchError("covariances of different shape")
  delta = mu1 - mu2
  root = linalg.sqrtm(sigma1 @ sigma2)   # (the reference passes disp=False, an argument newer scipy releases dropped)
  if not np.isfinite(root).all():
    warnings.warn(f"singular covariance product in the Frechet distance: adding {eps} to the diagonals")
    jitter = eps * np.eye(sigma1.shape[0])
    root = linalg.sqrtm((sigma1 + jitter) @ (sigma2 + jitter))
  if np.iscomplexobj(root):
    if not np.allclose(np.diagonal(root).imag, 0, atol=1e-3):
      raise ImaginaryComponentError(f"imaginary component {np.max(np.abs(root.imag))}")
    root = root.real
  return delta @ delta + np.trace(sigma1) + np.trace(sigma2) - 2.0 * np.trace(root)


def calculate_fid(pool1, pool2):
  """FID between two sets of Inception pool features [n, 2048] (tf_inception_utils.py:187-203): sample means and
  unbiased sample covariances (np.cov, rowvar=False)."""
  pool1, pool2 = np.asarray(pool1, np.float64), np.asarray(pool2, np.float64)
  return _calculate_frechet_distance(pool1.mean(0), np.cov(pool1, rowvar=False), pool2.mean(0),
                                     np.cov(pool2, rowvar=False))


def calculate_inception_score(pred, num_splits=10):
  """Inception score of class probabilities [n, classes] (tf_inception_utils.py:206-224): per split of n // num_splits
  rows exp(mean_x KL(p(y|x) || p(y))); returns (mean, std) over the splits. Rows beyond num_splits * (n // num_splits)
  are not used."""
  pred = np.asarray(pred, np.float64)
  per = pred.shape[0] // num_splits
  scores = []
  for s in range(num_splits):
    p = pred[s * per:(s + 1) * per]
    marginal = p.mean(0, keepdims=True)
    scores.append(np.exp(np.mean(np.sum(p * (np.log(p) - np.log(marginal)), axis=1))))
  return np.mean(scores), np.std(scores)
