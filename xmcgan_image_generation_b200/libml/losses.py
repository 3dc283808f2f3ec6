"""Counterparts of xmcgan/libml/losses.py on CUDA tensors (forward values; the training path uses the fused
forward+cotangent kernels through engine.py)."""
import torch

from .. import ops


def hinge_loss(real_logit, fake_logit):
  """losses.hinge_loss (losses.py:30-35) -> (discriminator_loss, generator_loss)."""
  r = torch.as_tensor(real_logit).to("cuda", torch.float32).reshape(-1)
  f = torch.as_tensor(fake_logit).to("cuda", torch.float32).reshape(-1)
  if r.numel() != f.numel():
    raise ValueError("real and fake logits must have the same batch size")
  logit = ops.empty(2 * r.numel(), ops.F32)
  logit[:r.numel()].copy_(r)
  logit[r.numel():].copy_(f)
  out = ops.empty(2, ops.F32)
  ops.hinge(logit, r.numel(), out[0:], out[1:])
  return out[0], out[1]


def hinge_loss_d(real_logit, fake_logit):
  return hinge_loss(real_logit, fake_logit)[0]


def hinge_loss_g(fake_logit):
  f = torch.as_tensor(fake_logit)
  return hinge_loss(torch.zeros_like(f), f)[1]


def symmetric_identity_cross_entropy(logits):
  """mean_i CE(row i, label i) + mean_j CE(col j, label j): the combination in which the reference uses
  tf_cross_entropy_loss_with_logits with one-hot identity labels (attention_lib.py:66-74,173-181)."""
  logits = torch.as_tensor(logits).to("cuda", torch.float32).contiguous()
  out = ops.empty(1, ops.F32)
  ops.ce_sym(logits, out, want_grad=False)
  return out[0]


def tf_cross_entropy_loss_with_logits(labels, logits):
  """losses.tf_cross_entropy_loss_with_logits (losses.py:47-51): -sum(labels * log_softmax(logits), axis=-1), one
  value per row, for arbitrary (soft or one-hot) labels."""
  labels = torch.as_tensor(labels).to("cuda", torch.float32).contiguous()
  logits = torch.as_tensor(logits).to("cuda", torch.float32).contiguous()
  if labels.shape != logits.shape:
    raise ValueError("labels and logits must have the same shape")
  n = logits.shape[-1]
  rows = logits.numel() // n
  out = ops.empty(rows, ops.F32)
  ops._call("xmc_softmax_xent", labels.data_ptr(), logits.data_ptr(), rows, n, out.data_ptr(), ops.stream())
  return out.view(logits.shape[:-1])
