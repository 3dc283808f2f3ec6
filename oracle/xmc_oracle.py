"""CPU oracle for XMC-GAN's train_step hot path — TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU (fp32) restatement of the reference algorithm. It is the checker used by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`; nothing in the product
package (`xmcgan_image_generation_b200/`) imports it and it must never be used as a fallback compute path.

PARITY, two tiers.
(a) PINNED against the reference's own code. Forward: tests/golden/make_reference_golden.py executes the reference's
losses.py, attention_lib.py, nets/xmc_net.py, nets/common.py, libml/layers.py and utils/resnet_v1.py from
/root/reference on numpy stand-ins for the `jax` entry points and the slice of `flax.linen` they use
(tests/golden/flax_stand_in.py); the results are committed as tests/golden/reference_libml.npz / reference_nets.npz
and tests/test_reference_golden.py holds this file to them — losses and attention at 2e-6, generator_apply /
discriminator_apply (train and inference mode, all statistics, state updates) and resnet50_apply at 2e-5 / 5e-5.
Backward (torch autograd over that forward): 19 directional derivatives of d_loss / g_loss agree to 1.4e-4 with
central differences of the reference's own loss_fn (taken out of xmcgan/xmc_gan.py, executed on the stand-ins in
float64, jax.lax.stop_gradient honoured by replaying the base run's stopped values).
(b) PARITY UNPINNED BY UPSTREAM is what is not array code of the reference: flax.optim.Adam.apply_gradient (restated
from flax 0.3.3), the collectives, jax.image.resize — the reference (JAX/Flax, not installable here) ships no golden
vectors or numeric tests for this path. Those are pinned by (1) analytic known-answer tests in
tests/test_oracle.py and (2) line-by-line citations below.

Every function cites the reference file:line it follows (paths relative to /root/reference).

Precision policies: `Policy("float32")` computes everything in fp32 (the reference's `config.dtype="float32"` path).
`Policy("bfloat16")` rounds activations and GEMM/conv weights to bf16 at the points where the reference's
`config.dtype="bfloat16"` path stores bf16 tensors (conv/dense outputs, BatchNorm outputs), keeping fp32
accumulation, fp32 statistics and fp32 losses; rounding uses a straight-through gradient. Where the B200 pipeline
keeps MORE precision than the reference's bf16 graph (fp32 logit head, fp32 pre-tanh accumulator) the bf16 policy
follows the pipeline; those spots are marked "B200 path" below.
"""
import math

import torch
import torch.nn.functional as F

LARGE_NUM = 1e9  # xmcgan/libml/attention_lib.py:20


# ----------------------------------------------------------------------------------------------------------------------
# precision policy
# ----------------------------------------------------------------------------------------------------------------------
class _RoundBoth(torch.autograd.Function):
  """bf16 rounding of the value in forward AND of the cotangent in backward (models bf16 storage of activation
  gradients, which both the reference's bf16 graph and the B200 pipeline do)."""

  @staticmethod
  def forward(ctx, x):
    return x.to(torch.bfloat16).to(torch.float32)

  @staticmethod
  def backward(ctx, g):
    return g.to(torch.bfloat16).to(torch.float32)


class Policy:
  def __init__(self, dtype="float32", round_grads=False):
    assert dtype in ("float32", "bfloat16")
    self.dtype = dtype
    self.round_grads = round_grads

  def q(self, x):
    """Storage rounding of an activation / cast of a weight to the compute dtype (straight-through gradient).
    With round_grads=True the cotangent is rounded to bf16 at the same points (noise model of bf16 gradient storage)."""
    if self.dtype == "float32":
      return x
    if self.round_grads and x.requires_grad:
      return _RoundBoth.apply(x)
    return x + (x.detach().to(torch.bfloat16).to(torch.float32) - x.detach())


FP32 = Policy("float32")


# ----------------------------------------------------------------------------------------------------------------------
# losses.py
# ----------------------------------------------------------------------------------------------------------------------
def hinge_loss(real_logit, fake_logit):
  """xmcgan/libml/losses.py:30-35."""
  generator_loss = -torch.mean(fake_logit)
  real_loss = F.relu(1.0 - real_logit)
  fake_loss = F.relu(1.0 + fake_logit)
  discriminator_loss = torch.mean(real_loss + fake_loss)
  return discriminator_loss, generator_loss


def tf_cross_entropy_loss_with_logits(labels, logits):
  """xmcgan/libml/losses.py:47-51."""
  logp = F.log_softmax(logits, dim=-1)
  return -torch.sum(labels * logp, dim=-1)


# ----------------------------------------------------------------------------------------------------------------------
# attention_lib.py
# ----------------------------------------------------------------------------------------------------------------------
def cosine_similarity(x1, x2):
  """attention_lib.py:23-27 (no epsilon)."""
  dist = torch.sum(x1 * x2, -1)
  return dist / (torch.linalg.norm(x1, dim=-1) * torch.linalg.norm(x2, dim=-1))


def l2_normalize(x, axis=-1, epsilon=1e-12):
  """attention_lib.py:30-33."""
  square_sum = torch.sum(x * x, dim=axis, keepdim=True)
  return x * torch.rsqrt(torch.clamp(square_sum, min=epsilon))


def get_statistics(logits, labels):
  """attention_lib.py:36-43."""
  prob = F.softmax(logits, dim=-1)
  entropy = -torch.mean(torch.sum(prob * torch.log(prob + 1e-8), dim=-1))
  label_acc = (torch.argmax(logits, dim=-1) == torch.argmax(labels, dim=-1)).float().mean()
  return label_acc, entropy


def contrastive_loss(image_feat, cond_feat, l2_norm=True, temperature=0.1, sync_match=False):
  """attention_lib.py:46-79. Negatives are the local (per-replica) batch; sync_match raises as in the reference."""
  if l2_norm:
    image_feat = l2_normalize(image_feat, -1)
    cond_feat = l2_normalize(cond_feat, -1)
  local_batch_size = image_feat.shape[0]
  if sync_match:
    raise NotImplementedError
  labels = torch.eye(local_batch_size, dtype=image_feat.dtype)
  logits_img2cond = image_feat @ cond_feat.t() / temperature
  logits_cond2img = cond_feat @ image_feat.t() / temperature
  loss_img2cond = tf_cross_entropy_loss_with_logits(labels, logits_img2cond).mean()
  loss_cond2img = tf_cross_entropy_loss_with_logits(labels, logits_cond2img).mean()
  loss = loss_img2cond + loss_cond2img
  accuracy1, entropy1 = get_statistics(logits_img2cond, labels)
  accuracy2, entropy2 = get_statistics(logits_cond2img, labels)
  return loss, 0.5 * (accuracy1 + accuracy2), 0.5 * (entropy1 + entropy2)


def attention(region_feat, word_feat, gamma, mask=None):
  """attention_lib.py:105-127: softmax over REGIONS (axis -2); context uses the normalised regions."""
  region_feat = l2_normalize(region_feat, -1)
  word_feat = l2_normalize(word_feat, -1)
  attn_matrix = region_feat @ word_feat.transpose(1, 2)
  attn_matrix = attn_matrix * gamma
  if mask is not None:
    attn_matrix = attn_matrix + mask * (-1e9)
  alpha = F.softmax(attn_matrix, dim=-2)
  region_context = alpha.transpose(1, 2) @ region_feat
  return region_context


def word_loss(image_feat, word_feat, max_len, gamma1=5, gamma2=5, gamma3=50):
  """attention_lib.py:130-191. image_feat [B,R,D], word_feat [B,L,D], max_len [B,1]."""
  batch_size, region_num, _ = image_feat.shape
  total_len = word_feat.shape[1]

  def my_func(max_len_i, word_feat_i):
    word_feat_i = word_feat_i[None].repeat(batch_size, 1, 1)
    max_len_i = max_len_i.repeat(region_num)
    mask = (torch.arange(total_len, dtype=torch.float32)[None, :] >= max_len_i[:, None]).float()
    mask = mask[None].repeat(batch_size, 1, 1)
    mask_2 = mask[:, 0, :]
    region_context = attention(image_feat, word_feat_i, gamma1, mask)
    row_sim = cosine_similarity(word_feat_i, region_context)
    row_sim = row_sim * gamma2
    row_sim = row_sim + mask_2 * (-1e9)
    row_sim = torch.logsumexp(row_sim, dim=-1, keepdim=True)
    return row_sim / gamma2

  similarities = torch.stack([my_func(max_len[j], word_feat[j]) for j in range(batch_size)])  # jax.vmap, :169
  similarities = similarities * gamma3
  similarities = similarities.reshape(batch_size, batch_size)  # jnp.squeeze of [B,B,1]
  similarities_transpose = similarities
  similarities = similarities_transpose.t()
  labels = torch.eye(batch_size)
  loss_0 = tf_cross_entropy_loss_with_logits(labels, similarities).mean()
  loss_1 = tf_cross_entropy_loss_with_logits(labels, similarities_transpose).mean()
  matching_loss = loss_0 + loss_1
  accuracy1, entropy1 = get_statistics(similarities, labels)
  accuracy2, entropy2 = get_statistics(similarities_transpose, labels)
  return matching_loss, 0.5 * (accuracy1 + accuracy2), 0.5 * (entropy1 + entropy2)


def attention_for_g(region_feat, word_feat, gamma, mask=None):
  """attention_lib.py:194-219: softmax over WORDS (last axis); context uses the normalised words."""
  region_feat = l2_normalize(region_feat, -1)
  word_feat = l2_normalize(word_feat, -1)
  attn_matrix = region_feat @ word_feat.transpose(1, 2)
  attn_matrix = attn_matrix * gamma
  if mask is not None:
    attn_matrix = attn_matrix + mask * (-1e9)
  attn = F.softmax(attn_matrix, dim=-1)
  region_context = attn @ word_feat
  return region_context, attn


# ----------------------------------------------------------------------------------------------------------------------
# flax.linen primitives (flax==0.3.3, not vendored) restated
# ----------------------------------------------------------------------------------------------------------------------
def conv2d(x, kernel, bias, policy=FP32, round_out=True):
  """flax nn.Conv, stride 1, padding SAME: x NHWC, kernel HWIO, both cast to dtype; output stored in dtype."""
  kh, kw = kernel.shape[0], kernel.shape[1]
  w = policy.q(kernel).permute(3, 2, 0, 1)
  y = F.conv2d(policy.q(x).permute(0, 3, 1, 2), w, padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1)
  if bias is not None:
    y = y + bias
  return policy.q(y) if round_out else y


def dense(x, kernel, bias, policy=FP32, round_out=True):
  """flax nn.Dense: kernel [in,out]."""
  y = policy.q(x) @ policy.q(kernel)
  if bias is not None:
    y = y + bias
  return policy.q(y) if round_out else y


def batch_norm(x, stats, train, momentum=0.9, epsilon=1e-5, policy=FP32):
  """flax nn.BatchNorm(use_scale=False, use_bias=False): fp32 statistics over (N,H,W), biased var = E[x^2]-E[x]^2.
  Returns (y, new_stats)."""
  x32 = x.float()
  if train:
    mean = x32.mean(dim=(0, 1, 2))
    mean2 = (x32 * x32).mean(dim=(0, 1, 2))
    var = mean2 - mean * mean
    new_stats = {
        "mean": momentum * stats["mean"] + (1 - momentum) * mean.detach(),
        "var": momentum * stats["var"] + (1 - momentum) * var.detach(),
    }
  else:
    mean, var = stats["mean"], stats["var"]
    new_stats = stats
  y = (x32 - mean) * torch.rsqrt(var + epsilon)
  return y, new_stats


# ----------------------------------------------------------------------------------------------------------------------
# layers.py
# ----------------------------------------------------------------------------------------------------------------------
def _l2_normalize_sn(x, eps):
  """layers.py:31-46 as used at :96-97 with axis=None: x * rsqrt(sum(x^2) + eps)."""
  return x * torch.rsqrt((x * x).sum() + eps)


def spectral_normalize(kernel, u0, eps=1e-10):
  """layers.py:94-101 / :211-221. kernel [..., out]; u0 [1,out]. Returns (kernel/(sigma+eps), new_u0)."""
  shape = kernel.shape
  w = kernel.reshape(-1, shape[-1])
  with torch.no_grad():
    v0 = _l2_normalize_sn(u0 @ w.t(), eps)
    u1 = _l2_normalize_sn(v0 @ w, eps)
  sigma = (v0 @ w @ u1.t())[0, 0]
  return (w / (sigma + eps)).reshape(shape), u1


def spectral_conv(x, p, st, train, policy=FP32, round_out=True):
  """layers.SpectralConv (layers.py:125-241). p = {kernel,bias}, st = {u0}. Returns (y, new_st)."""
  k, u1 = spectral_normalize(p["kernel"], st["u0"])
  y = conv2d(x, k, p.get("bias"), policy, round_out)
  return y, ({"u0": u1} if train else st)


def spectral_dense(x, p, st, train, policy=FP32, round_out=True):
  """layers.SpectralDense (layers.py:49-113)."""
  k, u1 = spectral_normalize(p["kernel"], st["u0"])
  y = dense(x, k, p.get("bias"), policy, round_out)
  return y, ({"u0": u1} if train else st)


class _Scope:
  """Minimal stand-in for a Flax module scope: reads params/collections by name, records mutated collections."""

  def __init__(self, variables, path=()):
    self.variables = variables
    self.path = path
    self.updates = {}

  def child(self, name):
    c = _Scope(self.variables, self.path + (name,))
    c.updates = self.updates
    return c

  def get(self, col):
    node = self.variables.get(col, {})
    for k in self.path:
      node = node.get(k, {}) if isinstance(node, dict) else {}
    return node

  def put(self, col, value):
    node = self.updates.setdefault(col, {})
    for k in self.path[:-1]:
      node = node.setdefault(k, {})
    node[self.path[-1]] = value


def _conv_fn(scope, name, x, spectral, train, policy, round_out=True):
  """conv_fn(...)(x) of xmc_net.py:66-80 / 176-191: SpectralConv or nn.Conv under Flax auto-name `name`."""
  s = scope.child(name)
  p = s.get("params")
  if spectral:
    y, st = spectral_conv(x, p, s.get("spectral_norm_stats"), train, policy, round_out)
    s.put("spectral_norm_stats", st)
    return y
  return conv2d(x, p["kernel"], p.get("bias"), policy, round_out)


def _dense_fn(scope, name, x, spectral, train, policy, round_out=True):
  s = scope.child(name)
  p = s.get("params")
  if spectral:
    y, st = spectral_dense(x, p, s.get("spectral_norm_stats"), train, policy, round_out)
    s.put("spectral_norm_stats", st)
    return y
  return dense(x, p["kernel"], p.get("bias"), policy, round_out)


def _bn(scope, x, train, policy):
  s = scope.child("BatchNorm_0")
  y, new_stats = batch_norm(x, s.get("batch_stats"), train, policy=policy)
  s.put("batch_stats", new_stats)
  return y


def conditional_batch_norm(scope, x, emb, spectral, train, policy):
  """layers.ConditionalBatchNorm (layers.py:244-258)."""
  filters = x.shape[-1]
  prefix = "SpectralDense" if spectral else "Dense"
  gamma = _dense_fn(scope, prefix + "_0", emb, spectral, train, policy).reshape(-1, 1, 1, filters)
  beta = _dense_fn(scope, prefix + "_1", emb, spectral, train, policy).reshape(-1, 1, 1, filters)
  x = _bn(scope, x, train, policy)
  return policy.q(x * (gamma + 1.0) + beta)


def local_conditional_batch_norm(scope, x, emb, spectral, train, policy):
  """layers.LocalConditionalBatchNorm (layers.py:261-273): per-pixel gamma/beta from 1x1 convs of emb."""
  prefix = "SpectralConv" if spectral else "Conv"
  gamma = _conv_fn(scope, prefix + "_0", emb, spectral, train, policy)
  beta = _conv_fn(scope, prefix + "_1", emb, spectral, train, policy)
  x = _bn(scope, x, train, policy)
  return policy.q(x * (gamma + 1.0) + beta)


# ----------------------------------------------------------------------------------------------------------------------
# nets/common.py
# ----------------------------------------------------------------------------------------------------------------------
def upsample(x, factor=2):
  """common.py:48-51 (jax.image.resize nearest: out[i] = in[i // factor])."""
  return x.repeat_interleave(factor, dim=1).repeat_interleave(factor, dim=2)


def dsample(x):
  """common.py:23-45,54-55: 2x2/2 mean (TF-style; denominator is exactly 4 for even sizes)."""
  n, h, w, c = x.shape
  return x.reshape(n, h // 2, 2, w // 2, 2, c).mean(dim=(2, 4))


def disc_block(scope, x, filters, downsample, spectral, train, policy):
  """common.DiscBlock (common.py:58-79)."""
  pre = "SpectralConv" if spectral else "Conv"
  needs_projection = downsample or x.shape[-1] != filters
  x0 = x
  x = F.relu(x)
  x = _conv_fn(scope, pre + "_0", x, spectral, train, policy)
  x = F.relu(x)
  x = _conv_fn(scope, pre + "_1", x, spectral, train, policy)
  if needs_projection:
    x0 = _conv_fn(scope, pre + "_2", x0, spectral, train, policy)
  if downsample:
    x = policy.q(dsample(x))
    x0 = policy.q(dsample(x0))
  return policy.q(x0 + x)


def disc_optimized_block(scope, x, spectral, train, policy):
  """common.DiscOptimizedBlock (common.py:117-133)."""
  pre = "SpectralConv" if spectral else "Conv"
  x0 = x
  x = _conv_fn(scope, pre + "_0", x, spectral, train, policy)
  x = F.relu(x)
  x = _conv_fn(scope, pre + "_1", x, spectral, train, policy)
  x = policy.q(dsample(x))
  x0 = policy.q(dsample(x0))
  x0 = _conv_fn(scope, pre + "_2", x0, spectral, train, policy)
  return policy.q(x + x0)


def gen_block(scope, x, cond, spectral, train, policy):
  """common.GenBlock (common.py:136-160)."""
  pre = "SpectralConv" if spectral else "Conv"
  x0 = x
  x = conditional_batch_norm(scope.child("ConditionalBatchNorm_0"), x, cond, spectral, train, policy)
  x = F.relu(x)
  x = upsample(x)
  x = _conv_fn(scope, pre + "_0", x, spectral, train, policy)
  x = conditional_batch_norm(scope.child("ConditionalBatchNorm_1"), x, cond, spectral, train, policy)
  x = F.relu(x)
  x = _conv_fn(scope, pre + "_1", x, spectral, train, policy)
  x0 = upsample(x0)
  x0 = _conv_fn(scope, pre + "_2", x0, spectral, train, policy)
  return policy.q(x + x0)


def gen_spatial_block(scope, x, cond0, cond1, spectral, train, policy):
  """common.GenSpatialBlock (common.py:163-186)."""
  pre = "SpectralConv" if spectral else "Conv"
  x0 = x
  x = local_conditional_batch_norm(scope.child("LocalConditionalBatchNorm_0"), x, cond0, spectral, train, policy)
  x = F.relu(x)
  x = upsample(x)
  x = _conv_fn(scope, pre + "_0", x, spectral, train, policy)
  x = local_conditional_batch_norm(scope.child("LocalConditionalBatchNorm_1"), x, cond1, spectral, train, policy)
  x = F.relu(x)
  x = _conv_fn(scope, pre + "_1", x, spectral, train, policy)
  x0 = upsample(x0)
  x0 = _conv_fn(scope, pre + "_2", x0, spectral, train, policy)
  return policy.q(x + x0)


# ----------------------------------------------------------------------------------------------------------------------
# nets/xmc_net.py
# ----------------------------------------------------------------------------------------------------------------------
def _channel_dims_g(image_size):
  if image_size == 256:
    return [16, 8, 8, 4, 2, 1]
  if image_size == 128:
    return [16, 8, 4, 2, 1]
  raise ValueError(f"image_size {image_size} not supported")  # reference: NameError (xmc_net.py:202-205)


def _channel_dims_d(image_size):
  if image_size == 128:
    return [2, 4, 8, 16, 16], [True, True, True, True, False]
  if image_size == 256:
    return [2, 4, 8, 8, 16, 16], [True, True, True, True, True, False]
  raise ValueError(f"image_size {image_size} not supported")  # xmc_net.py:81-86


def generator_apply(variables, inputs, config, train, policy=FP32):
  """xmc_net.Generator.__call__ (xmc_net.py:160-248). Returns (image in [0,1], updated collections)."""
  cond_dict, z = inputs
  cond = cond_dict["sentence_embedding"]
  word_feat = cond_dict["embedding"]
  max_len = cond_dict["max_len"]
  embedding_dim = word_feat.shape[-1]
  batch_size = z.shape[0]
  sn = bool(config.g_spectral_norm)
  if config.batch_norm_group_size > 0:
    raise NotImplementedError("cross-replica BatchNorm is not restated in the single-process oracle")
  dpre = "SpectralDense" if sn else "Dense"
  cpre = "SpectralConv" if sn else "Conv"
  scope = _Scope(variables)
  channel_dims = _channel_dims_g(config.image_size)
  gf = config.gf_dim
  global_cond = _dense_fn(scope, dpre + "_0", cond, sn, train, policy)
  global_cond = torch.cat([global_cond, policy.q(z)], dim=-1)
  x = _dense_fn(scope, dpre + "_1", z, sn, train, policy)
  x = x.reshape(-1, 4, 4, gf * 16)
  for i in range(2):
    x = gen_block(scope.child(f"GenBlock_{i}"), x, global_cond, sn, train, policy)
  x_cond = _conv_fn(scope, cpre + "_0", x, sn, train, policy)
  spatial_size = x_cond.shape[1]
  total_region_size = spatial_size * spatial_size
  total_len = word_feat.shape[1]
  x_cond = x_cond.reshape(batch_size, total_region_size, embedding_dim)
  mask = (torch.arange(total_len, dtype=torch.float32)[None, :] >= max_len).float()
  mask = mask[:, None, :].repeat(1, total_region_size, 1)
  region_context, _ = attention_for_g(x_cond, word_feat, config.gamma_for_g, mask)
  region_context = policy.q(region_context)
  region_context = region_context.reshape(batch_size, spatial_size, spatial_size, embedding_dim)
  spatial_cond = global_cond.reshape(batch_size, 1, 1, -1).repeat(1, spatial_size, spatial_size, 1)
  spatial_cond = torch.cat([region_context, spatial_cond], dim=-1)
  for i in range(2, len(channel_dims)):
    spatial_cond_upsample = upsample(spatial_cond)
    x = gen_spatial_block(scope.child(f"GenSpatialBlock_{i - 2}"), x, spatial_cond, spatial_cond_upsample, sn, train,
                          policy)
    spatial_cond = spatial_cond_upsample
  x = local_conditional_batch_norm(scope.child("LocalConditionalBatchNorm_0"), x, spatial_cond, sn, train, policy)
  x = F.relu(x)
  x = _conv_fn(scope, cpre + "_1", x, sn, train, policy, round_out=False)  # B200 path: fp32 accumulator -> tanh
  x = torch.tanh(x)
  x = (x + 1.0) / 2.0
  return x, scope.updates


def discriminator_apply(variables, inputs, config, train, policy=FP32):
  """xmc_net.Discriminator.__call__ (xmc_net.py:45-142). Returns ((logit, stat_dict), updated collections)."""
  x, cond_dict = inputs
  cond = cond_dict["sentence_embedding"]
  word_feat = cond_dict["embedding"]
  max_len = cond_dict["max_len"]
  sn = bool(config.d_spectral_norm)
  cpre = "SpectralConv" if sn else "Conv"
  dpre = "SpectralDense" if sn else "Dense"
  zero = torch.zeros(())
  stats = {k: zero for k in (
      "fake_word_loss", "fake_word_acc", "fake_word_entropy", "real_word_loss", "real_word_acc", "real_word_entropy",
      "fake_sentence_loss", "fake_sentence_acc", "fake_sentence_entropy", "real_sentence_loss", "real_sentence_acc",
      "real_sentence_entropy", "image_contrastive_loss", "image_contrastive_acc", "image_contrastive_entropy")}
  channel_dims, downsamples = _channel_dims_d(config.image_size)
  df = config.df_dim
  scope = _Scope(variables)
  x = policy.q(x)
  x = disc_optimized_block(scope.child("DiscOptimizedBlock_0"), x, sn, train, policy)
  x_cond = None
  for i, (c_ratio, downsample) in enumerate(zip(channel_dims, downsamples)):
    x = disc_block(scope.child(f"DiscBlock_{i}"), x, df * c_ratio, downsample, sn, train, policy)
    if x.shape[1] == config.cond_size:
      x_cond = x
  x = F.relu(x)
  x_pool = torch.sum(x.float(), dim=(1, 2))
  # B200 path: the logit head stays in fp32 (more accurate than the reference's bf16 dense, never less)
  out = _dense_fn(scope, dpre + "_0", x_pool, sn, train, FP32)
  embedding = _dense_fn(scope, dpre + "_1", cond, sn, train, policy, round_out=False)
  sent_cond = embedding
  tile_num = x_pool.shape[0] // embedding.shape[0]
  embedding = embedding.repeat(tile_num, 1)
  out = out + torch.sum(x_pool * embedding, dim=1, keepdim=True)
  half = x_pool.shape[0] // 2
  if config.sentence_contrastive:
    real_feat, fake_feat = x_pool[:half], x_pool[half:]  # real images are the first half (xmc_net.py:106-107)
    (stats["fake_sentence_loss"], stats["fake_sentence_acc"],
     stats["fake_sentence_entropy"]) = contrastive_loss(fake_feat, sent_cond)
    (stats["real_sentence_loss"], stats["real_sentence_acc"],
     stats["real_sentence_entropy"]) = contrastive_loss(real_feat, sent_cond)
  if config.word_contrastive:
    embedding_dim = word_feat.shape[-1]
    x_cond = _conv_fn(scope, cpre + "_0", x_cond, sn, train, policy)
    total_region_size = config.cond_size * config.cond_size
    x_cond_reshape = x_cond.reshape(-1, total_region_size, embedding_dim)
    real_x_cond, fake_x_cond = x_cond_reshape[:half], x_cond_reshape[half:]
    (stats["fake_word_loss"], stats["fake_word_acc"],
     stats["fake_word_entropy"]) = word_loss(fake_x_cond, word_feat, max_len)
    (stats["real_word_loss"], stats["real_word_acc"],
     stats["real_word_entropy"]) = word_loss(real_x_cond, word_feat, max_len)
  if config.image_contrastive:
    real_feat, fake_feat = x_pool[:half], x_pool[half:]
    (stats["image_contrastive_loss"], stats["image_contrastive_acc"],
     stats["image_contrastive_entropy"]) = contrastive_loss(fake_feat, real_feat)
  return (out, stats), scope.updates


# ----------------------------------------------------------------------------------------------------------------------
# tree utilities, optimizer (flax.optim.Adam, flax==0.3.3)
# ----------------------------------------------------------------------------------------------------------------------
def tree_map(fn, *trees):
  t0 = trees[0]
  if isinstance(t0, dict):
    return {k: tree_map(fn, *[t[k] for t in trees]) for k in t0}
  return fn(*trees)


def tree_leaves(tree, prefix=()):
  if isinstance(tree, dict):
    out = []
    for k in sorted(tree):
      out += tree_leaves(tree[k], prefix + (k,))
    return out
  return [("/".join(prefix), tree)]


def merge_state(old, updates):
  """Flax `mutable=` semantics: returned collections replace the entries they contain."""
  if not isinstance(updates, dict):
    return updates
  out = dict(old) if isinstance(old, dict) else {}
  for k, v in updates.items():
    out[k] = merge_state(out.get(k, {}), v)
  return out


def adam_apply(params, opt_state, grads, lr, beta1, beta2, eps=1e-8):
  """flax.optim.Adam.apply_gradient (weight_decay=0). opt_state = {"step": int, "m": tree, "v": tree}."""
  t = opt_state["step"] + 1
  m = tree_map(lambda m_, g: beta1 * m_ + (1.0 - beta1) * g, opt_state["m"], grads)
  v = tree_map(lambda v_, g: beta2 * v_ + (1.0 - beta2) * g * g, opt_state["v"], grads)
  c1 = 1.0 - beta1 ** t
  c2 = 1.0 - beta2 ** t
  new_params = tree_map(lambda p, m_, v_: p - lr * (m_ / c1) / (torch.sqrt(v_ / c2) + eps), params, m, v)
  return new_params, {"step": t, "m": m, "v": v}


def adam_init(params):
  return {"step": 0, "m": tree_map(torch.zeros_like, params), "v": tree_map(torch.zeros_like, params)}


# ----------------------------------------------------------------------------------------------------------------------
# xmc_gan.py / train_utils.py
# ----------------------------------------------------------------------------------------------------------------------
def split_input_dict(input_dict, splits, axis=0):
  """train_utils.py:69-88."""
  out = [dict() for _ in range(splits)]
  for k, v in input_dict.items():
    for i, part in enumerate(torch.chunk(v, splits, dim=axis)):
      out[i][k] = part
  return out


def calculate_contrastive_loss(result_dict):
  """xmc_gan.py:58-71."""
  real_loss = result_dict["real_word_loss"] + result_dict["real_sentence_loss"]
  fake_loss = result_dict["fake_word_loss"] + result_dict["fake_sentence_loss"]
  return real_loss, fake_loss + result_dict["image_contrastive_loss"]


def _with_grad(params):
  return tree_map(lambda p: p.detach().clone().requires_grad_(True), params)


def _grads_of(loss, params, retain=False):
  leaves = [p for _, p in tree_leaves(params)]
  gs = torch.autograd.grad(loss, leaves, retain_graph=retain, allow_unused=True)
  it = iter(gs)
  names = [n for n, _ in tree_leaves(params)]
  gmap = {n: (g if g is not None else torch.zeros_like(p)) for n, p, g in zip(names, leaves, it)}

  def rebuild(tree, prefix=()):
    if isinstance(tree, dict):
      return {k: rebuild(tree[k], prefix + (k,)) for k in tree}
    return gmap["/".join(prefix)]

  return rebuild(params)


def d_losses_and_grads(state, batch, config, policy=FP32, pretrained_fn=None, want_g=True):
  """Forward of train_g_d's loss_fn (xmc_gan.py:127-160) and both pull-backs (:162-167) at the current params.
  Returns dict with losses, grads and the new mutable collections. With want_g=False this is train_d's loss_fn
  (xmc_gan.py:220-245): only d_loss = hinge_d + real_word + real_sentence and its gradient."""
  params_d = _with_grad(state["d_params"])
  params_g = _with_grad(state["g_params"])
  g_vars = dict(state["generator_state"], params=params_g)
  d_vars = dict(state["discriminator_state"], params=params_d)
  z = batch["z"]
  fake, new_g = generator_apply(g_vars, (batch, z), config, True, policy)
  all_images = torch.cat([batch["image"], fake], dim=0)
  (logit, result), new_d = discriminator_apply(d_vars, (all_images, batch), config, True, policy)
  logit = logit.float()
  half = logit.shape[0] // 2
  real_logit, fake_logit = logit[:half], logit[half:]
  d_hinge, g_hinge = hinge_loss(real_logit, fake_logit)
  c_loss_d, c_loss_g = calculate_contrastive_loss(result)
  out = {"fake": fake.detach(), "logit": logit.detach(), "result": {k: v.detach() for k, v in result.items()},
         "new_generator_state": tree_map(lambda t: t.detach(), merge_state(state["generator_state"], new_g)),
         "new_discriminator_state": tree_map(lambda t: t.detach(), merge_state(state["discriminator_state"], new_d))}
  d_loss = d_hinge + c_loss_d
  out["d_loss"] = d_loss.detach()
  out["c_loss_d"] = torch.as_tensor(c_loss_d).detach()
  if not want_g:
    out["d_grad"] = _grads_of(d_loss, params_d)
    return out
  c_loss_g_pretrained = torch.zeros(())
  if config.pretrained_image_contrastive:
    if pretrained_fn is None:
      raise ValueError("pretrained_image_contrastive=True needs pretrained_fn")
    c_loss_g_pretrained = pretrained_fn(batch["image"], fake)
  g_loss = g_hinge + c_loss_g + c_loss_g_pretrained
  out["g_loss"] = g_loss.detach()
  out["c_loss_g"] = torch.as_tensor(c_loss_g).detach()
  out["c_loss_g_pretrained"] = c_loss_g_pretrained.detach()
  out["d_grad"] = _grads_of(d_loss, params_d, retain=True)
  out["g_grad"] = _grads_of(g_loss, params_g)
  return out


def train_d(state, batch, config, policy=FP32):
  """xmc_gan.train_d (xmc_gan.py:194-256), single replica (pmean over one device is the identity)."""
  r = d_losses_and_grads(state, batch, config, policy, want_g=False)
  new_params, new_opt = adam_apply(tree_map(lambda t: t.detach(), state["d_params"]), state["d_opt"], r["d_grad"],
                                   config.d_lr, config.beta1, config.beta2)
  new_state = dict(state)
  new_state["d_params"] = new_params
  new_state["d_opt"] = new_opt
  new_state["discriminator_state"] = r["new_discriminator_state"]
  return new_state, r


def train_g_d(state, batch, config, policy=FP32, pretrained_fn=None):
  """xmc_gan.train_g_d (xmc_gan.py:93-191)."""
  r = d_losses_and_grads(state, batch, config, policy, pretrained_fn=pretrained_fn, want_g=True)
  new_d, new_d_opt = adam_apply(tree_map(lambda t: t.detach(), state["d_params"]), state["d_opt"], r["d_grad"],
                                config.d_lr, config.beta1, config.beta2)
  new_g, new_g_opt = adam_apply(tree_map(lambda t: t.detach(), state["g_params"]), state["g_opt"], r["g_grad"],
                                config.g_lr, config.beta1, config.beta2)
  decay = config.polyak_decay
  new_ema = tree_map(lambda e, p: e * decay + (1 - decay) * p, state["ema_params"], new_g)
  new_state = dict(state)
  new_state.update(step=state["step"] + 1, d_params=new_d, d_opt=new_d_opt, g_params=new_g, g_opt=new_g_opt,
                   generator_state=r["new_generator_state"], discriminator_state=r["new_discriminator_state"],
                   ema_params=new_ema)
  metrics = {k: float(r[k]) for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g", "c_loss_g_pretrained")}
  return new_state, metrics, r


def train_step(state, batch, config, policy=FP32, pretrained_fn=None):
  """train_utils.train_step (train_utils.py:91-130)."""
  batches = split_input_dict(batch, config.d_step_per_g_step)
  for i in range(config.d_step_per_g_step - 1):
    state, _ = train_d(state, batches[i], config, policy)
  new_state, metrics, _ = train_g_d(state, batches[-1], config, policy, pretrained_fn)
  return new_state, metrics


def make_state(g_vars, d_vars):
  """create_train_state (train_utils.py:133-193) from already initialised variable collections."""
  g_vars = dict(g_vars)
  d_vars = dict(d_vars)
  g_params = g_vars.pop("params")
  d_params = d_vars.pop("params")
  return {
      "step": 0,
      "g_params": g_params, "g_opt": adam_init(g_params),
      "d_params": d_params, "d_opt": adam_init(d_params),
      "generator_state": g_vars, "discriminator_state": d_vars,
      "ema_params": tree_map(lambda t: t.clone(), g_params),
  }


# ----------------------------------------------------------------------------------------------------------------------
# frozen ResNet-50 feature branch: utils/resnet_v1.py, utils/pretrained_model_utils.py, xmc_gan.py:74-90
# ----------------------------------------------------------------------------------------------------------------------
RESNET50_STAGES = [3, 4, 6, 3]  # resnet_v1.py:179-180


def _same_pads(n, k, s):
  """XLA 'SAME': total = max((ceil(n/s)-1)*s + k - n, 0), low = total // 2, high = total - low."""
  out = -(-n // s)
  total = max((out - 1) * s + k - n, 0)
  return total // 2, total - total // 2


def _conv_same(x, kernel, stride, policy, scale=None):
  """flax nn.Conv(use_bias=False, strides, padding='SAME') on NHWC / HWIO. `scale`: per-output-channel factor folded
  into the kernel before the cast (B200 path: eval BatchNorm folded into the bf16 weights)."""
  kh, kw = kernel.shape[0], kernel.shape[1]
  w = kernel if scale is None else kernel * scale
  pt, pb = _same_pads(x.shape[1], kh, stride)
  pl, pr = _same_pads(x.shape[2], kw, stride)
  xp = F.pad(policy.q(x).permute(0, 3, 1, 2), (pl, pr, pt, pb))
  return F.conv2d(xp, policy.q(w).permute(3, 2, 0, 1), stride=stride).permute(0, 2, 3, 1)


def _bn_eval_affine(p, st, eps=1e-5):
  """flax nn.BatchNorm(use_running_average=True) as y = x*s + b."""
  s = p["scale"] * torch.rsqrt(st["var"] + eps)
  return s, p["bias"] - st["mean"] * s


def _conv_bn(x, params, stats, conv, bn, stride, policy, relu, residual=None):
  s, b = _bn_eval_affine(params[bn], stats[bn])
  if policy.dtype == "float32":
    y = _conv_same(x, params[conv]["kernel"], stride, policy) * s + b
  else:
    y = _conv_same(x, params[conv]["kernel"], stride, policy, scale=s) + b
  if residual is not None:
    y = y + residual
  if relu:
    y = F.relu(y)
  return policy.q(y)


def bottleneck_block(x, p, s, stride, policy=FP32):
  """resnet_v1.BottleneckResNetBlock (resnet_v1.py:60-86): 1x1 -> 3x3 (strided) -> 1x1, projection shortcut when the
  shape changes, relu(residual + x)."""
  residual = x
  y = _conv_bn(x, p, s, "conv1", "bn1", 1, policy, relu=True)
  y = _conv_bn(y, p, s, "conv2", "bn2", stride, policy, relu=True)
  if "proj_conv" in p:
    residual = _conv_bn(x, p, s, "proj_conv", "proj_bn", stride, policy, relu=False)
  return _conv_bn(y, p, s, "conv3", "bn3", 1, policy, relu=True, residual=residual)


def resnet_stem(variables, x, policy=FP32):
  """init_conv 7x7/2 + init_bn (no ReLU) + max_pool 3x3/2 SAME (resnet_v1.py:146-154). Returns (stem, pooled)."""
  P, S = variables["params"], variables["batch_stats"]
  stem = _conv_bn(x, P, S, "init_conv", "init_bn", 2, policy, relu=False)
  pt, pb = _same_pads(stem.shape[1], 3, 2)
  xp = F.pad(stem.permute(0, 3, 1, 2), (pt, pb, pt, pb), value=float("-inf"))
  return stem, F.max_pool2d(xp, 3, 2).permute(0, 2, 3, 1)


def resnet50_apply(variables, x, policy=FP32):
  """resnet_v1.ResNet.__call__ (resnet_v1.py:129-172), train=False, BottleneckResNetBlock (:60-86).
  Note: no ReLU after init_bn (:146-154); the stride sits on the 3x3 conv (:79). Returns (pool, logits)."""
  P, S = variables["params"], variables["batch_stats"]
  _, x = resnet_stem(variables, x, policy)
  for si, nblocks in enumerate(RESNET50_STAGES):
    for bi in range(nblocks):
      p, s = P[f"stage{si + 1}"][f"block{bi + 1}"], S[f"stage{si + 1}"][f"block{bi + 1}"]
      x = bottleneck_block(x, p, s, 2 if (si > 0 and bi == 0) else 1, policy)
  pool = x
  feat = pool.float().mean(dim=(1, 2))
  logits = policy.q(feat) @ policy.q(P["head"]["kernel"]) + P["head"]["bias"]
  return pool, logits


def get_pretrained_embs(variables, images, policy=FP32, size=224):
  """pretrained_model_utils.get_pretrained_embs (pretrained_model_utils.py:102-127): jax.image.resize "bilinear" to 224
  (half-pixel centres; identical to torch align_corners=False when up-sampling; when down-sampling — the 256 px
  configuration — jax widens the triangle kernel by the scale factor and renormalises, which is torch's
  antialias=True), then the frozen network."""
  if images.dim() != 4 or images.shape[3] != 3:
    raise ValueError("images should be of shape (H, W, 3).")
  if images.shape[1] != size and images.shape[2] != size:
    images = F.interpolate(images.permute(0, 3, 1, 2), size=(size, size), mode="bilinear", align_corners=False,
                           antialias=images.shape[1] > size).permute(0, 2, 3, 1)
  return resnet50_apply(variables, images, policy)


def calculate_contrastive_loss_on_pretrained(variables, real_images, fake_images, policy=FP32):
  """xmc_gan.calculate_contrastive_loss_on_pretrained (xmc_gan.py:74-90)."""
  _, real_outputs = get_pretrained_embs(variables, real_images, policy)
  _, fake_outputs = get_pretrained_embs(variables, fake_images, policy)
  loss, _, _ = contrastive_loss(real_outputs, fake_outputs)
  return loss


def resnet50_param_shapes(num_classes=1000, width=64):
  """Variable tree shapes of resnet_v1.ResNet50 (names as in resnet_v1.py); 25 557 032 parameters."""
  params, stats = {}, {}

  def bn(c):
    return {"scale": (c,), "bias": (c,)}, {"mean": (c,), "var": (c,)}

  params["init_conv"] = {"kernel": (7, 7, 3, width)}
  params["init_bn"], stats["init_bn"] = bn(width)
  cin = width
  for si, nblocks in enumerate(RESNET50_STAGES):
    f = width * 2 ** si
    sp, ss = {}, {}
    for bi in range(nblocks):
      p, s = {}, {}
      p["conv1"] = {"kernel": (1, 1, cin, f)}
      p["bn1"], s["bn1"] = bn(f)
      p["conv2"] = {"kernel": (3, 3, f, f)}
      p["bn2"], s["bn2"] = bn(f)
      p["conv3"] = {"kernel": (1, 1, f, 4 * f)}
      p["bn3"], s["bn3"] = bn(4 * f)
      if cin != 4 * f or (si > 0 and bi == 0):
        p["proj_conv"] = {"kernel": (1, 1, cin, 4 * f)}
        p["proj_bn"], s["proj_bn"] = bn(4 * f)
      sp[f"block{bi + 1}"], ss[f"block{bi + 1}"] = p, s
      cin = 4 * f
    params[f"stage{si + 1}"], stats[f"stage{si + 1}"] = sp, ss
  params["head"] = {"kernel": (cin, num_classes), "bias": (num_classes,)}
  return params, stats


def resnet50_random_variables(seed=0, head_scale=0.05, residual_scale=0.3):
  """Synthetic frozen weights (the reference's data/resnet_pretrained.npy is not shipped, README.md:60-63): He-normal
  kernels, BatchNorm scale ~ 1, small random bias / mean, var ~ 1, and a NON-zero head (the reference initialises the
  head to zeros, resnet_v1.py:171, which would make the loss the constant 2 log B)."""
  g = torch.Generator().manual_seed(seed)
  pshapes, sshapes = resnet50_param_shapes()

  def fill(tree, path=()):
    out = {}
    for k, v in tree.items():
      if isinstance(v, dict):
        out[k] = fill(v, path + (k,))
      else:
        shape = v
        if k == "kernel" and len(shape) == 4:
          fan_in = shape[0] * shape[1] * shape[2]
          out[k] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif k == "kernel":
          out[k] = torch.randn(shape, generator=g) * head_scale
        elif k == "scale":
          out[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
          if path and path[-1] == "bn3":
            out[k] = out[k] * residual_scale  # keeps the residual stream O(1) so that features stay image-dependent
        elif k == "var":
          out[k] = 1.0 + 0.2 * torch.rand(shape, generator=g)
        else:  # bias, mean
          out[k] = 0.1 * torch.randn(shape, generator=g)
    return out

  return {"params": fill(pshapes), "batch_stats": fill(sshapes)}


# ----------------------------------------------------------------------------------------------------------------------
# input contract: COCODataset.preprocess (xmcgan/libml/coco_dataset.py:127-167) for decoded, already resized examples
# ----------------------------------------------------------------------------------------------------------------------
def preprocess_batch(features, flip, sentence_idx, z):
  """Deterministic part of coco_dataset.py:131-166 given the three random draws (flip [N] bool, sentence_idx [N], z).
  image: flip_left_right then clip_by_value(0, 1) (:135-136); sentence_feat = reduce_sum(embedding, -2) / max_len over
  ALL word slots (:139-142); the chosen caption's embedding / max_len / sentence_embedding (:154-159)."""
  img = features["image"].float()
  emb = features["caption/embedding"].float()
  lens = features["caption/max_len"].float()[..., None]          # [N, M, 1]
  img = torch.where(torch.as_tensor(flip).bool()[:, None, None, None], img.flip(2), img).clamp(0.0, 1.0)
  sent = emb.sum(dim=-2) / lens                                     # [N, M, E]
  n = torch.arange(img.shape[0])
  idx = torch.as_tensor(sentence_idx).long()
  return {"image": img, "embedding": emb[n, idx], "max_len": lens[n, idx], "sentence_embedding": sent[n, idx],
          "z": torch.as_tensor(z).float()}
