"""Counterparts of xmcgan/libml/attention_lib.py with the same signatures, on CUDA tensors. Forward values only;
train_step uses the same kernels plus their backward through engine.py. Every loss returns the reference's
(loss, accuracy, entropy) triple (attention_lib.py:75-78,183-190); train_step itself never reads accuracy / entropy
(XLA removes them there) and skips them."""
import torch

from .. import engine as _engine
from .. import ops

LARGE_NUM = 1e9


def _dev(x):
  return torch.as_tensor(x).to("cuda", torch.float32).contiguous()


def l2_normalize(x, axis=-1, epsilon=1e-12):
  """attention_lib.l2_normalize (attention_lib.py:30-33); only the last axis is supported."""
  x = _dev(x)
  if axis not in (-1, x.dim() - 1):
    raise NotImplementedError("l2_normalize is built for the last axis only (the only use on the hot path)")
  y, _ = ops.l2norm_rows(x.reshape(-1, x.shape[-1]), eps=epsilon)
  return y.view(x.shape)


def contrastive_loss(image_feat, cond_feat, l2_norm=True, temperature=0.1, sync_match=False):
  """attention_lib.contrastive_loss (attention_lib.py:46-79): InfoNCE over the LOCAL batch."""
  if sync_match:
    raise NotImplementedError  # same as the reference (attention_lib.py:58-59)
  if not l2_norm:
    raise NotImplementedError("l2_norm=False is not used on the hot path and is not built")
  slot = ops.empty(1, ops.F32)
  stats = ops.empty(2, ops.F32)
  _engine.Contrastive(_dev(image_feat), _dev(cond_feat), slot, temperature, stats=stats)
  return slot[0], stats[0], stats[1]


def attention(region_feat, word_feat, gamma, mask=None):
  """attention_lib.attention (attention_lib.py:105-127): l2-normalise both, alpha = softmax over REGIONS (axis -2) of
  gamma * R^ W^^T (+ mask * -1e9), region_context = alpha^T R^ -> [batch, words, feat]. Kernel program: l2norm_rows,
  one tcgen05 GEMM for the scores, the tiled region-softmax kernel, one batched tcgen05 A^T B GEMM for the context.
  `mask`: None or the reference's word-padding mask (1 for padded words, constant over regions). The reference adds
  -1e9 to a padded word's whole column BEFORE the softmax over regions; in fp32, gamma * cos - 1e9 rounds to exactly
  -1e9 for every region (|gamma * cos| < 32 = half an ulp of 1e9), so a padded word gets the UNIFORM attention 1 / R
  and its context is the mean of the normalised regions. Reproduced here by clearing those score columns (a fill)
  before the softmax kernel. (word_loss drops these words again through its word mask, attention_lib.py:164-165.)"""
  r = _dev(region_feat).to(torch.bfloat16)
  w = _dev(word_feat)
  B, R, D = r.shape
  L = w.shape[1]
  if w.shape[0] != B:
    raise ValueError("region_feat and word_feat must have the same batch size")
  rh, _ = ops.l2norm_rows(r.reshape(B * R, D), out_dtype=ops.BF16)
  wh, _ = ops.l2norm_rows(w.reshape(B * L, D), out_dtype=ops.BF16)
  ldS = (L + 7) // 8 * 8
  S = ops.zeros((B, R, ldS), ops.F32)
  # scores of image b against ITS OWN words: one GEMM per batch entry (batched B operand)
  ops.conv_fwd(rh.view(B, 1, R, D), wh.view(B, L, D), 1, L, ldb=D, batched=True, stride_b=L * D,
               out=S.view(B, 1, R, ldS)[..., :L])
  if mask is not None:
    if abs(float(gamma)) >= 32.0:
      raise NotImplementedError("the -1e9 shift only absorbs the scores exactly for |gamma| < 32")
    pad_from = (L - _dev(mask)[:, 0, :].sum(-1)).long().tolist()   # index glue: the mask is [arange(L) >= max_len]
    for b, ml in enumerate(pad_from):
      if ml < L:
        S[b, :, ml:L].zero_()   # fill: equal scores -> uniform softmax over regions, as the reference's fp32 gives
  alpha = ops.empty((B, R, ldS), ops.BF16)
  alphaT = ops.zeros((B, ldS, R), ops.BF16)
  # the region-softmax kernel works on [images][R][columns]; here every "image" has its own L columns
  for b in range(B):   # B small launches (functional API; the fused word_loss handles all pairs in one)
    ops._call("xmc_wl_softmax", S[b].data_ptr(), 1, R, L, ldS, float(gamma), alpha[b].data_ptr(), alphaT[b].data_ptr(), 0,
              ops.stream())
  ctx = ops.empty((B, ldS, D), ops.F32)
  ops.wgrad(alpha.view(B, 1, R, ldS), rh.view(B, 1, R, D), 1, ctx, out_mode=1, batched=True, ld_out=D, tap_stride=0,
            batch_stride=ldS * D)
  return ctx[:, :L]


def word_loss(image_feat, word_feat, max_len, gamma1=5, gamma2=5, gamma3=50):
  """attention_lib.word_loss (attention_lib.py:130-191)."""
  if (gamma1, gamma2, gamma3) != (5, 5, 50):
    raise NotImplementedError("only the reference defaults gamma1=gamma2=5, gamma3=50 are built")
  img = _dev(image_feat).to(torch.bfloat16)  # plumbing cast of the caller's tensor; the kernels consume bf16 regions
  ws = _engine.WordShared(_dev(word_feat), _dev(max_len))
  slot = ops.empty(1, ops.F32)
  stats = ops.empty(2, ops.F32)
  _engine.WordLoss(img, ws, slot, stats=stats)
  return slot[0], stats[0], stats[1]


def attention_for_g(region_feat, word_feat, gamma, mask=None):
  """attention_lib.attention_for_g (attention_lib.py:194-219). `mask` must be the reference's word-padding mask
  (1 for w >= max_len, constant over regions) or None."""
  q = _dev(region_feat).to(torch.bfloat16)
  w = _dev(word_feat)
  B, R, D = q.shape
  L = w.shape[1]
  if mask is None:
    max_len = torch.full((B,), float(L), device="cuda")
  else:
    max_len = (L - _dev(mask)[:, 0, :].sum(-1)).contiguous()  # index glue: the mask is [arange(L) >= max_len]
  what, _ = ops.l2norm_rows(w.reshape(B * L, D))
  ctx = ops.empty((B * R, D))
  attn = ops.attention_g_fwd(q, what.view(B, L, D), max_len, gamma, ctx)
  return ctx.view(B, R, D), attn.view(B, R, L)
