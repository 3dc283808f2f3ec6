"""GPU-side producer of the train_step input contract: the per-example part of
COCODataset.preprocess (xmcgan/libml/coco_dataset.py:127-167) for examples that are already decoded and resized to
config.image_size — random left-right flip, clip to [0,1], caption choice, sentence_embedding = sum(words)/len, z.
TFRecord parsing, PNG decoding, the bilinear resize to image_size and `augmentation.augment` (whose output `image_aug`
train_step never reads) stay outside; the reference does them in tf.data.

Random numbers: the reference derives (flip, sentence index, z) from a per-example stateless TF seed; that stream cannot
be reproduced without TF, so they come from a seeded torch CUDA generator here (same distributions), or are passed in."""
import torch

from .. import _lib
from .. import ops


def preprocess(features, rng=0, z_dim=128, training=True, return_text=False, flip=None, sentence_idx=None, z=None):
  """features: {"image": [N,S,S,3] fp32 in [0,1], "caption/embedding": [N,M,L,E] fp32, "caption/max_len": [N,M] int}.
  Returns the batch dict train_step consumes: image, embedding [N,L,E], max_len [N,1] (float), sentence_embedding
  [N,E], z [N,z_dim]. `training=False` keeps the reference behaviour (preprocess is applied identically to the
  evaluation split); `return_text=True` picks the SHORTEST caption instead of a random one (coco_dataset.py:150-152)."""
  img = torch.as_tensor(features["image"]).to("cuda", torch.float32).contiguous()
  emb = torch.as_tensor(features["caption/embedding"]).to("cuda", torch.float32).contiguous()
  lens = torch.as_tensor(features["caption/max_len"]).to("cuda", torch.int32).contiguous()
  N, H, W, _ = img.shape
  _, M, L, E = emb.shape
  g = torch.Generator(device="cuda").manual_seed(int(rng))
  if flip is None:      # tf.image.stateless_random_flip_left_right: p = 0.5
    flip = torch.rand(N, device="cuda", generator=g) < 0.5
  flip = torch.as_tensor(flip).to("cuda", torch.uint8).contiguous()
  if sentence_idx is None:
    if return_text:     # argsort(max_len, DESCENDING)[-1]: the shortest caption (an index op)
      sentence_idx = torch.argsort(lens, dim=1, descending=True, stable=True)[:, -1]
    else:               # stateless_uniform([], minval=0, maxval=sentence_num, int32)
      sentence_idx = torch.randint(0, M, (N,), device="cuda", generator=g)
  sentence_idx = torch.as_tensor(sentence_idx).to("cuda", torch.int32).contiguous()
  if z is None:
    z = torch.randn(N, z_dim, device="cuda", generator=g)
  out_img = ops.empty((N, H, W, 3), ops.F32)
  ops._call("xmc_prep_image", img.data_ptr(), flip.data_ptr(), N, H, W, out_img.data_ptr(), _lib.stream())
  emb_out = ops.empty((N, L, E), ops.F32)
  len_out = ops.empty((N, 1), ops.F32)
  sent = ops.empty((N, E), ops.F32)
  ops._call("xmc_prep_caption", emb.data_ptr(), lens.data_ptr(), sentence_idx.data_ptr(), N, M, L, E,
            emb_out.data_ptr(), len_out.data_ptr(), sent.data_ptr(), _lib.stream())
  return {"image": out_img, "embedding": emb_out, "max_len": len_out, "sentence_embedding": sent,
          "z": torch.as_tensor(z).to("cuda", torch.float32)}
