"""Drop-in counterparts of xmc_net.Generator / xmc_net.Discriminator (xmcgan/nets/xmc_net.py:28-248) with the Flax
module call surface the reference's callers use:

    generator = functools.partial(Generator, config=config, dtype=dtype)
    variables = generator(train=False).init(rng, (batch, z))            # {"params", "batch_stats"}
    image, new_vars = generator(train=True).apply(variables, (batch, z), mutable=["batch_stats"])

Tensors are CUDA torch tensors (NHWC); variable collections are `FlatTree`s — nested-dict views (Flax auto-names,
SURVEY.md Appendix A) over one flat fp32 device buffer. Plain nested dicts of arrays are accepted as well.
All compute runs in libxmc.so (engine.py); there is no CPU path."""
import collections.abc

import torch

from .. import engine as _engine
from .. import ops

_ENGINES = {}


def _cfg_key(config, kind, embedding_dim):
  keys = ("image_size", "gf_dim", "df_dim", "z_dim", "g_spectral_norm", "d_spectral_norm", "batch_norm_group_size",
          "gamma_for_g", "word_contrastive", "sentence_contrastive", "image_contrastive", "cond_size", "dtype")
  return (kind, embedding_dim) + tuple(getattr(config, k) for k in keys)


def get_engine(config, kind, embedding_dim=768):
  key = _cfg_key(config, kind, embedding_dim)
  if key not in _ENGINES:
    cls = _engine.GeneratorEngine if kind == "g" else _engine.DiscriminatorEngine
    _ENGINES[key] = cls(config, embedding_dim)
  return _ENGINES[key]


class FlatTree(collections.abc.Mapping):
  """A variable collection: nested-dict view (Flax names) of one flat fp32 CUDA buffer."""

  def __init__(self, layout, buf):
    self.layout = layout
    self.buf = buf
    self._tree = layout.tree(buf)

  def __getitem__(self, k):
    return self._tree[k]

  def __iter__(self):
    return iter(self._tree)

  def __len__(self):
    return len(self._tree)

  def clone(self):
    return FlatTree(self.layout, self.buf.clone())

  def to_cpu_tree(self):
    def rec(n):
      return {k: rec(v) for k, v in n.items()} if isinstance(n, dict) else n.detach().float().cpu().clone()
    return rec(self._tree)


def as_flat(layout, coll):
  if isinstance(coll, FlatTree):
    assert coll.layout is layout
    return coll.buf
  buf = torch.zeros(layout.total, device="cuda", dtype=torch.float32)
  layout.load_tree(buf, coll)
  return buf


def _seed_of(rng):
  if rng is None:
    return 0
  if isinstance(rng, int):
    return rng
  t = torch.as_tensor(rng).reshape(-1)
  return int(t[-1].item()) & 0x7FFFFFFF


def _to_dev(x, dtype=torch.float32):
  return torch.as_tensor(x).to(device="cuda", dtype=dtype).contiguous()


def batch_to_device(batch):
  return {k: _to_dev(v) for k, v in batch.items() if k in ("image", "embedding", "max_len", "sentence_embedding", "z")}


class Generator:
  """xmc_net.Generator (xmc_net.py:145-248)."""

  def __init__(self, config, train, dtype=torch.bfloat16, activation_fn=None):
    self.config, self.train, self.dtype = config, train, dtype

  def _engine(self, inputs):
    cond_dict, _ = inputs
    return get_engine(self.config, "g", int(cond_dict["embedding"].shape[-1]))

  def init(self, rng, inputs):
    eng = self._engine(inputs)
    p, s = eng.init_params(_seed_of(rng))
    out = {"params": FlatTree(eng.layout, p), "batch_stats": FlatTree(eng.stats_layout, s)}
    if eng.sn:
      out["spectral_norm_stats"] = FlatTree(eng.u_layout, eng.init_u0(_seed_of(rng) + 2))
    return out

  def apply(self, variables, inputs, mutable=False, rngs=None):
    eng = self._engine(inputs)
    cond_dict, z = inputs
    batch = batch_to_device(cond_dict)
    z = _to_dev(z)
    params = as_flat(eng.layout, variables["params"])
    stats = as_flat(eng.stats_layout, variables["batch_stats"])
    u0 = as_flat(eng.u_layout, variables["spectral_norm_stats"]) if eng.sn else None
    u0_new = torch.empty_like(u0) if eng.sn else None
    eng.prep_weights(params, u0, u0_new)
    want_state = bool(mutable) and self.train
    new_stats = torch.empty_like(stats) if want_state else None
    img, _ = eng.forward(params, stats, batch, z, train=self.train, new_stats=new_stats)
    if mutable is False:
      return img
    out_state = {"batch_stats": FlatTree(eng.stats_layout, new_stats if want_state else stats)}
    if eng.sn:  # u0 advances only in train mode (layers.py:98-99,215-216)
      out_state["spectral_norm_stats"] = FlatTree(eng.u_layout, u0_new if self.train else u0)
    return img, out_state


class Discriminator:
  """xmc_net.Discriminator (xmc_net.py:28-142). Returns (logit [2B,1], statistic_dict)."""

  def __init__(self, config, train, dtype=torch.bfloat16, activation_fn=None):
    self.config, self.train, self.dtype = config, train, dtype

  def _engine(self, inputs):
    _, cond_dict = inputs
    return get_engine(self.config, "d", int(cond_dict["embedding"].shape[-1]))

  def init(self, rng, inputs):
    eng = self._engine(inputs)
    p, u = eng.init_params(_seed_of(rng))
    out = {"params": FlatTree(eng.layout, p)}
    if eng.sn:
      out["spectral_norm_stats"] = FlatTree(eng.u_layout, u)
    return out

  def apply(self, variables, inputs, mutable=False, rngs=None):
    eng = self._engine(inputs)
    x, cond_dict = inputs
    batch = batch_to_device(cond_dict)
    x = _to_dev(x)
    n2, s = x.shape[0], x.shape[1]
    with ops.act_dtype(eng.act):
      images = ops.cast_to_bf16(x.reshape(n2 * s * s, 3)).view(n2, s, s, 3)
    params = as_flat(eng.layout, variables["params"])
    u0 = as_flat(eng.u_layout, variables["spectral_norm_stats"]) if eng.sn else None
    u0_new = torch.empty_like(u0) if eng.sn else None
    eng.prep_weights(params, u0, u0_new)
    losses = torch.zeros(16, device="cuda")
    # the module API returns the reference's full statistic dict (xmc_net.py:106-141), including the accuracy /
    # entropy side statistics that train_step never reads (XLA removes them there; the engine skips them there too)
    stats = torch.zeros(16, 2, device="cuda")
    logit, _ = eng.forward(params, images, batch, losses, need_g=True, stats=stats)
    S = _engine.LOSS_SLOTS
    stat = {}
    for name, slot in (("fake_word", "fake_word"), ("real_word", "real_word"), ("fake_sentence", "fake_sent"),
                       ("real_sentence", "real_sent"), ("image_contrastive", "image")):
      stat[name + "_loss"] = losses[S[slot]]
      stat[name + "_acc"] = stats[S[slot], 0]
      stat[name + "_entropy"] = stats[S[slot], 1]
    out = (logit.view(n2, 1), stat)
    if mutable is False:
      return out
    new_state = {}
    if eng.sn:
      new_state["spectral_norm_stats"] = FlatTree(eng.u_layout, u0_new if self.train else u0)
    return out, new_state
