"""Execution engines for the XMC-GAN generator and discriminator: parameter layout, weight preparation, and explicit
forward / backward programs made of libxmc.so kernel launches (no autograd, no torch arithmetic).

Reference semantics: xmcgan/nets/xmc_net.py:28-248, xmcgan/nets/common.py:58-186, xmcgan/libml/layers.py:49-273,
xmcgan/libml/attention_lib.py:46-219. Differences that are exact in real arithmetic and deliberate:
  * LocalConditionalBatchNorm's 1x1 gamma/beta convolutions run once at 16x16 as ONE concatenated GEMM and are read
    through (h>>s, w>>s) indexing (a 1x1 conv commutes with nearest upsampling);
  * generator shortcuts run conv1x1 before the upsample (same reason) and are added in the main conv's epilogue;
  * the generator-side pull-back through the discriminator only touches the fake half of the batch (the
    discriminator has no cross-example op, so the real half cannot reach the generator).
"""
import collections
import math

import numpy as np
import torch

from . import _lib
from . import ops
from .ops import BF16, F32

ConvRec = collections.namedtuple(
    "ConvRec", "path kh cin cout w_off b_off fwd_off ld_fwd dg_off ld_dg sn")

LOSS_SLOTS = dict(hinge_d=0, hinge_g=1, real_word=2, fake_word=3, real_sent=4, fake_sent=5, image=6, pretrained=7)


def _r4(n):
  return (n + 3) // 4 * 4


def _r8(n):
  return (n + 7) // 8 * 8


def act_dtype_of(config):
  """Activation dtype of config.dtype (train_utils.py:148-151): "bfloat16" (coco_xmc.py:45) or "float32"."""
  dt = getattr(config, "dtype", "bfloat16")
  if dt == "bfloat16":
    return BF16
  if dt == "float32":
    return F32
  raise ValueError(f"config.dtype {dt!r} is not supported (bfloat16 or float32)")


def as4(t2d):
  """[rows, C] (pitched) -> [1,1,rows,C] so that the GEMM sees `rows` pixels along W."""
  return t2d[None, None]


class Layout:
  """Flat fp32 buffer layout: path tuple -> (offset, shape)."""

  def __init__(self):
    self.entries = collections.OrderedDict()
    self.total = 0

  def add(self, path, shape):
    size = int(np.prod(shape))
    self.entries[path] = (self.total, tuple(shape))
    self.total += _r4(size)
    return self.entries[path][0]

  def off(self, path):
    return self.entries[path][0]

  def view(self, buf, path):
    off, shape = self.entries[path]
    return buf[off:off + int(np.prod(shape))].view(shape)

  def tree(self, buf):
    root = {}
    for path in self.entries:
      node = root
      for k in path[:-1]:
        node = node.setdefault(k, {})
      node[path[-1]] = self.view(buf, path)
    return root

  def load_tree(self, buf, tree):
    for path in self.entries:
      node = tree
      for k in path:
        node = node[k]
      self.view(buf, path).copy_(torch.as_tensor(node).to(buf.device, buf.dtype).reshape(self.entries[path][1]))


def _glorot_std(shape):
  rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
  fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
  return math.sqrt(2.0 / (fan_in + fan_out))


def init_flat(layout, seed, kind_of):
  """Random initialisation matching the reference's initialisers in distribution: glorot-normal kernels
  (xmc_net.py:70-80,181-191), zero biases, u0 ~ N(0, 0.01^2) (layers.py:86-91), BN stats (0, 1)."""
  g = torch.Generator(device="cpu").manual_seed(seed)
  buf = torch.zeros(layout.total, dtype=torch.float32)
  for path, (off, shape) in layout.entries.items():
    n = int(np.prod(shape))
    kind = kind_of(path)
    if kind == "kernel":
      buf[off:off + n] = torch.randn(n, generator=g) * _glorot_std(shape)
    elif kind == "u0":
      buf[off:off + n] = torch.randn(n, generator=g) * 0.01
    elif kind == "var":
      buf[off:off + n] = 1.0
  return buf


def _kind(path):
  leaf = path[-1]
  return leaf if leaf in ("kernel", "bias", "u0", "mean", "var") else "other"


class _PrepTable:
  """Device table for xmc_prep_weights (fp32 HWIO -> bf16 K-major forward / dgrad copies)."""

  def __init__(self):
    self.entries = []
    self.tiles = 0

  def add(self, w_off, taps, cin, cout, fwd_off, ld_fwd, dg_off, ld_dg, sn, cscale_off=-1, split=False,
          dg_part_stride=0):
    """split: fp32-activation mode, every weight stored as [hi | hi | lo] (XmcPrepEntry.split); the caller's offsets
    and pitches already account for the 3x wider rows."""
    e = _lib.PrepEntry()
    e.split, e.dg_part_stride = int(split), dg_part_stride
    e.w_off, e.wk_fwd_off, e.wk_dg_off = w_off, fwd_off, dg_off
    e.bias_off = e.bias_dst_off = -1
    e.cscale_off = cscale_off
    e.taps, e.cin, e.cout = taps, cin, cout
    e.ld_fwd, e.ld_dg, e.sn = ld_fwd, ld_dg, sn
    e.tile_begin = self.tiles
    self.tiles += ((taps * cin + 63) // 64) * ((cout + 63) // 64)   # 64 x 64 tiles of xmc_prep_weights
    self.entries.append(e)

  def upload(self):
    arr = (_lib.PrepEntry * len(self.entries))(*self.entries)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
    self.dev = raw.cuda()
    self.n = len(self.entries)


class _SnTable:
  """Device table + workspaces for the multi-tensor spectral-norm kernels (xmc_sn_forward / xmc_sn_backward): one
  entry per SpectralConv / SpectralDense (layers.py:49-113,125-241), u0 of every layer in one flat `u_layout`."""

  def __init__(self):
    self.entries = []
    self.u_layout = Layout()
    self._t = self._s = self._rb = self._ct = self._eb = 0
    self.dev = None

  def add(self, path, w_off, rows, cols):
    """Registers a [rows, cols] kernel at params[w_off:]; returns its slot."""
    slot = len(self.entries)
    e = _lib.SnEntry()
    e.w_off, e.t_off, e.s_off, e.u_off = w_off, self._t, self._s, self.u_layout.add(path + ("u0",), (1, cols))
    e.rows, e.cols = rows, cols
    e.row_block_begin, e.col_tile_begin, e.elem_block_begin = self._rb, self._ct, self._eb
    self._t += _r4(rows)
    self._s += _r4(((rows + 255) // 256) * cols)   # row-tile partials of s = t W (added in order by sn_finalize)
    self._rb += (rows + 7) // 8
    self._ct += ((rows + 255) // 256) * ((cols + 31) // 32)
    self._eb += (rows * cols + 2047) // 2048
    self.entries.append(e)
    return slot

  @property
  def n(self):
    return len(self.entries)

  def upload(self):
    arr = (_lib.SnEntry * self.n)(*self.entries)
    self.dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().cuda()
    self.t_ws = torch.zeros(self._t, device="cuda")
    self.s_ws = torch.zeros(self._s, device="cuda")
    self.scalars = torch.zeros(4 * self.n, device="cuda")
    self.dot_ws = torch.zeros(max(self._eb, 1), device="cuda")

  def forward(self, params, u0, u0_new):
    """One power-iteration step for every layer: u0_new, and 1/(sigma+eps) into scalars[2n + slot]."""
    ops._call("xmc_sn_forward", self.dev.data_ptr(), self.n, 1e-10, params.data_ptr(), u0.data_ptr(),
              u0_new.data_ptr(), self.t_ws.data_ptr(), self.s_ws.data_ptr(), self._s, self.scalars.data_ptr(),
              self._rb, self._ct, _lib.stream(), launches=3)

  def backward(self, params, grads, u0_new):
    """grads holds d/dW~ for every registered kernel -> d/dW in place (sigma is differentiable, u/v are not)."""
    ops._call("xmc_sn_backward", self.dev.data_ptr(), self.n, params.data_ptr(), grads.data_ptr(),
              self.t_ws.data_ptr(), u0_new.data_ptr(), self.scalars.data_ptr(), self.dot_ws.data_ptr(), self._eb,
              _lib.stream(), launches=3)

  def inv_sigma(self, slot):
    return self.scalars[2 * self.n + slot:]


# ======================================================================================================================
# Generator
# ======================================================================================================================
class GeneratorEngine:
  """xmc_net.Generator (xmc_net.py:145-248)."""

  def __init__(self, config, embedding_dim=768):
    if config.image_size == 256:
      channel_dims = [16, 8, 8, 4, 2, 1]
    elif config.image_size == 128:
      channel_dims = [16, 8, 4, 2, 1]
    else:
      raise ValueError(f"image_size {config.image_size} is not supported (reference: xmc_net.py:202-205)")
    self.config = config
    # activation dtype: bf16 (reference default) or fp32 (config.dtype = "float32": fp32 activations, every GEMM as
    # three bf16 passes over [hi | lo] splits of both operands). S = width factor of the split weight copies.
    self.act = act_dtype_of(config)
    self.S = S_ = 3 if self.act == F32 else 1
    # g_spectral_norm switches EVERY conv / dense of the generator, including the gamma / beta layers inside
    # (Local)ConditionalBatchNorm, to the spectral variants (xmc_net.py:176-191, layers.py:244-273)
    self.sn = bool(config.g_spectral_norm)
    self.cpre = cp = "SpectralConv" if self.sn else "Conv"
    self.dpre = dp = "SpectralDense" if self.sn else "Dense"
    self.sntab = _SnTable() if self.sn else None
    self.E = E = embedding_dim
    self.zd = zd = config.z_dim
    self.cd = cd = 2 * zd
    self.scd = scd = E + cd
    gf = config.gf_dim
    self.ch = ch = [gf * c for c in channel_dims]
    self.c0 = c0 = gf * 16
    self.n_spatial = len(channel_dims) - 2
    self.gamma = float(config.gamma_for_g)
    if any(c % 8 for c in ch + [c0, E, zd]):
      raise ValueError("channel counts must be multiples of 8")

    # ---- block descriptions -----------------------------------------------------------------------------------
    blocks = []  # (name, kind, cin, cout)
    cin = c0
    for i in range(2):
      blocks.append((f"GenBlock_{i}", "cbn", cin, ch[i]))
      cin = ch[i]
    for k in range(self.n_spatial):
      blocks.append((f"GenSpatialBlock_{k}", "lcbn", cin, ch[k + 2]))
      cin = ch[k + 2]
    self.blocks = blocks
    self.c_last = cin

    # ---- concatenated gamma/beta groups ----------------------------------------------------------------------
    self.cbn = []   # (path_prefix, C, goff, boff)
    self.lcbn = []
    off = 0
    for name, kind, bcin, bcout in blocks:
      if kind == "cbn":
        for j, C in enumerate((bcin, bcout)):
          self.cbn.append(((name, f"ConditionalBatchNorm_{j}"), C, off, off + C))
          off += 2 * C
    self.NC = off
    off = 0
    for name, kind, bcin, bcout in blocks:
      if kind == "lcbn":
        for j, C in enumerate((bcin, bcout)):
          self.lcbn.append(((name, f"LocalConditionalBatchNorm_{j}"), C, off, off + C))
          off += 2 * C
    self.lcbn.append((("LocalConditionalBatchNorm_0",), self.c_last, off, off + self.c_last))
    off += 2 * self.c_last
    self.NL = off

    # ---- parameter layout: concatenated biases first (so one bias vector / one colsum serves the whole group) --
    L = self.layout = Layout()
    self.lcbn_bias_off = L.total
    for prefix, C, goff, boff in self.lcbn:
      assert L.add(prefix + (cp + "_0", "bias"), (C,)) == self.lcbn_bias_off + goff
      assert L.add(prefix + (cp + "_1", "bias"), (C,)) == self.lcbn_bias_off + boff
    self.cbn_bias_off = L.total
    for prefix, C, goff, boff in self.cbn:
      assert L.add(prefix + (dp + "_0", "bias"), (C,)) == self.cbn_bias_off + goff
      assert L.add(prefix + (dp + "_1", "bias"), (C,)) == self.cbn_bias_off + boff
    for prefix, C, goff, boff in self.lcbn:
      L.add(prefix + (cp + "_0", "kernel"), (1, 1, scd, C))
      L.add(prefix + (cp + "_1", "kernel"), (1, 1, scd, C))
    for prefix, C, goff, boff in self.cbn:
      L.add(prefix + (dp + "_0", "kernel"), (cd, C))
      L.add(prefix + (dp + "_1", "kernel"), (cd, C))

    self.arena_size = 0
    self.prep = _PrepTable()
    self.convs = {}

    self.subpixel = {}  # path -> (wf_off, vd_off): sub-pixel form of the convs that follow a nearest 2x upsample

    def add_conv(path, kh, cin_, cout_, dense=False, subpixel=False):
      shape = (cin_, cout_) if dense else (kh, kh, cin_, cout_)
      w_off = L.add(path + ("kernel",), shape)
      b_off = L.add(path + ("bias",), (cout_,))
      taps = kh * kh
      slot = self.sntab.add(path, w_off, taps * cin_, cout_) if self.sn else -1
      ld_fwd, ld_dg = _r8(S_ * taps * cin_), _r8(S_ * taps * cout_)
      if subpixel:
        # conv3x3(upsample(x)) as four 2x2 convs on x: [4*Cout][4*Cin] forward and [Cin][16*Cout] dgrad matrices
        wf_off = self.arena_size
        self.arena_size += 4 * cout_ * 4 * cin_ * S_
        vd_off = self.arena_size
        self.arena_size += cin_ * 16 * cout_ * S_
        self.subpixel[path] = (wf_off, vd_off)
        self.convs[path] = ConvRec(path, kh, cin_, cout_, w_off, b_off, -1, ld_fwd, -1, ld_dg, slot)
        return
      fwd_off = self.arena_size
      self.arena_size += _r8(cout_) * ld_fwd
      dg_off = self.arena_size
      self.arena_size += _r8(cin_) * ld_dg
      self.prep.add(w_off, taps, cin_, cout_, fwd_off, ld_fwd, dg_off, ld_dg, slot, split=S_ == 3)
      self.convs[path] = ConvRec(path, kh, cin_, cout_, w_off, b_off, fwd_off, ld_fwd, dg_off, ld_dg, slot)

    # registration order = the reference's call order (xmc_net.py:209-246): Flax auto-numbers modules per class
    add_conv((dp + "_0",), 1, E, zd, dense=True)
    add_conv((dp + "_1",), 1, zd, c0 * 16, dense=True)
    for name, kind, bcin, bcout in blocks:
      add_conv((name, cp + "_0"), 3, bcin, bcout, subpixel=True)
      add_conv((name, cp + "_1"), 3, bcout, bcout)
      add_conv((name, cp + "_2"), 1, bcin, bcout)
    add_conv((cp + "_0",), 1, ch[1], E)
    add_conv((cp + "_1",), 3, self.c_last, 3)

    # concatenated matrices: forward [N*][K] rows per layer, dgrad [K][N*] column slices (split mode: [K][3][N*])
    self.cbn_fwd_off = self.arena_size
    self.arena_size += self.NC * cd * S_
    self.cbn_dg_off = self.arena_size
    self.arena_size += cd * self.NC * S_
    self.lcbn_fwd_off = self.arena_size
    self.arena_size += self.NL * scd * S_
    self.lcbn_dg_off = self.arena_size
    self.arena_size += scd * self.NL * S_
    for prefix, C, goff, boff in self.cbn:
      for leaf, o in ((dp + "_0", goff), (dp + "_1", boff)):
        w_off = L.off(prefix + (leaf, "kernel"))
        slot = self.sntab.add(prefix + (leaf,), w_off, cd, C) if self.sn else -1
        self.prep.add(w_off, 1, cd, C, self.cbn_fwd_off + o * cd * S_, cd * S_, self.cbn_dg_off + o, self.NC * S_, slot,
                      split=S_ == 3, dg_part_stride=self.NC)
    for prefix, C, goff, boff in self.lcbn:
      for leaf, o in ((cp + "_0", goff), (cp + "_1", boff)):
        w_off = L.off(prefix + (leaf, "kernel"))
        slot = self.sntab.add(prefix + (leaf,), w_off, scd, C) if self.sn else -1
        self.prep.add(w_off, 1, scd, C, self.lcbn_fwd_off + o * scd * S_, scd * S_, self.lcbn_dg_off + o,
                      self.NL * S_, slot, split=S_ == 3, dg_part_stride=self.NL)
    self.u_layout = self.sntab.u_layout if self.sn else Layout()

    # ---- batch_stats layout ------------------------------------------------------------------------------------
    S = self.stats_layout = Layout()
    for prefix, C, _, _ in self.cbn + self.lcbn:
      S.add(prefix + ("BatchNorm_0", "mean"), (C,))
      S.add(prefix + ("BatchNorm_0", "var"), (C,))
    self.bn_index = {prefix: (C, goff, boff) for prefix, C, goff, boff in self.cbn + self.lcbn}
    self.arena = None
    self.prepped_for = None  # (data_ptr, version) of the parameter buffer the bf16 arena currently mirrors

  # ------------------------------------------------------------------------------------------------------------
  def init_params(self, seed):
    return init_flat(self.layout, seed, _kind).cuda(), init_flat(self.stats_layout, seed + 1, _kind).cuda()

  def init_u0(self, seed):
    return init_flat(self.u_layout, seed, _kind).cuda()

  def _ensure(self):
    if self.arena is None:
      self.arena = torch.zeros(self.arena_size, device="cuda", dtype=BF16)
      self.prep.upload()
      if self.sn:
        self.sntab.upload()

  def prep_weights(self, params, u0=None, u0_new=None):
    """bf16 weight copies of `params`; with g_spectral_norm one power-iteration step first (u0 -> u0_new)."""
    self._ensure()
    if self.sn:
      if u0 is None or u0_new is None:
        raise ValueError("g_spectral_norm=True: the spectral_norm_stats collection (u0) is required")
      self.sntab.forward(params, u0, u0_new)
    ops._call("xmc_prep_weights", self.prep.dev.data_ptr(), self.prep.n, self.prep.tiles, params.data_ptr(),
              self.sntab.scalars.data_ptr() if self.sn else None, self.sntab.n if self.sn else 0,
              self.arena.data_ptr(), None, None, _lib.stream())
    for path, (wf_off, vd_off) in self.subpixel.items():
      rec = self.convs[path]
      scale = self.sntab.inv_sigma(rec.sn).data_ptr() if self.sn else None
      ops._call("xmc_subpixel_prep", params[rec.w_off:].data_ptr(), scale, rec.cin, rec.cout, int(self.S == 3),
                self.arena[wf_off:].data_ptr(), self.arena[vd_off:].data_ptr(), _lib.stream())
    self.prepped_for = self.prep_key(params, u0_new)

  def prep_key(self, params, u0_new=None):
    """Identity of what the bf16 arena mirrors: parameter buffer + its torch version (+ where the new u0 went).
    Kernels that update the parameters through raw pointers (xmc_adam) must reset `prepped_for` themselves."""
    return (params.data_ptr(), params._version, u0_new.data_ptr() if (self.sn and u0_new is not None) else 0)

  def sn_backward(self, params, grads, u0_new):
    if self.sn:
      self.sntab.backward(params, grads, u0_new)

  def _wk(self, rec):
    return self.arena[rec.fwd_off:]

  def _wd(self, rec):
    return self.arena[rec.dg_off:]

  def _bn_group(self, B, train):
    """Cross-replica BatchNorm group of this rank (train mode, batch_norm_group_size > 0, xmc_net.py:196-200)."""
    gbs = self.config.batch_norm_group_size
    if not train or gbs <= 0:
      return None
    from . import parallel
    return parallel.bn_group(gbs, B)

  def _bn(self, x, prefix, stats, new_stats, train, group=None):
    C = x.shape[-1]
    SL = self.stats_layout
    mean = SL.view(stats, prefix + ("BatchNorm_0", "mean"))
    var = SL.view(stats, prefix + ("BatchNorm_0", "var"))
    if train:
      sums, P = ops.bn_stats(x)
      if group is not None:
        # flax BatchNorm pmean's [mean, mean of squares] over the replica group; all replicas hold the same number of
        # elements, so summing the raw sums and dividing by P * group_size is the same
        from . import parallel
        parallel.all_reduce_sum_(sums, group=group[0])
        P *= group[1]
      nm = SL.view(new_stats, prefix + ("BatchNorm_0", "mean")) if new_stats is not None else None
      nv = SL.view(new_stats, prefix + ("BatchNorm_0", "var")) if new_stats is not None else None
      return ops.bn_finalize(sums, P, C, mean, var, nm, nv)
    return ops.bn_eval_stats(mean, var, C)

  def forward(self, params, stats, batch, z, train=True, new_stats=None, fake_bf16=None):
    """Returns (image fp32 [B,S,S,3] in [0,1], ctx). `fake_bf16`: optional [B,S,S,3] view in the activation dtype that
    also receives the image (the second half of the discriminator input). Weights must have been prepared."""
    with ops.act_dtype(self.act):
      return self._forward(params, stats, batch, z, train, new_stats, fake_bf16)

  def _forward(self, params, stats, batch, z, train, new_stats, fake_bf16):
    P = params
    S_ = self.S
    E, zd, cd, scd = self.E, self.zd, self.cd, self.scd
    B = z.shape[0]
    cond = batch["sentence_embedding"].reshape(B, E)
    words = batch["embedding"]
    Lw = words.shape[1]
    max_len = batch["max_len"].reshape(B).contiguous()
    grp = self._bn_group(B, train)
    ctx = {"B": B, "Lw": Lw, "train": train, "bn_group": grp}

    cond_bf = ops.cast_to_bf16(cond)
    gc = ops.empty((B, cd))
    r = self.convs[(self.dpre + "_0",)]
    ops.conv_fwd(as4(cond_bf), self._wk(r), 1, zd, bias=P[r.b_off:], out=as4(gc[:, :zd]), ldb=r.ld_fwd)
    ops.cast_to_bf16(z.reshape(B, zd), gc[:, zd:])
    r = self.convs[(self.dpre + "_1",)]
    x = ops.conv_fwd(as4(gc[:, zd:]), self._wk(r), 1, r.cout, bias=P[r.b_off:], ldb=r.ld_fwd)
    x = x.view(B, 4, 4, self.c0)
    gbC = ops.conv_fwd(as4(gc), self.arena[self.cbn_fwd_off:], 1, self.NC, bias=P[self.cbn_bias_off:],
                       ldb=cd * S_).view(B, self.NC)
    ctx.update(cond_bf=cond_bf, gc=gc, gbC=gbC, blocks=[])

    gb, Hc = gbC, 1
    for name, kind, bcin, bcout in self.blocks:
      if kind == "lcbn" and Hc == 1:
        # ---- word-region attention at 16x16 and the spatial condition (xmc_net.py:220-235) ---------------------
        r = self.convs[(self.cpre + "_0",)]
        xq = ops.conv_fwd(x, self._wk(r), 1, E, bias=P[r.b_off:], ldb=r.ld_fwd)
        R = xq.shape[1] * xq.shape[2]
        what, _ = ops.l2norm_rows(words.reshape(B * Lw, E))
        spatial = ops.empty((B * R, scd))
        attn = ops.attention_g_fwd(xq.view(B, R, E), what.view(B, Lw, E), max_len, self.gamma, spatial)
        ops.bcast_rows(gc, R, spatial[:, E:])
        gbL = ops.conv_fwd(as4(spatial), self.arena[self.lcbn_fwd_off:], 1, self.NL, bias=P[self.lcbn_bias_off:],
                           ldb=scd * S_).view(B * R, self.NL)
        ctx.update(x16=x, xq=xq, what=what, attn=attn, spatial=spatial, gbL=gbL, R=R, Hc=xq.shape[1])
        gb, Hc = gbL, xq.shape[1]
      bn0 = (name, ("ConditionalBatchNorm_0" if kind == "cbn" else "LocalConditionalBatchNorm_0"))
      bn1 = (name, ("ConditionalBatchNorm_1" if kind == "cbn" else "LocalConditionalBatchNorm_1"))
      _, g0, b0 = self.bn_index[bn0]
      _, g1, b1 = self.bn_index[bn1]
      r0, r1, r2 = (self.convs[(name, f"{self.cpre}_{i}")] for i in range(3))
      mr0 = self._bn(x, bn0, stats, new_stats, train, grp)
      # CBN -> relu at the block's input resolution; the nearest 2x upsample is folded into the next conv
      # (sub-pixel form: four 2x2 convs, 2.25x fewer FLOPs, the up-sampled tensor is never materialised)
      u = ops.bn_apply(x, mr0, gb, Hc, g0, b0, True, False)
      wf_off, _ = self.subpixel[(name, self.cpre + "_0")]
      c1 = ops.conv_fwd(u, self.arena[wf_off:], 2, bcout, bias=P[r0.b_off:], ldb=4 * bcin * S_, pad=1, subpixel=True)
      mr1 = self._bn(c1, bn1, stats, new_stats, train, grp)
      h2 = ops.bn_apply(c1, mr1, gb, Hc, g1, b1, True, False)
      sc = ops.conv_fwd(x, self._wk(r2), 1, bcout, bias=P[r2.b_off:], ldb=r2.ld_fwd)
      out = ops.conv_fwd(h2, self._wk(r1), 3, bcout, bias=P[r1.b_off:], residual=sc, res_shift=1, ldb=r1.ld_fwd)
      ctx["blocks"].append(dict(x=x, mr0=mr0, u=u, c1=c1, mr1=mr1, h2=h2, Hc=Hc))
      x = out
    bnf = ("LocalConditionalBatchNorm_0",)
    _, gf_, bf_ = self.bn_index[bnf]
    mrf = self._bn(x, bnf, stats, new_stats, train, grp)
    hf = ops.bn_apply(x, mrf, gbL, Hc, gf_, bf_, True, False)
    r = self.convs[(self.cpre + "_1",)]
    S = x.shape[1]
    img = ops.empty((B, S, S, 3), F32)
    ops._call("xmc_conv_c3_out", hf.data_ptr(), ops._f32(hf), self._wk(r).data_ptr(), r.ld_fwd, P[r.b_off:].data_ptr(),
              B, S, S, self.c_last, 3, 3, 1, 0, img.data_ptr(),
              fake_bf16.data_ptr() if fake_bf16 is not None else None, _lib.stream())
    ctx.update(x_last=x, mrf=mrf, hf=hf, img=img)
    return img, ctx

  # ------------------------------------------------------------------------------------------------------------
  def _conv_wgrad(self, rec, x_in, dy, grads):
    ops.wgrad(x_in, dy, rec.kh, grads[rec.w_off:], out_mode=1, ld_out=rec.cout, tap_stride=rec.cin * rec.cout)
    ops.colsum(dy, grads[rec.b_off:])

  def backward(self, ctx, d_img, params, grads):
    """Accumulates d(loss)/d(params) into the flat fp32 buffer `grads` given d(loss)/d(image) (fp32)."""
    with ops.act_dtype(self.act):
      self._backward(ctx, d_img, params, grads)

  def _backward(self, ctx, d_img, params, grads):
    S_ = self.S
    B, E, zd, cd, scd = ctx["B"], self.E, self.zd, self.cd, self.scd
    S = d_img.shape[1]
    gbC, gbL, R, Hc16 = ctx["gbC"], ctx["gbL"], ctx["R"], ctx["Hc"]
    grp = ctx["bn_group"]
    dgbC = ops.empty((B, self.NC), F32)  # every column is overwritten by its layer's bn_bwd
    dgbL = ops.empty((B * R, self.NL), F32)

    # output head: tanh, conv3x3 (C -> 3)
    dpre = ops.empty((B, S, S, 3))
    ops._call("xmc_tanh01_bwd", d_img.data_ptr(), ctx["img"].data_ptr(), d_img.numel(), dpre.data_ptr(), ops._f32(dpre),
              _lib.stream())
    r = self.convs[(self.cpre + "_1",)]
    C = self.c_last
    # the weight gradient of the 3-channel output conv runs on the tensor-core wgrad kernel over a zero-bordered
    # 8-channel copy of d(pre-tanh) (packed-window form: M = 3 kw x 8 channels per kh tap)
    if self.act == F32:   # fp32 activations: the CUDA-core kernel (fp32 FMAs)
      ops.wgrad_c3(dpre, ctx["hf"], 3, 1, C * 3, 1, 3, grads[r.w_off:])
    else:
      dpad = ops.c3_pad(dpre)
      ops.c3_wgrad(dpad, ctx["hf"], 1, C * 3, 1, 3, grads[r.w_off:])
    ops.colsum(dpre, grads[r.b_off:])
    # the input gradient stays on the CUDA-core kernel: as a GEMM it is a 3-k-iteration tile bound by the epilogue's
    # latency (measured 0.42 ms vs 0.32 ms per 112-image launch); ops.c3_conv is the tensor-core form
    dhf = ops.empty((B, S, S, C))
    ops._call("xmc_conv_c3_in", dpre.data_ptr(), ops._f32(dpre), self._wd(r).data_ptr(), r.ld_dg, None, B, S, S, C, 3, 3,
              0, dhf.data_ptr(), _lib.stream())
    _, gf_, bf_ = self.bn_index[("LocalConditionalBatchNorm_0",)]
    dout = ops.bn_bwd(dhf, ctx["x_last"], ctx["mrf"], gbL, dgbL, Hc16, gf_, bf_, True, False, group=grp)

    dx16_extra = None
    for (name, kind, bcin, bcout), sv in zip(reversed(self.blocks), reversed(ctx["blocks"])):
      gb, dgb = (gbC, dgbC) if kind == "cbn" else (gbL, dgbL)
      bn0 = (name, ("ConditionalBatchNorm_0" if kind == "cbn" else "LocalConditionalBatchNorm_0"))
      bn1 = (name, ("ConditionalBatchNorm_1" if kind == "cbn" else "LocalConditionalBatchNorm_1"))
      _, g0, b0 = self.bn_index[bn0]
      _, g1, b1 = self.bn_index[bn1]
      r0, r1, r2 = (self.convs[(name, f"{self.cpre}_{i}")] for i in range(3))
      if kind == "cbn" and dx16_extra is None:
        # leaving the spatial part: finish everything that hangs off the 16x16 tensor (attention, spatial cond)
        dx16_extra = True
        dout = self._backward_cond16(ctx, dgbL, dout, params, grads)
      # conv2 (Conv_1)
      self._conv_wgrad(r1, sv["h2"], dout, grads)
      dh2 = ops.conv_fwd(dout, self._wd(r1), 3, bcout, ldb=r1.ld_dg)
      dc1 = ops.bn_bwd(dh2, sv["c1"], sv["mr1"], gb, dgb, sv["Hc"], g1, b1, True, False, group=grp)
      # conv1 (Conv_0, sub-pixel form): weight gradient from the low-resolution input and the 2x gradient; input
      # gradient = 4x4 / stride-2 / pad-1 convolution over the gradient, directly at the input resolution
      ops.wgrad(sv["u"], dc1, 3, grads[r0.w_off:], out_mode=1, ld_out=bcout, tap_stride=bcin * bcout, subpixel=True)
      ops.colsum(dc1, grads[r0.b_off:])
      _, vd_off = self.subpixel[(name, self.cpre + "_0")]
      du = ops.conv_fwd(dc1, self.arena[vd_off:], 4, bcin, ldb=16 * bcout * S_, stride=2, pad=1)
      dxa = ops.bn_bwd(du, sv["x"], sv["mr0"], gb, dgb, sv["Hc"], g0, b0, True, False, group=grp)
      # shortcut (Conv_2 at low resolution)
      dsc = ops.pool2(dout, scale=1.0)
      self._conv_wgrad(r2, sv["x"], dsc, grads)
      dout = ops.conv_fwd(dsc, self._wd(r2), 1, bcin, residual=dxa, ldb=r2.ld_dg)

    # ---- ConditionalBatchNorm gamma/beta dense layers + global condition ----------------------------------------
    gc = ctx["gc"]
    dgbC_bf = ops.cast_to_bf16(dgbC)
    L = self.layout
    for prefix, Cc, goff, boff in self.cbn:
      for leaf, o in ((self.dpre + "_0", goff), (self.dpre + "_1", boff)):
        ops.wgrad(as4(gc), as4(dgbC_bf[:, o:o + Cc]), 1, grads[L.off(prefix + (leaf, "kernel")):], out_mode=0,
                  ld_out=Cc, tap_stride=cd * Cc)
    ops.colsum(dgbC_bf, grads[self.cbn_bias_off:])
    dgc = ops.conv_fwd(as4(dgbC_bf), self.arena[self.cbn_dg_off:], 1, cd, ldb=self.NC * S_, out_dtype=F32).view(B, cd)
    ops.sum_rows(ctx["dspatial"][:, E:], B, R, dgc, accumulate=True)
    dgc_bf = ops.cast_to_bf16(dgc)
    r = self.convs[(self.dpre + "_0",)]
    ops.wgrad(as4(ctx["cond_bf"]), as4(dgc_bf[:, :zd]), 1, grads[r.w_off:], out_mode=0, ld_out=zd, tap_stride=E * zd)
    ops.colsum(dgc_bf[:, :zd], grads[r.b_off:])
    # Dense_1 (z -> 4x4xC0); d(out) = input gradient of GenBlock_0
    r = self.convs[(self.dpre + "_1",)]
    dx0 = dout.view(B, r.cout)
    ops.wgrad(as4(gc[:, zd:]), as4(dx0), 1, grads[r.w_off:], out_mode=0, ld_out=r.cout, tap_stride=zd * r.cout)
    ops.colsum(dx0, grads[r.b_off:])

  def _backward_cond16(self, ctx, dgbL, dx16, params, grads):
    """Backward of everything attached to the 16x16 tensor: LCBN gamma/beta 1x1 convs (concatenated), spatial
    condition, word-region attention, Conv_0. Returns the total gradient wrt the GenBlock_1 output."""
    B, E, scd, R = ctx["B"], self.E, self.scd, ctx["R"]
    L = self.layout
    spatial = ctx["spatial"]
    dgbL_bf = ops.cast_to_bf16(dgbL)
    for prefix, Cc, goff, boff in self.lcbn:
      for leaf, o in ((self.cpre + "_0", goff), (self.cpre + "_1", boff)):
        ops.wgrad(as4(spatial), as4(dgbL_bf[:, o:o + Cc]), 1, grads[L.off(prefix + (leaf, "kernel")):], out_mode=0,
                  ld_out=Cc, tap_stride=scd * Cc)
    ops.colsum(dgbL_bf, grads[self.lcbn_bias_off:])
    dspatial = ops.conv_fwd(as4(dgbL_bf), self.arena[self.lcbn_dg_off:], 1, scd, ldb=self.NL * self.S).view(B * R, scd)
    ctx["dspatial"] = dspatial
    xq = ctx["xq"]
    dq = ops.attention_g_bwd(dspatial, xq.view(B, R, E), ctx["what"].view(B, ctx["Lw"], E), ctx["attn"], self.gamma)
    r = self.convs[(self.cpre + "_0",)]
    dq4 = dq.view(xq.shape)
    self._conv_wgrad(r, ctx["x16"], dq4, grads)
    return ops.conv_fwd(dq4, self._wd(r), 1, r.cin, residual=dx16, ldb=r.ld_dg)


# ======================================================================================================================
# Loss heads
# ======================================================================================================================
class Contrastive:
  """attention_lib.contrastive_loss (attention_lib.py:46-79) on fp32 features a (image_feat), b (cond_feat)."""

  def __init__(self, a, b, slot, temperature=0.1, stats=None):
    """stats: optional fp32[2] device view receiving (accuracy, entropy) of attention_lib.py:75-78."""
    self.inv_t = 1.0 / temperature
    self.ah, self.ainv = ops.l2norm_rows(a)
    self.bh, self.binv = ops.l2norm_rows(b)
    logits = ops.small_gemm_nt(self.ah, self.bh, self.inv_t)
    self.dlogits = ops.ce_sym(logits, slot)
    if stats is not None:
      ops.ce_stats(logits, stats)

  def bwd_a(self, out, accumulate=True):
    dah = ops.small_gemm_nn(self.dlogits, False, self.bh, self.inv_t)
    ops.l2norm_rows_bwd(dah, self.ah, self.ainv, out=out, accumulate=accumulate)

  def bwd_b(self, out, accumulate=True):
    dbh = ops.small_gemm_nn(self.dlogits, True, self.ah, self.inv_t)
    ops.l2norm_rows_bwd(dbh, self.bh, self.binv, out=out, accumulate=accumulate)


class WordShared:
  """Per-batch word tensors shared by the real and fake word_loss calls."""

  def __init__(self, words, max_len):
    B, Lw, E = words.shape
    self.B, self.Lw, self.E = B, Lw, E
    self.BL = B * Lw
    self.ldS = _r8(self.BL)
    self.words = words.reshape(self.BL, E).contiguous()
    self.what_bf, self.winv = ops.l2norm_rows(self.words, out_dtype=ops.ACT[0])
    self.max_len = max_len.reshape(B).contiguous()
    self._whatT = None

  def whatT(self):
    if self._whatT is None:
      self._whatT = ops.transpose_bf16(self.what_bf, self.ldS)
    return self._whatT


class WordLoss:
  """attention_lib.word_loss (attention_lib.py:130-191) with gamma1=gamma2=5, gamma3=50. R: [B, regions, E] bf16."""
  G1, G2, G3 = 5.0, 5.0, 50.0

  def __init__(self, R, ws, slot, stats=None):
    """stats: optional fp32[2] device view receiving (accuracy, entropy) of attention_lib.py:183-190."""
    B, Rn, E = R.shape
    self.ws, self.B, self.Rn, self.E = ws, B, Rn, E
    BL, ldS = ws.BL, ws.ldS
    padded = ldS != BL
    self.Rh, self.rinv = ops.l2norm_rows(R.reshape(B * Rn, E), out_dtype=ops.ACT[0])
    S = ops.empty((B * Rn, ldS), F32)
    ops.conv_fwd(as4(self.Rh), ws.what_bf, 1, BL, ldb=E, out=as4(S[:, :BL]))
    self.alpha = ops.empty((B * Rn, ldS))
    self.alphaT = ops.zeros((B, ldS, Rn), ops.ACT[0]) if padded else ops.empty((B, ldS, Rn))
    ops._call("xmc_wl_softmax", S.data_ptr(), B, Rn, BL, ldS, self.G1, self.alpha.data_ptr(), self.alphaT.data_ptr(),
              ops._f32(self.alpha), _lib.stream())
    self.ctx = ops.empty((B, ldS, E), F32)
    ops.wgrad(self.alpha.view(B, 1, Rn, ldS), self.Rh.view(B, 1, Rn, E), 1, self.ctx, out_mode=1, batched=True,
              ld_out=E, tap_stride=0, batch_stride=ldS * E)
    self.cos = ops.empty((B, BL), F32)
    self.cnorm = ops.empty((B, BL), F32)
    ops._call("xmc_wl_cos", self.ctx.data_ptr(), ldS * E, ws.words.data_ptr(), ws.winv.data_ptr(), B, ws.Lw, E,
              self.cos.data_ptr(), self.cnorm.data_ptr(), _lib.stream())
    self.sim = ops.empty((B, B), F32)
    self.pw = ops.empty((B, BL), F32)
    ops._call("xmc_wl_sim", self.cos.data_ptr(), ws.max_len.data_ptr(), B, ws.Lw, self.G2, self.G3,
              self.sim.data_ptr(), self.pw.data_ptr(), _lib.stream())
    self.dsim = ops.ce_sym(self.sim, slot)
    if stats is not None:
      ops.ce_stats(self.sim, stats)

  def bwd(self):
    """Returns d(loss)/dR as bf16 [B*regions, E]."""
    ws, B, Rn, E = self.ws, self.B, self.Rn, self.E
    BL, ldS = ws.BL, ws.ldS
    padded = ldS != BL
    dctx = ops.zeros((B, ldS, E), ops.ACT[0]) if padded else ops.empty((B, ldS, E))
    ops._call("xmc_wl_cos_bwd", self.dsim.data_ptr(), self.pw.data_ptr(), self.cos.data_ptr(), self.cnorm.data_ptr(),
              self.ctx.data_ptr(), ldS * E, ws.words.data_ptr(), ws.winv.data_ptr(), B, ws.Lw, E, self.G3,
              dctx.data_ptr(), ldS * E, ops._f32(dctx), _lib.stream())
    dalpha = ops.empty((B * Rn, ldS), F32)
    ops.conv_fwd(self.Rh.view(B, 1, Rn, E), dctx, 1, BL, ldb=E, batched=True, stride_b=ldS * E,
                 out=dalpha.view(B, 1, Rn, ldS)[..., :BL])
    dS = ops.empty((B * Rn, ldS))
    ops._call("xmc_wl_softmax_bwd", self.alpha.data_ptr(), dalpha.data_ptr(), B, Rn, BL, ldS, self.G1, dS.data_ptr(),
              ops._f32(dS), _lib.stream())
    tmp = ops.empty((B * Rn, E))
    ops.wgrad(self.alphaT.view(B, 1, ldS, Rn), dctx.view(B, 1, ldS, E), 1, tmp, out_mode=2, batched=True, ld_out=E,
              tap_stride=0, batch_stride=Rn * E)
    dRh = ops.conv_fwd(as4(dS), ws.whatT(), 1, E, ldb=ldS, residual=as4(tmp))
    return ops.l2norm_rows_bwd(dRh.view(B * Rn, E), self.Rh, self.rinv)


# ======================================================================================================================
# Discriminator
# ======================================================================================================================
class DiscriminatorEngine:
  """xmc_net.Discriminator (xmc_net.py:28-142)."""

  def __init__(self, config, embedding_dim=768):
    if config.image_size == 128:
      channel_dims, downs = [2, 4, 8, 16, 16], [True, True, True, True, False]
    elif config.image_size == 256:
      channel_dims, downs = [2, 4, 8, 8, 16, 16], [True, True, True, True, True, False]
    else:
      raise ValueError(f"image_size {config.image_size} is not supported (reference: xmc_net.py:81-86)")
    self.config = config
    self.act = act_dtype_of(config)          # see GeneratorEngine
    self.S = S_ = 3 if self.act == F32 else 1
    self.E = E = embedding_dim
    df = config.df_dim
    self.sn = bool(config.d_spectral_norm)
    self.cpre = "SpectralConv" if self.sn else "Conv"
    self.dpre = "SpectralDense" if self.sn else "Dense"
    self.cond_size = config.cond_size
    if df % 8:
      raise ValueError("df_dim must be a multiple of 8")
    L = self.layout = Layout()
    self.sntab = _SnTable() if self.sn else None
    self.u_layout = self.sntab.u_layout if self.sn else Layout()
    self.arena_size = 0
    self.prep = _PrepTable()
    self.convs = {}

    # convs whose output is 2x2-mean-pooled (the second conv of DiscOptimizedBlock and of every down-sampling
    # DiscBlock) run in the pool-fused form: one 4x4 / stride-2 convolution with summed weights (xmc_poolconv_prep),
    # 2.25x fewer FLOPs in forward, input gradient and weight gradient, no full-resolution output. path -> (wf4, wdg)
    self.poolconv = {}

    def add_conv(path, kh, cin_, cout_, dense=False, prep=True, pooled=False):
      shape = (cin_, cout_) if dense else (kh, kh, cin_, cout_)
      w_off = L.add(path + ("kernel",), shape)
      b_off = L.add(path + ("bias",), (cout_,))
      taps = kh * kh
      slot = self.sntab.add(path, w_off, taps * cin_, cout_) if self.sn else -1
      ld_fwd, ld_dg = _r8(S_ * taps * cin_), _r8(S_ * taps * cout_)
      fwd_off = dg_off = -1
      if pooled:
        wf4_off = self.arena_size
        self.arena_size += cout_ * 16 * cin_ * S_
        wdg_off = self.arena_size
        self.arena_size += 4 * cin_ * 4 * cout_ * S_
        self.poolconv[path] = (wf4_off, wdg_off)
        prep = False
      if prep:
        fwd_off = self.arena_size
        self.arena_size += _r8(cout_) * ld_fwd
        dg_off = self.arena_size
        self.arena_size += _r8(cin_) * ld_dg
        self.prep.add(w_off, taps, cin_, cout_, fwd_off, ld_fwd, dg_off, ld_dg, slot, split=S_ == 3)
      self.convs[path] = ConvRec(path, kh, cin_, cout_, w_off, b_off, fwd_off, ld_fwd, dg_off, ld_dg, slot)

    cp = self.cpre
    add_conv(("DiscOptimizedBlock_0", cp + "_0"), 3, 3, df)
    add_conv(("DiscOptimizedBlock_0", cp + "_1"), 3, df, df, pooled=True)
    add_conv(("DiscOptimizedBlock_0", cp + "_2"), 1, 3, df)
    self.blocks = []
    cin = df
    size = config.image_size // 2
    self.cond_channels = None
    for i, (cr, down) in enumerate(zip(channel_dims, downs)):
      cout = df * cr
      proj = down or cin != cout
      name = f"DiscBlock_{i}"
      add_conv((name, cp + "_0"), 3, cin, cout)
      add_conv((name, cp + "_1"), 3, cout, cout, pooled=down)
      if proj:
        add_conv((name, cp + "_2"), 1, cin, cout)
      if not down and i != len(channel_dims) - 1:
        raise NotImplementedError("a non-final DiscBlock without downsampling is not built")
      size = size // 2 if down else size
      is_cond = size == self.cond_size
      self.blocks.append((name, cin, cout, down, proj, is_cond))
      if is_cond:
        self.cond_channels = cout
      cin = cout
    self.c_last = cin
    add_conv((self.dpre + "_0",), 1, cin, 1, dense=True, prep=False)
    add_conv((self.dpre + "_1",), 1, E, cin, dense=True)
    if config.word_contrastive:
      if self.cond_channels is None:
        raise ValueError("no discriminator feature map of size cond_size")
      add_conv((cp + "_0",), 1, self.cond_channels, E)
    self.arena = None

  def init_params(self, seed):
    return init_flat(self.layout, seed, _kind).cuda(), init_flat(self.u_layout, seed + 1, _kind).cuda()

  def _ensure(self):
    if self.arena is None:
      self.arena = torch.zeros(self.arena_size, device="cuda", dtype=BF16)
      self.prep.upload()
      if self.sn:
        self.sntab.upload()

  def prep_weights(self, params, u0, u0_new):
    """Spectral normalisation (one power-iteration step, new u0 written to u0_new) + bf16 weight copies."""
    self._ensure()
    if self.sn:
      self.sntab.forward(params, u0, u0_new)
    ops._call("xmc_prep_weights", self.prep.dev.data_ptr(), self.prep.n, self.prep.tiles, params.data_ptr(),
              self.sntab.scalars.data_ptr() if self.sn else None, self.sntab.n if self.sn else 0,
              self.arena.data_ptr(), None, None, _lib.stream())
    for path, (wf4_off, wdg_off) in self.poolconv.items():
      rec = self.convs[path]
      scale = self.sntab.inv_sigma(rec.sn).data_ptr() if self.sn else None
      ops._call("xmc_poolconv_prep", params[rec.w_off:].data_ptr(), scale, rec.cin, rec.cout, int(self.S == 3),
                self.arena[wf4_off:].data_ptr(), self.arena[wdg_off:].data_ptr(), _lib.stream())

  def sn_backward(self, params, grads, u0_new):
    if self.sn:
      self.sntab.backward(params, grads, u0_new)

  def _wk(self, rec):
    return self.arena[rec.fwd_off:]

  def _wd(self, rec):
    return self.arena[rec.dg_off:]

  def _pooled_fwd(self, rec, x, bias, residual):
    """dsample(conv3x3(x) + bias) + residual as one 4x4 / stride-2 convolution; x full resolution, result half."""
    wf4_off, _ = self.poolconv[rec.path]
    return ops.conv_fwd(x, self.arena[wf4_off:], 4, rec.cout, bias=bias, residual=residual, ldb=16 * rec.cin * self.S,
                        stride=2, pad=1)

  def _pooled_dgrad(self, rec, dout, mask):
    """Input gradient of _pooled_fwd at full resolution from the half-resolution output gradient (sub-pixel form),
    multiplied by [mask > 0]."""
    _, wdg_off = self.poolconv[rec.path]
    return ops.conv_fwd(dout, self.arena[wdg_off:], 2, rec.cin, ldb=4 * rec.cout * self.S, pad=1, subpixel=True,
                        mask=mask)

  def _pooled_wgrad(self, rec, x, dout, grads):
    """Weight / bias gradient of _pooled_fwd: x full resolution, dout half resolution."""
    ops.wgrad(x, dout, 3, grads[rec.w_off:], out_mode=1, ld_out=rec.cout, tap_stride=rec.cin * rec.cout, alpha=0.25,
              subpixel=2)
    ops.colsum(dout, grads[rec.b_off:])   # d/db of the mean of four copies of b = column sum of the pooled gradient

  def _inv_sigma(self, rec):
    return self.sntab.inv_sigma(rec.sn) if self.sn else None

  # ------------------------------------------------------------------------------------------------------------
  def forward(self, params, images, batch, losses, need_g=True, stats=None):
    """images: [2B,S,S,3] in the activation dtype (real first). Fills `losses` (fp32[16]) slots, returns (logit fp32
    [2B], ctx). stats: optional fp32 [16,2] receiving (accuracy, entropy) per loss slot — the side statistics of
    attention_lib.get_statistics, dead on the train path and therefore off by default."""
    with ops.act_dtype(self.act):
      return self._forward(params, images, batch, losses, need_g, stats)

  def _forward(self, params, images, batch, losses, need_g, stats):
    st = (lambda name: stats[LOSS_SLOTS[name]]) if stats is not None else (lambda name: None)
    P = params
    cfg = self.config
    cp = self.cpre
    N2, S = images.shape[0], images.shape[1]
    B = N2 // 2
    E = self.E
    df = cfg.df_dim
    ctx = {"B": B, "need_g": need_g, "images": images}
    r0, r1, r2 = (self.convs[("DiscOptimizedBlock_0", cp + f"_{i}")] for i in range(3))
    xp = ops.empty((N2, S // 2, S // 2, 3))
    f32 = ops._f32(images)
    ops._call("xmc_pool2_small", images.data_ptr(), f32, N2, S // 2, S // 2, 3, 0.25, xp.data_ptr(), _lib.stream())
    sc = ops.empty((N2, S // 2, S // 2, df))
    ops._call("xmc_conv_c3_in", xp.data_ptr(), f32, self._wk(r2).data_ptr(), r2.ld_fwd, P[r2.b_off:].data_ptr(), N2,
              S // 2, S // 2, df, 1, 1, 0, sc.data_ptr(), _lib.stream())
    # first conv (3 -> df, 3x3): forward on the CUDA-core kernel (see GeneratorEngine.backward), its weight gradient
    # on the tensor-core wgrad kernel over the zero-bordered 8-channel image copy made here
    xpad = ops.c3_pad(images) if not f32 else None   # fp32 mode: the CUDA-core weight-gradient kernel reads `images`
    c1r = ops.empty((N2, S, S, df))
    ops._call("xmc_conv_c3_in", images.data_ptr(), f32, self._wk(r0).data_ptr(), r0.ld_fwd, P[r0.b_off:].data_ptr(), N2,
              S, S, df, 3, 3, 1, c1r.data_ptr(), _lib.stream())
    x = self._pooled_fwd(r1, c1r, P[r1.b_off:], sc)     # dsample(conv3x3(c1r)) + shortcut, never at full resolution
    xr = ops.relu(x)
    del sc
    ctx["b0"] = dict(xp=xp, c1r=c1r, xpad=xpad)
    ctx["blocks"] = []
    x_cond = None
    for name, cin, cout, down, proj, is_cond in self.blocks:
      q0, q1 = self.convs[(name, cp + "_0")], self.convs[(name, cp + "_1")]
      c1r = ops.conv_fwd(xr, self._wk(q0), 3, cout, bias=P[q0.b_off:], relu=True, ldb=q0.ld_fwd)
      if down:
        # dsample(conv1x1(x)) == conv1x1(dsample(x)) (a 1x1 convolution commutes with the mean): the shortcut runs at
        # half resolution; the main path's second conv absorbs its pooling (pool-fused form)
        q2 = self.convs[(name, cp + "_2")]
        xs = ops.pool2(x)
        scf = ops.conv_fwd(xs, self._wk(q2), 1, cout, bias=P[q2.b_off:], ldb=q2.ld_fwd)
        ctx["blocks"].append(dict(x=x, xr=xr, c1r=c1r, xs=xs))
        x = self._pooled_fwd(q1, c1r, P[q1.b_off:], scf)
        xr = ops.relu(x)
      else:
        if proj:
          q2 = self.convs[(name, cp + "_2")]
          scf = ops.conv_fwd(x, self._wk(q2), 1, cout, bias=P[q2.b_off:], ldb=q2.ld_fwd)
        else:
          scf = x
        ctx["blocks"].append(dict(x=x, xr=xr, c1r=c1r))
        x, xr = ops.conv_fwd(c1r, self._wk(q1), 3, cout, bias=P[q1.b_off:], residual=scf, ldb=q1.ld_fwd), None
      if is_cond:
        x_cond = x
    ctx["x_last"] = x
    xpool = ops.relu_sumhw(x)
    C = self.c_last
    cond_bf = ops.cast_to_bf16(batch["sentence_embedding"].reshape(B, E))
    rd0, rd1 = self.convs[(self.dpre + "_0",)], self.convs[(self.dpre + "_1",)]
    sent = ops.conv_fwd(as4(cond_bf), self._wk(rd1), 1, C, bias=P[rd1.b_off:], ldb=rd1.ld_fwd,
                        out_dtype=F32).view(B, C)
    logit = ops.empty(N2, F32)
    inv0 = self._inv_sigma(rd0)
    ops._call("xmc_proj_logit", xpool.data_ptr(), P[rd0.w_off:].data_ptr(), inv0.data_ptr() if inv0 is not None else None,
              P[rd0.b_off:].data_ptr(), sent.data_ptr(), N2, B, C, logit.data_ptr(), _lib.stream())
    dl_d, dl_g = ops.hinge(logit, B, losses[LOSS_SLOTS["hinge_d"]:], losses[LOSS_SLOTS["hinge_g"]:])
    ctx.update(xpool=xpool, cond_bf=cond_bf, sent=sent, logit=logit, dl_d=dl_d, dl_g=dl_g)
    real_feat, fake_feat = xpool[:B], xpool[B:]
    if cfg.sentence_contrastive:
      ctx["real_sent"] = Contrastive(real_feat, sent, losses[LOSS_SLOTS["real_sent"]:], stats=st("real_sent"))
      if need_g:
        ctx["fake_sent"] = Contrastive(fake_feat, sent, losses[LOSS_SLOTS["fake_sent"]:], stats=st("fake_sent"))
    if cfg.word_contrastive:
      rw = self.convs[(cp + "_0",)]
      xw = ops.conv_fwd(x_cond, self._wk(rw), 1, E, bias=P[rw.b_off:], ldb=rw.ld_fwd)
      Rn = xw.shape[1] * xw.shape[2]
      ws = WordShared(batch["embedding"], batch["max_len"])
      ctx["x_cond"] = x_cond
      ctx["xw_shape"] = xw.shape
      ctx["real_word"] = WordLoss(xw[:B].view(B, Rn, E), ws, losses[LOSS_SLOTS["real_word"]:], stats=st("real_word"))
      if need_g:
        ctx["fake_word"] = WordLoss(xw[B:].view(B, Rn, E), ws, losses[LOSS_SLOTS["fake_word"]:],
                                    stats=st("fake_word"))
    if cfg.image_contrastive and need_g:
      ctx["image"] = Contrastive(fake_feat, real_feat, losses[LOSS_SLOTS["image"]:], stats=st("image"))
    return logit, ctx

  # ------------------------------------------------------------------------------------------------------------
  def _wgrad(self, rec, x_in, dy, grads, dy_low=None):
    ops.wgrad(x_in, dy, rec.kh, grads[rec.w_off:], out_mode=1, ld_out=rec.cout, tap_stride=rec.cin * rec.cout)
    # the bias gradient of a conv whose output is average-pooled equals the column sum of the pooled gradient
    ops.colsum(dy if dy_low is None else dy_low, grads[rec.b_off:])

  def _backward_trunk(self, ctx, dout, sl, grads, d_xw, xw_sub, want_image_grad):
    with ops.act_dtype(self.act):
      return self._backward_trunk_impl(ctx, dout, sl, grads, d_xw, xw_sub, want_image_grad)

  def _backward_trunk_impl(self, ctx, dout, sl, grads, d_xw, xw_sub, want_image_grad):
    """Backward through the residual trunk for images `sl` (a slice of the 2B batch). dout: gradient wrt the last
    block output. d_xw: gradient wrt the word-feature map (bf16 [n,16,16,E]) of the images `xw_sub` (a slice relative
    to `sl`), or None. grads None -> dgrad only."""
    P = None
    cp = self.cpre
    wg = grads is not None
    for (name, cin, cout, down, proj, is_cond), sv in zip(reversed(self.blocks), reversed(ctx["blocks"])):
      x, xr, c1r = sv["x"][sl], sv["xr"][sl], sv["c1r"][sl]
      q0, q1 = self.convs[(name, cp + "_0")], self.convs[(name, cp + "_1")]
      if is_cond and d_xw is not None:
        rw = self.convs[(cp + "_0",)]
        if wg:
          self._wgrad(rw, ctx["x_cond"][sl][xw_sub], d_xw, grads)
        # in place on the sub-batch that has a word-loss gradient (elementwise read-then-write of the same address)
        ops.conv_fwd(d_xw, self._wd(rw), 1, rw.cin, residual=dout[xw_sub], out=dout[xw_sub], ldb=rw.ld_dg)
      if down:
        # pool-fused second conv and half-resolution shortcut (see _forward): nothing here touches a full-resolution
        # copy of the output gradient
        q2 = self.convs[(name, cp + "_2")]
        if wg:
          self._pooled_wgrad(q1, c1r, dout, grads)
          self._wgrad(q2, sv["xs"][sl], dout, grads)
        dc1 = self._pooled_dgrad(q1, dout, c1r)
        dxb = ops.unpool2(ops.conv_fwd(dout, self._wd(q2), 1, cin, ldb=q2.ld_dg), 0.25)
      else:
        if wg:
          self._wgrad(q1, c1r, dout, grads)
        dc1 = ops.conv_fwd(dout, self._wd(q1), 3, cout, mask=c1r, ldb=q1.ld_dg)
        if proj:
          q2 = self.convs[(name, cp + "_2")]
          if wg:
            self._wgrad(q2, x, dout, grads)
          dxb = ops.conv_fwd(dout, self._wd(q2), 1, cin, ldb=q2.ld_dg)
        else:
          dxb = dout
      if wg:
        self._wgrad(q0, xr, dc1, grads)
      dout = ops.conv_fwd(dc1, self._wd(q0), 3, cin, mask=xr, residual=dxb, ldb=q0.ld_dg)
    # ---- DiscOptimizedBlock_0 ------------------------------------------------------------------------------------
    r0, r1, r2 = (self.convs[("DiscOptimizedBlock_0", cp + f"_{i}")] for i in range(3))
    b0 = ctx["b0"]
    images = ctx["images"][sl]
    xp, c1r = b0["xp"][sl], b0["c1r"][sl]
    n, S = images.shape[0], images.shape[1]
    df = r1.cout
    if wg:
      self._pooled_wgrad(r1, c1r, dout, grads)
    dc1 = self._pooled_dgrad(r1, dout, c1r)
    if wg:
      if self.act == F32:
        ops.wgrad_c3(images, dc1, 3, 0, 3 * df, df, 1, grads[r0.w_off:])
      else:
        ops.c3_wgrad(b0["xpad"][sl], dc1, 0, 3 * df, df, 1, grads[r0.w_off:])
      ops.colsum(dc1, grads[r0.b_off:])
      ops.wgrad_c3(xp, dout, 1, 0, 3 * df, df, 1, grads[r2.w_off:])
      ops.colsum(dout, grads[r2.b_off:])
    if not want_image_grad:
      return None
    dimg = ops.empty((n, S, S, 3), F32)
    ops._call("xmc_conv_c3_out", dc1.data_ptr(), ops._f32(dc1), self._wd(r0).data_ptr(), r0.ld_dg, None, n, S, S, df, 3,
              3, 0, 0, dimg.data_ptr(), None, _lib.stream())
    dxp = ops.empty((n, S // 2, S // 2, 3), F32)
    ops._call("xmc_conv_c3_out", dout.data_ptr(), ops._f32(dout), self._wd(r2).data_ptr(), r2.ld_dg, None, n, S // 2,
              S // 2, df, 1, 1, 0, 0, dxp.data_ptr(), None, _lib.stream())
    ops._call("xmc_unpool2_add_f32", dxp.data_ptr(), n, S // 2, S // 2, 3, 0.25, dimg.data_ptr(), _lib.stream())
    return dimg

  def backward_d(self, ctx, params, grads):
    """d(d_loss)/d(params_d) accumulated into `grads` (holds d/dW~ for spectrally normalised kernels until
    sn_backward runs). d_loss = hinge_d + real_word + real_sentence (xmc_gan.py:146-153,237-241)."""
    with ops.act_dtype(self.act):
      self._backward_d(ctx, params, grads)

  def _backward_d(self, ctx, params, grads):
    P = params
    B, C, E = ctx["B"], self.c_last, self.E
    N2 = 2 * B
    rd0, rd1 = self.convs[(self.dpre + "_0",)], self.convs[(self.dpre + "_1",)]
    dxpool = ops.empty((N2, C), F32)
    dsent = ops.zeros((B, C), F32)
    inv0 = self._inv_sigma(rd0)
    ops._call("xmc_proj_logit_bwd", ctx["dl_d"].data_ptr(), ctx["xpool"].data_ptr(), P[rd0.w_off:].data_ptr(),
              inv0.data_ptr() if inv0 is not None else None, ctx["sent"].data_ptr(), 0, N2, B, C, dxpool.data_ptr(), 0,
              grads[rd0.w_off:].data_ptr(), grads[rd0.b_off:].data_ptr(), dsent.data_ptr(), _lib.stream())
    if "real_sent" in ctx:
      ctx["real_sent"].bwd_a(dxpool[:B])
      ctx["real_sent"].bwd_b(dsent)
    dsent_bf = ops.cast_to_bf16(dsent)
    ops.wgrad(as4(ctx["cond_bf"]), as4(dsent_bf), 1, grads[rd1.w_off:], out_mode=0, ld_out=C, tap_stride=E * C)
    ops.colsum_f32(dsent, grads[rd1.b_off:])
    dout = ops.relu_sumhw_bwd(ctx["x_last"], dxpool)
    d_xw = None
    if "real_word" in ctx:
      # only the real half has a word-loss gradient for d_loss
      shp = ctx["xw_shape"]
      d_xw = ctx["real_word"].bwd().view(B, shp[1], shp[2], E)
    self._backward_trunk(ctx, dout, slice(0, N2), grads, d_xw, slice(0, B), False)

  def backward_g(self, ctx, params):
    """d(g_loss)/d(fake images): fp32 [B,S,S,3]. g_loss's discriminator part = hinge_g + fake_word + fake_sentence +
    image_contrastive (xmc_gan.py:146-154). Only the fake half is pulled back, dgrad only."""
    with ops.act_dtype(self.act):
      return self._backward_g(ctx, params)

  def _backward_g(self, ctx, params):
    P = params
    B, C, E = ctx["B"], self.c_last, self.E
    rd0 = self.convs[(self.dpre + "_0",)]
    dxpool = ops.empty((2 * B, C), F32)
    inv0 = self._inv_sigma(rd0)
    ops._call("xmc_proj_logit_bwd", ctx["dl_g"].data_ptr(), ctx["xpool"].data_ptr(), P[rd0.w_off:].data_ptr(),
              inv0.data_ptr() if inv0 is not None else None, ctx["sent"].data_ptr(), B, B, B, C, dxpool.data_ptr(), 0,
              None, None, None, _lib.stream())
    dfake = dxpool[B:]
    if "fake_sent" in ctx:
      ctx["fake_sent"].bwd_a(dfake)
    if "image" in ctx:
      ctx["image"].bwd_a(dfake)
    sl = slice(B, 2 * B)
    dout = ops.relu_sumhw_bwd(ctx["x_last"][sl], dfake)
    d_xw = None
    if "fake_word" in ctx:
      shp = ctx["xw_shape"]
      d_xw = ctx["fake_word"].bwd().view(B, shp[1], shp[2], E)
    return self._backward_trunk(ctx, dout, sl, None, d_xw, slice(0, B), True)


# ======================================================================================================================
# Frozen ResNet-50 feature branch (pretrained_image_contrastive)
# ======================================================================================================================
RESNET50_STAGES = [3, 4, 6, 3]


class ResNetEngine:
  """resnet_v1.ResNet50 in eval mode (xmcgan/utils/resnet_v1.py:129-180) behind get_pretrained_embs
  (pretrained_model_utils.py:102-127): forward on [real; fake] images, input gradient for the fake half.
  Eval BatchNorm is folded into the weights / an fp32 bias once at construction.
  dtype: "float32" (default — the reference builds this network without a dtype, i.e. always fp32,
  pretrained_model_utils.py:87-91): fp32 activations, every convolution as three bf16 tensor-core passes over hi / lo
  splits of both operands (16 mantissa bits per operand, fp32 accumulation). "bfloat16": bf16
  activations and operands, 1/3 of the tensor work and half the activation traffic (a stated deviation)."""
  T, PAD_LO, TP = 224, 2, 229  # 7x7/2 SAME on 224: pad (2,3)

  def __init__(self, num_classes=1000, width=64, dtype="float32"):
    self.width, self.num_classes = width, num_classes
    if dtype not in ("float32", "bfloat16"):
      raise ValueError(f"dtype {dtype!r} is not supported (float32 or bfloat16)")
    self.act = F32 if dtype == "float32" else BF16
    self.S = S_ = 3 if self.act == F32 else 1
    L = self.layout = Layout()
    S = self.stats_layout = Layout()
    self.convs = collections.OrderedDict()  # path -> dict(kh, cin, cout, stride, bn)
    self.arena_size = 0
    self.prep = _PrepTable()
    self.cscale_size = 0

    def add_bn(path, c):
      L.add(path + ("scale",), (c,))
      L.add(path + ("bias",), (c,))
      S.add(path + ("mean",), (c,))
      S.add(path + ("var",), (c,))

    def add_conv(path, bn_path, kh, cin, cout, stride):
      w_off = L.add(path + ("kernel",), (kh, kh, cin, cout))
      add_bn(bn_path, cout)
      taps = kh * kh
      rec = dict(kh=kh, cin=cin, cout=cout, stride=stride, bn=bn_path, w_off=w_off, cs_off=self.cscale_size)
      self.cscale_size += _r4(cout)
      if cin >= 8:
        rec["ld_fwd"], rec["ld_dg"] = _r8(S_ * taps * cin), _r8(S_ * taps * cout)
        rec["fwd_off"] = self.arena_size
        self.arena_size += _r8(cout) * rec["ld_fwd"]
        rec["dg_off"] = self.arena_size
        self.arena_size += _r8(cin) * rec["ld_dg"]
        self.prep.add(w_off, taps, cin, cout, rec["fwd_off"], rec["ld_fwd"], rec["dg_off"], rec["ld_dg"], -1,
                      cscale_off=rec["cs_off"], split=S_ == 3)
      self.convs[path] = rec

    add_conv(("init_conv",), ("init_bn",), 7, 3, width, 2)
    self.blocks = []
    cin = width
    for si, nb in enumerate(RESNET50_STAGES):
      f = width * 2 ** si
      for bi in range(nb):
        pre = (f"stage{si + 1}", f"block{bi + 1}")
        stride = 2 if (si > 0 and bi == 0) else 1
        add_conv(pre + ("conv1",), pre + ("bn1",), 1, cin, f, 1)
        add_conv(pre + ("conv2",), pre + ("bn2",), 3, f, f, stride)
        add_conv(pre + ("conv3",), pre + ("bn3",), 1, f, 4 * f, 1)
        proj = cin != 4 * f or stride == 2
        if proj:
          add_conv(pre + ("proj_conv",), pre + ("proj_bn",), 1, cin, 4 * f, stride)
        self.blocks.append((pre, cin, f, stride, proj))
        cin = 4 * f
    self.c_last = cin
    self.head_w = L.add(("head", "kernel"), (cin, num_classes))
    self.head_b = L.add(("head", "bias"), (num_classes,))
    if num_classes % 8:
      raise ValueError("num_classes must be a multiple of 8")
    self.head_fwd = self.arena_size
    self.arena_size += num_classes * cin * S_
    self.head_dg = self.arena_size
    self.arena_size += cin * num_classes * S_
    self.prep.add(self.head_w, 1, cin, num_classes, self.head_fwd, cin * S_, self.head_dg, num_classes * S_, -1,
                  split=S_ == 3)
    self.stem_off = self.arena_size          # packed stem weights [width][7*56]: k = kh*56 + kw*8 + c
    self.arena_size += width * 392
    self.stem_lo_off = self.arena_size       # fp32 mode: the bf16 remainders of the stem weights, same packing
    self.arena_size += width * 392
    # transposed stem matrix for the input gradient: [160 rows = (kh*7+kw)*3+c, 147 used][K = width (x S)]
    self.stemT_rows = 160
    self.stemT_off = self.arena_size
    self.arena_size += self.stemT_rows * width * S_

  def random_variables(self, seed=0, head_scale=0.05, residual_scale=0.3):
    """Synthetic frozen weights for benchmarking (the reference's data/resnet_pretrained.npy is not shipped,
    README.md:60-63): He-normal kernels, BatchNorm scale ~ 1 (bn3 scaled down so the residual stream stays O(1)),
    small random bias / mean, var ~ 1 and a non-zero head (the reference's zero-initialised head, resnet_v1.py:171,
    would make the loss the constant 2 log B). Returns {"params","batch_stats"} nested dicts of CPU tensors."""
    g = torch.Generator().manual_seed(seed)

    def leaf(path, shape):
      k = path[-1]
      if k == "kernel" and len(shape) == 4:
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[0] * shape[1] * shape[2]))
      if k == "kernel":
        return torch.randn(shape, generator=g) * head_scale
      if k == "scale":
        t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        return t * residual_scale if path[-2] == "bn3" else t
      if k == "var":
        return 1.0 + 0.2 * torch.rand(shape, generator=g)
      return 0.1 * torch.randn(shape, generator=g)

    out = {}
    for name, layout in (("params", self.layout), ("batch_stats", self.stats_layout)):
      root = {}
      for path, (_, shape) in layout.entries.items():
        node = root
        for k in path[:-1]:
          node = node.setdefault(k, {})
        node[path[-1]] = leaf(path, shape)
      out[name] = root
    return out

  def load(self, variables):
    """variables: {"params": tree, "batch_stats": tree} (Flax names). One-time set-up: folds BatchNorm, writes the
    bf16 weight arena. (Set-up only — torch is used here for the tiny per-channel fold, never on the step path.)"""
    self.params = torch.zeros(self.layout.total, device="cuda")
    self.stats = torch.zeros(self.stats_layout.total, device="cuda")
    self.layout.load_tree(self.params, variables["params"])
    self.stats_layout.load_tree(self.stats, variables["batch_stats"])
    self.arena = torch.zeros(self.arena_size, device="cuda", dtype=BF16)
    self.cscale = torch.ones(self.cscale_size, device="cuda")
    self.fbias = torch.zeros(self.cscale_size, device="cuda")
    L, S = self.layout, self.stats_layout
    for path, rec in self.convs.items():
      bn = rec["bn"]
      s = L.view(self.params, bn + ("scale",)) * torch.rsqrt(S.view(self.stats, bn + ("var",)) + 1e-5)
      b = L.view(self.params, bn + ("bias",)) - S.view(self.stats, bn + ("mean",)) * s
      self.cscale[rec["cs_off"]:rec["cs_off"] + rec["cout"]] = s
      self.fbias[rec["cs_off"]:rec["cs_off"] + rec["cout"]] = b
    self.prep.upload()
    ops._call("xmc_prep_weights", self.prep.dev.data_ptr(), self.prep.n, self.prep.tiles, self.params.data_ptr(), None,
              0, self.arena.data_ptr(), None, self.cscale.data_ptr(), _lib.stream())
    # stem: [co][kh*56 + kw*8 + c] = W[kh,kw,c,co] * s[co]  (zeros for c >= 3): per output pixel and kh the 7 kw taps
    # x 8 padded channels are ONE contiguous 56-element run of the zero-bordered 8-channel image
    rec = self.convs[("init_conv",)]
    w = L.view(self.params, ("init_conv", "kernel")) * self.cscale[rec["cs_off"]:rec["cs_off"] + self.width]
    n_stem = self.width * 392
    wt = w.permute(3, 0, 1, 2)
    hi = wt.to(BF16)
    packed = torch.zeros(self.width, 7, 7, 8, device="cuda", dtype=BF16)
    packed[:, :, :, :3] = hi
    if self.act == F32:
      # fp32 mode: the resized image carries [hi(3) | lo(3) | 0 0] per pixel (xmc_resize_bilinear_pad split = 1).
      # pass 1 multiplies both parts with w_hi (channels 0-2 and 3-5), pass 2 adds hi * w_lo (channels 0-2 only).
      packed[:, :, :, 3:6] = hi
      lo = torch.zeros(self.width, 7, 7, 8, device="cuda", dtype=BF16)
      lo[:, :, :, :3] = (wt - hi.float()).to(BF16)
      self.arena[self.stem_lo_off:self.stem_lo_off + n_stem] = lo.reshape(-1)
    self.arena[self.stem_off:self.stem_off + n_stem] = packed.reshape(-1)
    # input gradient of the stem as a GEMM: rows (kh, kw, c), K = output channels; fp32 mode: [hi | hi | lo] along K
    wT = torch.zeros(self.stemT_rows, self.width, device="cuda")
    wT[:147] = w.reshape(147, self.width)
    hiT = wT.to(BF16)
    parts = [hiT] if self.act != F32 else [hiT, hiT, (wT - hiT.float()).to(BF16)]
    self.arena[self.stemT_off:self.stemT_off + self.stemT_rows * self.width * self.S] = torch.cat(parts, 1).reshape(-1)
    torch.cuda.synchronize()

  def _bias(self, rec):
    return self.fbias[rec["cs_off"]:]

  def stem_forward(self, images_f32):
    """bilinear resize to 224 -> 7x7/2 conv (+ folded init_bn, no ReLU) -> 3x3/2 max-pool. Returns (stem, pooled)."""
    N, S = images_f32.shape[0], images_f32.shape[1]
    T, TP, W0 = self.T, self.TP, self.width
    f32 = self.act == F32
    xpad = ops.empty((N, TP, TP, 8), BF16)
    ops._call("xmc_resize_bilinear_pad", images_f32.data_ptr(), N, S, T, TP, self.PAD_LO, int(f32), xpad.data_ptr(),
              _lib.stream())
    rec = self.convs[("init_conv",)]
    view = dict(Hout=T // 2, Wout=T // 2, KH=7, KW=1, strideH=2, strideW=1, Hin=TP, Win=T // 2, pitchW=16,
                pitchH=TP * 8, pitchN=TP * TP * 8)
    stem = ops.conv_fwd(xpad, self.arena[self.stem_off:], 7, W0, bias=self._bias(rec), ldb=392, c=56, view=view,
                        pre_split=f32)
    if f32:   # second pass: + image_hi * w_lo, added through the residual input (same element read then written)
      ops.conv_fwd(xpad, self.arena[self.stem_lo_off:], 7, W0, ldb=392, c=56, view=view, pre_split=True, residual=stem,
                   out=stem)
    x = ops.empty((N, T // 4, T // 4, W0), self.act)
    ops._call("xmc_maxpool3s2", stem.data_ptr(), ops._f32(stem), N, T // 2, W0, x.data_ptr(), _lib.stream())
    return stem, x

  def stem_backward(self, dpool, stem, pooled, S, d_images):
    """d(loss)/d(pooled) bf16 -> accumulates d(loss)/d(images) (fp32 [n,S,S,3]) through max-pool, stem and resize."""
    n = dpool.shape[0]
    T, W0 = self.T, self.width
    dstem = ops.empty((n, T // 2, T // 2, W0), self.act)
    ops._call("xmc_maxpool3s2_bwd", dpool.data_ptr(), stem.data_ptr(), pooled.data_ptr(), ops._f32(stem), n, T // 2, W0,
              dstem.data_ptr(), _lib.stream())
    # transposed 7x7/2 convolution = tensor-core GEMM (dy x W^T, 147 columns per stem pixel) + col2im gather
    with ops.act_dtype(self.act):
      cols = ops.conv_fwd(dstem, self.arena[self.stemT_off:], 1, self.stemT_rows, ldb=W0 * self.S, out_dtype=F32)
    d224 = ops.empty((n, T, T, 3), F32)
    ops._call("xmc_stem_col2im", cols.data_ptr(), n, T, T // 2, self.stemT_rows, self.PAD_LO, d224.data_ptr(),
              _lib.stream())
    ops._call("xmc_resize_bilinear_bwd", d224.data_ptr(), n, S, T, d_images.data_ptr(), _lib.stream())

  def block_forward(self, x, spec, x_pair=None, want_pair=False, want_f32=True):
    """One bottleneck block. bf16 mode: x -> out (bf16 tensors). fp32 mode: activations travel between the convolutions
    as two-part bf16 operands [.., hi | lo] written by the producing epilogue and consumed as A operand (K-chunk
    remap), residual (hi + lo) and mask (sign of hi); fp32 tensors are only materialised where asked (want_f32: the
    block output, e.g. for the pooling head). One split pass of the block input at most (none with x_pair).
    Returns (out, stash[, out_pair if want_pair]); the stash holds what the backward needs (pairs in fp32 mode)."""
    pre, cin, f, stride, proj = spec
    r1c, r2c, r3c = (self.convs[pre + (f"conv{i}",)] for i in (1, 2, 3))
    w = lambda rec, key="fwd_off": self.arena[rec[key]:]
    if self.act != F32:
      r1 = ops.conv_fwd(x, w(r1c), 1, f, bias=self._bias(r1c), relu=True, ldb=r1c["ld_fwd"])
      r2 = ops.conv_fwd(r1, w(r2c), 3, f, bias=self._bias(r2c), relu=True, ldb=r2c["ld_fwd"], stride=stride)
      sc = x
      if proj:
        pc = self.convs[pre + ("proj_conv",)]
        sc = ops.conv_fwd(x, w(pc), 1, 4 * f, bias=self._bias(pc), ldb=pc["ld_fwd"], stride=stride)
      out = ops.conv_fwd(r2, w(r3c), 1, 4 * f, bias=self._bias(r3c), residual=sc, relu=True, ldb=r3c["ld_fwd"])
      stash = dict(x=x, r1=r1, r2=r2)
      return (out, stash, None) if want_pair else (out, stash)
    if x_pair is None:
      x_pair = ops._split_nhwc(x, pair=True)
    pair_only = dict(want_pair=True, want_f32=False)
    _, r1p = ops.conv_fwd(None, w(r1c), 1, f, bias=self._bias(r1c), relu=True, ldb=r1c["ld_fwd"], x_pair=x_pair,
                          **pair_only)
    _, r2p = ops.conv_fwd(None, w(r2c), 3, f, bias=self._bias(r2c), relu=True, ldb=r2c["ld_fwd"], stride=stride,
                          x_pair=r1p, **pair_only)
    res_kw = dict(residual_pair=x_pair)
    if proj:
      pc = self.convs[pre + ("proj_conv",)]
      _, scp = ops.conv_fwd(None, w(pc), 1, 4 * f, bias=self._bias(pc), ldb=pc["ld_fwd"], stride=stride, x_pair=x_pair,
                            **pair_only)
      res_kw = dict(residual_pair=scp)
    out, out_pair = ops.conv_fwd(None, w(r3c), 1, 4 * f, bias=self._bias(r3c), relu=True, ldb=r3c["ld_fwd"], x_pair=r2p,
                                 want_pair=True, want_f32=want_f32, **res_kw)
    stash = dict(x=x_pair, r1=r1p, r2=r2p)
    return (out, stash, out_pair) if want_pair else (out, stash)

  def _zero_insert(self, t, c):
    """z[n,2h,2w,:] = t[n,h,w,:], zero elsewhere (the transpose of a stride-2 sampling), on a [.., c]-channel tensor."""
    n, h, w = t.shape[0], t.shape[1], t.shape[2]
    z = ops.empty((n, 2 * h, 2 * w, c), t.dtype)
    ops._call("xmc_zero_insert2", t.data_ptr(), ops._f32(t), n, h, w, c, z.data_ptr(), _lib.stream())
    return z

  def block_backward(self, g, x, r1, r2, spec, mask_input, g_pair=None, want_pair=False, want_f32=True):
    """g: gradient wrt the block output, already multiplied by [output > 0]; x, r1, r2: the forward's stash. Returns the
    gradient wrt the block input (multiplied by [input > 0] when mask_input: the input is the previous block's relu
    output). fp32 mode: gradients and masks are two-part operands as in block_forward; the zero insertion of the
    stride-2 transposes acts on the operand directly (a pure copy)."""
    pre, cin, f, stride, proj = spec
    r1c, r2c, r3c = (self.convs[pre + (f"conv{i}",)] for i in (1, 2, 3))
    w = lambda rec: self.arena[rec["dg_off"]:]
    if self.act != F32:
      dr2 = ops.conv_fwd(g, w(r3c), 1, f, mask=r2, ldb=r3c["ld_dg"])
      if stride == 2:
        dr1 = ops.conv_fwd(self._zero_insert(dr2, f), w(r2c), 3, f, mask=r1, ldb=r2c["ld_dg"], pad=2, alg_scale=0.25)
      else:
        dr1 = ops.conv_fwd(dr2, w(r2c), 3, f, mask=r1, ldb=r2c["ld_dg"])
      sg = g
      if proj:
        pc = self.convs[pre + ("proj_conv",)]
        g_in = self._zero_insert(g, 4 * f) if stride == 2 else g
        sg = ops.conv_fwd(g_in, w(pc), 1, cin, ldb=pc["ld_dg"], alg_scale=0.25 if stride == 2 else 1.0)
      dx = ops.conv_fwd(dr1, w(r1c), 1, cin, residual=sg, ldb=r1c["ld_dg"], mask=x if mask_input else None,
                        mask_last=True)
      return (dx, None) if want_pair else dx
    if g_pair is None:
      g_pair = ops._split_nhwc(g, pair=True)
    pair_only = dict(want_pair=True, want_f32=False)
    _, dr2p = ops.conv_fwd(None, w(r3c), 1, f, mask_pair=r2, ldb=r3c["ld_dg"], x_pair=g_pair, **pair_only)
    if stride == 2:
      _, dr1p = ops.conv_fwd(None, w(r2c), 3, f, mask_pair=r1, ldb=r2c["ld_dg"], pad=2, alg_scale=0.25,
                             x_pair=self._zero_insert(dr2p, 2 * f), **pair_only)
    else:
      _, dr1p = ops.conv_fwd(None, w(r2c), 3, f, mask_pair=r1, ldb=r2c["ld_dg"], x_pair=dr2p, **pair_only)
    res_kw = dict(residual_pair=g_pair)
    if proj:
      pc = self.convs[pre + ("proj_conv",)]
      gp = self._zero_insert(g_pair, 8 * f) if stride == 2 else g_pair
      _, sgp = ops.conv_fwd(None, w(pc), 1, cin, ldb=pc["ld_dg"], alg_scale=0.25 if stride == 2 else 1.0, x_pair=gp,
                            **pair_only)
      res_kw = dict(residual_pair=sgp)
    dx, dxp = ops.conv_fwd(None, w(r1c), 1, cin, ldb=r1c["ld_dg"], mask_pair=x if mask_input else None, mask_last=True,
                           x_pair=dr1p, want_pair=True, want_f32=want_f32, **res_kw)
    return (dx, dxp) if want_pair else dx

  def forward(self, images_f32):
    """images_f32: fp32 [N,S,S,3] in [0,1]. Returns (logits fp32 [N,num_classes], ctx)."""
    with ops.act_dtype(self.act):
      return self._forward(images_f32)

  def _forward(self, images_f32):
    N, S = images_f32.shape[0], images_f32.shape[1]
    stem, x = self.stem_forward(images_f32)
    ctx = {"N": N, "S": S, "stem": stem, "pool0": x, "blocks": []}
    xp = None
    for i, spec in enumerate(self.blocks):
      # fp32 mode: only the last block's output is materialised in fp32 (for the pooling head)
      x, sv, xp = self.block_forward(x, spec, x_pair=xp, want_pair=True, want_f32=i == len(self.blocks) - 1)
      ctx["blocks"].append(sv)
    ctx["x_last"] = x
    feat = ops.relu_sumhw(x)  # the block output is already >= 0: this is the plain spatial sum
    feat_bf = ops.cast_to_bf16(feat)
    hw = x.shape[1] * x.shape[2]
    logits = ops.conv_fwd(as4(feat_bf), self.arena[self.head_fwd:], 1, self.num_classes,
                          bias=self.params[self.head_b:], ldb=self.c_last * self.S, alpha=1.0 / hw,
                          out_dtype=F32).view(N, self.num_classes)
    ctx["hw"] = hw
    return logits, ctx

  def backward(self, ctx, dlogits, n0, d_images):
    """dlogits: fp32 [n, num_classes] for images [n0, n0+n). Accumulates d(loss)/d(images) into d_images fp32
    [n,S,S,3] (the 128-px images, i.e. through the bilinear resize as well)."""
    with ops.act_dtype(self.act):
      self._backward(ctx, dlogits, n0, d_images)

  def _backward(self, ctx, dlogits, n0, d_images):
    n = dlogits.shape[0]
    sl = slice(n0, n0 + n)
    ncp = self.num_classes
    dl_bf = ops.cast_to_bf16(dlogits.reshape(n, ncp))
    dfeat = ops.conv_fwd(as4(dl_bf), self.arena[self.head_dg:], 1, self.c_last, ldb=ncp * self.S, alpha=1.0 / ctx["hw"],
                         out_dtype=F32).view(n, self.c_last)
    dout = ops.relu_sumhw_bwd(ctx["x_last"][sl], dfeat)   # includes the relu mask of the last block output
    dpair = None
    for i in range(len(self.blocks) - 1, -1, -1):
      sv = ctx["blocks"][i]
      # the first block's input (max-pool output) is not a relu output (no ReLU after init_bn, resnet_v1.py:146-154)
      dout, dpair = self.block_backward(dout, sv["x"][sl], sv["r1"][sl], sv["r2"][sl], self.blocks[i],
                                        mask_input=i > 0, g_pair=dpair, want_pair=True, want_f32=i == 0)
    self.stem_backward(dout, ctx["stem"][sl], ctx["pool0"][sl], ctx["S"], d_images)
