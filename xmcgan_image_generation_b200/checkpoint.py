"""Checkpoint interchange with the reference: the reference saves
`flax.serialization.to_bytes(TrainState)` as `<workdir>/checkpoints/ckpt-<N>.flax` through clu.checkpoint
(xmcgan/train_utils.py:370-376,457-459; clu==0.0.3, flax==0.3.3 — neither is vendored, the format below is restated from
flax/serialization.py of that release).

Wire format: msgpack (use_bin_type, strict_types) of the state dict; every array leaf is ExtType(1, msgpack((shape,
dtype.name, C-order bytes))), numpy scalars are ExtType(3, same tuple). State dict of the TrainState dataclass
(train_utils.py:42-50): {"step", "g_optimizer", "d_optimizer", "generator_state", "discriminator_state",
"ema_params"}; a flax.optim.Optimizer serialises as {"target": params, "state": {"step": int32, "param_states": the
params tree with leaves {"grad_ema", "grad_sq_ema"}}} (flax/optim/base.py, flax/optim/adam.py).

Direction: a `ckpt-N.flax` written by the reference restores here; a directory written HERE holds only the
`ckpt-N.flax` files — the reference's clu.checkpoint locates checkpoints through tf.train.CheckpointManager (the
`checkpoint` state file + `ckpt-N.index/.data`), which are not written, so the reference's restore_or_initialize does
not see it unless pointed at the file (flax.serialization.from_bytes on its bytes works: same state dict). The
top-level `step` is written as an int32 array, the reference's initial TrainState.step is a Python int: restores
identically, not byte-identical.

Pure host code (numpy + msgpack): moves state between HBM buffers and files, never on the step path. PARITY UNPINNED by
upstream: the reference ships no checkpoint fixture; tests pin the byte layout of a small tree by hand."""
import os
import re

import msgpack
import numpy as np

EXT_NDARRAY, EXT_NATIVE_COMPLEX, EXT_NPSCALAR = 1, 2, 3


def _array_payload(arr):
  arr = np.asarray(arr)
  return msgpack.packb((arr.shape, arr.dtype.name, arr.tobytes("C")), use_bin_type=True)


def _ext_pack(x):
  if isinstance(x, np.ndarray):
    return msgpack.ExtType(EXT_NDARRAY, _array_payload(x))
  if isinstance(x, np.generic):
    return msgpack.ExtType(EXT_NPSCALAR, _array_payload(x))
  if isinstance(x, complex):
    return msgpack.ExtType(EXT_NATIVE_COMPLEX, msgpack.packb((x.real, x.imag)))
  raise TypeError(f"cannot serialise {type(x)}")


def _ext_unpack(code, data):
  if code in (EXT_NDARRAY, EXT_NPSCALAR):
    shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
    arr = np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()  # writable, owns its memory
    return arr if code == EXT_NDARRAY else arr[()]
  if code == EXT_NATIVE_COMPLEX:
    re_, im = msgpack.unpackb(data)
    return complex(re_, im)
  return msgpack.ExtType(code, data)


def msgpack_serialize(tree):
  """flax.serialization.msgpack_serialize: nested dicts of numpy arrays / scalars -> bytes."""
  return msgpack.packb(tree, default=_ext_pack, strict_types=True, use_bin_type=True)


def msgpack_restore(data):
  """flax.serialization.msgpack_restore: bytes -> nested dicts of numpy arrays."""
  return msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False)


# ----------------------------------------------------------------------------------------------------------------------
def _np_tree(flat_tree):
  """FlatTree (device) -> nested dict of numpy fp32 arrays."""
  def rec(n):
    return {k: rec(v) for k, v in n.items()} if isinstance(n, dict) else n.numpy()
  return rec(flat_tree.to_cpu_tree())


def _adam_tree(layout, m, v):
  root = {}
  mt, vt = layout.tree(m.detach().cpu()), layout.tree(v.detach().cpu())

  def rec(a, b):
    return {k: rec(a[k], b[k]) for k in a} if isinstance(a, dict) else {"grad_ema": a.numpy(), "grad_sq_ema": b.numpy()}
  return rec(mt, vt)


def _optimizer_state_dict(opt):
  return {"target": _np_tree(opt.target),
          "state": {"step": np.asarray(opt.step, np.int32), "param_states": _adam_tree(opt.target.layout, opt.m, opt.v)}}


def to_state_dict(state):
  """TrainState -> the reference's state dict (numpy leaves, Flax variable names)."""
  return {
      "step": np.asarray(state.step, np.int32),
      "g_optimizer": _optimizer_state_dict(state.g_optimizer),
      "d_optimizer": _optimizer_state_dict(state.d_optimizer),
      "generator_state": {k: _np_tree(v) for k, v in (state.generator_state or {}).items()},
      "discriminator_state": {k: _np_tree(v) for k, v in (state.discriminator_state or {}).items()},
      "ema_params": _np_tree(state.ema_params),
  }


def to_bytes(state):
  return msgpack_serialize(to_state_dict(state))


def _load(flat_tree, tree):
  flat_tree.layout.load_tree(flat_tree.buf, tree)


def _load_adam(opt, sd):
  _load(opt.target, sd["target"])
  opt.set_step(int(np.asarray(sd["state"]["step"])))  # host counter and, in graph mode, the device counter
  lay = opt.target.layout
  pick = lambda t, leaf: ({k: pick(v, leaf) for k, v in t.items()} if leaf not in t else t[leaf])
  lay.load_tree(opt.m, pick(sd["state"]["param_states"], "grad_ema"))
  lay.load_tree(opt.v, pick(sd["state"]["param_states"], "grad_sq_ema"))


def from_state_dict(state, sd):
  """Fills `state` (a TrainState of the right configuration, e.g. from create_train_state) in place and returns it,
  like flax.serialization.from_state_dict(target, state). Missing / unexpected leaves raise KeyError / ValueError."""
  _load_adam(state.g_optimizer, sd["g_optimizer"])
  _load_adam(state.d_optimizer, sd["d_optimizer"])
  for coll, tree in (state.generator_state or {}).items():
    _load(tree, sd["generator_state"][coll])
  for coll, tree in (state.discriminator_state or {}).items():
    _load(tree, sd["discriminator_state"][coll])
  _load(state.ema_params, sd["ema_params"])
  state.step = int(np.asarray(sd["step"]))
  return state


def from_bytes(state, data):
  return from_state_dict(state, msgpack_restore(data))


def save_checkpoint(checkpoint_dir, state):
  """clu.checkpoint naming: <dir>/ckpt-<N>.flax with N counting saved checkpoints from 1."""
  os.makedirs(checkpoint_dir, exist_ok=True)
  nums = [int(m.group(1)) for f in os.listdir(checkpoint_dir) for m in [re.fullmatch(r"ckpt-(\d+)\.flax", f)] if m]
  path = os.path.join(checkpoint_dir, f"ckpt-{max(nums, default=0) + 1}.flax")
  tmp = path + ".tmp"
  with open(tmp, "wb") as f:
    f.write(to_bytes(state))
  os.replace(tmp, path)
  return path


def latest_checkpoint(checkpoint_dir):
  if not os.path.isdir(checkpoint_dir):
    return None
  nums = [int(m.group(1)) for f in os.listdir(checkpoint_dir) for m in [re.fullmatch(r"ckpt-(\d+)\.flax", f)] if m]
  return os.path.join(checkpoint_dir, f"ckpt-{max(nums)}.flax") if nums else None


def restore_checkpoint(state, path):
  with open(path, "rb") as f:
    return from_bytes(state, f.read())
