"""Checkpoint interchange: the Flax msgpack wire format of flax==0.3.3 (the reference's pin) and
the TrainState state-dict structure of xmcgan/train_utils.py:42-50 + flax.optim. Host-only code: runs without a GPU on
CPU-resident state buffers."""
import numpy as np
import torch

from tests import helpers
from xmcgan_image_generation_b200 import checkpoint as ck
from xmcgan_image_generation_b200 import engine, train_utils
from xmcgan_image_generation_b200.nets import xmc_net


def test_msgpack_byte_layout_known_answer():
  """Hand-assembled from the published format: map{ "a": ext8(type 1){[shape],dtype-name,bin}, "s": ext8(type 3){...} }."""
  got = ck.msgpack_serialize({"a": np.arange(3, dtype=np.float32), "s": np.int32(7)})
  want = (b"\x82"                                            # fixmap, 2 entries
          b"\xa1a" b"\xc7\x19\x01"                           # "a": ext8, 25 payload bytes, type 1 = ndarray
          b"\x93" b"\x91\x03" b"\xa7float32" b"\xc4\x0c"     # (shape=[3], "float32", bin8 of 12 bytes)
          + np.arange(3, dtype="<f4").tobytes() +
          b"\xa1s" b"\xc7\x0e\x03"                           # "s": ext8, 14 payload bytes, type 3 = numpy scalar
          b"\x93" b"\x90" b"\xa5int32" b"\xc4\x04" + np.int32(7).tobytes())
  assert got == want
  back = ck.msgpack_restore(got)
  assert back["a"].dtype == np.float32 and back["a"].tolist() == [0.0, 1.0, 2.0]
  assert back["s"] == 7 and back["s"].dtype == np.int32


def _cpu_state(cfg, seed):
  g, d, g_vars, d_vars = helpers.cpu_variables(cfg, 64, seed)
  flat = lambda lay, tree: xmc_net.FlatTree(lay, _flat(lay, tree))
  gp, dp = flat(g.layout, g_vars["params"]), flat(d.layout, d_vars["params"])
  g_state = {"batch_stats": flat(g.stats_layout, g_vars["batch_stats"])}
  if g.sn:
    g_state["spectral_norm_stats"] = flat(g.u_layout, g_vars["spectral_norm_stats"])
  state = train_utils.TrainState(3, train_utils.Optimizer(gp, cfg.g_lr, cfg.beta1, cfg.beta2),
                                 train_utils.Optimizer(dp, cfg.d_lr, cfg.beta1, cfg.beta2), g_state,
                                 {"spectral_norm_stats": flat(d.u_layout, d_vars["spectral_norm_stats"])}, gp.clone())
  gen = torch.Generator().manual_seed(seed)
  for opt, t in ((state.g_optimizer, 3), (state.d_optimizer, 6)):
    opt.m.copy_(torch.randn(opt.m.shape, generator=gen))
    opt.v.copy_(torch.rand(opt.v.shape, generator=gen))
    opt.step = t
  return state


def _leaves(tree, prefix=""):
  for k, v in tree.items():
    if isinstance(v, dict):
      yield from _leaves(v, prefix + "/" + k)
    else:
      yield prefix + "/" + k, v


def _flat(layout, tree):
  buf = torch.zeros(layout.total)
  layout.load_tree(buf, tree)
  return buf


def test_train_state_round_trip_and_reference_tree_structure(tmp_path):
  cfg = helpers.small_config(g_spectral_norm=True)
  state = _cpu_state(cfg, 5)
  sd = ck.to_state_dict(state)
  # structure the reference's flax.serialization.to_state_dict(TrainState) has
  assert list(sd) == ["step", "g_optimizer", "d_optimizer", "generator_state", "discriminator_state", "ema_params"]
  assert set(sd["g_optimizer"]) == {"target", "state"} and set(sd["g_optimizer"]["state"]) == {"step", "param_states"}
  leaf = sd["d_optimizer"]["state"]["param_states"]["DiscBlock_0"]["SpectralConv_0"]["kernel"]
  assert set(leaf) == {"grad_ema", "grad_sq_ema"} and leaf["grad_ema"].shape == (3, 3, 16, 32)
  assert sd["d_optimizer"]["state"]["step"].dtype == np.int32 and int(sd["d_optimizer"]["state"]["step"]) == 6
  assert sd["generator_state"]["spectral_norm_stats"]["GenBlock_0"]["SpectralConv_0"]["u0"].shape == (1, 256)
  assert sd["ema_params"]["SpectralDense_0"]["kernel"].dtype == np.float32
  # bytes -> a differently initialised state of the same configuration -> identical buffers
  path = ck.save_checkpoint(str(tmp_path), state)
  assert path.endswith("ckpt-1.flax") and ck.save_checkpoint(str(tmp_path), state).endswith("ckpt-2.flax")
  assert ck.latest_checkpoint(str(tmp_path)).endswith("ckpt-2.flax")
  other = _cpu_state(cfg, 9)
  assert not torch.equal(other.g_optimizer.target.buf, state.g_optimizer.target.buf)
  ck.restore_checkpoint(other, path)
  assert (other.step, other.g_optimizer.step, other.d_optimizer.step) == (3, 3, 6)
  def leaves(opt, buf):  # the moments are compared leaf by leaf: the alignment padding between leaves is not state
    return torch.cat([t.reshape(-1) for _, t in sorted(_leaves(opt.target.layout.tree(buf)))])

  for a, b in ((other.g_optimizer.target.buf, state.g_optimizer.target.buf),
               (leaves(other.g_optimizer, other.g_optimizer.m), leaves(state.g_optimizer, state.g_optimizer.m)),
               (leaves(other.d_optimizer, other.d_optimizer.v), leaves(state.d_optimizer, state.d_optimizer.v)),
               (other.ema_params.buf, state.ema_params.buf),
               (other.generator_state["batch_stats"].buf, state.generator_state["batch_stats"].buf),
               (other.generator_state["spectral_norm_stats"].buf, state.generator_state["spectral_norm_stats"].buf),
               (other.discriminator_state["spectral_norm_stats"].buf,
                state.discriminator_state["spectral_norm_stats"].buf)):
    assert torch.equal(a, b)


def test_make_grid_is_the_reference_index_permutation():
  """image_utils.make_grid (image_utils.py:23-38): exact."""
  x = torch.arange(10 * 2 * 3 * 1, dtype=torch.float32).reshape(10, 2, 3, 1)
  g = train_utils.make_grid(x, show_num=6)          # h_num = 2, w_num = 3
  want = x[:6].numpy().reshape(2, 3, 2, 3, 1).swapaxes(1, 2).reshape(4, 9, 1)
  assert g.shape == (4, 9, 1) and np.array_equal(g.numpy(), want)
  assert train_utils.make_grid(x, show_num=64).shape == (3 * 2, 3 * 3, 1)  # cut to the batch: 10 -> 3 x 3
