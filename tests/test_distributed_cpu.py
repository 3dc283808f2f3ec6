"""world_size=2 gloo test of the data-parallel plumbing (parallel.py): the gradient exchange is a sum all-reduce with
the 1/world folded into the optimiser, metrics are replica means; InfoNCE negatives stay rank-local."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import xmc_oracle as orc
from tests import helpers


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from xmcgan_image_generation_b200 import parallel
  torch.set_num_threads(2)
  cfg = helpers.small_config(gf_dim=8, df_dim=8)
  _, _, g_vars, d_vars = helpers.cpu_variables(cfg, E=16, seed=3)
  state = orc.make_state(g_vars, d_vars)
  batch = helpers.make_batch(4, cfg, E=16, L=5, seed=9, min_len=2)
  shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in batch.items()}  # rank-local examples and negatives
  r = orc.d_losses_and_grads(state, shard, cfg, orc.FP32, want_g=False)
  flat = torch.cat([g.reshape(-1) for _, g in orc.tree_leaves(r["d_grad"])])
  local = flat.clone()
  assert parallel.world_size() == world and parallel.rank() == rank
  parallel.all_reduce_sum_(flat)
  mean = flat / world
  loss = torch.tensor([r["d_loss"].item()])
  parallel.all_reduce_sum_(loss)
  # cross-replica BatchNorm groups (device_utils.get_device_groups, xmc_net.py:196-200): a group spanning both ranks
  # sums the per-channel statistics over the group; a group of one replica means "no exchange" (None)
  grp = parallel.bn_group(2 * 4, 4)
  stats = torch.tensor([float(rank + 1), 10.0 * (rank + 1)])
  parallel.all_reduce_sum_(stats, group=grp[0])
  solo = parallel.bn_group(4, 4)
  out[rank] = (local, mean, loss / world, (grp[1], stats.tolist(), solo))
  dist.destroy_process_group()


def test_two_rank_gradient_and_metric_mean():
  world = 2
  port = _free_port()
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
  (l0, m0, s0, g0), (l1, m1, s1, g1) = out[0], out[1]
  assert g0 == g1 == (2, [3.0, 30.0], None)
  assert torch.allclose(m0, m1)
  assert torch.allclose(m0, (l0 + l1) / 2, rtol=1e-6, atol=1e-8)
  assert not torch.allclose(l0, l1)  # different shards -> different local gradients
  assert torch.allclose(s0, s1)
