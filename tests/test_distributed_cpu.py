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


def _sliced_worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from xmcgan_image_generation_b200 import xmc_gan
  n = 5000 + 4 * 37                      # a multiple of 4 (xmc_adam's granularity), not of the slice size
  g = torch.Generator().manual_seed(11 + rank)
  flat = torch.randn(n, generator=g)
  local = flat.clone()
  handles = xmc_gan._all_reduce_sliced(flat, 3)
  # consecutive, non-overlapping slices on 1024-element boundaries that cover the buffer exactly once
  assert handles[0][0] == 0 and handles[-1][1] == n
  for (lo, hi, _), (lo2, _, _) in zip(handles[:-1], handles[1:]):
    assert hi == lo2 and lo % 1024 == 0 and hi % 1024 == 0
  for lo, hi, work in handles:
    work.wait()
  other = torch.randn(n, generator=torch.Generator().manual_seed(11 + (1 - rank)))
  ok = torch.allclose(flat, local + other, atol=0, rtol=0)
  # one replica: no handles at all (the caller then applies Adam in one call)
  out[rank] = bool(ok)
  dist.barrier()
  dist.destroy_process_group()


def test_sliced_all_reduce_covers_the_buffer_once():
  """xmc_gan._all_reduce_sliced (the generator-gradient all-reduce issued in slices so that Adam of slice i can run
  under the all-reduce of slice i+1): on two gloo ranks every element is summed exactly once, bit-exact."""
  world, port = 2, _free_port()
  out = mp.Manager().dict()
  mp.spawn(_sliced_worker, args=(world, port, out), nprocs=world, join=True)
  assert out[0] and out[1]


def test_sm_reservation_bookkeeping():
  """ops.reserve_sms / release_sms (SMs left to a concurrent collective): the window closes after the announced amount
  of executed GEMM work or at release; pure host-side state on top of xmc_set_sm_limit."""
  from xmcgan_image_generation_b200 import ops
  ops.release_sms()
  assert ops._RESERVE[0] == 0.0
  ops.reserve_sms(16, 1.0)               # 1 TFLOP of GEMM work
  assert ops._RESERVE[0] == 1e12
  ops._spend(0.4e12)
  assert ops._RESERVE[0] == 0.6e12
  ops._spend(0.7e12)                     # past the window: released
  assert ops._RESERVE[0] == 0.0
  ops.reserve_sms(0, 1.0)                # n <= 0: nothing reserved
  assert ops._RESERVE[0] == 0.0
  ops.reserve_sms(8, 2.0)
  ops.release_sms()
  assert ops._RESERVE[0] == 0.0
