"""Cross-replica BatchNorm (config.batch_norm_group_size > 0; xmc_net.py:192-201, device_utils.py:18-26) on the CUDA
path. Two ranks share the one GPU of the test box (gloo moves the tiny per-channel statistic vectors; on a multi-GPU
job the same code runs over NCCL). Property checked — it needs no oracle: the generator has no other cross-example op
than BatchNorm, so two replicas of B examples each with a statistics group spanning both must reproduce ONE replica
running the same 2B examples with replica-local statistics: same images, same new running statistics, and parameter
gradients that sum to the single-replica gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers

gpu = pytest.mark.gpu


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.cuda.set_device(0)
  from xmcgan_image_generation_b200 import engine, ops, parallel
  from xmcgan_image_generation_b200.nets import xmc_net
  B, E = 3, 64
  cfg_sync = helpers.small_config(batch_norm_group_size=world * B)
  cfg_local = helpers.small_config()
  assert parallel.get_device_groups(world * B, B) == [[0, 1]]
  g_sync, g_local = engine.GeneratorEngine(cfg_sync, E), engine.GeneratorEngine(cfg_local, E)
  _, _, g_vars, _ = helpers.cpu_variables(cfg_local, E, seed=21)
  params = xmc_net.as_flat(g_local.layout, g_vars["params"])
  stats = xmc_net.as_flat(g_local.stats_layout, g_vars["batch_stats"])
  full = xmc_net.batch_to_device(helpers.make_batch(world * B, cfg_local, seed=22))
  torch.manual_seed(23)
  d_img_full = (torch.randn(world * B, 128, 128, 3) * 0.05).cuda()
  sl = slice(rank * B, (rank + 1) * B)
  shard = {k: v[sl].contiguous() for k, v in full.items()}

  def run(eng, batch, d_img):
    eng.prep_weights(params)
    new_stats = torch.empty_like(stats)
    img, ctx = eng.forward(params, stats, batch, batch["z"], train=True, new_stats=new_stats)
    grads = torch.zeros_like(params)
    eng.backward(ctx, d_img.contiguous(), params, grads)
    torch.cuda.synchronize()
    return img, new_stats, grads

  img_s, stats_s, grads_s = run(g_sync, shard, d_img_full[sl])
  parallel.all_reduce_sum_(grads_s)                       # sum over replicas of the per-replica gradients
  img_f, stats_f, grads_f = run(g_local, full, d_img_full)
  rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
  res = {"img": rel(img_s, img_f[sl]), "stats": rel(stats_s, stats_f), "grads": rel(grads_s, grads_f)}
  # and the group really changes the result: replica-local statistics on the shard differ from the full batch
  img_l, _, _ = run(g_local, shard, d_img_full[sl])
  res["local_differs"] = rel(img_l, img_f[sl])
  out[rank] = res
  dist.destroy_process_group()


@gpu
def test_grouped_batch_norm_over_two_replicas_equals_one_replica_on_the_joint_batch():
  """Tolerances: images 1e-2 rel-L2 (bf16 activations; the statistics differ only in summation order), running
  statistics 1e-4, summed parameter gradients 2e-2."""
  world = 2
  out = mp.Manager().dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  for r in range(world):
    res = out[r]
    assert res["img"] < 1e-2, res
    assert res["stats"] < 1e-4, res
    assert res["grads"] < 2e-2, res
    assert res["local_differs"] > 5 * res["img"], res
