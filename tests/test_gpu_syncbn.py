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


def _worker(rank, world, port, out, backend="gloo"):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  if backend == "nccl":     # one GPU per rank: the production configuration (NCCL over NVLink)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  else:                     # both ranks share the one GPU of the test box; gloo moves the small statistic vectors
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

  # ---- single op: statistics + backward of one LocalConditionalBatchNorm over the group vs the joint batch ---------
  torch.manual_seed(24)
  C, H, Hc = 32, 16, 16
  x_full = (torch.randn(world * B, H, H, C) * 2 + 0.5).cuda().to(torch.bfloat16)
  gb_full = (torch.randn(world * B * Hc * Hc, 2 * C) * 0.5).cuda().to(torch.bfloat16)
  dy_full = (torch.randn(world * B, H, H, C) * 0.1).cuda().to(torch.bfloat16)
  grp = parallel.bn_group(world * B, B)
  assert grp is not None and grp[1] == world

  def bn_op(x, gb, dy, group):
    sums, P = ops.bn_stats(x)
    if group is not None:
      parallel.all_reduce_sum_(sums, group=group[0])
      P *= group[1]
    zeros, ones = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    mr = ops.bn_finalize(sums, P, C, zeros, ones, None, None)
    y = ops.bn_apply(x, mr, gb, Hc, 0, C, True, False)
    dgb = torch.zeros(gb.shape, device="cuda")
    dx = ops.bn_bwd(dy, x, mr, gb, dgb, Hc, 0, C, True, False, group=group)
    torch.cuda.synchronize()
    return y, dx, dgb

  rows = slice(rank * B * Hc * Hc, (rank + 1) * B * Hc * Hc)
  y_s, dx_s, dgb_s = bn_op(x_full[sl].contiguous(), gb_full[rows].contiguous(), dy_full[sl].contiguous(), grp)
  y_f, dx_f, dgb_f = bn_op(x_full, gb_full, dy_full, None)
  rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
  op = {"y": rel(y_s, y_f[sl]), "dx": rel(dx_s, dx_f[sl]), "dgb": rel(dgb_s, dgb_f[rows])}

  img_s, stats_s, grads_s = run(g_sync, shard, d_img_full[sl])
  parallel.all_reduce_sum_(grads_s)                       # sum over replicas of the per-replica gradients
  img_f, stats_f, grads_f = run(g_local, full, d_img_full)
  res = {"op": op, "img": rel(img_s, img_f[sl]), "stats": rel(stats_s, stats_f), "grads": rel(grads_s, grads_f)}
  # and the group really changes the result: replica-local statistics on the shard differ from the full batch
  img_l, _, grads_l = run(g_local, shard, d_img_full[sl])
  parallel.all_reduce_sum_(grads_l)
  res["local_differs"] = rel(img_l, img_f[sl])
  res["local_grads_differ"] = rel(grads_l, grads_f)
  out[rank] = res
  dist.destroy_process_group()


def _check(out, world):
  for r in range(world):
    res = out[r]
    print("sync-BN result rank", r, res)
    assert max(res["op"].values()) < 2e-3, res
    assert res["img"] < 1e-2, res
    assert res["stats"] < 1e-4, res
    assert res["grads"] < 1.5e-1, res
    assert res["local_differs"] > 5 * res["img"], res
    assert res["local_grads_differ"] > 3 * res["grads"], res


@gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_grouped_batch_norm_over_nccl_one_gpu_per_rank():
  """Same property as below with the production plumbing: one process per GPU, NCCL all-reduce of the statistics."""
  world = 2
  out = mp.Manager().dict()
  mp.spawn(_worker, args=(world, _free_port(), out, "nccl"), nprocs=world, join=True)
  _check(out, world)


@gpu
def test_grouped_batch_norm_over_two_replicas_equals_one_replica_on_the_joint_batch():
  """Tolerances: one BatchNorm op (forward, dx, dgamma/dbeta) 2e-3 — only the summation order of the statistics
  differs (measured 1e-5); whole generator: images 1e-2 rel-L2 (measured 2e-3), running statistics 1e-4, summed
  parameter gradients 1.5e-1 rel-L2: they pass through 11 normalisation layers with bf16 activation gradients and
  relu masks that flip on last-bit differences of the statistics, and the split-K / dgamma reductions are atomics, so
  the value moves between 0.045 and 0.056 from run to run; the same two replicas with replica-local statistics are at
  0.84, which is what the test separates (and it requires a 3x margin between the two)."""
  world = 2
  out = mp.Manager().dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  _check(out, world)
