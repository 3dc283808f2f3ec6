"""Cross-replica plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on B200, gloo in CPU
tests). The path is pure data parallelism: the only exchanges are the gradient mean (xmc_gan.py:170-171,251 —
jax.lax.pmean) and the 5-scalar metric mean (xmc_gan.py:185-190). InfoNCE negatives, BatchNorm statistics and
word-loss pairs stay rank-local exactly as in the reference (attention_lib.py:58-62; coco_xmc.py:44)."""
import torch
import torch.distributed as dist


def world_size():
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
  return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def all_reduce_sum_(flat, async_op=False, group=None):
  """In-place sum all-reduce of a flat buffer; the 1/world of pmean is folded into the Adam kernel's grad_scale."""
  if world_size() == 1:
    return None
  return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op, group=group)


def get_device_groups(group_batch_size, device_batch_size, device_count=None):
  """device_utils.get_device_groups (xmcgan/utils/device_utils.py:18-26): contiguous replica groups."""
  device_count = world_size() if device_count is None else device_count
  assert group_batch_size % device_batch_size == 0
  group_size = group_batch_size // device_batch_size
  assert device_count % group_size == 0
  return [list(range(i, i + group_size)) for i in range(0, device_count, group_size)]


_BN_GROUPS = {}


def bn_group(group_batch_size, device_batch_size):
  """Process group of this rank for cross-replica BatchNorm (config.batch_norm_group_size > 0, xmc_net.py:192-201):
  returns (process_group, group_size) for the contiguous replica group of get_device_groups, or None when statistics
  stay replica-local (group of one). Every rank creates every group (torch.distributed.new_group is collective)."""
  groups = get_device_groups(group_batch_size, device_batch_size)
  if len(groups[0]) == 1:
    return None
  key = (group_batch_size, device_batch_size, world_size())
  if key not in _BN_GROUPS:
    mine = None
    for ranks in groups:
      pg = dist.new_group(ranks)
      if rank() in ranks:
        mine = (pg, len(ranks))
    _BN_GROUPS[key] = mine
  return _BN_GROUPS[key]
