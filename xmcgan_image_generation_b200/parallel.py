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


def all_reduce_sum_(flat, async_op=False):
  """In-place sum all-reduce of a flat buffer; the 1/world of pmean is folded into the Adam kernel's grad_scale."""
  if world_size() == 1:
    return None
  return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)


def get_device_groups(group_batch_size, device_batch_size, device_count=None):
  """device_utils.get_device_groups (xmcgan/utils/device_utils.py:18-26): contiguous replica groups."""
  device_count = world_size() if device_count is None else device_count
  assert group_batch_size % device_batch_size == 0
  group_size = group_batch_size // device_batch_size
  assert device_count % group_size == 0
  return [list(range(i, i + group_size)) for i in range(0, device_count, group_size)]
