"""Cross-replica plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on B200, gloo in CPU
tests). The path is pure data parallelism: the only exchanges are the gradient mean (xmc_gan.py:170-171,251 —
jax.lax.pmean) and the 5-scalar metric mean (xmc_gan.py:185-190). InfoNCE negatives, BatchNorm statistics and
word-loss pairs stay rank-local exactly as in the reference (attention_lib.py:58-62; coco_xmc.py:44)."""
import os

import torch
import torch.distributed as dist

# SMs left to NCCL's thread blocks while a gradient all-reduce overlaps the persistent tensor-core GEMMs
# (ops.reserve_sms / xmc_set_sm_limit); NCCL is held to the same number of thread blocks (NCCL_MAX_CTAS, read when the
# communicator is created, i.e. after this module is imported). Opt-in (XMC_RESERVED_SMS=n): on 2 GPUs the hidden
# all-reduces cost ~0.05 ms each as it is, and holding NCCL to 8 / 16 / 32 thread blocks made the step 1.4 / 0.5 / 0.5 ms
# slower (`profiles/r02_nccl_overlap.md`).
RESERVED_SMS = int(os.environ.get("XMC_RESERVED_SMS", "0"))
if RESERVED_SMS > 0:
  os.environ.setdefault("NCCL_MAX_CTAS", str(RESERVED_SMS))
# Slices of the generator-gradient all-reduce (xmc_gan._all_reduce_sliced): Adam of slice i overlaps NCCL on slice i+1.
# Opt-in as well: an NCCL thread block needs nearly a whole register file, so it only runs on SMs nothing else occupies,
# and Adam (HBM-bound on all SMs) and the all-reduce end up taking turns: 4 slices measured +0.4 ms on 8 GPUs.
G_SLICES = int(os.environ.get("XMC_G_SLICES", "1"))


def reserve_tflop(nbytes):
  """Executed GEMM work (TFLOP) issued on the reduced grid after an all-reduce of `nbytes` starts: about the
  all-reduce's duration (>= 250 GB/s algorithmic bandwidth at RESERVED_SMS thread blocks) at ~800 TFLOP/s of GEMMs."""
  return nbytes / 250e9 * 800.0


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
