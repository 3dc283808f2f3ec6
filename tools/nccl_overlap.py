"""GPU debugging aid (run under torchrun, >= 2 GPUs): where do the gradient all-reduces of a train_step sit on rank 0's
timeline? CUPTI kernel trace (torch.profiler) of 2 eager steps: for every NCCL kernel its start, duration, the compute
kernels' busy time inside its interval, and the compute-idle time before the next compute kernel after it ends.
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_overlap.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench
from xmcgan_image_generation_b200 import engine, train_utils, xmc_gan

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 56
config = bench.make_config(128, True)
config.batch_size = B * dist.get_world_size()
host = bench.synth_batch(2 * B, config, 42 + rank)
dev = {k: v.cuda() for k, v in host.items()}
gen, disc, state = train_utils.create_train_state(config, 42, host)
add = xmc_gan.create_additional_data(config, variables=engine.ResNetEngine().random_variables(7))
for _ in range(3):
  state, m = train_utils.train_step(None, state, dev, xmc_gan, gen, disc, config, add)
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(2):
    state, m = train_utils.train_step(None, state, dev, xmc_gan, gen, disc, config, add)
  torch.cuda.synchronize()
dist.barrier()
if rank == 0:
  ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
  iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
  t0 = iv[0][0]
  comp = [(a, b, n) for a, b, n in iv if "nccl" not in n.lower()]
  nccl = [(a, b, n) for a, b, n in iv if "nccl" in n.lower()]
  print(f"span {(iv[-1][1]-t0)/1e3:.2f} ms for 2 steps; compute busy {sum(b-a for a,b,_ in comp)/1e3:.2f} ms")
  for a, b, n in nccl:
    inside = sum(min(b, cb) - max(a, ca) for ca, cb, _ in comp if cb > a and ca < b)
    during = [(cb - ca, cn) for ca, cb, cn in comp if cb > a and ca < b]
    nxt = min((ca for ca, _, _ in comp if ca >= b), default=b)
    last_before = max((cb for _, cb, _ in comp if cb <= nxt), default=a)
    print(f"  nccl @{(a-t0)/1e3:8.2f} ms  dur {(b-a)/1e3:6.3f} ms  compute busy inside {inside/1e3:6.3f} ms "
          f"({len(during)} kernels)  compute gap around its end {max(0.0, nxt - max(last_before, a))/1e3:6.3f} ms  {n[:40]}")
  # compute-idle gaps > 100 us
  gaps = sorted(((comp[i + 1][0] - comp[i][1], (comp[i][1] - t0) / 1e3, comp[i][2][:40], comp[i + 1][2][:40])
                 for i in range(len(comp) - 1)), reverse=True)
  for g in gaps[:8]:
    print(f"  compute gap {g[0]/1e3:.3f} ms @{g[1]:.2f} ms  {g[2]} -> {g[3]}")
sys.stdout.flush()
dist.barrier()
os._exit(0)
