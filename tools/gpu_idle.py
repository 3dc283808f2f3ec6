"""GPU debugging aid: how much of a train_step is the GPU idle (host launch-bound)? Uses the CUPTI kernel trace of
torch.profiler over 2 steps: sum of kernel durations vs the span they cover, and the largest gaps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from xmcgan_image_generation_b200 import engine, train_utils, xmc_gan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 56
RESNET_DTYPE = sys.argv[2] if len(sys.argv) > 2 else "float32"
config = bench.make_config(128, True)
config.batch_size = B
host = bench.synth_batch(2 * B, config, 42)
dev = {k: v.cuda() for k, v in host.items()}
gen, disc, state = train_utils.create_train_state(config, 42, host)
add = xmc_gan.create_additional_data(config, variables=engine.ResNetEngine().random_variables(7),
                                     image_model_dtype=RESNET_DTYPE)
for _ in range(3):
  state, m = train_utils.train_step(None, state, dev, xmc_gan, gen, disc, config, add)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(2):
    state, m = train_utils.train_step(None, state, dev, xmc_gan, gen, disc, config, add)
  torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
busy = sum(b - a for a, b, _ in iv)
span = iv[-1][1] - iv[0][0]
gaps = sorted(((iv[i + 1][0] - iv[i][1], iv[i][2][:50], iv[i + 1][2][:50]) for i in range(len(iv) - 1)), reverse=True)
print(f"kernels {len(iv)}  busy {busy/1e3:.2f} ms  span {span/1e3:.2f} ms  idle {100*(span-busy)/span:.1f}%")
print("gaps > 20us:", sum(1 for g in gaps if g[0] > 20), " total gap us:", sum(g[0] for g in gaps if g[0] > 0))
for g in gaps[:6]:
  print(f"  {g[0]:.1f} us between {g[1]} -> {g[2]}")
import collections
import re
agg = collections.defaultdict(lambda: [0, 0.0])
for a, b, name in iv:
  k = re.sub(r"\(.*", "", name)[:70]
  agg[k][0] += 1
  agg[k][1] += (b - a) / 2e3   # ms per step (2 steps profiled)
print("| ms/step | share | launches/step | kernel |")
print("|---:|---:|---:|---|")
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
  print(f"| {v[1]:.3f} | {100*v[1]/tot:.1f}% | {v[0]//2} | {k} |")
