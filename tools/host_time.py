"""GPU debugging aid: host enqueue time of a train_step (is the step host-bound?). Times K steps on the host clock
without synchronising, then the device time of the same steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from xmcgan_image_generation_b200 import engine, train_utils, xmc_gan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 56
config = bench.make_config(128, True)
config.batch_size = B
host = bench.synth_batch(2 * B, config, 42)
dev = {k: v.cuda() for k, v in host.items()}
gen, disc, state = train_utils.create_train_state(config, 42, host)
add = xmc_gan.create_additional_data(config, variables=engine.ResNetEngine().random_variables(7))
for _ in range(3):
  state, m = train_utils.train_step(None, state, dev, xmc_gan, gen, disc, config, add)
torch.cuda.synchronize()
K = 10
t0 = time.perf_counter()
for _ in range(K):
  state, m = train_utils.train_step(None, state, dev, xmc_gan, gen, disc, config, add)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/K:.2f} ms/step; until device idle {1e3*(t2-t0)/K:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
  state, m = train_utils.train_step(None, state, dev, xmc_gan, gen, disc, config, add)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
