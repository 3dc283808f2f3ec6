"""Generates tests/golden/oracle_tiny.npz from the CPU oracle (fp32 policy): losses, gradient norms and updated
state of one train_step at a tiny configuration with fixed seeds. Run: python -m tests.golden.make_golden"""
import os

import numpy as np
import torch

from oracle import xmc_oracle as orc
from tests import helpers


def compute():
  torch.manual_seed(0)
  cfg = helpers.small_config(gf_dim=8, df_dim=8)
  _, _, g_vars, d_vars = helpers.cpu_variables(cfg, E=16, seed=11)
  state = orc.make_state(g_vars, d_vars)
  batch = helpers.make_batch(4, cfg, E=16, L=6, seed=5, min_len=2)
  b0, b1 = orc.split_input_dict(batch, 2)
  r = orc.d_losses_and_grads(state, b1, cfg, orc.FP32, want_g=True)
  new_state, metrics = orc.train_step(state, batch, cfg, orc.FP32)
  out = {"losses": np.array([r["d_loss"].item(), r["g_loss"].item(), r["c_loss_d"].item(), r["c_loss_g"].item()]),
         "metrics": np.array([metrics[k] for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g")]),
         "fake_mean": np.array([r["fake"].mean().item(), r["fake"].std().item()]),
         "logit": r["logit"].reshape(-1).numpy()}
  out["d_grad_norms"] = np.array([g.norm().item() for _, g in orc.tree_leaves(r["d_grad"])])
  out["g_grad_norms"] = np.array([g.norm().item() for _, g in orc.tree_leaves(r["g_grad"])])
  out["new_u0_first"] = orc.tree_leaves(new_state["discriminator_state"])[0][1].reshape(-1).numpy()
  out["new_g_param_norms"] = np.array([p.norm().item() for _, p in orc.tree_leaves(new_state["g_params"])])
  return out


if __name__ == "__main__":
  path = os.path.join(os.path.dirname(__file__), "oracle_tiny.npz")
  np.savez(path, **compute())
  print("wrote", path)
