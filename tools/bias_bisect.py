"""Bisect of the one-sided g_loss offset of test_train_step_with_pretrained_image_contrastive:
CUDA 18.10 vs bf16-policy oracle 18.21, always the same sign.

g_loss is taken in train_g_d AFTER train_d's Adam step on the discriminator. At t = 1 Adam's update is
-lr * g / (|g| + eps): every parameter moves by +-lr with the sign of its gradient, so hinge_g = -mean(fake logit)
jumps from -8.4 to +7.4 in this configuration (oracle, CPU) and any parameter whose gradient sits at rounding-noise
level contributes +-2*lr*dlogit/dp depending on a coin flip of the implementation's rounding.

This tool takes the discriminator gradient of the CUDA path for train_d, feeds it to the ORACLE's Adam and lets the
ORACLE run train_g_d on the result ("hybrid"): if the hybrid reproduces the CUDA g_loss, the whole offset is the
sign-step amplifying last-bit gradient differences, not a defect in the train_g_d path (ResNet branch, zero-insert
dgrad, stem dgrad ...). It also reports the share of parameters whose gradient sign differs, per leaf.
Run on the GPU box:  python tools/bias_bisect.py > gpurun_out/r02_bias_bisect.log"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xmc_oracle as orc  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_parity import _build  # noqa: E402
from xmcgan_image_generation_b200 import train_utils, xmc_gan  # noqa: E402
from xmcgan_image_generation_b200.nets import xmc_net  # noqa: E402


def main():
  cfg = helpers.small_config(pretrained_image_contrastive=True)
  B = 3
  variables = orc.resnet50_random_variables(2)
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=8)
  batch = helpers.make_batch(2 * B, cfg, seed=9)
  pol = orc.Policy("bfloat16")
  pre = lambda real, fake: orc.calculate_contrastive_loss_on_pretrained(variables, real, fake, pol)
  batches = orc.split_input_dict(batch, 2)

  # ---- CUDA: train_d's discriminator gradient (before Adam), then the full train_step ---------------------------
  def fresh():
    g_vars_, d_vars_, gp, gs, dp, du = _build(cfg, seed=8)
    return train_utils.TrainState(0, train_utils.Optimizer(gp, cfg.g_lr, cfg.beta1, cfg.beta2),
                                  train_utils.Optimizer(dp, cfg.d_lr, cfg.beta1, cfg.beta2),
                                  {"batch_stats": gs}, {"spectral_norm_stats": du}, gp.clone())

  state = fresh()
  dev = xmc_net.batch_to_device(batches[0])
  g_eng, d_eng = xmc_gan._engines(cfg, dev)
  ws = xmc_gan._workspace(state, g_eng, d_eng)
  losses = torch.zeros(16, device="cuda")
  gctx, dctx, state = xmc_gan._forward_both(state, dev, cfg, ws, g_eng, d_eng, losses, keep_g_state=False, need_g=False)
  ws.d_grads.zero_()
  d_eng.backward_d(dctx, state.d_optimizer.target.buf, ws.d_grads)
  d_eng.sn_backward(state.d_optimizer.target.buf, ws.d_grads, ws.u0_alt)
  torch.cuda.synchronize()
  cuda_dgrad = xmc_net.FlatTree(d_eng.layout, ws.d_grads.clone()).to_cpu_tree()

  additional = xmc_gan.create_additional_data(cfg, variables=variables)
  state = fresh()
  state, metrics = train_utils.train_step(None, state, batch, xmc_gan, None, None, cfg, additional)
  cuda_metrics = metrics.compute()

  # ---- oracle ------------------------------------------------------------------------------------------------
  ostate = orc.make_state(g_vars, d_vars)
  st1, r_d = orc.train_d(ostate, batches[0], cfg, pol)
  _, want, _ = orc.train_g_d(st1, batches[1], cfg, pol, pre)

  # ---- hybrid: oracle everything, except that train_d's Adam consumes the CUDA gradient ---------------------------
  new_params, new_opt = orc.adam_apply(ostate["d_params"], ostate["d_opt"], cuda_dgrad, cfg.d_lr, cfg.beta1, cfg.beta2)
  hyb = dict(st1)
  hyb["d_params"], hyb["d_opt"] = new_params, new_opt
  _, hybrid, _ = orc.train_g_d(hyb, batches[1], cfg, pol, pre)

  # ---- oracle with its own gradient perturbed at the bf16-rounding level (sensitivity of the quantity itself) -------
  st32, _ = orc.train_d(ostate, batches[0], cfg, orc.FP32)
  mix = dict(st1)
  mix["d_params"], mix["d_opt"] = st32["d_params"], st32["d_opt"]
  _, mixed, _ = orc.train_g_d(mix, batches[1], cfg, pol, pre)

  print("metric                 cuda        oracle(bf16)  hybrid(cuda D-grad -> oracle)  oracle fwd + fp32-oracle D-step")
  for k in ("d_loss", "g_loss", "c_loss_d", "c_loss_g", "c_loss_g_pretrained"):
    print(f"{k:22s} {cuda_metrics[k]:11.5f} {want[k]:11.5f} {hybrid[k]:11.5f} {mixed[k]:11.5f}")
  print("\nper-leaf: rel-L2 of the CUDA gradient vs oracle, share of elements with a different sign, and the share of "
        "|g| mass those elements hold")
  tot = flips = 0
  for (path, g), (_, r) in zip(orc.tree_leaves(cuda_dgrad), orc.tree_leaves(r_d["d_grad"])):
    g, r = g.float().reshape(-1), r.reshape(-1)
    diff = torch.sign(g) != torch.sign(r)
    tot += g.numel()
    flips += int(diff.sum())
    print(f"{path:60s} n={g.numel():7d} rel={helpers.rel(g, r):.2e} flips={diff.float().mean().item():.4f} "
          f"mass={(r.abs() * diff).sum().item() / max(r.abs().sum().item(), 1e-30):.2e}")
  print(f"total sign flips: {flips} of {tot} = {flips / tot:.4%}")


if __name__ == "__main__":
  main()
