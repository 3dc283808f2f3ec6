"""Sharp single-op GPU checks added in round 2: Adam + EMA (host-side and device-side step count), determinism."""
import pytest
import torch

from oracle import xmc_oracle as orc
from tests import helpers

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("use_step_dev", [False, True])
def test_adam_and_ema_single_op_matches_oracle(use_step_dev):
  """xmc_adam (flax.optim.Adam.apply_gradient, created train_utils.py:181-186, applied xmc_gan.py:172-173, + the
  polyak EMA of :174-177) for steps t = 1..4 on one flat buffer, with gradient scaling 1/world folded in, vs
  orc.adam_apply: parameters / moments 2e-6 rel-L2 (fp32 arithmetic, different operation order), EMA 1e-6. With
  step_dev the bias corrections come from the device-side step count (the CUDA-graph path), host values ignored."""
  from xmcgan_image_generation_b200 import _lib, ops
  torch.manual_seed(7)
  n = 4096 + 36
  p0 = torch.randn(n) * 0.05
  lr, b1, b2, eps, decay, world = 4e-4, 0.5, 0.999, 1e-8, 0.999, 4
  p = p0.cuda()
  m, v, ema = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), p0.cuda()
  step_dev = torch.zeros(1, dtype=torch.int32, device="cuda") if use_step_dev else None
  o_p, o_opt, o_ema = {"w": p0.clone()}, orc.adam_init({"w": p0}), p0.clone()
  for t in range(1, 5):
    g = torch.randn(n) * (10.0 ** -t)            # magnitudes from 1e-1 down to 1e-4
    g[::7] = 0.0                                 # exact zeros: update must be exactly -lr*m_hat/(sqrt(v_hat)+eps)
    gsum = (g * world).cuda()                    # what the sum all-reduce leaves in the buffer
    bogus = 0.123 if use_step_dev else None      # host bias corrections must be ignored when step_dev is given
    # odd steps: the step applied in two calls over sub-ranges (as xmc_gan does, slice by slice of the all-reduce),
    # the device step count advancing on the last call only
    cuts = [0, (n // 8) * 4, n] if t % 2 else [0, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
      ops._call("xmc_adam", p.data_ptr() + 4 * lo, gsum.data_ptr() + 4 * lo, m.data_ptr() + 4 * lo,
                v.data_ptr() + 4 * lo, hi - lo, lr, b1, b2, eps, bogus or 1.0 - b1 ** t, bogus or 1.0 - b2 ** t,
                1.0 / world, ema.data_ptr() + 4 * lo, decay, step_dev.data_ptr() if use_step_dev else None,
                int(hi == n), _lib.stream())
    o_p, o_opt = orc.adam_apply(o_p, o_opt, {"w": g}, lr, b1, b2)
    o_ema = o_ema * decay + (1 - decay) * o_p["w"]
    # the update itself, not only the parameters (whose relative change per step is tiny)
    assert helpers.rel(p.cpu() - p0, o_p["w"] - p0) < 1e-4, t
    assert helpers.rel(p, o_p["w"]) < 2e-6 and helpers.rel(m, o_opt["m"]["w"]) < 2e-6
    assert helpers.rel(v, o_opt["v"]["w"]) < 2e-6 and helpers.rel(ema, o_ema) < 1e-6
  if use_step_dev:
    assert int(step_dev.item()) == 4


def _state(cfg, seed):
  from tests.test_gpu_parity import _build
  from xmcgan_image_generation_b200 import train_utils
  g_vars, d_vars, g_params, g_stats, d_params, d_u = _build(cfg, seed=seed)
  return train_utils.TrainState(0, train_utils.Optimizer(g_params, cfg.g_lr, cfg.beta1, cfg.beta2),
                                train_utils.Optimizer(d_params, cfg.d_lr, cfg.beta1, cfg.beta2),
                                {"batch_stats": g_stats}, {"spectral_norm_stats": d_u}, g_params.clone())


@gpu
@pytest.mark.parametrize("pretrained", [False, True])
def test_train_step_is_bit_reproducible(pretrained):
  """The reference (one XLA executable) is run-to-run deterministic. So is this path: the weight-gradient split-K goes
  through an fp32 workspace and a fixed-order second stage, BatchNorm / bias / spectral-norm reductions are two-stage
  without atomics, the bilinear-resize transpose is a gather. Two runs of two train_steps from the same state on the
  same batches must agree in EVERY bit of the metrics and of the new state (with and without the ResNet branch)."""
  from xmcgan_image_generation_b200 import train_utils, xmc_gan
  cfg = helpers.small_config(pretrained_image_contrastive=pretrained)
  additional = xmc_gan.create_additional_data(cfg, variables=orc.resnet50_random_variables(2)) if pretrained else {}
  batches = [helpers.make_batch(6, cfg, seed=60 + i) for i in range(2)]
  runs = []
  for _ in range(2):
    state = _state(cfg, 19)
    ms = []
    for b in batches:
      state, m = train_utils.train_step(None, state, b, xmc_gan, None, None, cfg, additional)
      ms.append(m.compute())
    runs.append((state, ms))
  (s1, m1), (s2, m2) = runs
  assert m1 == m2, (m1, m2)
  for name, a, b in (("g", s1.g_optimizer.target.buf, s2.g_optimizer.target.buf),
                     ("d", s1.d_optimizer.target.buf, s2.d_optimizer.target.buf),
                     ("g.m", s1.g_optimizer.m, s2.g_optimizer.m), ("g.v", s1.g_optimizer.v, s2.g_optimizer.v),
                     ("d.m", s1.d_optimizer.m, s2.d_optimizer.m), ("d.v", s1.d_optimizer.v, s2.d_optimizer.v),
                     ("ema", s1.ema_params.buf, s2.ema_params.buf)):
    assert torch.equal(a, b), name
  for key, coll in (("generator_state", "batch_stats"), ("discriminator_state", "spectral_norm_stats")):
    for (path, a), (_, b) in zip(orc.tree_leaves(getattr(s1, key)[coll].to_cpu_tree()),
                                 orc.tree_leaves(getattr(s2, key)[coll].to_cpu_tree())):
      assert torch.equal(a, b), (coll, path)


@gpu
def test_wgrad_split_k_is_deterministic_and_matches_single_pass():
  """The layer with the deepest K split (3x3, 96 -> 96 at 128x128: 1 output tile per tap, ~49 splits) twice on the
  same inputs: bit-identical; and accumulate semantics: a second launch into the same buffer doubles it exactly."""
  from xmcgan_image_generation_b200 import ops
  torch.manual_seed(3)
  N, S, C = 4, 128, 96
  x = (torch.randn(N, S, S, C, device="cuda") * 0.5).to(torch.bfloat16)
  dy = (torch.randn(N, S, S, C, device="cuda") * 0.1).to(torch.bfloat16)
  outs = []
  for _ in range(2):
    dw = torch.zeros(9 * C * C, device="cuda")
    ops.wgrad(x, dy, 3, dw, out_mode=0, ld_out=C, tap_stride=C * C)
    outs.append(dw)
  assert torch.equal(outs[0], outs[1])
  ops.wgrad(x, dy, 3, outs[0], out_mode=0, ld_out=C, tap_stride=C * C)
  assert torch.equal(outs[0], 2 * outs[1])


@gpu
def test_deferred_train_d_equals_plain_train_d():
  """xmc_gan.train_d_deferred (the multi-GPU schedule: train_d's gradient all-reduce stays in flight and its Adam / u0
  hand-over complete inside the next call, behind the generator forward) gives bit-identical states and metrics to the
  plain train_d -> train_g_d sequence."""
  from xmcgan_image_generation_b200 import train_utils, xmc_gan
  from xmcgan_image_generation_b200.nets import xmc_net
  cfg = helpers.small_config()
  batch = xmc_net.batch_to_device(helpers.make_batch(6, cfg, seed=70))
  b0, b1 = train_utils.split_input_dict(batch, 2)
  out = []
  for td in (xmc_gan.train_d, xmc_gan.train_d_deferred):
    state = _state(cfg, 23)
    old_d = state.d_optimizer.target.buf.clone()
    state = td(None, state, b0, None, None, cfg)
    if td is xmc_gan.train_d_deferred:
      assert torch.equal(state.d_optimizer.target.buf, old_d)      # the update is still pending
    state, m = xmc_gan.train_g_d(None, state, b1, None, None, cfg, {})
    out.append((state, m.compute()))
  (s1, m1), (s2, m2) = out
  assert m1 == m2
  assert (s2.step, s2.d_optimizer.step, s2.g_optimizer.step) == (1, 2, 1)
  for a, b in ((s1.d_optimizer.target.buf, s2.d_optimizer.target.buf), (s1.g_optimizer.target.buf, s2.g_optimizer.target.buf),
               (s1.d_optimizer.v, s2.d_optimizer.v), (s1.ema_params.buf, s2.ema_params.buf),
               (s1.discriminator_state["spectral_norm_stats"].buf, s2.discriminator_state["spectral_norm_stats"].buf)):
    assert torch.equal(a, b)


@gpu
@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
@pytest.mark.parametrize("N,H,C,Cout", [(2, 16, 64, 64), (3, 8, 96, 96), (1, 64, 32, 32), (2, 4, 192, 192)])
def test_pool_fused_conv_equals_conv_then_dsample(N, H, C, Cout, dtype):
  """dsample(conv3x3(x) + b) (the tail of DiscBlock / DiscOptimizedBlock, common.py:66-78,125-132) computed as ONE
  4x4 / stride-2 convolution with summed weights (xmc_poolconv_prep), its input gradient in sub-pixel form from the
  half-resolution output gradient (with the relu mask of x), and its weight gradient (XmcWgradDesc.subpixel = 2),
  against the oracle's conv2d -> dsample and autograd. bf16: forward 5e-3 (summed weights rounded once), gradients
  1e-2; fp32 mode (3 x bf16 split): 2e-5 / 1e-4."""
  from xmcgan_image_generation_b200 import _lib, ops
  fp32 = dtype == "float32"
  adt = torch.float32 if fp32 else torch.bfloat16
  q = (lambda t: t) if fp32 else (lambda t: t.to(torch.bfloat16).float())
  S = 3 if fp32 else 1
  torch.manual_seed(H * 7 + C)
  x = q(torch.relu(torch.randn(N, H, H, C))).requires_grad_(True)      # a relu output, as in the blocks
  kern = (torch.randn(3, 3, C, Cout) * 0.05).requires_grad_(True)
  bias = torch.randn(Cout)
  res = q(torch.randn(N, H // 2, H // 2, Cout) * 0.3)
  want = orc.dsample(orc.conv2d(x, kern, bias, orc.FP32, round_out=False)) + res
  dy = q(torch.randn(N, H // 2, H // 2, Cout) * 0.1)
  (want * dy).sum().backward()
  wf4 = ops.empty((Cout, 16 * C * S), torch.bfloat16)
  wdg = ops.empty((4 * C, 4 * Cout * S), torch.bfloat16)
  kd = kern.detach().cuda().contiguous()
  ops._call("xmc_poolconv_prep", kd.data_ptr(), None, C, Cout, int(fp32), wf4.data_ptr(), wdg.data_ptr(), _lib.stream())
  xd = x.detach().cuda().to(adt)
  with ops.act_dtype(adt):
    got = ops.conv_fwd(xd, wf4, 4, Cout, bias=bias.cuda(), residual=res.cuda().to(adt), ldb=16 * C * S, stride=2, pad=1,
                       out_dtype=torch.float32)
    assert got.shape == (N, H // 2, H // 2, Cout)
    e_fwd = helpers.rel(got, want)
    dyd = dy.cuda().to(adt)
    dx = ops.conv_fwd(dyd, wdg, 2, C, ldb=4 * Cout * S, pad=1, subpixel=True, mask=xd, out_dtype=torch.float32)
    e_dx = helpers.rel(dx, x.grad * (x.detach() > 0))
    dw = torch.zeros(9 * C * Cout, device="cuda")
    ops.wgrad(xd, dyd, 3, dw, out_mode=0, ld_out=Cout, tap_stride=C * Cout, alpha=0.25, subpixel=2)
    e_dw = helpers.rel(dw.view(3, 3, C, Cout), kern.grad)
  print(f"\\n[pool-fused {dtype} N{N} H{H} C{C}] fwd {e_fwd:.2e} dx {e_dx:.2e} dw {e_dw:.2e}")
  t_fwd, t_bwd = (2e-5, 1e-4) if fp32 else (5e-3, 1e-2)
  assert e_fwd < t_fwd and e_dx < t_bwd and e_dw < t_bwd


@gpu
@pytest.mark.parametrize("mode,N,H,Ca,Cb", [
    # plain 3x3: tg = 4 with one A slab (Ca <= 64); tg = 3 (96 -> 96: three groups of three taps); tg = 2 at BN = 192
    # (five groups, the last with a single tap); BN = 256 stays ungrouped; ragged maps (partial pixel chunks)
    (0, 3, 16, 48, 64), (0, 2, 16, 96, 96), (0, 2, 8, 192, 192), (0, 2, 8, 160, 176), (0, 1, 8, 64, 256),
    (0, 3, 12, 96, 32), (0, 37, 16, 96, 96),
    # sub-pixel parity taps: groups of four (per (a, b)) and of two (per (a, dh, b))
    (1, 2, 8, 64, 96), (1, 2, 8, 192, 192), (1, 3, 6, 96, 48),
    # pool-fused taps
    (2, 2, 8, 96, 96), (2, 2, 8, 192, 192), (2, 3, 6, 64, 128)])
def test_wgrad_tap_groups_match_oracle(mode, N, H, Ca, Cb):
  """gemm_wgrad_kernel's tap groups (several taps per work item sharing one B chunk, see plan_wgrad) in all three
  tap geometries, against the oracle's autograd: plain conv3x3 (ops.py wgrad), the sub-pixel form (weight gradient of
  conv3x3(upsample2(x)), common.py:148-157) and the pool-fused form (dsample(conv3x3(x)), common.py:66-78). 1e-4 rel
  on identical bf16 operands; the K split (N = 37) goes through the workspace and the fixed-order second stage."""
  from xmcgan_image_generation_b200 import ops
  torch.manual_seed(mode * 1000 + Ca + Cb + H)
  q = lambda t: t.to(torch.bfloat16).float()
  kern = (torch.randn(3, 3, Ca, Cb) * 0.05).requires_grad_(True)
  if mode == 0:
    x, dy = q(torch.randn(N, H, H, Ca)), q(torch.randn(N, H, H, Cb) * 0.1)
    y = orc.conv2d(x, kern, None, orc.FP32, round_out=False)
  elif mode == 1:
    x, dy = q(torch.randn(N, H, H, Ca)), q(torch.randn(N, 2 * H, 2 * H, Cb) * 0.1)
    y = orc.conv2d(orc.upsample(x), kern, None, orc.FP32, round_out=False)
  else:
    x, dy = q(torch.randn(N, 2 * H, 2 * H, Ca)), q(torch.randn(N, H, H, Cb) * 0.1)
    y = orc.dsample(orc.conv2d(x, kern, None, orc.FP32, round_out=False))
  (y * dy).sum().backward()
  dw = torch.zeros(9 * Ca * Cb, device="cuda")
  ops.wgrad(x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16), 3, dw, out_mode=0, ld_out=Cb, tap_stride=Ca * Cb,
            alpha=0.25 if mode == 2 else 1.0, subpixel=mode)
  err = helpers.rel(dw.view(3, 3, Ca, Cb), kern.grad)
  assert err < 1e-4, err
  # accumulation into a non-zero destination, bit-identical on a second run
  dw2 = torch.zeros(9 * Ca * Cb, device="cuda")
  ops.wgrad(x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16), 3, dw2, out_mode=0, ld_out=Cb,
            tap_stride=Ca * Cb, alpha=0.25 if mode == 2 else 1.0, subpixel=mode)
  assert torch.equal(dw, dw2)
  # store mode (out_mode 1: dw = ..., no read of the destination) overwrites whatever the buffer held, same bits
  dw3 = torch.full((9 * Ca * Cb,), 123.0, device="cuda")
  ops.wgrad(x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16), 3, dw3, out_mode=1, ld_out=Cb,
            tap_stride=Ca * Cb, alpha=0.25 if mode == 2 else 1.0, subpixel=mode)
  assert torch.equal(dw, dw3)
