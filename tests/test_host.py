"""CPU-side tests: the C-ABI library loads and exports every declared symbol, parameter layouts reproduce the
reference's variable trees, host logic (config surface, splitting, device groups, error conventions)."""
import ctypes

import pytest
import torch

from tests import helpers
from xmcgan_image_generation_b200 import _lib, engine, parallel, train_utils
from xmcgan_image_generation_b200.configs import coco_xmc


def test_library_exports_every_declared_symbol():
  decl = _lib.declared_functions()
  assert len(decl) >= 45
  L = _lib.lib()
  for name in decl:
    assert hasattr(L, name), name
  assert L.xmc_version() >= 101
  # struct mirrors: the loader refuses a layout mismatch; spot-check two sizes against the header by hand
  assert L.xmc_sizeof(0) == ctypes.sizeof(_lib.ConvDesc) and L.xmc_sizeof(1) == ctypes.sizeof(_lib.WgradDesc)
  assert L.xmc_sizeof(2) == ctypes.sizeof(_lib.BnDesc) == 12 * 4 and L.xmc_sizeof(99) == -1
  assert L.xmc_strerror(-1).decode().startswith("invalid")


def test_abi_rejects_bad_descriptors_without_touching_the_gpu():
  L = _lib.lib()
  d = _lib.ConvDesc()
  assert L.xmc_conv2d_fwd(ctypes.byref(d), None, None, None, None, None, None, None, None) == -1
  w = _lib.WgradDesc()
  assert L.xmc_conv2d_wgrad(ctypes.byref(w), None, None, None, None, 0, None) == -1
  need = ctypes.c_longlong(-1)
  assert L.xmc_conv2d_wgrad_workspace_bytes(ctypes.byref(w), ctypes.byref(need)) == -1   # empty descriptor
  assert L.xmc_adam(None, None, None, None, 16, 0.1, 0.5, 0.9, 1e-8, 0.5, 0.1, 1.0, None, 0.0, None, 1, None) == -1


def test_parameter_counts_match_reference():
  """G 78 507 779 / D 87 911 713 (128 px); 92 865 539 / 99 415 585 (256 px)."""
  import numpy as np
  for size, ng, nd in ((128, 78507779, 87911713), (256, 92865539, 99415585)):
    c = coco_xmc.get_config()
    c.image_size = size
    g, d = engine.GeneratorEngine(c), engine.DiscriminatorEngine(c)
    assert sum(int(np.prod(s)) for _, s in g.layout.entries.values()) == ng
    assert sum(int(np.prod(s)) for _, s in d.layout.entries.values()) == nd


def test_variable_tree_names_follow_flax_auto_naming():
  c = coco_xmc.get_config()
  g, d = engine.GeneratorEngine(c), engine.DiscriminatorEngine(c)
  gp = g.layout.entries
  assert gp[("Dense_0", "kernel")][1] == (768, 128)
  assert gp[("Dense_1", "kernel")][1] == (128, 24576)
  assert gp[("GenBlock_0", "ConditionalBatchNorm_0", "Dense_1", "kernel")][1] == (256, 1536)
  assert gp[("GenBlock_1", "Conv_0", "kernel")][1] == (3, 3, 1536, 768)
  assert gp[("GenBlock_1", "Conv_2", "kernel")][1] == (1, 1, 1536, 768)
  assert gp[("Conv_0", "kernel")][1] == (1, 1, 768, 768)
  assert gp[("GenSpatialBlock_2", "LocalConditionalBatchNorm_1", "Conv_0", "kernel")][1] == (1, 1, 1024, 96)
  assert gp[("LocalConditionalBatchNorm_0", "Conv_1", "bias")][1] == (96,)
  assert gp[("Conv_1", "kernel")][1] == (3, 3, 96, 3)
  assert len(g.stats_layout.entries) == 22  # 11 BatchNorm layers x (mean, var)
  dp = d.layout.entries
  assert dp[("DiscOptimizedBlock_0", "SpectralConv_0", "kernel")][1] == (3, 3, 3, 96)
  assert dp[("DiscBlock_3", "SpectralConv_2", "kernel")][1] == (1, 1, 768, 1536)
  assert ("DiscBlock_4", "SpectralConv_2", "kernel") not in dp
  assert dp[("SpectralDense_0", "kernel")][1] == (1536, 1)
  assert dp[("SpectralDense_1", "kernel")][1] == (768, 1536)
  assert dp[("SpectralConv_0", "kernel")][1] == (1, 1, 384, 768)
  assert d.u_layout.entries[("SpectralConv_0", "u0")][1] == (1, 768)
  assert d.sntab.n == 20
  # g_spectral_norm=True renames every conv / dense of the generator (also inside the conditional BatchNorms) and adds
  # a spectral_norm_stats collection with one u0 per kernel (xmc_net.py:176-191, layers.py:86-91,203-208)
  c.g_spectral_norm = True
  gs = engine.GeneratorEngine(c)
  sp = gs.layout.entries
  assert sp[("SpectralDense_0", "kernel")][1] == (768, 128)
  assert sp[("GenBlock_0", "ConditionalBatchNorm_0", "SpectralDense_1", "kernel")][1] == (256, 1536)
  assert sp[("GenSpatialBlock_2", "LocalConditionalBatchNorm_1", "SpectralConv_0", "kernel")][1] == (1, 1, 1024, 96)
  assert sp[("SpectralConv_1", "kernel")][1] == (3, 3, 96, 3)
  assert not any(k.startswith(("Conv_", "Dense_")) for path in sp for k in path)
  assert gs.layout.total == g.layout.total
  assert gs.u_layout.entries[("GenBlock_1", "SpectralConv_2", "u0")][1] == (1, 768)
  assert gs.u_layout.entries[("LocalConditionalBatchNorm_0", "SpectralConv_0", "u0")][1] == (1, 96)
  assert gs.sntab.n == 2 + 15 + 2 + 8 + 14  # dense, block convs, attention / output conv, CBN dense, LCBN conv


def test_layout_tree_roundtrip_and_alignment():
  c = helpers.small_config()
  g = engine.GeneratorEngine(c, 64)
  buf = engine.init_flat(g.layout, 3, engine._kind)
  tree = g.layout.tree(buf)
  buf2 = torch.zeros_like(buf)
  g.layout.load_tree(buf2, tree)
  assert torch.equal(buf, buf2)
  assert all(off % 4 == 0 for off, _ in g.layout.entries.values())
  # concatenated-bias blocks are contiguous (one bias vector / one column sum per group)
  assert g.layout.off(g.lcbn[0][0] + ("Conv_0", "bias")) == g.lcbn_bias_off
  assert g.cbn_bias_off == g.lcbn_bias_off + g.NL


def test_config_surface():
  c = coco_xmc.get_config()
  for k in ("dtype", "z_dim", "d_step_per_g_step", "polyak_decay", "pretrained_image_contrastive", "image_size",
            "gf_dim", "df_dim", "g_spectral_norm", "d_spectral_norm", "batch_norm_group_size", "gamma_for_g",
            "word_contrastive", "sentence_contrastive", "image_contrastive", "cond_size", "g_lr", "d_lr", "beta1",
            "beta2", "architecture", "model_name"):
    assert hasattr(c, k), k
  assert (c.batch_size, c.image_size, c.gf_dim, c.df_dim, c.z_dim) == (56, 128, 96, 96, 128)
  assert (c.g_lr, c.d_lr, c.beta1, c.beta2, c.polyak_decay) == (1e-4, 4e-4, 0.5, 0.999, 0.999)
  t = coco_xmc.get_test_config()
  assert (t.batch_size, t.gf_dim, t.df_dim, t.z_dim) == (2, 16, 16, 8)
  c.batch_size = 8
  assert c["batch_size"] == 8


def test_split_input_dict_is_an_exact_index_op():
  batch = {"a": torch.arange(24).reshape(6, 4), "b": torch.arange(6)}
  parts = train_utils.split_input_dict(batch, 2)
  assert torch.equal(parts[0]["a"], batch["a"][:3]) and torch.equal(parts[1]["a"], batch["a"][3:])
  assert torch.equal(parts[1]["b"], torch.tensor([3, 4, 5]))
  with pytest.raises(ValueError):
    train_utils.split_input_dict({"a": torch.zeros(5, 2)}, 2)


def test_error_conventions():
  c = helpers.small_config(image_size=64)
  with pytest.raises(ValueError):
    engine.GeneratorEngine(c, 64)
  with pytest.raises(ValueError):
    engine.DiscriminatorEngine(c, 64)
  c = helpers.small_config(architecture="resnet")
  with pytest.raises(ValueError):
    train_utils.create_train_state(c, 0, helpers.make_batch(2, c))
  from xmcgan_image_generation_b200.libml import attention_lib
  with pytest.raises(NotImplementedError):
    attention_lib.contrastive_loss(torch.ones(2, 4), torch.ones(2, 4), sync_match=True)


def test_device_groups():
  assert parallel.get_device_groups(16, 8, device_count=8) == [[0, 1], [2, 3], [4, 5], [6, 7]]
  assert parallel.get_device_groups(8, 8, device_count=2) == [[0], [1]]
  with pytest.raises(AssertionError):
    parallel.get_device_groups(12, 8, device_count=8)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_has_no_cpu_fallback():
  """Without a CUDA device the product must fail loudly, never compute on the CPU."""
  c = helpers.small_config()
  with pytest.raises(Exception):
    train_utils.create_train_state(c, 0, helpers.make_batch(2, c))


def test_bench_algorithmic_work_matches_survey():
  """algorithmic TFLOP per train_step per device — 26.93 (128 px, B=56), 25.55 without the ResNet
  branch, 3.800 (B=8), 37.73 (256 px, B=24). bench.py's model-FLOP figure is derived from these."""
  import importlib.util
  import os
  spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
  bench = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(bench)
  assert abs(bench.algorithmic_tflop_per_step(56) - 26.93) < 0.01
  assert abs(bench.algorithmic_tflop_per_step(56, pretrained=False) - 25.55) < 0.01
  assert abs(bench.algorithmic_tflop_per_step(8) - 3.800) < 0.005
  assert abs(bench.algorithmic_tflop_per_step(24, True, 256) - 37.73) < 0.01
  cfg = bench.make_config(256, False, False)
  assert (cfg.image_size, cfg.pretrained_image_contrastive, cfg.word_contrastive) == (256, False, False)
  b = bench.synth_batch(4, cfg, 1)
  assert b["image"].shape == (4, 256, 256, 3) and b["max_len"].min() >= 3 and b["max_len"].max() <= 17
  assert torch.allclose(b["sentence_embedding"], b["embedding"].sum(1) / b["max_len"])


def test_wgrad_planner_accepts_every_layer_shape_and_bounds_its_workspace():
  """plan_wgrad (K split, tap groups, ring depth) through its CPU-callable face xmc_conv2d_wgrad_workspace_bytes: every
  weight-gradient shape of the BASELINE configuration (gf = df = 96, B = 56 / 2B = 112, 128 px) in its three tap
  geometries plans without error, a workspace exists exactly when an output element has several producers (the
  sub-pixel / pool-fused forms always, plain taps only with a K split), holds a whole number of partial weight
  tensors, and stays below 256 MB (the cost model charges the round trip, so wide layers are not split)."""
  L = _lib.lib()

  def plan(N, H, Ca, Cb, sub=0, k=3):
    d = _lib.WgradDesc()
    d.N, d.H, d.W, d.Ca, d.Cb, d.ldA, d.ldB = N, H, H, Ca, Cb, Ca, Cb
    d.KH = d.KW = k
    d.pad_h = d.pad_w = k // 2
    d.out_mode, d.ldOut, d.out_tap_stride, d.alpha, d.subpixel = 0, Cb, Ca * Cb, 1.0, sub
    need = ctypes.c_longlong(-1)
    assert L.xmc_conv2d_wgrad_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == 0, (N, H, Ca, Cb, sub)
    return need.value

  chans = [(96, 96), (96, 192), (192, 192), (192, 384), (384, 384), (384, 768), (768, 768), (768, 1536), (1536, 1536)]
  for N in (56, 112):
    for ca, cb in chans:
      for H in (4, 8, 16, 32, 64, 128):
        if H * H * N * max(ca, cb) > 112 * 128 * 128 * 96:
          continue
        for sub in (0, 1, 2):
          need = plan(N, H, ca, cb, sub)
          taps = 16 if sub else 9
          per_split = 4 * taps * ca * cb
          assert need % per_split == 0 and need <= 256 << 20, (N, H, ca, cb, sub, need)
          if sub:
            assert need >= per_split
  # a 1x1 conv of few pixels has one producer per element: no workspace
  assert plan(2, 4, 64, 64, 0, k=1) == 0
  # channel counts that are not multiples of 8 are rejected, not mis-planned
  d = _lib.WgradDesc()
  d.N, d.H, d.W, d.Ca, d.Cb, d.ldA, d.ldB, d.KH, d.KW = 1, 8, 8, 12, 16, 16, 16, 3, 3
  need = ctypes.c_longlong(0)
  assert L.xmc_conv2d_wgrad_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == -1


def test_sm_limit_switch_is_a_plain_setter():
  """xmc_set_sm_limit caps the persistent GEMM grids (host-side state, no GPU needed to set or clear it)."""
  L = _lib.lib()
  assert L.xmc_set_sm_limit(132) == 0 and L.xmc_set_sm_limit(0) == 0 and L.xmc_set_sm_limit(-5) == 0
