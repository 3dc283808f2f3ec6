#!/usr/bin/env python
"""Benchmark of XMC-GAN's train_step hot path (BASELINE.json metric: images/sec, G+D step, 128 px, bs=56 per GPU).

  python bench.py --gpus N --steps K --warmup W              # B200 arm (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on the host CPU cores

One *step* = one train_step = train_d on B examples + train_g_d on B examples (train_utils.py:91-130), i.e. 2B real
images consumed per device; images/sec = world * 2B / t_step. Data: synthetic COCO-shaped batches (BERT-sized
embeddings + noise), random-init weights. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

E_DIM, N_WORDS = 768, 17
PER_GPU_B = 56  # per-device sub-batch of train_d and of train_g_d (coco_xmc.py:49 with one device)


def make_config(image_size=128, pretrained=True, word_contrastive=True):
  from xmcgan_image_generation_b200.configs import coco_xmc
  c = coco_xmc.get_config()  # reference defaults, incl. pretrained_image_contrastive=True (coco_xmc.py:65)
  c.update(dict(image_size=image_size, pretrained_image_contrastive=bool(pretrained),
                word_contrastive=bool(word_contrastive)))
  return c


def synth_batch(n, config, seed):
  """image U[0,1), embedding N(0,0.5^2), max_len U{3..17}, sentence = sum/len, z N(0,1)."""
  g = torch.Generator().manual_seed(seed)
  S = config.image_size
  emb = torch.randn(n, N_WORDS, E_DIM, generator=g) * 0.5
  max_len = torch.randint(3, N_WORDS + 1, (n, 1), generator=g).float()
  return {"image": torch.rand(n, S, S, 3, generator=g), "embedding": emb, "max_len": max_len,
          "sentence_embedding": emb.sum(1) / max_len, "z": torch.randn(n, config.z_dim, generator=g)}


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons during the timed region."""
  Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.rows, self.proc, self.index = [], None, index

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    self.thread.join(timeout=2)
    sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
    mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
            "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def algorithmic_tflop_per_step(B, pretrained=True, image_size=128, word_contrastive=True):
  """BASELINE.md §3 / SURVEY.md §8(d): 26.93 TFLOP at 128 px, B=56 (25.55 with the ResNet branch off); 37.73 TFLOP
  at 256 px, B=24. The word loss works on 16x16 regions at both resolutions."""
  wl = 0.01337e-3 * B * B if word_contrastive else 0.0  # TFLOP per word_loss forward call
  g, d, r = 43.40e-3, 21.23e-3, 8.18e-3  # TFLOP per image forward (algorithmic G, D, ResNet-50 @224)
  if image_size == 256:
    g, d = 147.03e-3, 73.58e-3
  train_d = g * B + d * 2 * B * 3 + 3 * wl
  train_g_d = 3 * g * B + d * 2 * B * 3 + d * B + 6 * wl + (r * 3 * B if pretrained else 0.0)
  return train_d + train_g_d


class GemmTimer:
  """CUDA-event timing of every tcgen05 GEMM launch (on the launching stream) + its algorithmic FLOPs."""

  def __init__(self):
    self.records = []  # (kernel, flops, start_event, end_event)
    self.bytes = {}    # kernel -> algorithmic HBM bytes over all timed launches

  def install(self, ops):
    self.ops = ops
    self._fwd, self._wgrad = ops.conv_fwd, ops.wgrad
    timer = self

    def conv_fwd(x, wk, kh, cout, **kw):
      xp = kw.get("x_pair")
      if xp is not None:   # fp32 mode: bf16 [.., hi | lo] operand from the producing epilogue; x may be None
        c = xp.shape[3] // 2
        x = x if x is not None else xp[..., :c]
      else:
        c = kw.get("c") or x.shape[3]
      st = kw.get("stride", 1)
      ho, wo, taps = x.shape[1] // st, x.shape[2] // st, kh * kh
      view = kw.get("view")
      if view is not None:  # packed-window convs (ResNet stem 7x7x3, image convs 3x3x3): count the real taps/channels
        ho, wo, taps, c = view["Hout"], view["Wout"], view.get("real_taps", 49), view.get("real_c", 3)
      if kw.get("subpixel"):  # four parities x 2x2 taps per input pixel
        taps = 16
      if kw.get("pre_split"):   # fp32 stem: the packed window carries the hi and lo image parts, count the image once
        c = 3
      # algorithmic FLOPs: operands counted once whatever the number of bf16 passes (fp32 mode runs three), and the
      # transposes of stride-2 convolutions at their real work (alg_scale = 0.25 over a zero-inserted gradient)
      flops = 2.0 * x.shape[0] * ho * wo * taps * c * cout * kw.get("alg_scale", 1.0)
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      out = timer._fwd(x, wk, kh, cout, **kw)
      e.record()
      # algorithmic HBM bytes: the input and the output once (+ residual / mask), the weights once
      outs = [t for t in (out if isinstance(out, tuple) else (out,)) if t is not None]
      nbytes = x.numel() * 4.0 if xp is not None else x.numel() * x.element_size()
      nbytes += sum(t.numel() * t.element_size() for t in outs) + 2.0 * taps * c * cout
      for extra in (kw.get("residual"), kw.get("mask"), kw.get("residual_pair"), kw.get("mask_pair")):
        if extra is not None:
          nbytes += extra.numel() * extra.element_size()
      timer.bytes["gemm_fwd_kernel"] = timer.bytes.get("gemm_fwd_kernel", 0.0) + nbytes
      # bf16 tensor-core passes the launch executes: 3 for fp32 operands (hi*hi + lo*hi + hi*lo), 2 for the fp32 stem
      passes = 3 if (xp is not None or x.dtype == torch.float32) else 2 if kw.get("pre_split") else 1
      timer.records.append(("gemm_fwd_kernel", flops, s, e,
                            (x.shape[0], x.shape[1], x.shape[2], c, kh, cout, int(kw.get("batched", False)),
                             int(kw.get("subpixel", 0))), passes))
      return out

    def wgrad(xa, xb, kh, out, **kw):
      ca = kw.get("ca") or xa.shape[3]
      cb = kw.get("cb") or xb.shape[3]
      taps = 16 if kw.get("subpixel") else kh * kh
      if kw.get("view_a") is not None:  # packed-window 3-channel image wgrad: 9 taps x 3 real channels
        taps, ca = 9, 3
      # pixel grid of the reduction (pool-fused mode: the low-resolution gradient's)
      g = xb if (kw.get("view_a") is not None or kw.get("subpixel") == 2) else xa
      flops = 2.0 * g.shape[0] * g.shape[1] * g.shape[2] * taps * ca * cb
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      r = timer._wgrad(xa, xb, kh, out, **kw)
      e.record()
      timer.records.append(("gemm_wgrad_kernel", flops, s, e,
                            (g.shape[0], g.shape[1], g.shape[2], ca, kh, cb, int(kw.get("batched", False)),
                             int(kw.get("subpixel", 0))), 3 if xa.dtype == torch.float32 else 1))
      return r

    ops.conv_fwd, ops.wgrad = conv_fwd, wgrad

  def uninstall(self):
    self.ops.conv_fwd, self.ops.wgrad = self._fwd, self._wgrad

  def summary(self):
    agg = {}
    for k, fl, s, e, _, passes in self.records:
      a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, {}])
      ms = s.elapsed_time(e)
      a[0] += 1
      a[1] += fl
      a[2] += ms
      a[3] += fl * passes
      b = a[4].setdefault("bf16 operands" if passes == 1 else "fp32 operands (3 bf16 passes)", [0, 0.0, 0.0])
      b[0] += 1
      b[1] += fl
      b[2] += ms
    return {k: {"launches": v[0], "tflop": v[1] / 1e12, "ms": v[2], "executed_tflop": v[3] / 1e12,
                "by_operand": v[4]} for k, v in agg.items()}

  def per_shape(self):
    agg = {}
    for k, fl, s, e, shp, _ in self.records:
      a = agg.setdefault((k,) + shp, [0, 0.0, 0.0])
      a[0] += 1
      a[1] += fl
      a[2] += s.elapsed_time(e)
    rows = [{"kernel": k[0], "N": k[1], "H": k[2], "W": k[3], "C": k[4], "k": k[5], "Cout": k[6], "batched": k[7], "subpixel": k[8],
             "launches": v[0], "ms": round(v[2], 3), "tflops": round(v[1] / 1e12 / (v[2] / 1e3), 1)}
            for k, v in agg.items()]
    return sorted(rows, key=lambda r: -r["ms"])


def run_b200(args):
  from xmcgan_image_generation_b200 import ops, train_utils, xmc_gan
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  from xmcgan_image_generation_b200 import engine
  pretrained = not args.no_pretrained
  config = make_config(args.image_size, pretrained, not args.no_word_contrastive)
  B = args.batch
  config.batch_size = B * world
  host = synth_batch(2 * B, config, 42 + rank)
  pinned = {k: v.pin_memory() for k, v in host.items()}
  dev = {k: v.cuda() for k, v in host.items()}
  generator, discriminator, state = train_utils.create_train_state(config, 42, host)
  # frozen ResNet-50 with synthetic weights (the reference's checkpoint is not shipped): same seed on every rank
  additional = xmc_gan.create_additional_data(
      config, variables=engine.ResNetEngine().random_variables(7) if pretrained else None,
      image_model_dtype=args.resnet_dtype)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(ms):
    if world == 1:
      return ms
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

  # ---- warm-up ------------------------------------------------------------------------------------------------------
  metrics = None
  for _ in range(args.warmup):
    state, metrics = train_utils.train_step(None, state, dev, xmc_gan, generator, discriminator, config, additional)
  barrier()
  # ---- roofline pass: K eager steps with every tcgen05 GEMM launch bracketed by CUDA events on its stream. Kept out of
  # the timed pass below (~500 event records per step cost a few per cent of step time) and run BEFORE the graph is
  # built: eager launches issued after a capture allocate from a second memory pool and were measured 2x slower.
  timer = GemmTimer()
  timer.install(ops)
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    state, metrics = train_utils.train_step(None, state, dev, xmc_gan, generator, discriminator, config, additional)
  e1.record()
  barrier()
  ms_instr = e0.elapsed_time(e1)
  gemm = timer.summary()
  timer_bytes = dict(timer.bytes)
  if args.dump_gemm and rank == 0:
    with open(args.dump_gemm, "w") as f:
      json.dump({"steps": args.steps, "rows": timer.per_shape()}, f, indent=0)
  timer.uninstall()

  eager_step = lambda batch: train_utils.train_step(None, state, batch, xmc_gan, generator, discriminator, config,
                                                    additional)
  launches_per_step = None
  # --graph 1 (default): the step replayed from one CUDA graph at every N (captured NCCL all-reduces included).
  # Tearing the process group down while a graph still holds captured NCCL kernels hung at exit in round 1; the ranks
  # now leave through a final barrier + os._exit (below) instead of destroy_process_group. --graph 0 = eager.
  use_graph = args.graph >= 1
  if use_graph:
    # the public graphed entry point: the whole train_step replayed from one CUDA graph (train_utils.GraphedTrainStep)
    n0 = ops.LAUNCHES[0]
    graphed = train_utils.GraphedTrainStep(state, dev, xmc_gan, generator, discriminator, config, additional, warmup=1)
    launches_per_step = (ops.LAUNCHES[0] - n0) // 2   # one eager warm-up step + the captured step
    state = graphed.state
    step_fn = graphed
    barrier()
  else:
    step_fn = eager_step

  # ---- timed: inputs resident in HBM -----------------------------------------------------------------------------------
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  launches0 = ops.LAUNCHES[0]
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for _ in range(args.steps):
    state, metrics = step_fn(dev)
  e1.record()
  barrier()
  ms_dev = max_over_ranks(e0.elapsed_time(e1))
  launches = ops.LAUNCHES[0] - launches0 if launches_per_step is None else launches_per_step * args.steps
  clocks = sampler.stop() if rank == 0 else None
  last = metrics.compute()

  # ---- timed: end to end through the public API with host buffers ------------------------------------------------------
  h2d = sum(v.numel() * v.element_size() for v in pinned.values())
  barrier()
  e0.record()
  for _ in range(args.steps):
    if use_graph:   # the graphed step copies the pinned host batch into its static device buffers itself
      state, metrics = step_fn(pinned)
    else:
      step_in = {k: v.cuda(non_blocking=True) for k, v in pinned.items()}
      state, metrics = eager_step(step_in)
    last = metrics.compute()  # device -> host read of the step's result
  e1.record()
  barrier()
  ms_e2e = max_over_ranks(e0.elapsed_time(e1))

  if world > 1:
    # every rank is past its last collective (max_over_ranks above); with a graph alive the communicator is not torn
    # down (see --graph): flush and leave the process directly
    barrier()
    if use_graph and rank != 0:
      sys.stdout.flush()
      os._exit(0)
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  peaks = {}
  ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(ppath):
    peaks = json.load(open(ppath))
  peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
  peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
  dom = max(gemm, key=lambda k: gemm[k]["ms"]) if gemm else None
  roofline = None
  if dom:
    ach = gemm[dom]["tflop"] / (gemm[dom]["ms"] / 1e3)
    # measured DRAM traffic of the dominant kernel: one ncu capture of a train_step of this configuration
    # (tools/summarize_ncu.py traffic), per launch like `achieved`; next to it the algorithmic bytes per launch
    traffic, traffic_src = None, None
    tname = f"r02_gemm_fwd_traffic_{args.resnet_dtype if pretrained else 'noresnet'}.json"
    tpath = os.path.join(ROOT, "profiles", tname)
    if dom == "gemm_fwd_kernel" and os.path.exists(tpath) and args.image_size == 128 and B == PER_GPU_B:
      tj = json.load(open(tpath))
      traffic, traffic_src = round(tj["dram_bytes_per_launch"]), f"profiles/{tname} ({tj['launches']} launches of one step)"
    alg_bytes = timer_bytes.get(dom, 0.0) / max(gemm[dom]["launches"], 1)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": round(ach, 1), "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(ach / peak_tf, 4), "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": round(alg_bytes),
                "traffic_note": "average over the launches (many shapes) of one step; the forward figure includes "
                                "conv3x3_resident_kernel launches. Per-shape captures: profiles/r01_gemm_shape_classes.md, "
                                "r01_resident_full.md",
                "peak_source": peak_src,
                # `achieved` counts every multiply once whatever the operand precision; the frozen ResNet branch runs
                # fp32 operands as three bf16 tensor-core passes, so the tensor pipe executes `executed` TFLOP/s
                "executed": {"achieved": round(gemm[dom]["executed_tflop"] / (gemm[dom]["ms"] / 1e3), 1),
                             "frac": round(gemm[dom]["executed_tflop"] / (gemm[dom]["ms"] / 1e3) / peak_tf, 4)},
                "by_operand": {k: {"launches_per_step": v[0] // args.steps, "ms_per_step": round(v[2] / args.steps, 2),
                                   "achieved": round(v[1] / 1e12 / (v[2] / 1e3), 1),
                                   "frac": round(v[1] / 1e12 / (v[2] / 1e3) / peak_tf, 4)}
                               for k, v in gemm[dom]["by_operand"].items()},
                "launches_timed": gemm[dom]["launches"],
                "share_of_step": round(gemm[dom]["ms"] / ms_instr, 3),
                "timed_in": "second pass of the same K steps with per-launch CUDA events "
                            f"({round(ms_instr / args.steps, 3)} ms/step incl. event overhead)",
                "all_gemm": {k: {"tflops": round(v["tflop"] / (v["ms"] / 1e3), 1), "ms_per_step":
                                 round(v["ms"] / args.steps, 2)} for k, v in gemm.items()}}
  imgs = world * 2 * B * args.steps
  alg_tf = algorithmic_tflop_per_step(B, pretrained, config.image_size, config.word_contrastive)
  out = {
      "metric": f"images/sec (G+D train_step, {config.image_size}px, bs={B} per GPU)",
      "value": round(imgs / (ms_dev / 1e3), 2),
      "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "bf16", "data": "synthetic",
      "config": {"workload": f"coco_xmc.py {config.image_size}px, per-GPU sub-batch B={B} (2B real images per step), "
                             "train_d + train_g_d, Adam, EMA, grad all-reduce",
                 "global_batch": B * world, "parallelism": f"dp{world}",
                 "pretrained_image_contrastive": pretrained, "resnet_dtype": args.resnet_dtype if pretrained else None, "word_contrastive": bool(config.word_contrastive),
                 "cuda_graph": bool(use_graph),
                 "l2": "per-step working set (several GB of activations) >> 126 MB L2; no explicit flush",
                 "algorithmic_tflop_per_step_per_gpu": round(alg_tf, 2),
                 "model_tflops_per_gpu": round(alg_tf / (ms_dev / args.steps / 1e3), 1)},
      "e2e": {"value": round(imgs / (ms_e2e / 1e3), 2), "unit": "images/sec", "h2d_bytes_per_step": h2d,
              "d2h_bytes_per_step": 20},
      "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "losses": last,
  }
  if world == 1 and not args.no_cpu_baseline:
    out["cpu_baseline"] = cpu_baseline(args, quick=True)
  print(json.dumps(out), flush=True)
  if world > 1:
    if use_graph:   # see --graph: leave without tearing the communicator down under a live graph
      sys.stdout.flush()
      os._exit(0)
    dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
def _oracle_state(config, seed=1):
  from oracle import xmc_oracle as orc
  from xmcgan_image_generation_b200 import engine
  g = engine.GeneratorEngine(config, E_DIM)
  d = engine.DiscriminatorEngine(config, E_DIM)
  gp = engine.init_flat(g.layout, seed, engine._kind)
  gs = engine.init_flat(g.stats_layout, seed + 1, engine._kind)
  dp = engine.init_flat(d.layout, seed + 2, engine._kind)
  du = engine.init_flat(d.u_layout, seed + 3, engine._kind)
  return orc.make_state({"params": g.layout.tree(gp), "batch_stats": g.stats_layout.tree(gs)},
                        {"params": d.layout.tree(dp), "spectral_norm_stats": d.u_layout.tree(du)})


CPU_B = 8   # BASELINE.json configs[0]: "coco_xmc.py 128px bs=8 world_size=1 on ... CPU backend" -> per-device sub-batch 8


def cpu_baseline(args, quick):
  """The reference algorithm (oracle restatement, torch-CPU fp32 ops, all host threads) on BASELINE config 1
  (128 px, per-device sub-batch B = 8, i.e. 16 real images per train_step, full-width networks, ResNet branch on).
  quick (the default bench run): one warm-up train_step at B = 2 (thread pool, oneDNN primitives), then ONE timed step
  at B = 8 — the bounded sample. --impl reference: one warm-up step at B = 8, then as many timed steps (<= --steps) as
  fit in ~150 s, median."""
  from oracle import xmc_oracle as orc
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  pretrained = not args.no_pretrained
  config = make_config(args.image_size, pretrained)
  pre = None
  if pretrained:
    rvars = orc.resnet50_random_variables(7)
    pre = lambda real, fake: orc.calculate_contrastive_loss_on_pretrained(rvars, real, fake, orc.FP32)
  state = _oracle_state(config)
  Bc = CPU_B
  batch = synth_batch(2 * Bc, config, 42)

  def step(st, bt):
    t0 = time.time()
    st, _ = orc.train_step(st, bt, config, orc.FP32, pretrained_fn=pre)
    return st, time.time() - t0

  if quick:
    step(state, synth_batch(4, config, 41))       # warm-up at B = 2 (result discarded)
    state, dt = step(state, batch)
    times, note = [dt], "1 timed train_step after a warm-up step at B=2"
  else:
    state, first = step(state, batch)
    n = max(2, min(max(1, args.steps), int(150.0 / max(first, 1e-3))))
    times = []
    for _ in range(n):
      state, dt = step(state, batch)
      times.append(dt)
    note = f"median of {n} timed train_steps after 1 warm-up step (requested --steps {args.steps}, capped to ~150 s)"
  dt = sorted(times)[len(times) // 2]
  return {"value": round(2 * Bc / dt, 4), "unit": "images/sec", "cores": cores, "kind": "port",
          "sample": f"{note}; torch-CPU fp32 restatement of the reference at per-device sub-batch B={Bc} "
                    f"({2 * Bc} real images per step, BASELINE config 1), {args.image_size}px, full-width networks",
          "sec_per_step": round(dt, 2), "batch": Bc}


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cb = cpu_baseline(args, quick=False)
  out = {"impl": "reference", "metric": "images/sec (G+D train_step, 128px, bs=56 per GPU)", "value": cb["value"],
         "unit": "images/sec", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
         "ms_per_step": round(cb["sec_per_step"] * 1e3, 1), "higher_is_better": True, "scaling": "weak",
         "vs_baseline": None, "dtype": "f32", "data": "synthetic",
         "config": {"workload": f"coco_xmc.py 128px train_step, reference algorithm (CPU restatement: JAX/Flax are not "
                                f"installable here); bounded sample: per-device sub-batch B={cb['batch']} instead of 56",
                    "sample_batch": cb["batch"], "pretrained_image_contrastive": not args.no_pretrained},
         "cpu_baseline": cb,
         "e2e": {"value": cb["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
  print(json.dumps(out), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--batch", type=int, default=PER_GPU_B, help="per-GPU sub-batch B")
  ap.add_argument("--image-size", type=int, default=128)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-pretrained", action="store_true",
                  help="switch the frozen ResNet-50 image-image InfoNCE branch off (reference default: on)")
  ap.add_argument("--no-word-contrastive", action="store_true",
                  help="BASELINE config 5 (attention ablation): discriminator-side word_loss off (xmc_net.py:112)")
  ap.add_argument("--resnet-dtype", default="float32", choices=["float32", "bfloat16"],
                  help="precision of the frozen ResNet-50 branch: float32 = the reference's (default), bfloat16 = faster")
  ap.add_argument("--graph", type=int, default=1, choices=[0, 1, 2],
                  help="1: train_utils.GraphedTrainStep (train_step replayed from a CUDA graph) on one GPU, eager "
                       "train_step under torchrun; 2: graph replay for any N; 0: eager everywhere")
  ap.add_argument("--dump-gemm", default=None, help="write per-shape GEMM timings (JSON) to this file")
  args = ap.parse_args()
  if args.warmup < 3 and args.impl == "b200":
    args.warmup = 3
  if args.impl == "reference":
    run_reference(args)
  else:
    run_b200(args)


if __name__ == "__main__":
  main()
