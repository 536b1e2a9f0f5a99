"""Benchmark of the SigLIP ViT-B/16 two-tower training step (BASELINE.json metric:
image-text pairs/sec; config 4 at weak scaling: 1024 pairs per GPU, global batch 1024*N).

  python bench.py --gpus 1 --steps 8 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference's algorithm on the host cores (oracle port)

A "step" = update_fn: two-tower forward, pairwise sigmoid loss over all-gathered text
embeddings, backward, gradient all-reduce, fused Adam.  Rank 0 prints ONE JSON line.
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

MODEL_KW = dict(
    image=dict(variant="B/16", pool_type="map"),
    text=dict(variant="B", vocab_size=32_000),
    out_dim=(None, 768), temperature_init=10.0, bias_init=-10.0)
OPT_CONFIG = dict(optax_name="scale_by_adam", optax=dict(b2=0.95, mu_dtype="bfloat16"), lr=1e-3,
                  wd=1e-4, grad_clip_norm=1.0, schedule=dict(decay_type="cosine", warmup_steps=10))
RES, TXT_LEN = 224, 64
# algorithmic training FLOPs per image-text pair (3 x forward; SURVEY.md 8d / BASELINE.md 3)
FLOPS_PER_PAIR = 139.3e9
WORKLOAD = ("SigLIP two_towers ViT-B/16 (map pool) + text-B (64 tok, vocab 32000), 224x224, full update_fn "
            "(fwd, sigmoid loss over gathered ztxt, bwd, grad all-reduce, Adam)")
# ncu-measured DRAM traffic per GEMM launch (all GEMM launches of one bench run, see profiles/)
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = 0.927e9


def measured_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return d, "measured"
  return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.gpu_index, self.rows, self.proc = gpu_index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
           "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      if len(r) < 8:
        continue
      try:
        sm.append(float(r[1]))
        mx.append(float(r[2]))
      except ValueError:
        continue
      for nm, v in zip(names, r[4:8]):
        if v.lower().startswith("active"):
          reasons.add(nm)
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(n, seed):
  import numpy as np
  rng = np.random.default_rng(seed)
  image = rng.uniform(-1, 1, size=(n, RES, RES, 3)).astype(np.float32)
  text = np.ones((n, TXT_LEN), dtype=np.int32)
  lens = rng.integers(4, TXT_LEN, size=n)
  for i in range(n):
    text[i, :lens[i]] = rng.integers(2, 32_000, size=lens[i])
  return image, text


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_port_step(pairs, threads=None):
  """One fwd+bwd of the SigLIP B/16 loss on `pairs` pairs with the CPU oracle; returns seconds."""
  import numpy as np
  import torch
  from oracle import bv_oracle as O
  from big_vision_b200.models import vit
  from big_vision_b200.models.proj.image_text import two_towers
  if threads:
    torch.set_num_threads(threads)
  if not hasattr(cpu_port_step, "_state"):
    model = two_towers.Model(**MODEL_KW)
    P = model.init(0, (pairs, RES, RES, 3), (pairs, TXT_LEN), device="cpu")
    tree = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in P.numpy_tree("f").items()}
    cfg = {"image": dict(depth=12, num_heads=12, pool_type="map", posemb="learn", rep_size=False, num_classes=None),
           "text": dict(depth=12, num_heads=12, pool_type="last", num_classes=768)}
    cpu_port_step._state = (tree, cfg)
  tree, cfg = cpu_port_step._state
  image, text = synthetic_batch(pairs, 0)
  t0 = time.perf_counter()
  for v in tree.values():
    v.grad = None
  # the port in fp32 (the reference's CPU default dtype) -- same algorithm, float32 arithmetic
  O.F64 = torch.float32
  zimg, ztxt, ex = O.two_towers_forward(tree, torch.from_numpy(image), torch.from_numpy(text), cfg, "float32")
  loss = O.siglip_loss(zimg, ztxt, ex["t"], ex["b"])
  loss.backward()
  O.F64 = torch.float64
  return time.perf_counter() - t0


def usable_host_threads():
  """Threads the CPU legs may use: torch's default intra-op pool, capped by the scheduler affinity
  and the cgroup CPU quota.  More threads than runnable CPUs makes the OpenMP barriers of these
  small-batch ops spin against each other (observed: minutes per step instead of seconds)."""
  import math
  import torch
  n = torch.get_num_threads()
  try:
    n = min(n, len(os.sched_getaffinity(0)))
  except AttributeError:
    pass
  try:
    quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
    if quota != "max":
      n = min(n, max(1, math.ceil(int(quota) / int(period))))
  except (OSError, ValueError):
    pass
  return max(1, n)


def run_reference(args):
  import torch
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  threads = usable_host_threads()
  # Each step is a bounded sample of the workload: `pairs` image-text pairs through fwd + bwd of the
  # oracle port.  K and W are honoured; the sample shrinks (8 -> 4 -> 2 -> 1 pairs) if the first
  # step shows that W + K steps would not finish within ~3 minutes on this host.
  pairs, budget_s = 8, 180.0
  warmup = max(1, args.warmup)
  steps = max(1, args.steps)
  t_first = cpu_port_step(pairs, threads)          # warm-up step 1 (allocations, MKL plans)
  est = t_first * (warmup - 1 + steps)
  while est > budget_s and pairs > 1:
    pairs //= 2
    est /= 2
  if est > budget_s:                               # pathological host: keep the run bounded anyway
    steps = max(1, int(budget_s / (est / (warmup - 1 + steps))) - (warmup - 1))
  for _ in range(warmup - 1):
    cpu_port_step(pairs, threads)
  t = sum(cpu_port_step(pairs, threads) for _ in range(steps))
  val = pairs * steps / t
  line = {
      "impl": "reference", "metric": "siglip_vit_b16_pairs_per_sec", "value": val, "unit": "pairs/s",
      "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t / steps,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {"workload": WORKLOAD, "global_batch": pairs, "parallelism": f"cpu{threads}",
                 "sample": "fwd+bwd of the pairwise sigmoid loss through both towers on a bounded "
                           "sample of the batch (no optimizer step)"},
      "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port",
                       "sample": f"{steps} steps x {pairs} pairs, oracle port (torch-CPU fp32, no optimizer)"},
      "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import torch.distributed as dist
  from big_vision_b200 import lib as L
  from big_vision_b200 import ops
  from big_vision_b200 import optax as bv_optax
  from big_vision_b200.models.proj.image_text import two_towers
  from big_vision_b200.trainers.proj.image_text import siglip

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a GPU (there is no CPU fallback for the kernels)")
  torch.cuda.set_device(local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  if L.load().bv_device_supported() != 1:
    raise SystemExit("bench.py needs a compute-capability 10.x device")
  n = args.per_gpu_batch
  model = two_towers.Model(**MODEL_KW)
  P = model.init(0, (n, RES, RES, 3), (n, TXT_LEN), device="cuda")
  tx, _ = bv_optax.make(OPT_CONFIG, P, sched_kw=dict(total_steps=10_000, batch_size=n * world,
                                                     data_size=10_000_000))
  state = {"params": P, "opt": tx.init(P)}
  update_fn = siglip.make_update_fn(model, tx, OPT_CONFIG)
  image_h, text_h = synthetic_batch(n, seed=rank)
  image_pin = torch.from_numpy(image_h).pin_memory()
  text_pin = torch.from_numpy(text_h).pin_memory()
  image_d, text_d = image_pin.cuda(), text_pin.cuda()
  batch = {"image": image_d, "labels": text_d}

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident timing -----------------------------------------------------------------
  for _ in range(args.warmup):
    state, m = update_fn(state, None, batch)
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  ops_gemm = ops.gemm
  gemm_events, gemm_flops, gemm_bytes = [], [0.0], [0.0]

  def timed_gemm(a, b, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops_gemm(a, b, **kw)
    e1.record()
    a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
    M = a.shape[1] if a_mn else a.shape[0]
    K = a.shape[0] if a_mn else a.shape[1]
    N = b.shape[1] if b_mn else b.shape[0]
    gemm_flops[0] += 2.0 * M * N * K
    # algorithmic bytes of this launch: both operands once, every output once, the epilogue operand
    o = out[0] if isinstance(out, tuple) else out
    nbytes = 2.0 * K * (M + N) + M * N * o.element_size() * (2 if isinstance(out, tuple) else 1)
    if kw.get("aux") is not None:
      nbytes += 2.0 * M * N if not kw.get("aux_row_mod") else 2.0 * kw["aux_row_mod"] * N
    gemm_bytes[0] += nbytes
    gemm_events.append((e0, e1))
    return out

  launches0 = L.LAUNCHES[0]
  ops.gemm = timed_gemm
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  ev0.record()
  for _ in range(args.steps):
    state, m = update_fn(state, None, batch)
  ev1.record()
  barrier()
  ops.gemm = ops_gemm
  launches = L.LAUNCHES[0] - launches0
  ms = ev0.elapsed_time(ev1)
  gemm_ms = sum(a.elapsed_time(b) for a, b in gemm_events)
  clocks = sampler.stop() if rank == 0 else None
  loss = float(m["training_loss"])

  # ---- end to end: host buffers in, loss out, every step --------------------------------------
  # The public input API (input_pipeline.start_input_pipeline, the reference's prefetch-to-device
  # iterator) uploads step i+1's batch from pinned host memory on a side stream while step i
  # computes; every step's loss is copied back to pinned host memory.  All of it is inside the
  # timed region, which ends after the last step's loss has landed on the host.
  from big_vision_b200 import input_pipeline

  def host_batches(k=None):
    for _ in range(args.steps if k is None else k):
      yield {"image": image_pin, "labels": text_pin}

  # untimed warm-up of the end-to-end path itself (side stream, device slots of the prefetcher:
  # a first-use cudaMalloc would otherwise synchronise the device inside the timed region)
  if os.environ.get("BV_E2E") != "serial":
    for dev_batch in input_pipeline.start_input_pipeline(
        host_batches(2), n_prefetch=int(os.environ.get("BV_E2E_PREFETCH", "1"))):
      state, m = update_fn(state, None, dev_batch)
    del dev_batch

  loss_host = torch.empty(args.steps, dtype=torch.float32).pin_memory()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  if os.environ.get("BV_E2E") == "serial":     # A/B switch: copy, step, blocking read, in order
    for i in range(args.steps):
      image_d.copy_(image_pin, non_blocking=True)
      text_d.copy_(text_pin, non_blocking=True)
      state, m = update_fn(state, None, {"image": image_d, "labels": text_d})
      loss_host[i] = m["training_loss"].item()
  else:
    n_pre = int(os.environ.get("BV_E2E_PREFETCH", "1"))
    for i, dev_batch in enumerate(input_pipeline.start_input_pipeline(host_batches(), n_prefetch=n_pre)):
      state, m = update_fn(state, None, dev_batch)
      loss_host[i:i + 1].copy_(m["training_loss"].reshape(1), non_blocking=True)   # device -> host
  e1.record()
  barrier()
  ms_e2e = e0.elapsed_time(e1)
  assert bool(torch.isfinite(loss_host).all()), loss_host

  if args.profile_calls:      # every rank runs the extra step (collectives); rank 0 prints
    import collections
    L.PROFILE = []
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pe0.record()
    state, m = update_fn(state, None, batch)
    pe1.record()
    torch.cuda.synchronize()
    prof, L.PROFILE = L.PROFILE, None
    tot, cnt = collections.defaultdict(float), collections.Counter()
    flops = collections.defaultdict(float)
    for name, a, b, fl in prof:
      dt = a.elapsed_time(b)
      tot[name] += dt
      cnt[name] += 1
      flops[name] += fl
      if name.startswith("bv_gemm "):
        tot["bv_gemm (all)"] += dt
        cnt["bv_gemm (all)"] += 1
        flops["bv_gemm (all)"] += fl
    step_ms = pe0.elapsed_time(pe1)
    ssum = sum(v for k, v in tot.items() if k != "bv_gemm (all)")
    if rank == 0:
      print(f"[profile-calls] step {step_ms:.2f} ms, sum of kernel spans {ssum:.2f} ms, "
            f"gap {step_ms - ssum:.2f} ms", file=sys.stderr)
      for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        tf = f"  {flops[k] / v * 1e-9:7.1f} TFLOP/s" if flops[k] else ""
        print(f"[profile-calls]   {k:44s} {v:8.2f} ms  n={cnt[k]:4d}{tf}", file=sys.stderr)

  t = torch.tensor([ms, ms_e2e, gemm_ms], dtype=torch.float64, device="cuda")
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms, ms_e2e, gemm_ms = (float(x) for x in t.tolist())

  if rank == 0:
    peaks, peak_src = measured_peaks()
    pairs = n * world * args.steps
    value = pairs / (ms * 1e-3)
    e2e_val = pairs / (ms_e2e * 1e-3)
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    gemm_tf = gemm_flops[0] / (gemm_ms * 1e-3) / 1e12
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
      steps_cpu, pairs_cpu = 2, 8
      nthr = usable_host_threads()
      t_warm = cpu_port_step(pairs_cpu, nthr)
      steps_cpu = max(1, min(steps_cpu, int(60.0 / max(t_warm, 1e-3))))
      tt = sum(cpu_port_step(pairs_cpu, nthr) for _ in range(steps_cpu))
      cpu = {"value": pairs_cpu * steps_cpu / tt, "unit": "pairs/s", "cores": nthr,
             "kind": "port", "sample": f"{steps_cpu} steps x {pairs_cpu} pairs fwd+bwd, oracle port in "
                                       "torch-CPU fp32 (no optimizer step)"}
    line = {
        "metric": "siglip_vit_b16_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": n * world, "per_gpu_batch": n, "seq_len": 196 + TXT_LEN,
                   "parallelism": f"dp{world}", "l2_policy": "inputs (616 MB images/step) and "
                   "activations (>70 GB) exceed the 126 MB L2; no explicit flush",
                   "final_loss": loss},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": image_pin.numel() * 4 + text_pin.numel() * 4,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_kernel (tcgen05 persistent GEMM)",
                     "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf,
                     "peak_source": f"{peak_src} bf16_tflops_sustained",
                     # per launch, averaged over the step's GEMM launches (shapes differ)
                     "launches_per_step": len(gemm_events) // args.steps,
                     "flop_per_launch": gemm_flops[0] / len(gemm_events),
                     "algorithmic_bytes_per_launch": gemm_bytes[0] / len(gemm_events),
                     "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum over the GEMM "
                                       "launches of profiles/r01_final_launch_summary.md",
                     "gemm_share_of_step": gemm_ms / ms,
                     "step_mfu": value / world * FLOPS_PER_PAIR / 1e12 / peak_tf},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=8)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--per-gpu-batch", type=int, default=1024)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--profile-calls", action="store_true",
                  help="time every C-ABI call of one extra step with CUDA events; breakdown on stderr")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    if args.warmup < 3:
      args.warmup = 3
    run_ours(args)


if __name__ == "__main__":
  main()
