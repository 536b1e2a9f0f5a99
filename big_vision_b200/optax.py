"""Optimizer factory -- the chain `big_vision/optax.py:75-149` builds, executed as fused CUDA
launches over the flat parameter buffer:

  clip_by_global_norm(grad_clip_norm)          (norm over the NON-frozen gradients, :104-113)
  -> inner transform `config.optax_name`       scale_by_adam (fused, `bv_adam_step`) | scale (`bv_scale_step`)
  -> scale(lr) [* lr_mults, first match]       (:120-129)
  -> add_decayed_weights(wd * wd_mults)        (first match; default mask ".*/kernel$", :136-145)
  -> scale_by_schedule, one per config.schedule pattern (first match; None = frozen, :79-101)
  -> scale(-1), applied to the parameters in place (optax.apply_updates).

Every stored parameter gets (schedule index | frozen, lr multiplier, weight decay) from the regex
lists exactly like `u.make_mask_trees` assigns them (FIRST matching pattern wins, full match on the
reference name "a/b/c"); neighbours in the flat layout with the same setting are merged into one
launch -- two launches for the default config (decayed kernels | everything else).  Frozen ranges
get no launch and no optimizer state (optax_test.py:301-317) and do not count towards the clipping
norm or `l2_grads` (optax_test.py:206-299, siglip.py:315-321).

Not built: BV-Adafactor (`big_vision.scale_by_adafactor`, optax.py:187-214) and per-example clipping
raise instead of silently doing something else.
"""
import re

import torch

from big_vision_b200 import ops
from big_vision_b200 import utils as u


def _first_match(patterns, names):
  """Index of the first pattern that fully matches ALL reference names of one stored parameter
  (a fused q|k|v kernel carries three names); None if no pattern matches any of them.  Names of one
  storage that would be assigned differently cannot be honoured and raise."""
  hits = set()
  for name in names:
    hit = next((i for i, p in enumerate(patterns) if p.fullmatch(name)), None)
    hits.add(hit)
  if len(hits) > 1:
    raise NotImplementedError(f"{sorted(names)} share one fused storage tensor but match different "
                              "optimizer patterns")
  return hits.pop()


class Chain:
  """tx-like object: init(P) -> opt state, update(P, opt, ...) applies one step in place."""

  def __init__(self, config, P, sched_kw):
    name = config.get("optax_name", "scale_by_adam")
    kw = dict(config.get("optax", {}) or {})
    if name == "scale_by_adam":
      self.b1, self.b2, self.eps = kw.pop("b1", 0.9), kw.pop("b2", 0.999), kw.pop("eps", 1e-8)
      mu_dtype = kw.pop("mu_dtype", None)
      self.mu_dtype = torch.bfloat16 if mu_dtype in ("bfloat16", torch.bfloat16) else torch.float32
      if kw.pop("eps_root", 0.0):
        raise NotImplementedError("eps_root")
      self.step_size = None
    elif name == "scale":
      self.step_size = float(kw.pop("step_size"))
    else:
      raise NotImplementedError(f"optax_name={name}: built are scale_by_adam and scale")
    if kw:
      raise NotImplementedError(f"{name} options {sorted(kw)}")
    if config.get("grad_clip_per_example"):
      raise NotImplementedError("grad_clip_per_example")
    if not config.get("weight_decay_decouple", True):
      raise AssertionError("Coupled weight decay not supported anymore.")
    self.inner = name
    self.lr = float(config.get("lr", 1e-3))
    self.clip = float(config.get("grad_clip_norm", 0.0) or 0.0)

    # ---- schedules (first match; None = frozen) ------------------------------------------------
    schedule = config.get("schedule", {})
    if not isinstance(schedule, (tuple, list)):
      schedule = [(".*", schedule)]
    sched_pat = u.check_and_compile_patterns([p for p, _ in schedule])
    self.sched_fns, sched_slot = [], []
    for _, sc in schedule:
      if sc is None:
        sched_slot.append(None)
      else:
        sc = dict(sc)
        if "base" in sc:
          raise AssertionError(sc)
        sched_slot.append(len(self.sched_fns))
        self.sched_fns.append(u.create_learning_rate_schedule(base=sc.pop("mult", 1.0), **sched_kw, **sc))
    lr_mults = list(config.get("lr_mults") or [])
    if not all(m > 0 for _, m in lr_mults):
      raise AssertionError(f"Use schedule=None for parameter freezing instead of lr_mults={lr_mults}")
    lr_pat = u.check_and_compile_patterns([p for p, _ in lr_mults]) if lr_mults else []
    wd = float(config.get("wd", 0.0) or 0.0)
    wd_mults = list(config.get("wd_mults", [(".*/kernel$", 1.0)])) if wd else []
    wd_pat = u.check_and_compile_patterns([p for p, _ in wd_mults]) if wd_mults else []

    # ---- per stored parameter -> merged launch ranges over the flat layout -----------------------
    names_of = {}
    for a in P.aliases.values():
      names_of.setdefault(a.storage, []).append(a.name)
    uncovered, self.ranges = [], []      # ranges: [lo, hi, sched slot | None, lr mult, wd]
    for storage, (off, shape) in sorted(P.offsets.items(), key=lambda kv: kv[1][0]):
      names = names_of.get(storage, [storage])
      si = _first_match(sched_pat, names)
      if si is None:
        uncovered += names
        continue
      li = _first_match(lr_pat, names) if lr_pat else None
      wi = _first_match(wd_pat, names) if wd_pat else None
      key = (sched_slot[si], 1.0 if li is None else float(lr_mults[li][1]),
             0.0 if wi is None else wd * float(wd_mults[wi][1]))
      n = 1
      for dim in shape:
        n *= dim
      hi = off + (n + 7) // 8 * 8          # engine.ALIGN: the padding belongs to its parameter
      if self.ranges and tuple(self.ranges[-1][2:]) == key and self.ranges[-1][1] == off:
        self.ranges[-1][1] = hi
      else:
        self.ranges.append([off, hi, *key])
    if uncovered:
      raise AssertionError(f"All params must be covered (use `None` for freezing): {uncovered}")
    # optimizer state only for what is trained, packed in range order
    self.state_off, n_state = [], 0
    for lo, hi, slot, _, _ in self.ranges:
      self.state_off.append(n_state if slot is not None else None)
      if slot is not None:
        n_state += hi - lo
    self.n_state = n_state

  def init(self, P):
    dev = P.flat.device
    state = {"count": 0, "scalars": torch.zeros(4, dtype=torch.float32, device=dev)}   # [gnorm_sq, upd_sq, param_sq]
    if self.inner == "scale_by_adam":
      state["mu"] = torch.zeros(self.n_state, dtype=self.mu_dtype, device=dev)
      state["nu"] = torch.zeros(self.n_state, dtype=torch.float32, device=dev)
    return state

  def update(self, P, opt, grad_mult=1.0):
    """Applies one step in place; returns the device tensor [gnorm_sq, upd_sq, param_sq, 0]."""
    sc = opt["scalars"]
    sc.zero_()
    trained = [(r, so) for r, so in zip(self.ranges, self.state_off) if r[2] is not None]
    for (lo, hi, *_), _ in trained:                      # norm over the non-frozen gradients only
      ops.sumsq(P.grad[lo:hi], sc[0:1])
    scheds = [fn(opt["count"]) for fn in self.sched_fns]   # evaluated at the pre-increment count
    step = opt["count"] + 1
    for (lo, hi, slot, lr_mult, wd), so in trained:
      sched = scheds[slot]
      common = dict(wd_eff=sched * wd, grad_mult=grad_mult, clip_norm=self.clip, gnorm_sq=sc[0:1],
                    upd_sq=sc[1:2], param_sq=sc[2:3])
      if self.inner == "scale_by_adam":
        ops.adam_step(P.flat[lo:hi], P.grad[lo:hi], opt["mu"][so:so + hi - lo], opt["nu"][so:so + hi - lo],
                      P.half[lo:hi], lr_eff=sched * self.lr * lr_mult, b1=self.b1, b2=self.b2, eps=self.eps,
                      step=step, **common)
      else:
        ops.scale_step(P.flat[lo:hi], P.grad[lo:hi], P.half[lo:hi],
                       lr_eff=sched * self.lr * lr_mult * self.step_size, **common)
    frozen = [r for r in self.ranges if r[2] is None]
    for lo, hi, *_ in frozen:                            # l2_params covers every parameter
      ops.sumsq(P.flat[lo:hi], sc[2:3])
    opt["count"] = step
    return sc


FusedAdam = Chain     # name used by round-1 callers


def make(config, params, *, sched_kw):
  """Returns (tx, schedule_fns) like optax.py:75 `make`; `params` is the model's FlatParams."""
  if "optim" in config:
    raise AssertionError("Deprecated option, use config.optax.")
  if "weight_decay" in config:
    raise AssertionError("Deprecated option. Use wd and schedule.")
  tx = Chain(config, params, sched_kw)
  return tx, list(tx.sched_fns)
