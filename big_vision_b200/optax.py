"""Optimizer factory -- mirror of big_vision/optax.py:75-149 for the chain the hot path uses:

  clip_by_global_norm(grad_clip_norm) -> scale_by_adam(**config.optax) -> scale(lr)
  -> add_decayed_weights(wd, mask=".*/kernel$") -> scale_by_schedule -> scale(-1)

executed as ONE fused CUDA launch per weight-decay group over the flat parameter buffer
(bv_adam_step), plus one sum-of-squares launch for the global gradient norm.
Not built yet (SURVEY.md 8f "next" #1): BV-Adafactor, per-pattern schedules / frozen
params, lr_mults.  Those configurations raise instead of silently doing something else.
"""
import math

import torch

from big_vision_b200 import ops
from big_vision_b200 import utils as u


class FusedAdam:
  """tx-like object: init() -> opt state, update(P, opt, ...) applies the step in place."""

  def __init__(self, config, sched_fn):
    name = config.get("optax_name", "scale_by_adam")
    if name != "scale_by_adam":
      raise NotImplementedError(f"optax_name={name}: only scale_by_adam is built on this path")
    if config.get("lr_mults"):
      raise NotImplementedError("lr_mults")
    kw = dict(config.get("optax", {}))
    self.b1 = kw.pop("b1", 0.9)
    self.b2 = kw.pop("b2", 0.999)
    self.eps = kw.pop("eps", 1e-8)
    mu_dtype = kw.pop("mu_dtype", None)
    self.mu_dtype = torch.bfloat16 if mu_dtype in ("bfloat16", torch.bfloat16) else torch.float32
    if kw.pop("eps_root", 0.0):
      raise NotImplementedError("eps_root")
    if kw:
      raise NotImplementedError(f"scale_by_adam options {sorted(kw)}")
    self.lr = float(config.get("lr", 1e-3))
    self.wd = float(config.get("wd", 0.0) or 0.0)
    self.clip = float(config.get("grad_clip_norm", 0.0) or 0.0)
    self.sched_fn = sched_fn

  def init(self, P):
    dev = P.flat.device
    return {
        "mu": torch.zeros(P.total, dtype=self.mu_dtype, device=dev),
        "nu": torch.zeros(P.total, dtype=torch.float32, device=dev),
        "count": 0,
        # [gnorm_sq, upd_sq, param_sq]
        "scalars": torch.zeros(4, dtype=torch.float32, device=dev),
    }

  def update(self, P, opt, grad_mult=1.0):
    """Applies one step in place; returns the device tensor [gnorm_sq, upd_sq, param_sq, 0]."""
    sc = opt["scalars"]
    sc.zero_()
    ops.sumsq(P.grad, sc[0:1])
    sched = self.sched_fn(opt["count"])   # schedule evaluated at the pre-increment count
    step = opt["count"] + 1
    groups = [(0, P.n_decay, self.wd), (P.n_decay, P.total, 0.0)]
    for lo, hi, wd in groups:
      if hi <= lo:
        continue
      ops.adam_step(P.flat[lo:hi], P.grad[lo:hi], opt["mu"][lo:hi], opt["nu"][lo:hi], P.half[lo:hi],
                    lr_eff=sched * self.lr, b1=self.b1, b2=self.b2, eps=self.eps,
                    wd_eff=sched * wd, step=step, grad_mult=grad_mult, clip_norm=self.clip,
                    gnorm_sq=sc[0:1], upd_sq=sc[1:2], param_sq=sc[2:3])
    opt["count"] = step
    return sc


def make(config, params, *, sched_kw):
  """Returns (tx, [schedule_fn]) like optax.py:75 `make` (single global schedule only)."""
  schedule = config.get("schedule", {})
  if isinstance(schedule, (tuple, list)):
    if len(schedule) != 1 or schedule[0][0] != ".*":
      raise NotImplementedError("per-pattern schedules / frozen parameters")
    schedule = schedule[0][1]
  if schedule is None:
    raise NotImplementedError("schedule=None (all parameters frozen)")
  schedule = dict(schedule)
  mult = schedule.pop("mult", 1.0)
  sched_fn = u.create_learning_rate_schedule(base=mult, **sched_kw, **schedule)
  wd_mults = config.get("wd_mults", [(".*/kernel$", 1.0)])
  if list(map(tuple, wd_mults)) != [(".*/kernel$", 1.0)]:
    raise NotImplementedError("custom wd_mults")
  return FusedAdam(config, sched_fn), [sched_fn]
