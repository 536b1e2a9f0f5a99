"""Host-side utilities mirrored from big_vision/utils.py that the update step needs:
`steps` (utils.py:1002-1067, reduced to the keys used here) and
`create_learning_rate_schedule` (utils.py:1070-1143).  Pure Python/NumPy (runs on the host
once per step, like the reference's `sched_fns_cpu`)."""
import math

import numpy as np


def steps(prefix, config, data_size=None, batch_size=None, total_steps=None, default=ValueError):
  """Gets duration named `prefix` out of `config` and converts it to steps (utils.py:1002-1067):
  `{prefix}_{steps,examples,epochs,percent}`, negative entries ignored, rounded to nearest
  with a floor of one step unless zero was asked for."""
  suffixes = ("steps", "examples", "epochs", "percent")
  matches = set()
  for s in suffixes:
    x = config.get(f"{prefix}_{s}")
    if x is not None and x >= 0:
      matches.add(f"{prefix}_{s}")
  assert len(matches) <= 1, f"Only one of '{matches}' should be defined."

  if f"{prefix}_steps" in matches:
    return config[f"{prefix}_steps"]

  def to_integer(x):
    return max(1, round(x)) if x else 0

  if batch_size and f"{prefix}_examples" in matches:
    return to_integer(config[f"{prefix}_examples"] / batch_size)
  if batch_size and data_size and f"{prefix}_epochs" in matches:
    return to_integer(config[f"{prefix}_epochs"] * (data_size / batch_size))
  if total_steps and f"{prefix}_percent" in matches:
    pct = config[f"{prefix}_percent"]
    assert 0.0 <= pct <= 1.0, f"Percents should lie in [0.0, 1.0], but {prefix}_percent is {pct}"
    return to_integer(pct * total_steps)
  if default is ValueError:
    raise ValueError(f"Cannot convert {prefix} to steps, due to missing batch_size ({batch_size}), "
                     f"data_size ({data_size}), total_steps ({total_steps}), or config entry")
  return default


def create_learning_rate_schedule(total_steps, batch_size=None, data_size=None, base=1.0,
                                  decay_type="stair", scale_with_batchsize=False, **kw):
  """Same semantics as utils.py:1070-1143; returns step -> float."""

  def to_steps(name, default=0):
    return steps(name, kw, data_size, batch_size, total_steps, default=default)

  warmup_steps = to_steps("warmup")
  cooldown_steps = to_steps("cooldown")
  assert (total_steps <= 1) or (warmup_steps < total_steps), "warmup_steps is >= total_steps"

  def step_fn(step):
    lr = base
    if scale_with_batchsize:
      lr = lr * batch_size / 256.0
    progress = (step - warmup_steps) / float(total_steps - warmup_steps)
    progress = float(np.clip(progress, 0.0, 1.0))
    if decay_type in ("linear", "polynomial"):
      power = kw.get("power", 1)
      zero = kw.get("end", kw.get("linear_end", 0))
      lr = zero + (lr - zero) * (1.0 - progress) ** power
    elif decay_type == "cosine":
      lr = lr * 0.5 * (1.0 + math.cos(math.pi * progress))
    elif decay_type == "rsqrt":
      t = to_steps("timescale", default=kw.get("timescale", 10_000))
      shift = to_steps("shift", default=kw.get("shift", 0))
      if warmup_steps <= step:
        lr = lr / math.sqrt(1 + (step + shift - warmup_steps) / t)
      else:
        lr = lr / math.sqrt(1 + shift / t)
    elif decay_type == "stair":
      i = int(np.searchsorted(np.array(kw.get("steps", [])), step + 1))
      lr = lr * ([1.0] + list(kw.get("mults", [])))[i]
    else:
      raise ValueError(f"Unknown lr type {decay_type}")
    if warmup_steps:
      lr = lr * min(1.0, step / warmup_steps)
    if cooldown_steps:
      lr = lr * min(1.0, (total_steps - step) / cooldown_steps)
    return float(np.float32(lr))

  return step_fn
