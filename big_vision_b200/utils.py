"""Host-side helpers of the update step and of checkpoint interchange, written against the
CONTRACT of big_vision/utils.py (behaviour pinned by the reference's own known-answer tests,
utils_test.py:228-281 for durations / schedules, and by the on-disk .npz format), not against its text:

  steps(prefix, config, ...)               duration keys -> optimizer steps      (utils.py:1002-1067)
  create_learning_rate_schedule(...)       step -> multiplier                    (utils.py:1070-1143)
  tree_flatten_with_names / tree_get / recover_tree / load_params / save_checkpoint_np
                                           flat "a/b/c" names <-> nested trees   (utils.py:133-228,616-752,827-862)

Pure Python / NumPy; runs on the host (the reference evaluates its schedules on the host too,
`sched_fns_cpu`).
"""
import math
import os

import numpy as np

# ----------------------------------------------------------------------------------------------
# durations and schedules
# ----------------------------------------------------------------------------------------------
_DURATION_UNITS = ("steps", "examples", "epochs", "percent")


def _whole_steps(x):
  """Nearest whole number of steps, but never round a non-zero duration down to nothing."""
  return max(1, round(x)) if x else 0


def steps(prefix, config, data_size=None, batch_size=None, total_steps=None, default=ValueError):
  """Duration `prefix` of `config` in optimizer steps.

  The duration may be given in one of four units, as `{prefix}_steps`, `{prefix}_examples`,
  `{prefix}_epochs` or `{prefix}_percent` (of `total_steps`); entries that are None or negative count
  as absent, more than one present is an error.  Conversions need `batch_size` (examples),
  `batch_size` and `data_size` (epochs) or `total_steps` (percent); when the unit present cannot be
  converted, or none is present, `default` is returned (raised, if it is ValueError)."""
  present = {}
  for unit in _DURATION_UNITS:
    amount = config.get(f"{prefix}_{unit}")
    if amount is not None and amount >= 0:
      present[unit] = amount
  if len(present) > 1:
    raise AssertionError(f"Only one of '{ {f'{prefix}_{u}' for u in present} }' should be defined.")
  if "steps" in present:
    return present["steps"]
  if "examples" in present and batch_size:
    return _whole_steps(present["examples"] / batch_size)
  if "epochs" in present and batch_size and data_size:
    return _whole_steps(present["epochs"] * (data_size / batch_size))
  if "percent" in present and total_steps:
    pct = present["percent"]
    if not 0.0 <= pct <= 1.0:
      raise AssertionError(f"Percents should lie in [0.0, 1.0], but {prefix}_percent is {pct}")
    return _whole_steps(pct * total_steps)
  if default is ValueError:
    raise ValueError(f"Cannot convert {prefix} to steps, due to missing batch_size ({batch_size}), "
                     f"data_size ({data_size}), total_steps ({total_steps}), or config entry")
  return default


class _Schedule:
  """step -> learning-rate multiplier: base value (optionally scaled with batch_size / 256), one of
  the decay shapes below over the post-warm-up progress in [0, 1], then linear warm-up from zero and
  linear cool-down to zero.  Evaluates to a Python float holding a float32 value."""

  def __init__(self, total_steps, batch_size, data_size, base, decay_type, scale_with_batchsize, kw):
    self.total, self.kw, self.decay_type = total_steps, kw, decay_type
    self.peak = base * batch_size / 256.0 if scale_with_batchsize else base
    dur = lambda name, default=0: steps(name, kw, data_size, batch_size, total_steps, default=default)
    self.warmup, self.cooldown = dur("warmup"), dur("cooldown")
    if total_steps > 1 and not self.warmup < total_steps:
      raise AssertionError("warmup_steps is >= total_steps")
    if decay_type == "rsqrt":
      self.timescale = dur("timescale", kw.get("timescale", 10_000))
      self.shift = dur("shift", kw.get("shift", 0))
    if decay_type not in self._SHAPES:
      raise ValueError(f"Unknown lr type {decay_type}")

  # each shape: (self, step, progress) -> value before warm-up / cool-down
  def _poly(self, step, progress):
    floor = self.kw.get("end", self.kw.get("linear_end", 0))
    return floor + (self.peak - floor) * (1.0 - progress) ** self.kw.get("power", 1)

  def _cosine(self, step, progress):
    return self.peak * 0.5 * (1.0 + math.cos(math.pi * progress))

  def _rsqrt(self, step, progress):
    past = max(step - self.warmup, 0)          # constant until the warm-up is over
    return self.peak / math.sqrt(1 + (past + self.shift) / self.timescale)

  def _stair(self, step, progress):
    boundaries = self.kw.get("steps", [])
    passed = sum(1 for b in boundaries if b <= step)   # boundaries reached so far (sorted input)
    return self.peak * ([1.0] + list(self.kw.get("mults", [])))[passed]

  _SHAPES = {"linear": _poly, "polynomial": _poly, "cosine": _cosine, "rsqrt": _rsqrt, "stair": _stair}

  def __call__(self, step):
    span = float(self.total - self.warmup)
    progress = min(max((step - self.warmup) / span, 0.0), 1.0)
    value = self._SHAPES[self.decay_type](self, step, progress)
    if self.warmup:
      value *= min(1.0, step / self.warmup)
    if self.cooldown:
      value *= min(1.0, (self.total - step) / self.cooldown)
    return float(np.float32(value))


def create_learning_rate_schedule(total_steps, batch_size=None, data_size=None, base=1.0,
                                  decay_type="stair", scale_with_batchsize=False, **kw):
  """Schedule factory with the reference's signature; durations in `kw` (`warmup_*`, `cooldown_*`,
  `timescale_*`, `shift_*`) go through `steps`.  Returns a callable step -> float."""
  return _Schedule(total_steps, batch_size, data_size, base, decay_type, scale_with_batchsize, kw)


# ----------------------------------------------------------------------------------------------
# Checkpoint interchange: the reference's .npz format.  A checkpoint is a tree (nested dicts; tuples
# and lists are addressed by position) whose leaves are arrays; on disk every leaf is one .npz entry
# named by its path joined with "/", and the entries appear in sorted-key order.
# ----------------------------------------------------------------------------------------------
def _children(node):
  """(key as str, child) pairs of an inner node in canonical order, or None for a leaf."""
  if isinstance(node, dict):
    return [(k, node[k]) for k in sorted(node)]
  if isinstance(node, (list, tuple)):
    return [(str(i), c) for i, c in enumerate(node)]
  return None


def _walk(tree, inner=False):
  """Depth-first (path, node) pairs; leaves always, inner nodes too when `inner` (after their
  children, the root under the empty path).  None sub-trees are empty."""
  stack = [("", tree, False)]
  while stack:
    path, node, expanded = stack.pop()
    if node is None:
      continue
    kids = _children(node)
    if kids is None:
      yield path, node
    elif expanded:
      yield path, node
    else:
      if inner:
        stack.append((path, node, True))
      for key, child in reversed(kids):
        stack.append((f"{path}/{key}" if path else key, child, False))


def tree_flatten_with_names(tree):
  """([(name, leaf), ...], None): the leaves in canonical order with their "a/b/c" names (for trees
  of dicts / tuples / lists this is also jax's flattening order, so positions line up with
  jax.tree.leaves on the reference side)."""
  return list(_walk(tree)), None


def tree_get(tree, name):
  """The leaf -- or whole sub-tree -- stored under the flat name `name`."""
  node = tree
  for part in (name.split("/") if name else []):
    kids = _children(node)
    nxt = dict(kids).get(part, None) if kids is not None else None
    if nxt is None and not (kids is not None and part in dict(kids)):
      known = [p for p, _ in _walk(tree, inner=True)]
      raise KeyError("\n".join([name, "Available keys:", *known, ""]))
    node = nxt
  return node


def recover_tree(keys, values):
  """Inverse of the flattening for dict trees: {"a/b": 1, "a/c": 2, "d": 3} -> {"a": {"b": 1, "c": 2},
  "d": 3}.  Key order of the result follows first appearance."""
  root = {}
  for name, value in zip(keys, values):
    *parents, leaf = name.split("/")
    node = root
    for part in parents:
      node = node.setdefault(part, {})
    node[leaf] = value
  return root


def recover_dtype(a):
  """numpy has no bfloat16: a bf16 array round-trips through .npz as 2-byte void records.  Such
  leaves are widened to float32 here (exact -- bf16 is the top half of an fp32 word)."""
  if getattr(a, "dtype", None) is not None and a.dtype.kind == "V":
    if a.itemsize != 2:
      raise AssertionError("Unknown dtype!")
    return (a.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
  return a


def npload(fname):
  """An array (.npy) or a dict of arrays (.npz)."""
  data = np.load(fname, allow_pickle=False)
  return data if isinstance(data, np.ndarray) else {k: data[k] for k in data.files}


def load_checkpoint_np(npz):
  """Tree of a .npz checkpoint given by path or as an already-loaded mapping of flat names."""
  flat = npload(npz) if isinstance(npz, str) else npz
  names = list(flat.keys())
  return recover_tree(names, [flat[k] for k in names])


def _tree_map(fn, tree):
  kids = _children(tree)
  if kids is None:
    return fn(tree)
  if isinstance(tree, dict):
    return {k: _tree_map(fn, tree[k]) for k in tree}
  return type(tree)(_tree_map(fn, c) for c in tree)


def load_params(ckpt):
  """Model parameters out of a checkpoint.  `ckpt`: an already-loaded mapping, or a path
  "dir/file.npz" optionally followed by ":sub/tree" to select part of the parameters.  The parameters
  sit under "params" (train-state checkpoints), under "opt/target" (older optimizer-state
  checkpoints) or are the whole file (bare parameter dumps)."""
  subtree = None
  if isinstance(ckpt, str):
    path, sep, rest = ckpt.rpartition(":")
    if sep and "/" in path and all(ch.isalnum() or ch in "_/" for ch in rest) and rest:
      ckpt, subtree = path, rest
    if "/" not in ckpt:
      raise ValueError(f"Weird ckpt path: {ckpt} ; Maybe prepend ./ ?")
    if ".npz" not in ckpt:
      raise ValueError("only the .npz checkpoint format is supported here")
  tree = _tree_map(recover_dtype, load_checkpoint_np(ckpt))
  if "params" in tree:
    tree = tree["params"]
  elif "opt" in tree:
    tree = tree["opt"]["target"]
  return tree if subtree is None else tree_get(tree, subtree)


def save_checkpoint_np(checkpoint, path):
  """Writes a tree in the reference's .npz layout (one entry per leaf, flat names), via a temporary
  file and a rename so that a reader never sees a half-written checkpoint."""
  entries = {name: np.asarray(leaf) for name, leaf in _walk(checkpoint)}
  tmp = path + "-TEMPORARY.npz"
  with open(tmp, "wb") as f:
    np.savez(f, **entries)
  os.replace(tmp, path)


def check_and_compile_patterns(patterns):
  """One regex string or a sequence of them -> list of compiled patterns."""
  import re
  if isinstance(patterns, str):
    patterns = [patterns]
  if not isinstance(patterns, (list, tuple)):
    raise AssertionError(f"Must be a sequence of regex strings: {patterns!r}")
  return [re.compile(p) for p in patterns]


# ----------------------------------------------------------------------------------------------
# mixup (utils.py:1146-1158)
# ----------------------------------------------------------------------------------------------
def get_mixup(rng, p):
  """Mirror of `get_mixup(rng, p)`: draws a ~ Beta(p, p), a = max(a, 1 - a), and returns
  `_mixup(*things, **more_things) -> (rng, things, more_things)` mixing every thing with its roll
  by one along the batch axis (`bv_mixup`).  `rng` is a numpy Generator (the reference's is a jax
  PRNG key; the stream of random numbers is necessarily a different one, the arithmetic is not)."""
  a = float(rng.beta(p, p))
  a = max(a, 1.0 - a)

  def _mixup(*things, **more_things):
    from big_vision_b200 import ops
    mix = lambda thing: ops.mixup(thing, a)
    return rng, tuple(mix(t) for t in things), {k: mix(v) for k, v in more_things.items()}

  _mixup.a = a
  return _mixup


def mixup(rng, *things, p, **more_things):
  return get_mixup(rng, p)(*things, **more_things)
