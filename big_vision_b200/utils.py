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


# ----------------------------------------------------------------------------------------------
# Checkpoint interchange: the reference's .npz format (flat "a/b/c" keys), SURVEY 8f rank 3.
# Trees are nested dicts (tuples / lists by index) of numpy arrays; everything here is host code.
# ----------------------------------------------------------------------------------------------
import collections as _collections
import os as _os
import re as _re


def _traverse_with_names(tree, with_inner_nodes=False):
  """(name, leaf) pairs in sorted-key order -- utils.py:616-640."""
  if tree is None:
    return
  if isinstance(tree, dict):
    for key in sorted(tree.keys()):
      for path, v in _traverse_with_names(tree[key], with_inner_nodes):
        yield (key + "/" + path).rstrip("/"), v
    if with_inner_nodes:
      yield "", tree
  elif isinstance(tree, (list, tuple)):
    for idx in range(len(tree)):
      for path, v in _traverse_with_names(tree[idx], with_inner_nodes):
        yield (str(idx) + "/" + path).rstrip("/"), v
    if with_inner_nodes:
      yield "", tree
  else:
    yield "", tree


def tree_flatten_with_names(tree):
  """[(name, value), ...] -- utils.py:642-670 for dict / tuple / list trees (for those, jax's
  flattening order is the sorted-key order produced here).  Returns (names_and_vals, None)."""
  return list(_traverse_with_names(tree)), None


def tree_get(tree, name):
  """Entry of a tree by flattened key, e.g. 'a/b/c' or an inner node 'a/b' -- utils.py:726-752."""
  flattened = dict(_traverse_with_names(tree, with_inner_nodes=True))
  try:
    return flattened[name]
  except KeyError:
    raise KeyError("\n".join([name, "Available keys:", *flattened, ""])) from None


def recover_tree(keys, values):
  """Nested dict from flat '/'-separated names -- utils.py:836-862."""
  tree = {}
  sub_trees = _collections.defaultdict(list)
  for k, v in zip(keys, values):
    if "/" not in k:
      tree[k] = v
    else:
      k_left, k_right = k.split("/", 1)
      sub_trees[k_left].append((k_right, v))
  for k, kv_pairs in sub_trees.items():
    k_subtree, v_subtree = zip(*kv_pairs)
    tree[k] = recover_tree(k_subtree, v_subtree)
  return tree


def recover_dtype(a):
  """numpy stores bfloat16 as a 2-byte void type (utils.py:827-833); there is no numpy bfloat16
  here, so such arrays come back as float32 (exact: bf16 is the top half of fp32)."""
  if hasattr(a, "dtype") and a.dtype.type is np.void:
    assert a.itemsize == 2, "Unknown dtype!"
    return (a.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
  return a


def npload(fname):
  """np.ndarray (np.save file) or dict of arrays (np.savez file) -- utils.py:133-149."""
  loaded = np.load(fname, allow_pickle=False)
  return loaded if isinstance(loaded, np.ndarray) else dict(loaded)


def load_checkpoint_np(npz):
  """Nested tree from a .npz path or dict-like of flat names -- utils.py:152-170."""
  if isinstance(npz, str):
    npz = npload(npz)
  keys, values = zip(*list(npz.items()))
  return recover_tree(keys, values)


def _tree_map(fn, tree):
  if isinstance(tree, dict):
    return {k: _tree_map(fn, v) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return type(tree)(_tree_map(fn, v) for v in tree)
  return fn(tree)


def load_params(ckpt):
  """Parameters of a big_vision .npz checkpoint -- utils.py:173-228 (npz branch).  `ckpt` may be
  '/path/file.npz:img/head' to take a sub-tree, or an already-loaded dict-like.  Handles the three
  containers the reference does: {'params': ...}, {'opt': {'target': ...}}, or the bare tree."""
  key = None
  if isinstance(ckpt, str):
    match = _re.match(r"^(.*?/.*?)(?::([\w/]+))?$", ckpt)
    if not match:
      raise ValueError(f"Weird ckpt path: {ckpt} ; Maybe prepend ./ ?")
    ckpt, key = match.groups()
    if ".npz" not in ckpt:
      raise ValueError("only the .npz checkpoint format is supported here")
  checkpoint = _tree_map(recover_dtype, load_checkpoint_np(ckpt))
  if "params" in checkpoint:
    params = checkpoint["params"]
  elif "opt" in checkpoint:
    params = checkpoint["opt"]["target"]
  else:
    params = checkpoint
  if key is not None:
    params = tree_get(params, key)
  return params


def save_checkpoint_np(checkpoint, path):
  """Writes a tree as the reference's .npz: one array per leaf under its flat name, atomically
  (temporary file + rename), like the reference's npz writer."""
  names_and_vals, _ = tree_flatten_with_names(checkpoint)
  tmp = path + "-TEMPORARY.npz"
  with open(tmp, "wb") as f:
    np.savez(f, **{k: np.asarray(v) for k, v in names_and_vals})
  _os.replace(tmp, path)


def check_and_compile_patterns(patterns):
  """utils.py: a str or a sequence of regex strs -> compiled patterns."""
  if isinstance(patterns, str):
    patterns = [patterns]
  assert isinstance(patterns, (list, tuple)), patterns
  return [_re.compile(p) for p in patterns]


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
