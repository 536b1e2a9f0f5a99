"""Flat parameter storage shared by the models, the backward pass and the optimizer.

All parameters of a model live in ONE fp32 buffer (`flat`), their gradients in a second
buffer of the same layout (`grad`) and a bf16 shadow copy (`half`) that the tensor-core
GEMMs read.  One flat layout means: one fused optimizer launch per parameter group, one
bucketed NCCL all-reduce over `grad`, one zero-fill per step.  Parameters are addressed
by the reference's tree names (`img/Transformer/encoderblock_0/...`);
fused storage (q|k|v in one [d, 3d] matrix) is exposed through strided views so the
reference names and shapes still resolve.
"""
import re
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch


@dataclass
class ParamSpec:
  name: str                      # storage name ("a/b/c")
  shape: Tuple[int, ...]         # storage shape
  init: Callable                 # init(rng: np.random.Generator, shape) -> np.ndarray (fp32)


@dataclass
class Alias:
  """A reference-named view into a stored parameter."""
  name: str
  storage: str
  view: Callable                 # torch storage tensor -> torch view with the reference shape


ALIGN = 8  # elements: 32 B in fp32, 16 B in bf16 (TMA base alignment)


class FlatParams:
  """fp32 master params + grads + bf16 shadow in three flat device buffers."""

  def __init__(self, specs: List[ParamSpec], aliases: List[Alias], device, decay_regex=r".*/kernel$"):
    self.specs = specs
    self.aliases = {a.name: a for a in aliases}
    self.device = torch.device(device)
    # decayed parameters first, so weight decay is one contiguous launch range
    # (optax.py:133 default mask `.*/kernel$`).
    rx = re.compile(decay_regex) if decay_regex else None
    alias_by_storage: Dict[str, List[str]] = {}
    for a in aliases:
      alias_by_storage.setdefault(a.storage, []).append(a.name)

    def decayed(s):
      if rx is None:
        return False
      names = alias_by_storage.get(s.name, [s.name])
      return any(rx.match(n) for n in names)

    order = [s for s in specs if decayed(s)] + [s for s in specs if not decayed(s)]
    self.offsets: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
    off = 0
    for s in order:
      self.offsets[s.name] = (off, tuple(s.shape))
      n = int(np.prod(s.shape))
      off += (n + ALIGN - 1) // ALIGN * ALIGN
      if decayed(s):
        self.n_decay = off
    if not hasattr(self, "n_decay"):
      self.n_decay = 0
    self.total = off
    self.flat = torch.zeros(self.total, dtype=torch.float32, device=self.device)
    self.grad = torch.zeros(self.total, dtype=torch.float32, device=self.device)
    self.half = torch.zeros(self.total, dtype=torch.bfloat16, device=self.device)
    self._views = {}

  # ---- raw views ------------------------------------------------------------------
  def _view(self, buf, name):
    off, shape = self.offsets[name]
    n = int(np.prod(shape))
    return buf[off:off + n].view(shape)

  def f(self, name):
    """fp32 master view of storage parameter `name`."""
    key = ("f", name)
    if key not in self._views:
      self._views[key] = self._view(self.flat, name)
    return self._views[key]

  def g(self, name):
    key = ("g", name)
    if key not in self._views:
      self._views[key] = self._view(self.grad, name)
    return self._views[key]

  def h(self, name):
    key = ("h", name)
    if key not in self._views:
      self._views[key] = self._view(self.half, name)
    return self._views[key]

  # ---- init / interchange ---------------------------------------------------------
  def init(self, seed=0):
    rng = np.random.default_rng(seed)
    host = np.zeros(self.total, dtype=np.float32)
    for s in self.specs:
      off, shape = self.offsets[s.name]
      n = int(np.prod(shape))
      host[off:off + n] = np.asarray(s.init(rng, tuple(shape)), dtype=np.float32).reshape(-1)
    self.flat.copy_(torch.from_numpy(host))
    self.sync_half()
    return self

  def sync_half(self):
    """Refreshes the bf16 shadow from the fp32 master (the optimizer kernel does this itself)."""
    if self.flat.is_cuda:
      from big_vision_b200 import ops
      ops.cast(self.flat, self.half)
    else:
      self.half.copy_(self.flat)

  def tree(self, which="f"):
    """dict: reference name -> tensor view (params 'f', grads 'g')."""
    out = {}
    aliased = {a.storage for a in self.aliases.values()}
    get = {"f": self.f, "g": self.g, "h": self.h}[which]
    for s in self.specs:
      if s.name not in aliased:
        out[s.name] = get(s.name)
    for a in self.aliases.values():
      out[a.name] = a.view(get(a.storage))
    return out

  def load_tree(self, tree: Dict[str, "np.ndarray"]):
    """Copies a reference-named tree (numpy / torch, reference shapes) into the flat buffer."""
    views = self.tree("f")
    missing = [k for k in views if k not in tree]
    if missing:
      raise KeyError(f"missing parameters: {missing[:5]}{'...' if len(missing) > 5 else ''}")
    for k, v in views.items():
      src = torch.as_tensor(np.asarray(tree[k], dtype=np.float32)).to(self.device)
      v.copy_(src.reshape(v.shape))
    self.sync_half()
    return self

  def numpy_tree(self, which="f"):
    return {k: v.detach().float().cpu().numpy().copy() for k, v in self.tree(which).items()}

  def zero_grad(self):
    self.grad.zero_()


# ---- initialisers (numpy; same distributions as the reference's) -------
def xavier_uniform(fan_in, fan_out):
  def init(rng, shape):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape)
  return init


def lecun_normal(fan_in):
  def init(rng, shape):
    # flax lecun_normal = variance_scaling(1.0, "fan_in", "truncated_normal")
    std = np.sqrt(1.0 / fan_in) / 0.87962566103423978
    x = rng.standard_normal(size=shape)
    bad = np.abs(x) > 2
    while bad.any():
      x[bad] = rng.standard_normal(size=int(bad.sum()))
      bad = np.abs(x) > 2
    return x * std
  return init


def normal(std):
  return lambda rng, shape: rng.standard_normal(size=shape) * std


def zeros(rng, shape):
  return np.zeros(shape, dtype=np.float32)


def ones(rng, shape):
  return np.ones(shape, dtype=np.float32)


def constant(v):
  return lambda rng, shape: np.full(shape, v, dtype=np.float32)
