"""ViT on the B200 kernels -- host-side mirror of big_vision/models/vit.py.

Same factory (`Model(num_classes, variant=..., **kw)`), same fields, same parameter-tree
names/shapes as the reference (models/vit.py:186-281, param names SURVEY.md 8b); the
computation is an explicit forward + hand-written backward over the C-ABI kernels
(bv_gemm / bv_attention_* / bv_layernorm_* ...) instead of flax modules under jax.grad.

dtype flow (reference with dtype_mm="bfloat16"): residual stream bf16,
LayerNorm statistics fp32, every matmul bf16 x bf16 -> fp32 accumulate, parameters and
their gradients fp32.  Unlike the reference, the MAP head / heads also run their matmuls
in bf16 (fp32 accumulate, fp32 outputs); see DESIGN.md "numerics".
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from big_vision_b200 import engine as E
from big_vision_b200 import lib as L
from big_vision_b200 import ops


def posemb_sincos_2d(h, w, width, temperature=10_000.0):
  """Fixed 2-D sine/cosine position table [h*w, width] in the MoCo-v3 channel layout the reference
  uses (models/vit.py:34-44): the width is cut into four equal bands holding sin(x w_k), cos(x w_k),
  sin(y w_k), cos(y w_k) for the token at column x, row y (row-major token order), with frequencies
  w_k = temperature^(-k/(width/4 - 1))."""
  if width % 4:
    raise AssertionError("Width must be mult of 4 for sincos posemb")
  bands = width // 4
  freq = 1.0 / temperature ** (np.arange(bands) / (bands - 1))
  col = np.tile(np.arange(w), h)          # x of token t = t % w
  row = np.repeat(np.arange(h), w)        # y of token t = t // w
  ax, ay = np.outer(col, freq), np.outer(row, freq)
  return np.concatenate([np.sin(ax), np.cos(ax), np.sin(ay), np.cos(ay)], axis=1).astype(np.float32)


# name: (width, depth, mlp_dim, num_heads) -- the size table of models/vit.py:284-303
_VARIANTS = {
    "mu": (32, 1, 128, 2), "Ti": (192, 12, 768, 3), "S": (384, 12, 1536, 6), "M": (512, 12, 2048, 8),
    "B": (768, 12, 3072, 12), "L": (1024, 24, 4096, 16), "So400m": (1152, 27, 4304, 16),
    "H": (1280, 32, 5120, 16), "g": (1408, 40, 6144, 16), "g-opt": (1536, 40, 6144, 16),
    "G": (1664, 48, 8192, 16), "G-opt": (1536, 48, 8192, 16), "e": (1792, 56, 15360, 16),
}


def decode_variant(variant):
  """"B" / "B/16" -> dict(width, depth, mlp_dim, num_heads[, patch_size]); None -> {}."""
  if variant is None:
    return {}
  name, _, patch = variant.partition("/")
  width, depth, mlp_dim, num_heads = _VARIANTS[name]
  out = dict(width=width, depth=depth, mlp_dim=mlp_dim, num_heads=num_heads)
  if patch:
    out["patch_size"] = (int(patch), int(patch))
  return out


# ------------------------------------------------------------------------------------------
# building blocks (each: specs(), fwd(), bwd()); `P` is an engine.FlatParams
# ------------------------------------------------------------------------------------------
def _stacked(init, stack):
  """Initialiser of a scan-stacked parameter: `stack` independent draws along a new leading axis."""
  if not stack:
    return init
  return lambda rng, shape: np.stack([np.asarray(init(rng, tuple(shape[1:]))) for _ in range(shape[0])])


def _shape(shape, stack):
  return ((stack,) + tuple(shape)) if stack else tuple(shape)


def ln_specs(p, d, stack=0):
  return [E.ParamSpec(p + "scale", _shape((d,), stack), _stacked(E.ones, stack)),
          E.ParamSpec(p + "bias", _shape((d,), stack), _stacked(E.zeros, stack))]


def mlp_specs(p, d, m, stack=0):
  """MlpBlock (models/vit.py:57-78): xavier_uniform kernels, normal(1e-6) biases."""
  return [
      E.ParamSpec(p + "Dense_0/kernel", _shape((d, m), stack), _stacked(E.xavier_uniform(d, m), stack)),
      E.ParamSpec(p + "Dense_0/bias", _shape((m,), stack), _stacked(E.normal(1e-6), stack)),
      E.ParamSpec(p + "Dense_1/kernel", _shape((m, d), stack), _stacked(E.xavier_uniform(m, d), stack)),
      E.ParamSpec(p + "Dense_1/bias", _shape((d,), stack), _stacked(E.normal(1e-6), stack)),
  ]


class Scope:
  """A parameter sub-tree of a FlatParams: `S.f("Dense_0/kernel")` is the fp32 master of
  `prefix + "Dense_0/kernel"` (g: gradient, h: bf16 shadow).  With `index` the stored tensors carry
  a leading stack axis (the reference's scan=True layout, models/vit.py:129-148: one `encoderblock`
  sub-tree whose leaves have a leading `depth` axis) and the scope addresses slice `index` of it."""

  def __init__(self, P, prefix, index=None):
    self.P, self.prefix, self.index = P, prefix, index

  def _get(self, kind, name):
    full = self.prefix + name
    if self.index is None:
      return getattr(self.P, kind)(full)
    key = (kind, full, self.index)
    v = self.P._views.get(key)   # pylint: disable=protected-access
    if v is None:
      v = getattr(self.P, kind)(full)[self.index]
      self.P._views[key] = v     # pylint: disable=protected-access
    return v

  def f(self, name):
    return self._get("f", name)

  def g(self, name):
    return self._get("g", name)

  def h(self, name):
    return self._get("h", name)

  def sub(self, rel):
    return Scope(self.P, self.prefix + rel, self.index)


def mlp_fwd(S, y, resid, out_dtype=torch.bfloat16):
  """resid + Dense_1(gelu(Dense_0(y))) with S the MlpBlock's Scope.  Returns (out, saved)."""
  act, pre = ops.gemm(y, S.h("Dense_0/kernel"), b_mn=True, bias=S.f("Dense_0/bias"),
                      epilogue=L.EPI_BIAS_GELU)
  out = ops.gemm(act, S.h("Dense_1/kernel"), b_mn=True, bias=S.f("Dense_1/bias"),
                 aux=resid, epilogue=L.EPI_BIAS_RESID if resid is not None else L.EPI_BIAS,
                 out_dtype=out_dtype)
  return out, (y, act, pre)


def mlp_bwd(S, dout, saved, want_bias2_grad=True):
  """dout: bf16 [M,d] gradient of the block output.  Returns d(y) (bf16).

  The bias gradient of Dense_1 is colsum(dout); callers that already have that column sum
  from the LayerNorm-backward kernel pass want_bias2_grad=False."""
  y, act, pre = saved
  if want_bias2_grad:
    ops.colsum(dout, S.g("Dense_1/bias"))
  ops.gemm(act, dout, a_mn=True, b_mn=True, out=S.g("Dense_1/kernel"), reduce_out=True)
  # the Dense_0 bias gradient (column sums of dpre) is accumulated by the same GEMM's epilogue
  dpre = ops.gemm(dout, S.h("Dense_1/kernel"), aux=pre, epilogue=L.EPI_DGELU,
                  colsum=S.g("Dense_0/bias"))
  ops.gemm(y, dpre, a_mn=True, b_mn=True, out=S.g("Dense_0/kernel"), reduce_out=True)
  return ops.gemm(dpre, S.h("Dense_0/kernel"))


def mha_specs(p, d, heads, fuse_qkv=True, stack=0):
  """flax MultiHeadDotProductAttention params: query/key/value kernels [d,h,dh] + bias [h,dh],
  out kernel [h,dh,d] + bias [d]; kernel_init xavier_uniform (models/vit.py:95,177), zero biases.
  Stored fused ([d,3d] or q:[d,d] + kv:[d,2d]) and aliased to the reference names.  `stack` > 0
  adds the leading scan axis to every stored tensor and every alias."""
  dh = d // heads
  xav = E.xavier_uniform(d, d)
  ax = 1 if stack else 0          # axis of the `d` input features in the stored kernels

  def fused_init(k):
    return _stacked(lambda rng, shape: np.concatenate([xav(rng, (d, d)) for _ in range(k)], axis=1), stack)

  specs, aliases = [], []

  def alias_cols(store, names):
    for i, nm in enumerate(names):
      aliases.append(E.Alias(p + nm + "/kernel", p + store + "/kernel",
                             lambda t, i=i: t.narrow(ax + 1, i * d, d).unflatten(ax + 1, (heads, dh))))
      aliases.append(E.Alias(p + nm + "/bias", p + store + "/bias",
                             lambda t, i=i: t.narrow(ax, i * d, d).unflatten(ax, (heads, dh))))

  if fuse_qkv:
    specs += [E.ParamSpec(p + "qkv/kernel", _shape((d, 3 * d), stack), fused_init(3)),
              E.ParamSpec(p + "qkv/bias", _shape((3 * d,), stack), _stacked(E.zeros, stack))]
    alias_cols("qkv", ["query", "key", "value"])
  else:
    specs += [E.ParamSpec(p + "q/kernel", _shape((d, d), stack), _stacked(xav, stack)),
              E.ParamSpec(p + "q/bias", _shape((d,), stack), _stacked(E.zeros, stack)),
              E.ParamSpec(p + "kv/kernel", _shape((d, 2 * d), stack), fused_init(2)),
              E.ParamSpec(p + "kv/bias", _shape((2 * d,), stack), _stacked(E.zeros, stack))]
    alias_cols("q", ["query"])
    alias_cols("kv", ["key", "value"])
  specs += [E.ParamSpec(p + "out_proj/kernel", _shape((d, d), stack), _stacked(xav, stack)),
            E.ParamSpec(p + "out/bias", _shape((d,), stack), _stacked(E.zeros, stack))]
  aliases.append(E.Alias(p + "out/kernel", p + "out_proj/kernel",
                         lambda t: t.unflatten(ax, (heads, dh))))
  return specs, aliases


class EncoderBlock:
  """Encoder1DBlock (models/vit.py:81-112): x + MHSA(LN(x)); x + MLP(LN(x)).
  `index` = position in the scan-stacked `encoderblock` sub-tree (None: own `encoderblock_{i}`)."""

  def __init__(self, prefix, d, m, heads, index=None):
    self.p, self.d, self.m, self.heads, self.index = prefix, d, m, heads, index

  def specs(self, stack=0):
    att = self.p + "MultiHeadDotProductAttention_0/"
    s, a = mha_specs(att, self.d, self.heads, stack=stack)
    return (ln_specs(self.p + "LayerNorm_0/", self.d, stack) + s + ln_specs(self.p + "LayerNorm_1/", self.d, stack)
            + mlp_specs(self.p + "MlpBlock_0/", self.d, self.m, stack)), a

  def scope(self, P):
    return Scope(P, self.p, self.index)

  def fwd(self, P, x, n, N):
    d = self.d
    S = self.scope(P)
    A = S.sub("MultiHeadDotProductAttention_0/")
    ln1, mean1, rstd1 = ops.layernorm_fwd(x, S.f("LayerNorm_0/scale"), S.f("LayerNorm_0/bias"))
    qkv = ops.gemm(ln1, A.h("qkv/kernel"), b_mn=True, bias=A.f("qkv/bias"))
    qkv3 = qkv.view(n, N, 3 * d)
    o, lse = ops.attention_fwd(qkv3[:, :, 0:d], qkv3[:, :, d:2 * d], qkv3[:, :, 2 * d:], self.heads)
    x1 = ops.gemm(o.view(n * N, d), A.h("out_proj/kernel"), b_mn=True,
                  bias=A.f("out/bias"), aux=x, epilogue=L.EPI_BIAS_RESID)
    ln2, mean2, rstd2 = ops.layernorm_fwd(x1, S.f("LayerNorm_1/scale"), S.f("LayerNorm_1/bias"))
    x2, mlp_saved = mlp_fwd(S.sub("MlpBlock_0/"), ln2, x1)
    return x2, (x, ln1, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, mlp_saved)

  def dense1_bias_grad(self, P):
    """Receives colsum(d block-output): the gradient of this block's MlpBlock Dense_1 bias."""
    return self.scope(P).g("MlpBlock_0/Dense_1/bias")

  def bwd(self, P, dx2, saved, n, N, dx_colsum_out):
    """dx2: bf16 [M,d] grad of block output; colsum(dx2) has ALREADY been accumulated into
    this block's Dense_1 bias grad by whoever produced dx2.  Returns dx (grad of block input);
    colsum(dx) is accumulated into `dx_colsum_out` (the upstream bias/posemb gradient)."""
    d = self.d
    S = self.scope(P)
    A = S.sub("MultiHeadDotProductAttention_0/")
    x, ln1, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, mlp_saved = saved
    dln2 = mlp_bwd(S.sub("MlpBlock_0/"), dx2, mlp_saved, want_bias2_grad=False)
    dx1 = ops.layernorm_bwd(dln2, x1, S.f("LayerNorm_1/scale"), mean2, rstd2, dres=dx2,
                            dscale=S.g("LayerNorm_1/scale"), dbias=S.g("LayerNorm_1/bias"),
                            dx_colsum=A.g("out/bias"))
    del dln2
    o2 = o.view(n * N, d)
    ops.gemm(o2, dx1, a_mn=True, b_mn=True, out=A.g("out_proj/kernel"), reduce_out=True)
    do = ops.gemm(dx1, A.h("out_proj/kernel"))
    qkv3 = qkv.view(n, N, 3 * d)
    dqkv = torch.empty_like(qkv)
    dqkv3 = dqkv.view(n, N, 3 * d)
    gb = A.g("qkv/bias")     # q|k|v bias gradients come out of the attention backward
    ops.attention_bwd(do.view(n, N, d), qkv3[:, :, 0:d], qkv3[:, :, d:2 * d], qkv3[:, :, 2 * d:],
                      o, lse, self.heads, dq=dqkv3[:, :, 0:d], dk=dqkv3[:, :, d:2 * d],
                      dv=dqkv3[:, :, 2 * d:], dq_colsum=gb[0:d], dk_colsum=gb[d:2 * d],
                      dv_colsum=gb[2 * d:])
    del do
    ops.gemm(ln1, dqkv, a_mn=True, b_mn=True, out=A.g("qkv/kernel"), reduce_out=True)
    dln1 = ops.gemm(dqkv, A.h("qkv/kernel"))
    del dqkv
    dx = ops.layernorm_bwd(dln1, x, S.f("LayerNorm_0/scale"), mean1, rstd1, dres=dx1,
                           dscale=S.g("LayerNorm_0/scale"), dbias=S.g("LayerNorm_0/bias"),
                           dx_colsum=dx_colsum_out)
    return dx


class Encoder:
  """vit.Encoder (models/vit.py:115-160): depth blocks + LayerNorm("encoder_norm").

  scan=False: a Python loop over `encoderblock_{i}` (models/vit.py:151-158); everything the backward
  needs is kept.  scan=True: the reference's nn.scan over ONE `encoderblock` whose parameters carry
  a leading depth axis, each iteration wrapped in nn.remat with policy `nothing_saveable`
  (models/vit.py:129-148): only the block INPUT survives the forward and the block is recomputed
  in the backward, which is what makes L/14@336 at 2048 pairs per GPU fit in HBM."""

  def __init__(self, prefix, depth, d, m, heads, scan=False, remat_policy="nothing_saveable"):
    self.prefix, self.depth, self.d, self.scan = prefix, depth, d, scan
    if scan and remat_policy not in ("nothing_saveable", None):
      raise NotImplementedError(f"remat_policy={remat_policy!r}: only nothing_saveable (recompute the "
                                "whole block) is built")
    if scan:
      self.blocks = [EncoderBlock(f"{prefix}encoderblock/", d, m, heads, index=i) for i in range(depth)]
    else:
      self.blocks = [EncoderBlock(f"{prefix}encoderblock_{i}/", d, m, heads) for i in range(depth)]

  def specs(self):
    specs, aliases = [], []
    if self.scan:
      specs, aliases = self.blocks[0].specs(stack=self.depth)
    else:
      for b in self.blocks:
        s, a = b.specs()
        specs += s
        aliases += a
    specs = specs + ln_specs(self.prefix + "encoder_norm/", self.d)
    return specs, aliases

  def fwd(self, P, x, n, N):
    saved = []
    for b in self.blocks:
      x_in = x
      x, s = b.fwd(P, x, n, N)
      saved.append(x_in if self.scan else s)      # remat: keep the block input only
    return x, saved   # pre-encoder_norm activations; the caller applies encoder_norm

  def last_bias_grad(self, P):
    """Gradient buffer that must receive colsum(d x_out): the last block's Dense_1 bias."""
    return self.blocks[-1].dense1_bias_grad(P)

  def bwd(self, P, dx, saved, n, N, dx_colsum_out):
    for i in reversed(range(self.depth)):
      cs = self.blocks[i - 1].dense1_bias_grad(P) if i > 0 else dx_colsum_out
      s = saved[i]
      if self.scan:                               # recompute the block from its input
        x_out, s = self.blocks[i].fwd(P, s, n, N)
        del x_out
      dx = self.blocks[i].bwd(P, dx, s, n, N, cs)
      saved[i] = s = None
      # every parameter from this block on (in spec order) now has its final gradient: lets a
      # data-parallel trainer start reducing them while the earlier blocks are still running
      hook = getattr(P, "on_ready", None)
      if hook is not None and not self.scan:
        hook(self.blocks[i].p + "LayerNorm_0/scale")
    return dx


class MAPHead:
  """Multihead attention pooling (models/vit.py:163-183)."""

  def __init__(self, prefix, d, m, heads):
    self.p, self.d, self.m, self.heads = prefix, d, m, heads
    self.att = prefix + "MultiHeadDotProductAttention_0/"

  def specs(self):
    d = self.d
    s, a = mha_specs(self.att, d, self.heads, fuse_qkv=False)
    probe = E.ParamSpec(self.p + "probe", (1, 1, d), E.xavier_uniform(1, d))
    return ([probe] + s + ln_specs(self.p + "LayerNorm_0/", d)
            + mlp_specs(self.p + "MlpBlock_0/", d, self.m)), a

  def fwd(self, P, enc, n, N):
    d = self.d
    q1 = ops.gemm(P.h(self.p + "probe").view(1, d), P.h(self.att + "q/kernel"), b_mn=True,
                  bias=P.f(self.att + "q/bias"))
    qn = ops.broadcast_row(q1, n)
    kv = ops.gemm(enc, P.h(self.att + "kv/kernel"), b_mn=True, bias=P.f(self.att + "kv/bias"))
    kv3 = kv.view(n, N, 2 * d)
    o, lse = ops.attention_fwd(qn.view(n, 1, d), kv3[:, :, 0:d], kv3[:, :, d:], self.heads)
    a = ops.gemm(o.view(n, d), P.h(self.att + "out_proj/kernel"), b_mn=True, bias=P.f(self.att + "out/bias"))
    y, mean, rstd = ops.layernorm_fwd(a, P.f(self.p + "LayerNorm_0/scale"), P.f(self.p + "LayerNorm_0/bias"))
    out, mlp_saved = mlp_fwd(Scope(P, self.p + "MlpBlock_0/"), y, a, out_dtype=torch.float32)
    return out, (enc, qn, kv, o, lse, a, mean, rstd, mlp_saved)

  def bwd(self, P, dout, saved, n, N):
    """dout fp32 [n,d] -> d(enc) bf16 [n*N, d]."""
    d = self.d
    enc, qn, kv, o, lse, a, mean, rstd, mlp_saved = saved
    dout16 = ops.cast(dout, torch.empty_like(dout, dtype=torch.bfloat16))
    dy = mlp_bwd(Scope(P, self.p + "MlpBlock_0/"), dout16, mlp_saved, want_bias2_grad=True)
    da = ops.layernorm_bwd(dy, a, P.f(self.p + "LayerNorm_0/scale"), mean, rstd, dres=dout16,
                           dscale=P.g(self.p + "LayerNorm_0/scale"), dbias=P.g(self.p + "LayerNorm_0/bias"),
                           dx_colsum=P.g(self.att + "out/bias"))
    ops.gemm(o.view(n, d), da, a_mn=True, b_mn=True, out=P.g(self.att + "out_proj/kernel"), reduce_out=True)
    do = ops.gemm(da, P.h(self.att + "out_proj/kernel"))
    kv3 = kv.view(n, N, 2 * d)
    dkv = torch.empty_like(kv)
    dkv3 = dkv.view(n, N, 2 * d)
    dq = torch.empty_like(qn)
    ops.attention_bwd(do.view(n, 1, d), qn.view(n, 1, d), kv3[:, :, 0:d], kv3[:, :, d:], o, lse,
                      self.heads, dq=dq.view(n, 1, d), dk=dkv3[:, :, 0:d], dv=dkv3[:, :, d:])
    ops.colsum(dkv, P.g(self.att + "kv/bias"))
    ops.gemm(enc, dkv, a_mn=True, b_mn=True, out=P.g(self.att + "kv/kernel"), reduce_out=True)
    denc = ops.gemm(dkv, P.h(self.att + "kv/kernel"))
    # the single probe query is shared by the batch: its gradient is the batch sum of dq
    dq1 = torch.zeros(d, dtype=torch.float32, device=dq.device)
    ops.colsum(dq, dq1)
    ops.axpby(P.g(self.att + "q/bias"), dq1, 1.0, 1.0, out=P.g(self.att + "q/bias"))
    dq1h = ops.cast(dq1, torch.empty(d, dtype=torch.bfloat16, device=dq.device)).view(1, d)
    ops.gemm(P.h(self.p + "probe").view(1, d), dq1h, a_mn=True, b_mn=True,
             out=P.g(self.att + "q/kernel"), reduce_out=True)
    ops.gemm(dq1h, P.h(self.att + "q/kernel"), out=P.g(self.p + "probe").view(1, d), reduce_out=True)
    return denc


# ------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------
@dataclass
class _Model:
  """ViT model; fields as in models/vit.py:186-204."""
  num_classes: Optional[int] = None
  patch_size: Sequence[int] = (16, 16)
  width: int = 768
  depth: int = 12
  mlp_dim: Optional[int] = None
  num_heads: int = 12
  posemb: str = "learn"
  rep_size: Union[int, bool] = False
  dropout: float = 0.0
  pool_type: str = "gap"
  head_zeroinit: bool = True
  scan: bool = False
  remat_policy: str = "nothing_saveable"
  dtype_mm: str = "bfloat16"
  name: str = ""

  def __post_init__(self):
    if self.dropout:
      raise NotImplementedError("dropout > 0 is not on the benchmarked path (reference configs use 0)")
    if self.width % self.num_heads or self.width // self.num_heads != 64:
      raise NotImplementedError("the attention kernels are built for head dim 64")
    self.mlp = self.mlp_dim or 4 * self.width
    self.prefix = (self.name + "/") if self.name else ""
    self.encoder = Encoder(self.prefix + "Transformer/", self.depth, self.width, self.mlp, self.num_heads,
                           scan=self.scan, remat_policy=self.remat_policy)
    self.map_head = (MAPHead(self.prefix + "MAPHead_0/", self.width, self.mlp, self.num_heads)
                     if self.pool_type == "map" else None)
    self._geom = None

  # ---- parameters ------------------------------------------------------------------------
  def setup(self, image_hw):
    ph, pw = self.patch_size
    H, W = image_hw
    self._geom = (H // ph, W // pw)
    return self

  def specs(self, image_hw=None, in_ch=3):
    if image_hw is not None:
      self.setup(image_hw)
    gh, gw = self._geom
    d, p = self.width, self.prefix
    ph, pw = self.patch_size
    K = ph * pw * in_ch
    Kp = (K + 7) // 8 * 8
    lec = E.lecun_normal(K)   # flax Conv default kernel_init, fan_in = ph*pw*C
    specs = [
        E.ParamSpec(p + "embedding/kernel_flat", (Kp, d),
                    lambda rng, shape: np.concatenate([lec(rng, (K, d)), np.zeros((Kp - K, d))], 0)),
        E.ParamSpec(p + "embedding/bias", (d,), E.zeros),
    ]
    aliases = [E.Alias(p + "embedding/kernel", p + "embedding/kernel_flat",
                       lambda t: t[:K].unflatten(0, (ph, pw, in_ch)))]
    if self.posemb == "learn":
      specs.append(E.ParamSpec(p + "pos_embedding", (1, gh * gw, d), E.normal(1 / math.sqrt(d))))
    if self.pool_type == "tok":
      specs.append(E.ParamSpec(p + "cls", (1, 1, d), E.zeros))
    s, a = self.encoder.specs()
    specs += s
    aliases += a
    if self.map_head is not None:
      s, a = self.map_head.specs()
      specs += s
      aliases += a
    if self.rep_size:
      rep = d if self.rep_size is True else self.rep_size
      specs += [E.ParamSpec(p + "pre_logits/kernel", (d, rep), E.lecun_normal(d)),
                E.ParamSpec(p + "pre_logits/bias", (rep,), E.zeros)]
    if self.num_classes:
      rep = (d if self.rep_size is True else self.rep_size) if self.rep_size else d
      kinit = E.zeros if self.head_zeroinit else E.lecun_normal(rep)
      specs += [E.ParamSpec(p + "head/kernel", (rep, self.num_classes), kinit),
                E.ParamSpec(p + "head/bias", (self.num_classes,), E.zeros)]
    self._in_ch, self._K, self._Kp = in_ch, K, Kp
    return specs, aliases

  def init(self, seed, image_shape, device="cuda"):
    """Counterpart of model.init(rng, zeros_image)["params"] (train.py:195-205)."""
    specs, aliases = self.specs(image_shape[1:3], image_shape[3])
    return E.FlatParams(specs, aliases, device).init(seed)

  # ---- forward / backward ----------------------------------------------------------------
  def _posemb16(self, P):
    if self.posemb == "learn":
      return P.h(self.prefix + "pos_embedding").view(-1, self.width)
    if getattr(self, "_sincos", None) is None or self._sincos.device != P.device:
      gh, gw = self._geom
      self._sincos = torch.from_numpy(posemb_sincos_2d(gh, gw, self.width)).to(P.device).bfloat16()
    return self._sincos

  def fwd(self, P, image):
    """image [n,H,W,C] fp32 in [-1,1] -> (x fp32 [n, out], saved)."""
    if self._geom is None:
      self.setup(image.shape[1:3])
    n = image.shape[0]
    gh, gw = self._geom
    N0, d, p = gh * gw, self.width, self.prefix
    patches = ops.patchify(image, self.patch_size[0])
    x = ops.gemm(patches, P.h(p + "embedding/kernel_flat"), b_mn=True, bias=P.f(p + "embedding/bias"),
                 aux=self._posemb16(P), aux_row_mod=N0, epilogue=L.EPI_BIAS_RESID)
    N = N0
    if self.pool_type == "tok":
      # cls token is prepended AFTER the position embedding was added (models/vit.py:223-225)
      x = ops.concat_cls(x, P.f(p + "cls").view(d), n, N0)
      N = N0 + 1
    x, enc_saved = self.encoder.fwd(P, x, n, N)
    en = self.prefix + "Transformer/encoder_norm/"
    saved = {"patches": patches, "enc": enc_saved, "n": n, "N": N}
    if self.pool_type == "map":
      encd, mean, rstd = ops.layernorm_fwd(x, P.f(en + "scale"), P.f(en + "bias"))
      saved["norm"] = (x, mean, rstd)
      out, saved["map"] = self.map_head.fwd(P, encd, n, N)
    elif self.pool_type == "gap":
      encd, mean, rstd = ops.layernorm_fwd(x, P.f(en + "scale"), P.f(en + "bias"))
      saved["norm"] = (x, mean, rstd)
      out = ops.pool_fwd(encd, n, N, 0, out_dtype=torch.float32)
    elif self.pool_type in ("0", "tok"):
      # LayerNorm is per token, so LN(x)[:, 0] == LN(x[:, 0]): select first, normalise one row
      x0 = ops.pool_fwd(x, n, N, 1, tok=0)
      out, mean, rstd = ops.layernorm_fwd(x0, P.f(en + "scale"), P.f(en + "bias"), out_dtype=torch.float32)
      saved["norm"] = (x0, mean, rstd)
    elif self.pool_type == "none":
      # no pooling (models/vit.py:252-253): pre_logits / head run on every token, out is [n, N, .]
      out, mean, rstd = ops.layernorm_fwd(x, P.f(en + "scale"), P.f(en + "bias"))
      saved["norm"] = (x, mean, rstd)
    else:
      raise ValueError(f"Unknown pool type: '{self.pool_type}'")
    if self.rep_size:
      pre = ops.gemm(self._to16(out), P.h(p + "pre_logits/kernel"), b_mn=True,
                     bias=P.f(p + "pre_logits/bias"), out_dtype=torch.float32)
      saved["rep_in"] = out
      out = ops.tanh_fwd(pre)
      saved["rep_out"] = out
    if self.num_classes:
      saved["head_in"] = out
      out = ops.gemm(self._to16(out), P.h(p + "head/kernel"), b_mn=True, bias=P.f(p + "head/bias"),
                     out_dtype=torch.float32)
    if self.pool_type == "none":
      if out.dtype != torch.float32:
        out = ops.cast(out, torch.empty_like(out, dtype=torch.float32))
      out = out.view(n, N, -1)
    return out, saved

  @staticmethod
  def _to16(x):
    if x.dtype == torch.bfloat16:
      return x
    return ops.cast(x, torch.empty_like(x, dtype=torch.bfloat16))

  def bwd(self, P, dout, saved):
    """dout: fp32 [n, out] ([n, N, out] without pooling).  Accumulates parameter gradients into P.grad."""
    p, d = self.prefix, self.width
    n, N = saved["n"], saved["N"]
    en = self.prefix + "Transformer/encoder_norm/"
    if self.pool_type == "none":
      dout = dout.reshape(n * N, -1)
    if self.num_classes:
      d16 = self._to16(dout)
      ops.colsum(dout, P.g(p + "head/bias"))
      ops.gemm(self._to16(saved["head_in"]), d16, a_mn=True, b_mn=True, out=P.g(p + "head/kernel"), reduce_out=True)
      dout = ops.gemm(d16, P.h(p + "head/kernel"), out_dtype=torch.float32)
    if self.rep_size:
      dpre = ops.tanh_bwd(dout, saved["rep_out"])
      d16 = self._to16(dpre)
      ops.colsum(dpre, P.g(p + "pre_logits/bias"))
      ops.gemm(self._to16(saved["rep_in"]), d16, a_mn=True, b_mn=True, out=P.g(p + "pre_logits/kernel"), reduce_out=True)
      dout = ops.gemm(d16, P.h(p + "pre_logits/kernel"), out_dtype=torch.float32)
    last_b = self.encoder.last_bias_grad(P)
    if self.pool_type == "map":
      denc = self.map_head.bwd(P, dout, saved["map"], n, N)
      x, mean, rstd = saved["norm"]
      dx = ops.layernorm_bwd(denc, x, P.f(en + "scale"), mean, rstd, dscale=P.g(en + "scale"),
                             dbias=P.g(en + "bias"), dx_colsum=last_b)
    elif self.pool_type in ("gap", "none"):
      denc = ops.pool_bwd(dout, n, N, 0) if self.pool_type == "gap" else self._to16(dout)
      x, mean, rstd = saved["norm"]
      dx = ops.layernorm_bwd(denc, x, P.f(en + "scale"), mean, rstd, dscale=P.g(en + "scale"),
                             dbias=P.g(en + "bias"), dx_colsum=last_b)
    else:
      x0, mean, rstd = saved["norm"]
      dx0 = ops.layernorm_bwd(dout, x0, P.f(en + "scale"), mean, rstd, dscale=P.g(en + "scale"),
                              dbias=P.g(en + "bias"), dx_colsum=last_b)
      dx = ops.pool_bwd(dx0, n, N, 1, tok=0)
    # encoder: the column sum of the gradient reaching the embedding output is the patch-embed
    # bias gradient (models/vit.py:212-214)
    if self.pool_type == "tok":
      dx = self.encoder.bwd(P, dx, saved["enc"], n, N, None)
      # batch-sum of the gradient at every token position: row 0 is d cls, the rest d pos_embedding;
      # the patch-embed bias gradient is the sum of the latter over positions
      N0 = N - 1
      tmp = torch.zeros(N * d, dtype=torch.float32, device=dx.device)
      ops.colsum(dx.view(n, N * d), tmp)
      gcls = P.g(p + "cls").view(d)
      ops.axpby(gcls, tmp[:d], 1.0, 1.0, out=gcls)
      if self.posemb == "learn":
        gpos = P.g(p + "pos_embedding").view(N0 * d)
        ops.axpby(gpos, tmp[d:], 1.0, 1.0, out=gpos)
      ops.colsum(tmp[d:].view(N0, d), P.g(p + "embedding/bias"))
      dx = ops.drop_cls(dx, n, N0)
    else:
      dx = self.encoder.bwd(P, dx, saved["enc"], n, N, P.g(p + "embedding/bias"))
      if self.posemb == "learn":
        ops.colsum(dx.view(n, N * d), P.g(p + "pos_embedding").view(N * d))
    ops.gemm(saved["patches"], dx, a_mn=True, b_mn=True, out=P.g(p + "embedding/kernel_flat"), reduce_out=True)

  # ---- reference-style entry points --------------------------------------------------------
  def apply(self, variables, image, *, train=False):
    """(x, out) like flax apply (models/vit.py:206-276); `out` holds what this path keeps."""
    P = variables["params"]
    x, saved = self.fwd(P, image)
    out = {"head_input": x} if not (self.rep_size or self.num_classes) else {}
    out["pre_logits" if not self.num_classes else "logits"] = x
    return x, out


def Model(num_classes=None, *, variant=None, **kw):  # pylint: disable=invalid-name
  """Factory, same signature as big_vision.models.vit.Model (models/vit.py:279-281)."""
  return _Model(num_classes, **{**decode_variant(variant), **kw})


# ----------------------------------------------------------------------------------------------
# Checkpoint interchange (host side; nested dicts of numpy arrays under the reference's names).
# Contract of models/vit.py:306-433: accept every on-disk generation of ViT checkpoints the
# reference accepts and deliver a tree shaped like the model's freshly initialised parameters.
# ----------------------------------------------------------------------------------------------
def resample_posemb(old, new):
  """Position embeddings for a different input resolution ("high-res finetuning"): the square
  [1, g*g, d] grid `old` is bilinearly resized (order-1 spline, scipy.ndimage.zoom) to the grid size
  of `new`; returned unchanged when the shapes already agree."""
  old = np.asarray(old)
  want = tuple(new.shape)
  if old.shape == want:
    return old
  import scipy.ndimage
  side_from, side_to = (int(np.sqrt(shape[1])) for shape in (old.shape, want))
  ratio = side_to / side_from
  resized = scipy.ndimage.zoom(old.reshape(side_from, side_from, -1), (ratio, ratio, 1), order=1)
  return resized.reshape(1, side_to * side_to, -1)


# Older checkpoint generations, oldest quirk first; each entry rewrites the top-level tree in place.
def _posemb_out_of_encoder(tree):
  """The position embedding used to be a parameter of the encoder ("Transformer/pos_embedding", and
  before that of a "posembed_input" sub-module); today it is a top-level parameter."""
  enc = tree.get("Transformer")
  if not isinstance(enc, dict):
    return
  enc = tree["Transformer"] = dict(enc)
  if "posembed_input" in enc:
    tree["pos_embedding"] = enc.pop("posembed_input")["pos_embedding"]
  if "pos_embedding" in enc:
    tree["pos_embedding"] = enc.pop("pos_embedding")


def _cls_out_of_posemb(tree):
  """[cls] used to be concatenated BEFORE the position embedding was added, so old tables have
  g*g + 1 rows; the extra first row is folded into the cls parameter."""
  table = tree.get("pos_embedding")
  if table is None:
    return
  rows = int(table.shape[1])
  grid = int(np.sqrt(rows))
  if grid * grid + 1 != rows:
    return
  tree["pos_embedding"] = table[:, 1:]
  if "cls" in tree:
    tree["cls"] = tree["cls"] + table[:, :1]


def _map_head_into_module(tree):
  """The MAP head was written inline at first; its four parameter groups now live in "MAPHead_0"."""
  if "probe" not in tree:
    return
  moved = ("probe", "MlpBlock_0", "MultiHeadDotProductAttention_0", "LayerNorm_0")
  tree["MAPHead_0"] = {name: tree.pop(name) for name in moved}


_CHECKPOINT_FIXES = (_posemb_out_of_encoder, _cls_out_of_posemb, _map_head_into_module)


def fix_old_checkpoints(params):
  """Brings a ViT parameter tree of any older generation to today's layout (a new top-level dict;
  the input is not modified).  Pre-linen checkpoints cannot occur in .npz files written by
  linen-era code and are not handled."""
  tree = dict(params)
  for fix in _CHECKPOINT_FIXES:
    fix(tree)
  return tree


def _block_names(encoder_tree):
  names = [k for k in encoder_tree if k.startswith("encoderblock_")]
  return sorted(names, key=lambda k: int(k.rsplit("_", 1)[1]))


def _zip_trees(fn, trees):
  """fn over corresponding leaves of identically shaped dict trees."""
  first = trees[0]
  if isinstance(first, dict):
    return {k: _zip_trees(fn, [t[k] for t in trees]) for k in first}
  return fn(trees)


def pyloop_to_scan(params_pyloop, encoder="Transformer"):
  """Per-layer sub-trees "encoderblock_0..L-1" of the Python-loop encoder -> the single
  "encoderblock" of the scanned encoder, every leaf stacked along a new leading layer axis."""
  out = dict(params_pyloop)
  enc = dict(out[encoder])
  layers = _block_names(enc)
  if [int(k.rsplit("_", 1)[1]) for k in layers] != list(range(len(layers))):
    raise ValueError(f"encoder blocks are not numbered 0..{len(layers) - 1}: {layers}")
  enc["encoderblock"] = _zip_trees(np.stack, [enc.pop(k) for k in layers])
  out[encoder] = enc
  return out


def scan_to_pyloop(params_scan, encoder="Transformer"):
  """The inverse: slice l of every stacked leaf becomes layer "encoderblock_l"."""
  out = dict(params_scan)
  enc = dict(out[encoder])
  stacked = enc.pop("encoderblock")
  depth = len(stacked["LayerNorm_0"]["bias"])
  for layer in range(depth):
    enc[f"encoderblock_{layer}"] = _zip_trees(lambda leaves, layer=layer: leaves[0][layer], [stacked])
  out[encoder] = enc
  return out


def load(init_params, init_file, model_cfg, dont_load=()):
  """Parameters for `model_cfg` initialised from checkpoint `init_file` ("path.npz[:sub/tree]"):
  older layouts are modernised, the encoder is (un)stacked to match `model_cfg["scan"]`, names
  matching `dont_load` keep their fresh value from `init_params`, and the position embedding is
  resampled when the checkpoint was trained at another resolution.  Trees are nested dicts under the
  reference names (`utils.recover_tree(*zip(*P.numpy_tree().items()))` / `P.load_tree(dict(flat))`
  convert from and to a FlatParams)."""
  from big_vision_b200 import utils
  from big_vision_b200.models import common
  tree = fix_old_checkpoints(utils.load_params(init_file))
  stored_scanned = "encoderblock" in tree["Transformer"]
  if bool(model_cfg.get("scan")) != stored_scanned:
    tree = scan_to_pyloop(tree) if stored_scanned else pyloop_to_scan(tree)
  tree = common.merge_params(tree, init_params, dont_load)
  if init_params and "pos_embedding" in init_params:
    tree["pos_embedding"] = resample_posemb(old=tree["pos_embedding"], new=init_params["pos_embedding"])
  return tree
