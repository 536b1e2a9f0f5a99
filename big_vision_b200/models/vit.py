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
  """MoCo v3 layout [sin x, cos x, sin y, cos y] (models/vit.py:34-44). Returns [h*w, width]."""
  y, x = np.mgrid[:h, :w]
  assert width % 4 == 0, "Width must be mult of 4 for sincos posemb"
  omega = np.arange(width // 4) / (width // 4 - 1)
  omega = 1.0 / (temperature ** omega)
  y = np.einsum("m,d->md", y.flatten(), omega)
  x = np.einsum("m,d->md", x.flatten(), omega)
  pe = np.concatenate([np.sin(x), np.cos(x), np.sin(y), np.cos(y)], axis=1)
  return pe.astype(np.float32)


def decode_variant(variant):
  """Converts a string like "B" or "B/32" into a params dict (models/vit.py:284-303)."""
  if variant is None:
    return {}
  v, patch = variant, {}
  if "/" in variant:
    v, patch = variant.split("/")
    patch = {"patch_size": (int(patch), int(patch))}
  return {
      "width": {"mu": 32, "Ti": 192, "S": 384, "M": 512, "B": 768, "L": 1024, "So400m": 1152, "H": 1280, "g": 1408, "g-opt": 1536, "G": 1664, "G-opt": 1536, "e": 1792}[v],
      "depth": {"mu": 1, "Ti": 12, "S": 12, "M": 12, "B": 12, "L": 24, "So400m": 27, "H": 32, "g": 40, "g-opt": 40, "G": 48, "G-opt": 48, "e": 56}[v],
      "mlp_dim": {"mu": 128, "Ti": 768, "S": 1536, "M": 2048, "B": 3072, "L": 4096, "So400m": 4304, "H": 5120, "g": 6144, "g-opt": 6144, "G": 8192, "G-opt": 8192, "e": 15360}[v],
      "num_heads": {"mu": 2, "Ti": 3, "S": 6, "M": 8, "B": 12, "L": 16, "So400m": 16, "H": 16, "g": 16, "g-opt": 16, "G": 16, "G-opt": 16, "e": 16}[v],
      **patch
  }


# ------------------------------------------------------------------------------------------
# building blocks (each: specs(), fwd(), bwd()); `P` is an engine.FlatParams
# ------------------------------------------------------------------------------------------
def ln_specs(p, d):
  return [E.ParamSpec(p + "scale", (d,), E.ones), E.ParamSpec(p + "bias", (d,), E.zeros)]


def mlp_specs(p, d, m):
  """MlpBlock (models/vit.py:57-78): xavier_uniform kernels, normal(1e-6) biases."""
  return [
      E.ParamSpec(p + "Dense_0/kernel", (d, m), E.xavier_uniform(d, m)),
      E.ParamSpec(p + "Dense_0/bias", (m,), E.normal(1e-6)),
      E.ParamSpec(p + "Dense_1/kernel", (m, d), E.xavier_uniform(m, d)),
      E.ParamSpec(p + "Dense_1/bias", (d,), E.normal(1e-6)),
  ]


def mlp_fwd(P, p, y, resid, out_dtype=torch.bfloat16):
  """resid + Dense_1(gelu(Dense_0(y))).  Returns (out, saved)."""
  act, pre = ops.gemm(y, P.h(p + "Dense_0/kernel"), b_mn=True, bias=P.f(p + "Dense_0/bias"),
                      epilogue=L.EPI_BIAS_GELU)
  out = ops.gemm(act, P.h(p + "Dense_1/kernel"), b_mn=True, bias=P.f(p + "Dense_1/bias"),
                 aux=resid, epilogue=L.EPI_BIAS_RESID if resid is not None else L.EPI_BIAS,
                 out_dtype=out_dtype)
  return out, (y, act, pre)


def mlp_bwd(P, p, dout, saved, want_bias2_grad=True):
  """dout: bf16 [M,d] gradient of the block output.  Returns d(y) (bf16).

  The bias gradient of Dense_1 is colsum(dout); callers that already have that column sum
  from the LayerNorm-backward kernel pass want_bias2_grad=False."""
  y, act, pre = saved
  if want_bias2_grad:
    ops.colsum(dout, P.g(p + "Dense_1/bias"))
  ops.gemm(act, dout, a_mn=True, b_mn=True, out=P.g(p + "Dense_1/kernel"), reduce_out=True)
  # the Dense_0 bias gradient (column sums of dpre) is accumulated by the same GEMM's epilogue
  dpre = ops.gemm(dout, P.h(p + "Dense_1/kernel"), aux=pre, epilogue=L.EPI_DGELU,
                  colsum=P.g(p + "Dense_0/bias"))
  ops.gemm(y, dpre, a_mn=True, b_mn=True, out=P.g(p + "Dense_0/kernel"), reduce_out=True)
  return ops.gemm(dpre, P.h(p + "Dense_0/kernel"))


def mha_specs(p, d, heads, fuse_qkv=True):
  """flax MultiHeadDotProductAttention params: query/key/value kernels [d,h,dh] + bias [h,dh],
  out kernel [h,dh,d] + bias [d]; kernel_init xavier_uniform (models/vit.py:95,177), zero biases.
  Stored fused ([d,3d] or q:[d,d] + kv:[d,2d]) and aliased to the reference names."""
  dh = d // heads
  xav = E.xavier_uniform(d, d)

  def fused_init(k):
    return lambda rng, shape: np.concatenate([xav(rng, (d, d)) for _ in range(k)], axis=1)

  specs, aliases = [], []

  def alias_cols(store, names, width):
    for i, nm in enumerate(names):
      aliases.append(E.Alias(p + nm + "/kernel", p + store + "/kernel",
                             lambda t, i=i: t[:, i * d:(i + 1) * d].unflatten(1, (heads, dh))))
      aliases.append(E.Alias(p + nm + "/bias", p + store + "/bias",
                             lambda t, i=i: t[i * d:(i + 1) * d].unflatten(0, (heads, dh))))

  if fuse_qkv:
    specs += [E.ParamSpec(p + "qkv/kernel", (d, 3 * d), fused_init(3)),
              E.ParamSpec(p + "qkv/bias", (3 * d,), E.zeros)]
    alias_cols("qkv", ["query", "key", "value"], 3 * d)
  else:
    specs += [E.ParamSpec(p + "q/kernel", (d, d), xav), E.ParamSpec(p + "q/bias", (d,), E.zeros),
              E.ParamSpec(p + "kv/kernel", (d, 2 * d), fused_init(2)),
              E.ParamSpec(p + "kv/bias", (2 * d,), E.zeros)]
    alias_cols("q", ["query"], d)
    alias_cols("kv", ["key", "value"], 2 * d)
  specs += [E.ParamSpec(p + "out_proj/kernel", (d, d), xav), E.ParamSpec(p + "out/bias", (d,), E.zeros)]
  aliases.append(E.Alias(p + "out/kernel", p + "out_proj/kernel",
                         lambda t: t.unflatten(0, (heads, dh))))
  return specs, aliases


class EncoderBlock:
  """Encoder1DBlock (models/vit.py:81-112): x + MHSA(LN(x)); x + MLP(LN(x))."""

  def __init__(self, prefix, d, m, heads):
    self.p, self.d, self.m, self.heads = prefix, d, m, heads
    self.att = prefix + "MultiHeadDotProductAttention_0/"

  def specs(self):
    s, a = mha_specs(self.att, self.d, self.heads)
    return (ln_specs(self.p + "LayerNorm_0/", self.d) + s + ln_specs(self.p + "LayerNorm_1/", self.d)
            + mlp_specs(self.p + "MlpBlock_0/", self.d, self.m)), a

  def fwd(self, P, x, n, N):
    p, d = self.p, self.d
    ln1, mean1, rstd1 = ops.layernorm_fwd(x, P.f(p + "LayerNorm_0/scale"), P.f(p + "LayerNorm_0/bias"))
    qkv = ops.gemm(ln1, P.h(self.att + "qkv/kernel"), b_mn=True, bias=P.f(self.att + "qkv/bias"))
    qkv3 = qkv.view(n, N, 3 * d)
    o, lse = ops.attention_fwd(qkv3[:, :, 0:d], qkv3[:, :, d:2 * d], qkv3[:, :, 2 * d:], self.heads)
    x1 = ops.gemm(o.view(n * N, d), P.h(self.att + "out_proj/kernel"), b_mn=True,
                  bias=P.f(self.att + "out/bias"), aux=x, epilogue=L.EPI_BIAS_RESID)
    ln2, mean2, rstd2 = ops.layernorm_fwd(x1, P.f(p + "LayerNorm_1/scale"), P.f(p + "LayerNorm_1/bias"))
    x2, mlp_saved = mlp_fwd(P, p + "MlpBlock_0/", ln2, x1)
    return x2, (x, ln1, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, mlp_saved)

  def bwd(self, P, dx2, saved, n, N, dx_colsum_out):
    """dx2: bf16 [M,d] grad of block output; colsum(dx2) has ALREADY been accumulated into
    this block's Dense_1 bias grad by whoever produced dx2.  Returns dx (grad of block input);
    colsum(dx) is accumulated into `dx_colsum_out` (the upstream bias/posemb gradient)."""
    p, d = self.p, self.d
    x, ln1, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, mlp_saved = saved
    dln2 = mlp_bwd(P, p + "MlpBlock_0/", dx2, mlp_saved, want_bias2_grad=False)
    dx1 = ops.layernorm_bwd(dln2, x1, P.f(p + "LayerNorm_1/scale"), mean2, rstd2, dres=dx2,
                            dscale=P.g(p + "LayerNorm_1/scale"), dbias=P.g(p + "LayerNorm_1/bias"),
                            dx_colsum=P.g(self.att + "out/bias"))
    del dln2
    o2 = o.view(n * N, d)
    ops.gemm(o2, dx1, a_mn=True, b_mn=True, out=P.g(self.att + "out_proj/kernel"), reduce_out=True)
    do = ops.gemm(dx1, P.h(self.att + "out_proj/kernel"))
    qkv3 = qkv.view(n, N, 3 * d)
    dqkv = torch.empty_like(qkv)
    dqkv3 = dqkv.view(n, N, 3 * d)
    gb = P.g(self.att + "qkv/bias")     # q|k|v bias gradients come out of the attention backward
    ops.attention_bwd(do.view(n, N, d), qkv3[:, :, 0:d], qkv3[:, :, d:2 * d], qkv3[:, :, 2 * d:],
                      o, lse, self.heads, dq=dqkv3[:, :, 0:d], dk=dqkv3[:, :, d:2 * d],
                      dv=dqkv3[:, :, 2 * d:], dq_colsum=gb[0:d], dk_colsum=gb[d:2 * d],
                      dv_colsum=gb[2 * d:])
    del do
    ops.gemm(ln1, dqkv, a_mn=True, b_mn=True, out=P.g(self.att + "qkv/kernel"), reduce_out=True)
    dln1 = ops.gemm(dqkv, P.h(self.att + "qkv/kernel"))
    del dqkv
    dx = ops.layernorm_bwd(dln1, x, P.f(p + "LayerNorm_0/scale"), mean1, rstd1, dres=dx1,
                           dscale=P.g(p + "LayerNorm_0/scale"), dbias=P.g(p + "LayerNorm_0/bias"),
                           dx_colsum=dx_colsum_out)
    return dx


class Encoder:
  """vit.Encoder (models/vit.py:115-160): depth blocks + LayerNorm("encoder_norm").
  Python loop over blocks (the reference's scan=False path); param names encoderblock_{i}."""

  def __init__(self, prefix, depth, d, m, heads):
    self.prefix, self.depth, self.d = prefix, depth, d
    self.blocks = [EncoderBlock(f"{prefix}encoderblock_{i}/", d, m, heads) for i in range(depth)]

  def specs(self):
    specs, aliases = [], []
    for b in self.blocks:
      s, a = b.specs()
      specs += s
      aliases += a
    specs += ln_specs(self.prefix + "encoder_norm/", self.d)
    return specs, aliases

  def fwd(self, P, x, n, N):
    saved = []
    for b in self.blocks:
      x, s = b.fwd(P, x, n, N)
      saved.append(s)
    return x, saved   # pre-encoder_norm activations; the caller applies encoder_norm

  def last_bias_grad(self, P):
    """Gradient buffer that must receive colsum(d x_out): the last block's Dense_1 bias."""
    return P.g(self.blocks[-1].p + "MlpBlock_0/Dense_1/bias")

  def bwd(self, P, dx, saved, n, N, dx_colsum_out):
    for i in reversed(range(self.depth)):
      cs = (P.g(self.blocks[i - 1].p + "MlpBlock_0/Dense_1/bias") if i > 0 else dx_colsum_out)
      dx = self.blocks[i].bwd(P, dx, saved[i], n, N, cs)
      saved[i] = None
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
    out, mlp_saved = mlp_fwd(P, self.p + "MlpBlock_0/", y, a, out_dtype=torch.float32)
    return out, (enc, qn, kv, o, lse, a, mean, rstd, mlp_saved)

  def bwd(self, P, dout, saved, n, N):
    """dout fp32 [n,d] -> d(enc) bf16 [n*N, d]."""
    d = self.d
    enc, qn, kv, o, lse, a, mean, rstd, mlp_saved = saved
    dout16 = ops.cast(dout, torch.empty_like(dout, dtype=torch.bfloat16))
    dy = mlp_bwd(P, self.p + "MlpBlock_0/", dout16, mlp_saved, want_bias2_grad=True)
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
    self.encoder = Encoder(self.prefix + "Transformer/", self.depth, self.width, self.mlp, self.num_heads)
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
    return out, saved

  @staticmethod
  def _to16(x):
    if x.dtype == torch.bfloat16:
      return x
    return ops.cast(x, torch.empty_like(x, dtype=torch.bfloat16))

  def bwd(self, P, dout, saved):
    """dout: fp32 [n, out].  Accumulates parameter gradients into P.grad."""
    p, d = self.prefix, self.width
    n, N = saved["n"], saved["N"]
    en = self.prefix + "Transformer/encoder_norm/"
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
    elif self.pool_type == "gap":
      denc = ops.pool_bwd(dout, n, N, 0)
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
# Checkpoint loading (host side; nested dicts of numpy arrays under the reference's names).
# Mirrors models/vit.py:306-433.
# ----------------------------------------------------------------------------------------------
def resample_posemb(old, new):
  """'High-res finetuning': bilinear (order-1 spline) rescale of the [1, N, d] grid of position
  embeddings to the shape of `new` -- models/vit.py:306-322."""
  import scipy.ndimage
  old = np.asarray(old)
  if old.shape == tuple(new.shape):
    return old
  gs_old = int(np.sqrt(old.shape[1]))
  gs_new = int(np.sqrt(new.shape[1]))
  grid = old.reshape(gs_old, gs_old, -1)
  zoom = (gs_new / gs_old, gs_new / gs_old, 1)
  grid = scipy.ndimage.zoom(grid, zoom, order=1)
  return grid.reshape(1, gs_new * gs_new, -1)


def fix_old_checkpoints(params):
  """Small backward incompatibilities of old ViT checkpoints -- models/vit.py:325-365 (the
  pre-linen conversion of the reference is not applicable to .npz trees written by linen code)."""
  params = {k: (dict(v) if isinstance(v, dict) else v) for k, v in params.items()}
  t = params.get("Transformer", {})
  if "posembed_input" in t:                       # original ViT paper variant: posemb in a module
    params["pos_embedding"] = t.pop("posembed_input")["pos_embedding"]
  if "pos_embedding" in t:                        # pre-2022: posemb inside the Encoder
    params["pos_embedding"] = t.pop("pos_embedding")
  if "pos_embedding" in params:                   # old: [cls] concatenated before adding posemb
    pe = params["pos_embedding"]
    if int(np.sqrt(pe.shape[1])) ** 2 + 1 == int(pe.shape[1]):
      pe_cls, params["pos_embedding"] = pe[:, :1], pe[:, 1:]
      if "cls" in params:
        params["cls"] = params["cls"] + pe_cls
  if "probe" in params:                           # MAP head inlined during ViT-G development
    params["MAPHead_0"] = {k: params.pop(k) for k in
                           ["probe", "MlpBlock_0", "MultiHeadDotProductAttention_0", "LayerNorm_0"]}
  return params


def _map_leaves(fn, *trees):
  if isinstance(trees[0], dict):
    return {k: _map_leaves(fn, *[t[k] for t in trees]) for k in trees[0]}
  return fn(*trees)


def pyloop_to_scan(params_pyloop, encoder="Transformer"):
  """encoderblock_{i} sub-trees -> one 'encoderblock' with a leading depth axis -- vit.py:368-390."""
  params = dict(params_pyloop)
  t = dict(params[encoder])
  blocks = {k for k in t if k.startswith("encoderblock_")}
  depth = 1 + max(int(k.split("_")[-1]) for k in blocks)
  t["encoderblock"] = _map_leaves(lambda *v: np.stack(v), *[t[f"encoderblock_{i}"] for i in range(depth)])
  for i in range(depth):
    del t[f"encoderblock_{i}"]
  params[encoder] = t
  return params


def scan_to_pyloop(params_scan, encoder="Transformer"):
  """The inverse of pyloop_to_scan -- vit.py:393-409."""
  params = dict(params_scan)
  t = dict(params[encoder])
  depth = len(t["encoderblock"]["LayerNorm_0"]["bias"])
  for i in range(depth):
    t[f"encoderblock_{i}"] = _map_leaves(lambda x, i=i: x[i], t["encoderblock"])
  del t["encoderblock"]
  params[encoder] = t
  return params


def load(init_params, init_file, model_cfg, dont_load=()):
  """Init from a checkpoint, old formats included, + hi-res posemb -- models/vit.py:412-433.
  `init_params` / the result are nested dicts under the reference names (use
  `utils.recover_tree(*zip(*P.numpy_tree().items()))` and `P.load_tree(dict(flatten))` to go from
  and to a FlatParams).  This implementation runs the blocks as a Python loop (`scan=False`)."""
  from big_vision_b200 import utils
  from big_vision_b200.models import common
  restored = utils.load_params(init_file)
  restored = fix_old_checkpoints(restored)
  if model_cfg.get("scan") and "encoderblock" not in restored["Transformer"]:
    restored = pyloop_to_scan(restored)
  if not model_cfg.get("scan") and "encoderblock" in restored["Transformer"]:
    restored = scan_to_pyloop(restored)
  restored = common.merge_params(restored, init_params, dont_load)
  if init_params and "pos_embedding" in init_params:
    restored["pos_embedding"] = resample_posemb(old=restored["pos_embedding"],
                                                new=init_params["pos_embedding"])
  return restored
