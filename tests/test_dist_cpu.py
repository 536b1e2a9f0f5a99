"""CPU, world_size 2 over gloo: the host-side sharding logic of the SigLIP loss
(trainers/proj/image_text/siglip.py in this repo): which columns hold a rank's positives,
all-gather order, reduce-scatter of d ztxt, SUM (not mean) of per-rank partial losses.

The CUDA kernels cannot run here, so the three ops the function calls are replaced by
oracle-backed test doubles (this is the test harness, not a product fallback: the product's
ops refuse CPU tensors, see tests/test_abi.py)."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret, loss_fn="sigmoid"):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from big_vision_b200 import ops
  from big_vision_b200.trainers.proj.image_text import siglip
  from oracle import bv_oracle as O

  # ---- test doubles for the three kernels used by sigmoid_loss_fwd_bwd --------------------
  def cast(src, dst):
    return src          # keep fp32: this test is about the sharding algebra, not bf16 rounding

  def gemm(a, b, a_mn=False, b_mn=False, out_dtype=torch.float32, out=None, reduce_out=False, **kw):
    A = a.double().T if a_mn else a.double()
    Bm = b.double() if b_mn else b.double().T
    res = A @ Bm
    if out is None:
      return res.to(out_dtype)
    if reduce_out:
      out += res.to(out.dtype)
    else:
      out.copy_(res.to(out.dtype))
    return out

  def siglip_loss(dots, row_offset, t_param, b_param, global_b, loss, dt, db):
    d = dots.double().requires_grad_(True)
    t = t_param.double().requires_grad_(True)
    b = b_param.double().requires_grad_(True)
    n, B = d.shape
    m = -torch.ones(n, B, dtype=torch.float64)
    cols = row_offset + torch.arange(n)
    ok = (cols >= 0) & (cols < B)                 # a block without this rank's positives: none
    m[torch.arange(n)[ok], cols[ok]] = 1
    l = -(torch.nn.functional.logsigmoid(m * (d * t.exp() + b))).sum() / global_b
    l.backward()
    loss += l.detach().float()
    dt += t.grad.float()
    db += b.grad.float()
    return d.grad.float()

  def softmax_contrastive_loss(dots, row_offset, t_param, global_b, weight, loss, dt, ncorrect):
    d = dots.double().requires_grad_(True)
    t = t_param.double().requires_grad_(True)
    n = d.shape[0]
    x = d * t.exp()
    idx = torch.arange(n)
    l = weight * (torch.logsumexp(x, 1) - x[idx, row_offset + idx]).sum() / global_b
    l.backward()
    loss += l.detach().float()
    dt += t.grad.float()
    ncorrect += float((x.argmax(1) == row_offset + idx).sum())
    return d.grad.float()

  def axpby(x, y, a=1.0, b=1.0, out=None):
    return a * x + b * y

  ops.cast, ops.gemm, ops.siglip_loss = cast, gemm, siglip_loss
  ops.softmax_contrastive_loss, ops.axpby = softmax_contrastive_loss, axpby

  class FakeP:
    offsets = {"t": 0, "b": 1}
    _f = {"t": torch.tensor([math.log(10.0)]), "b": torch.tensor([-10.0])}
    _g = {"t": torch.zeros(1), "b": torch.zeros(1)}
    def f(self, k): return self._f[k]
    def g(self, k): return self._g[k]

  g = torch.Generator().manual_seed(0)
  B, D = 12, 16
  n = B // world
  zi = O.l2_normalize(torch.randn(B, D, generator=g).double())
  zt = O.l2_normalize(torch.randn(B, D, generator=g).double())
  P = FakeP()
  scal = torch.zeros(4)
  fn = siglip._loss_fn({"loss_fn": loss_fn})
  dzimg, dztxt = fn(P, zi[rank * n:(rank + 1) * n].float(), zt[rank * n:(rank + 1) * n].float(),
                    siglip.Dist(), scal)
  dist.all_reduce(scal)
  dist.all_reduce(P.g("t"))
  dist.all_reduce(P.g("b"))
  # oracle on the GLOBAL batch
  zir, ztr = zi.clone().requires_grad_(True), zt.clone().requires_grad_(True)
  tr = torch.tensor(math.log(10.0), dtype=torch.float64, requires_grad=True)
  br = torch.tensor(-10.0, dtype=torch.float64, requires_grad=True)
  if loss_fn == "softmax":
    l, _ = O.softmax_contrastive_loss(zir, ztr, tr.exp())
    (l + 0.0 * br).backward()
  else:
    l = O.siglip_loss(zir, ztr, tr.exp(), br)
    l.backward()
  errs = {
      "loss": abs(float(scal[0]) - float(l.detach())),
      "dzimg": float((dzimg.double() - zir.grad[rank * n:(rank + 1) * n]).abs().max()),
      "dztxt": float((dztxt.double() - ztr.grad[rank * n:(rank + 1) * n]).abs().max()),
      "dt": abs(float(P.g("t")) - float(tr.grad)),
      "db": abs(float(P.g("b")) - float(br.grad)),
  }
  ret[rank] = errs
  dist.destroy_process_group()


@pytest.mark.parametrize("world,loss_fn", [(2, "sigmoid"), (3, "sigmoid"), (2, "chunked_sigmoid"),
                                           (3, "chunked_sigmoid"), (2, "softmax"), (3, "softmax")])
def test_sharded_sigmoid_loss_equals_global_loss(world, loss_fn):
  """Both DP forms of the loss -- all-gather ([n,B] slab) and the paper's chunked rounds ([n,n]
  blocks, broadcast from / reduce to the chunk's owner) -- reproduce the global-batch loss and
  gradients (SURVEY 8e invariant)."""
  port = 29500 + os.getpid() % 1000 + world + {"sigmoid": 0, "chunked_sigmoid": 10, "softmax": 20}[loss_fn]
  ctx = mp.get_context("spawn")
  ret = ctx.Manager().dict()
  procs = [ctx.Process(target=_worker, args=(r, world, port, ret, loss_fn)) for r in range(world)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(120)
    assert p.exitcode == 0
  for r in range(world):
    errs = ret.get(r)
    assert errs is not None, f"rank {r} returned nothing"
    assert all(v < 2e-5 for v in errs.values()), (r, dict(errs))


def _bucket_worker(rank, world, port, ret):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "tests"))
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  import common
  from big_vision_b200.models.proj.image_text import two_towers
  from big_vision_b200.trainers.proj.image_text import siglip
  model = two_towers.Model(**common.TINY)
  specs, aliases = model.specs(common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE)
  from big_vision_b200 import engine
  P = engine.FlatParams(specs, aliases, "cpu")
  rng = np.random.default_rng(100 + rank)
  full = torch.from_numpy(rng.standard_normal(P.total).astype(np.float32))
  expect = full.clone()
  dist.all_reduce(expect)
  red = siglip.BucketedGradAllReduce(P, siglip.Dist(), bucket_elems=3000)
  launched = []
  orig = red._launch
  red._launch = lambda lo, hi: (launched.append((lo, hi)), orig(lo, hi))[1]
  red.begin()
  # the backward, as far as the reducer can tell: gradients become final in REVERSE spec order and
  # every encoder block reports once its first parameter (LayerNorm_0/scale) is done
  for spec in reversed(P.specs):
    off, shape = P.offsets[spec.name]
    n = int(np.prod(shape))
    P.grad[off:off + n] = full[off:off + n]
    if spec.name.endswith("LayerNorm_0/scale") and "encoderblock" in spec.name:
      P.on_ready(spec.name)
  red.finish()
  # padding between parameters is never written by a backward: compare the parameter slots only
  ok = True
  for name, (off, shape) in P.offsets.items():
    n = int(np.prod(shape))
    ok &= bool(torch.equal(P.grad[off:off + n], expect[off:off + n]))
  spans = sorted(launched)
  ret[rank] = {"ok": ok, "buckets": len(launched),
               "disjoint": all(a[1] <= b[0] for a, b in zip(spans, spans[1:])),
               "covered": sum(hi - lo for lo, hi in spans) == P.total}
  dist.destroy_process_group()


def test_bucketed_gradient_all_reduce_equals_one_all_reduce():
  """BucketedGradAllReduce (C3 overlapped with the backward): slices are launched only once final,
  are disjoint, cover the whole flat buffer, and the result equals a single all-reduce."""
  world = 2
  port = 29300 + os.getpid() % 500
  ctx = mp.get_context("spawn")
  ret = ctx.Manager().dict()
  procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, ret)) for r in range(world)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(120)
    assert p.exitcode == 0
  for r in range(world):
    res = dict(ret[r])
    assert res["ok"] and res["disjoint"] and res["covered"], res
    assert res["buckets"] >= 4, res          # the backward really was overlapped in several slices
