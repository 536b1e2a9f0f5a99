"""The optimizer chain against the reference's own known-answer tests (big_vision/optax_test.py:
test_make_simple :103, test_make_wd :130, test_make_clip_norm :171, test_make_multi :206,
test_frozen_no_state :301), ported to the flat-buffer implementation: same configs, same parameter
names and values, the expected updates computed by the same closed forms."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(tree):
  from big_vision_b200 import engine as E
  specs = [E.ParamSpec(k, (1,) if np.isscalar(v) else tuple(np.shape(v)),
                       E.constant(v) if np.isscalar(v) else (lambda rng, shape, v=v: np.asarray(v)))
           for k, v in tree.items()]
  return E.FlatParams(specs, [], "cuda", decay_regex=None).init(0)


# The applied update is read back as (parameter after - parameter before) in fp32, so it carries the
# rounding of the fp32 parameter itself (1 ulp of |p| ~ 1e-7 |p|) on top of the update's own.
RTOL, ATOL = 1e-5, 1.5e-6


def _step(tx, state, P, grads):
  """One tx.update with the given gradient tree; returns the applied update per parameter."""
  before = P.numpy_tree("f")
  P.zero_grad()
  for k, g in grads.items():
    P.g(k).fill_(g)
  tx.update(P, state)
  after = P.numpy_tree("f")
  return {k: after[k] - before[k] for k in before}, before


def test_make_simple():
  from big_vision_b200 import optax as bv_optax
  P = _params({"Dense_0/kernel": 1.0, "Dense_0/bias": 2.0})
  config = dict(lr=0.01, schedule=dict(decay_type="linear"), optax_name="scale", optax=dict(step_size=0.5))
  total = 10
  tx, (sched_fn,) = bv_optax.make(config, P, sched_kw=dict(global_batch_size=1, total_steps=total))
  state = tx.init(P)
  for step in range(total):
    upd, _ = _step(tx, state, P, {k: 1.0 for k in P.offsets})
    assert state["count"] == step + 1
    sched = sched_fn(step)
    np.testing.assert_almost_equal(sched, 1.0 / total * (total - step))
    for k, v in upd.items():
      np.testing.assert_allclose(v, -sched * 0.01 * 0.5 * 1.0, rtol=RTOL, atol=ATOL, err_msg=k)


def test_make_wd():
  from big_vision_b200 import optax as bv_optax
  P = _params({"Dense_0/kernel": 1.0, "Dense_0/bias": 2.0, "Dense_0/other": 3.0})
  wds = {"Dense_0/kernel": 2e-3, "Dense_0/bias": 5e-4, "Dense_0/other": 0.0}
  config = dict(lr=0.01, wd=1e-3, wd_mults=[(".*/kernel", 2.0), (".*/bias", 0.5)],
                schedule=dict(decay_type="linear"), optax_name="scale", optax=dict(step_size=0.5))
  total = 10
  tx, (sched_fn,) = bv_optax.make(config, P, sched_kw=dict(global_batch_size=1, total_steps=total))
  state = tx.init(P)
  for step in range(total):
    upd, p = _step(tx, state, P, {k: 1.0 for k in P.offsets})
    sched = sched_fn(step)
    for k in upd:
      np.testing.assert_allclose(upd[k], -sched * (0.01 * 0.5 * 1.0 + p[k] * wds[k]), rtol=RTOL, atol=ATOL, err_msg=k)


def test_make_clip_norm():
  from big_vision_b200 import optax as bv_optax
  P = _params({"Dense_0/kernel": 1.0, "Dense_0/bias": 2.0, "Dense_0/other": 3.0})
  config = dict(lr=0.01, schedule=dict(decay_type="linear"), optax_name="scale", grad_clip_norm=1.0,
                optax=dict(step_size=0.5))
  total = 10
  tx, (sched_fn,) = bv_optax.make(config, P, sched_kw=dict(global_batch_size=1, total_steps=total))
  state = tx.init(P)
  factor = min(1.0, 1.0 / np.sqrt(3.0))
  for step in range(total):
    upd, _ = _step(tx, state, P, {k: 1.0 for k in P.offsets})
    sched = sched_fn(step)
    for k in upd:
      np.testing.assert_allclose(upd[k], -sched * 0.01 * 0.5 * factor, rtol=RTOL, atol=ATOL, err_msg=k)
    assert float(state["scalars"][0].sqrt()) == pytest.approx(np.sqrt(3.0), rel=1e-6)


def test_make_multi():
  from big_vision_b200 import optax as bv_optax
  vals = {f"Dense_{i}/{n}": float(3 * i + j + 1) for i in range(4) for j, n in enumerate(["kernel", "bias", "other"])}
  P = _params(vals)
  lrb, lr1, lr2, wdb, wd1, wd2, mult1, mult2 = 0.01, 2.0, 0.5, 1e-3, 10.0, 0.1, 1.0, 0.1
  lr_mults = {k: {"0": lr1, "1": lr2}.get(k[6], 1.0) for k in vals}
  wds = {k: 0.0 if k.startswith("Dense_3") else {"kernel": wd1 * wdb, "bias": wd2 * wdb, "other": 0.0}[k.split("/")[1]]
         for k in vals}
  config = dict(lr=lrb, lr_mults=[("Dense_0/.*", lr1), ("Dense_1/.*", lr2)], wd=wdb,
                wd_mults=[(".*/kernel", wd1), (".*/bias", wd2)],
                schedule=[("Dense_0/.*", dict(decay_type="linear", mult=mult1, linear_end=mult1)),
                          ("Dense_[12]/.*", dict(decay_type="linear", mult=mult2)), (".*", None)],
                optax_name="scale", grad_clip_norm=1.0, optax=dict(step_size=0.5))
  total = 10
  tx, (fn1, fn2) = bv_optax.make(config, P, sched_kw=dict(global_batch_size=1, total_steps=total))
  state = tx.init(P)
  sched_of = {k: {"0": fn1, "1": fn2, "2": fn2, "3": (lambda _: 0.0)}[k[6]] for k in vals}
  factor = min(1.0, 1.0 / np.sqrt(9.0))        # frozen Dense_3 does not count towards the norm
  for step in range(total):
    upd, p = _step(tx, state, P, {k: 1.0 for k in vals})
    np.testing.assert_almost_equal(fn1(step), mult1)
    np.testing.assert_almost_equal(fn2(step), mult2 * (total - step) / total)
    for k in vals:
      want = -sched_of[k](step) * (lrb * lr_mults[k] * 0.5 * factor + p[k] * wds[k])
      np.testing.assert_allclose(upd[k], want, rtol=RTOL, atol=ATOL, err_msg=f"{k} step {step}")


def test_frozen_no_state_and_uncovered_params():
  from big_vision_b200 import optax as bv_optax
  P = _params({"small": np.zeros(1, np.float32), "large": np.zeros(1000, np.float32)})
  config = dict(lr=0.01, schedule=[("small", dict(decay_type="cosine")), ("large", None)], optax_name="scale_by_adam")
  tx, fns = bv_optax.make(config, P, sched_kw=dict(global_batch_size=1, total_steps=1))
  state = tx.init(P)
  nbytes = sum(v.numel() * v.element_size() for k, v in state.items() if k in ("mu", "nu"))
  assert nbytes < 1_000 and len(fns) == 1
  P.g("large").fill_(1.0)
  P.g("small").fill_(1.0)
  tx.update(P, state)
  assert float(P.f("large").abs().max()) == 0.0 and float(P.f("small").abs().max()) > 0.0
  with pytest.raises(AssertionError):
    bv_optax.make(dict(lr=0.01, schedule=[("small", dict(decay_type="cosine"))]), P,
                  sched_kw=dict(total_steps=1))
  with pytest.raises(NotImplementedError):
    bv_optax.make(dict(lr=0.01, optax_name="lion"), P, sched_kw=dict(total_steps=1))


def test_adafactor_state_size_known_answer():
  """optax_test.py:320-341: a [1024, 1024] kernel keeps 2 * 1024 second-moment statistics (+ scalars)."""
  from big_vision_b200 import optax as bv_optax
  P = _params({"Dense_0/kernel": np.zeros((1024, 1024), np.float32)})
  tx, _ = bv_optax.make(dict(optax_name="big_vision.scale_by_adafactor", lr=0.01, schedule=dict(decay_type="linear")),
                        P, sched_kw=dict(global_batch_size=1, total_steps=1))
  state = tx.init(P)
  (st,) = state["af"]
  assert st["red_h"].numel() + st["red_l"].numel() == 2 * 1024 and "vfull" not in st


@pytest.mark.parametrize("momentum", [0.9, 0.0])
def test_adafactor_matches_oracle(momentum):
  """BV-Adafactor on every tensor geometry of the path -- Dense [in,out] with either axis largest,
  DenseGeneral q/k/v [d,h,dh] and out [h,dh,d] as views of fused storage, scan-stacked kernels,
  unfactored small / 1-D tensors -- against the numpy restatement of optax's factored rms, 4 steps,
  with weight decay, gradient clipping and a schedule in the chain."""
  from big_vision_b200 import engine as E, optax as bv_optax
  from oracle import bv_oracle as O
  rng = np.random.default_rng(0)
  d, h, dh = 96, 3, 32
  rnd = lambda rng_, shape: rng_.standard_normal(shape).astype(np.float32)
  specs = [E.ParamSpec("a/kernel", (64, 96), rnd), E.ParamSpec("b/kernel", (96, 64), rnd),
           E.ParamSpec("att/qkv/kernel", (d, 3 * d), rnd), E.ParamSpec("att/out_proj/kernel", (d, d), rnd),
           E.ParamSpec("stk/kernel", (2, 48, 40), rnd), E.ParamSpec("small/kernel", (16, 20), rnd),
           E.ParamSpec("a/bias", (100,), rnd), E.ParamSpec("pos_embedding", (1, 50, 64), rnd)]
  aliases = [E.Alias(f"att/{nm}/kernel", "att/qkv/kernel", lambda t, i=i: t[:, i * d:(i + 1) * d].unflatten(1, (h, dh)))
             for i, nm in enumerate(["query", "key", "value"])]
  aliases.append(E.Alias("att/out/kernel", "att/out_proj/kernel", lambda t: t.unflatten(0, (h, dh))))
  P = E.FlatParams(specs, aliases, "cuda").init(1)
  config = dict(optax_name="big_vision.scale_by_adafactor", optax=dict(momentum=momentum), lr=0.05, wd=0.01,
                grad_clip_norm=1.0, schedule=dict(decay_type="cosine", warmup_steps=1))
  tx, (sched_fn,) = bv_optax.make(config, P, sched_kw=dict(total_steps=10))
  state = tx.init(P)
  modes = {t.name: t.mode for t, *_ in tx.tensors}
  assert modes["a/kernel"] == 1 and modes["b/kernel"] == 2 and modes["att/query/kernel"] == 2
  assert modes["att/out/kernel"] == 1 and modes["small/kernel"] == 0 and modes["a/bias"] == 0
  ref_p = {k: v.astype(np.float64) for k, v in P.numpy_tree("f").items()}
  ref_state = {k: {} for k in ref_p}
  for step in range(4):
    P.zero_grad()
    grads = {}
    for k, v in P.tree("g").items():
      gk = rng.standard_normal(tuple(v.shape)).astype(np.float32) * (0.1 + step)
      v.copy_(torch.from_numpy(gk))
      grads[k] = gk.astype(np.float64)
    gnorm = np.sqrt(sum((g * g).sum() for g in grads.values()))
    factor = 1.0 if gnorm < 1.0 else 1.0 / gnorm
    tx.update(P, state)
    sched = sched_fn(step)
    for k in ref_p:
      wd = 0.01 if k.endswith("/kernel") else 0.0
      ref_p[k], ref_state[k] = O.adafactor_reference(ref_p[k], grads[k] * factor, ref_state[k], step, lr=0.05,
                                                     wd=wd, sched=sched, momentum=momentum)
    got = P.numpy_tree("f")
    for k in ref_p:
      np.testing.assert_allclose(got[k], ref_p[k], rtol=3e-3 if momentum else 2e-5, atol=1e-5,
                                 err_msg=f"{k} step {step}")
  # the bf16 shadow follows the master copy
  np.testing.assert_allclose(P.numpy_tree("h")["a/kernel"], got["a/kernel"], rtol=1e-2, atol=1e-3)
