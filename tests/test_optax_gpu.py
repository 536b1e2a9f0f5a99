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
      np.testing.assert_allclose(v, -sched * 0.01 * 0.5 * 1.0, rtol=2e-6, atol=1e-9, err_msg=k)


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
      np.testing.assert_allclose(upd[k], -sched * (0.01 * 0.5 * 1.0 + p[k] * wds[k]), rtol=3e-6, atol=1e-9, err_msg=k)


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
      np.testing.assert_allclose(upd[k], -sched * 0.01 * 0.5 * factor, rtol=3e-6, atol=1e-9, err_msg=k)
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
      np.testing.assert_allclose(upd[k], want, rtol=4e-6, atol=1e-9, err_msg=f"{k} step {step}")


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
    bv_optax.make(dict(lr=0.01, optax_name="big_vision.scale_by_adafactor"), P, sched_kw=dict(total_steps=1))
