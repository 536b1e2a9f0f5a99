"""CPU: host logic pinned by the reference's own known-answer tables
(big_vision/utils_test.py:228-281)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from big_vision_b200 import optax as bv_optax
from big_vision_b200 import utils as u


@pytest.mark.parametrize("data_size,batch_size,total,cfg,expected", [
    (1000, None, None, dict(foo_steps=3), 3),
    (1000, 100, None, dict(foo_epochs=3), 30),
    (None, 100, None, dict(foo_examples=300), 3),
    (None, None, 10, dict(foo_percent=0.30), 3),
    (1000, 100, 10, dict(foo_steps=-1, foo_epochs=-1, foo_examples=-1, foo_percent=0.30), 3),
    (None, None, 10, dict(foo_percent=0.0), 0),
    (1001, None, None, dict(foo_steps=3), 3),
    (1001, 100, None, dict(foo_epochs=3), 30),
    (None, 101, None, dict(foo_examples=300), 3),
    (None, None, 11, dict(foo_percent=0.30), 3),
])
def test_steps(data_size, batch_size, total, cfg, expected):   # utils_test.py:230-256
  assert u.steps("foo", cfg, data_size=data_size, batch_size=batch_size, total_steps=total) == expected
  with pytest.raises(ValueError):
    u.steps("bar", cfg, data_size=data_size, batch_size=batch_size, total_steps=total)
  assert u.steps("bar", cfg, data_size=data_size, batch_size=batch_size, total_steps=total,
                 default=1234) == 1234


@pytest.mark.parametrize("decay_type,extra,step,expected", [
    ("linear", {}, 13, .5),
    ("polynomial", {"end": .1, "power": 2}, 13, .325),
    ("cosine", {}, 13, .5),
    ("rsqrt", {"timescale": 1}, 13, 0.3333333),
    ("stair", {"steps": [10], "mults": [.5]}, 5, 1.),
    ("stair", {"steps": [10], "mults": [.5]}, 10, .5),
    ("rsqrt", {"timescale": 1}, 3, .6),
    ("rsqrt", {"timescale": 1}, 20, .05),
])
def test_schedule(decay_type, extra, step, expected):          # utils_test.py:260-281
  lr_fn = u.create_learning_rate_schedule(total_steps=21, batch_size=512, base=.5,
                                          decay_type=decay_type, scale_with_batchsize=True,
                                          warmup_steps=5, cooldown_steps=5, **extra)
  assert lr_fn(step) == pytest.approx(expected, abs=1e-6)


def _tiny_params():
  import common
  from big_vision_b200.models.proj.image_text import two_towers
  model = two_towers.Model(**common.TINY)
  return model.init(0, common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE, device="cpu")


def test_make_rejects_unbuilt_configurations():
  P = _tiny_params()
  with pytest.raises(NotImplementedError):
    bv_optax.make(dict(optax_name="big_vision.scale_by_adafactor", optax=dict(clipping_threshold=1.0),
                       schedule={}), P, sched_kw=dict(total_steps=10, batch_size=8, data_size=100))
  with pytest.raises(NotImplementedError):
    bv_optax.make(dict(optax_name="lion", schedule={}), P, sched_kw=dict(total_steps=10))
  # BV-Adafactor on the two-tower tree: q/k/v kernels are factored per head over (d, dh), Dense
  # kernels over (in, out); biases, LayerNorm parameters and the 16x16x3xd patch kernel are not
  tx, _ = bv_optax.make(dict(optax_name="big_vision.scale_by_adafactor", schedule={}), P,
                        sched_kw=dict(total_steps=10, batch_size=8, data_size=100))
  modes = {t.name: (t.mode, t.dims) for t, *_ in tx.tensors}
  assert modes["img/Transformer/encoderblock_0/MultiHeadDotProductAttention_0/query/kernel"] == (1, (1, 64, 1, 64))   # d == dh tie: argsort order
  assert modes["img/Transformer/encoderblock_0/MlpBlock_0/Dense_0/kernel"] == (1, (1, 64, 1, 128))
  assert modes["img/embedding/kernel"][0] == 0 and modes["img/Transformer/encoderblock_0/LayerNorm_0/scale"][0] == 0
  tx, fns = bv_optax.make(dict(optax_name="scale_by_adam", optax=dict(b2=0.95), lr=1e-3, wd=1e-4,
                               grad_clip_norm=1.0, schedule=dict(decay_type="cosine", warmup_steps=2)),
                          P, sched_kw=dict(total_steps=10, batch_size=8, data_size=100))
  assert tx.b2 == 0.95 and tx.clip == 1.0 and fns[0](0) == 0.0 and fns[0](2) == pytest.approx(1.0)
  # default config: two launches -- the decayed kernels (stored first) and everything else
  assert [(r[0], r[1], r[2], r[4] > 0) for r in tx.ranges] == [(0, P.n_decay, 0, True), (P.n_decay, P.total, 0, False)]
  assert tx.n_state == P.total


def test_make_assigns_first_matching_pattern_and_merges_ranges():
  """optax.py:79-145 with utils.make_mask_trees (first match wins): frozen image tower
  (configs/proj/image_text/siglip_lit_coco.py:102 style), lr multiplier on the text head, custom
  weight-decay multipliers; adjacent parameters with the same setting share a launch."""
  P = _tiny_params()
  config = dict(optax_name="scale_by_adam", lr=1e-3, wd=1e-4, lr_mults=[("txt/head/.*", 2.0)],
                wd_mults=[(".*/kernel", 1.0), ("txt/Embed_0/embedding", 0.5)],
                schedule=[("img/.*", None), (".*", dict(decay_type="cosine", warmup_steps=2))])
  tx, fns = bv_optax.make(config, P, sched_kw=dict(total_steps=10, batch_size=8, data_size=100))
  assert len(fns) == 1
  names_of = {}
  for a in P.aliases.values():
    names_of.setdefault(a.storage, []).append(a.name)
  covered = 0
  for storage, (off, shape) in P.offsets.items():
    (r,) = [r for r in tx.ranges if r[0] <= off < r[1]]
    name = names_of.get(storage, [storage])[0]
    assert (r[2] is None) == name.startswith("img/"), name
    assert r[3] == (2.0 if name.startswith("txt/head/") else 1.0), name
    want_wd = 1e-4 if name.endswith("/kernel") else 0.5e-4 if name == "txt/Embed_0/embedding" else 0.0
    assert r[4] == pytest.approx(want_wd), name
  for a, b in zip(tx.ranges, tx.ranges[1:]):
    assert a[1] <= b[0] and (a[1] < b[0] or tuple(a[2:]) != tuple(b[2:]))   # disjoint, maximally merged
    covered += a[1] - a[0]
  assert covered + tx.ranges[-1][1] - tx.ranges[-1][0] == P.total
  assert tx.n_state == sum(r[1] - r[0] for r in tx.ranges if r[2] is not None) < P.total
  with pytest.raises(AssertionError):      # every parameter must be covered by a schedule pattern
    bv_optax.make(dict(config, schedule=[("img/.*", None)]), P, sched_kw=dict(total_steps=10))


def test_get_mixup_draws_a_at_least_one_half():
  """utils.py:1146-1150: a ~ Beta(p, p) then a = max(a, 1 - a)."""
  import numpy as np
  from big_vision_b200 import utils as u
  for seed in range(20):
    a = u.get_mixup(np.random.default_rng(seed), 0.2).a
    assert 0.5 <= a <= 1.0
    b = np.random.default_rng(seed).beta(0.2, 0.2)
    assert a == max(b, 1.0 - b)


def test_siglip_loss_fn_selection():
  """config.loss_fn as in _deprecated_contrastive.py:322-331; "softmax" is the CLIP loss of :80-101."""
  import pytest
  from big_vision_b200.trainers.proj.image_text import siglip
  assert siglip._loss_fn(None) is siglip.sigmoid_loss_fwd_bwd
  assert siglip._loss_fn({"loss_fn": "chunked_sigmoid"}) is siglip.chunked_sigmoid_loss_fwd_bwd
  assert siglip._loss_fn({"loss_fn": "softmax"}) is siglip.softmax_loss_fwd_bwd
  with pytest.raises(NotImplementedError):
    siglip._loss_fn({"loss_fn": "hinge"})
