"""GPU parity of the classification step (train.py update_fn: ViT with cls token / gap, MLP-Mixer;
sigmoid_xent / softmax_xent) against the float64 oracle on tiny seeded configurations."""
import numpy as np
import pytest
import torch

from oracle import bv_oracle as O

pytestmark = pytest.mark.gpu


def _randomize_zero_inits(tree, seed):
  """Zero-initialised heads / cls would make most gradients vanish: give them small values
  ('zero-init heads replaced by small normal for parity runs')."""
  rng = np.random.default_rng(seed)
  out = {}
  for k, v in tree.items():
    if not np.any(v):
      out[k] = (rng.standard_normal(v.shape) * 0.05).astype(np.float32)
    else:
      out[k] = v
  return out


def _check(model, oracle_fwd, image_shape, loss_name, num_classes, seed=0, tokens=None, **fwd_kw):
  from big_vision_b200 import train
  P = model.init(seed, image_shape, device="cuda")
  tree = _randomize_zero_inits(P.numpy_tree("f"), seed + 1)
  P.load_tree(tree)
  rng = np.random.default_rng(seed + 2)
  image = rng.uniform(-1, 1, size=image_shape).astype(np.float32)
  lshape = image_shape[0] if tokens is None else (image_shape[0], tokens)     # per-token labels without pooling
  labels = np.eye(num_classes, dtype=np.float32)[rng.integers(0, num_classes, size=lshape)]
  loss, logits = train.loss_and_grads(model, P, torch.from_numpy(image).cuda(),
                                      torch.from_numpy(labels).cuda(), loss_name, **fwd_kw)
  # oracle
  p64 = O.to_f64_tree(tree, requires_grad=True)
  ref_logits_bf16 = oracle_fwd(O.to_f64_tree(tree), torch.from_numpy(image), "bfloat16")
  ref_logits = oracle_fwd(p64, torch.from_numpy(image), "float32")
  ref_loss = getattr(O, loss_name)(ref_logits, torch.from_numpy(labels).double())
  ref_loss.backward()
  scale = float(ref_logits.abs().max())
  assert float((logits.double().cpu() - ref_logits_bf16).abs().max()) <= 2e-2 * scale
  assert float((logits.double().cpu() - ref_logits.detach()).abs().max()) <= 6e-2 * scale
  assert float(loss) == pytest.approx(float(ref_loss), rel=2e-2)
  grads = P.numpy_tree("g")
  gmax = max(float(v.grad.abs().max()) for v in p64.values() if v.grad is not None)
  bad = {}
  for k, g in grads.items():
    ref = p64[k].grad.numpy() if p64[k].grad is not None else np.zeros_like(g)
    err = float(np.abs(g.astype(np.float64) - ref).max())
    tol = 6e-2 * float(np.abs(ref).max()) + 3e-3 * gmax
    if err > tol:
      bad[k] = (err, tol)
  assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0])[:8]


@pytest.mark.parametrize("pool,posemb,rep,loss", [("tok", "learn", True, "sigmoid_xent"),
                                                  ("gap", "sincos2d", True, "softmax_xent"),
                                                  ("0", "learn", False, "sigmoid_xent"),
                                                  ("none", "learn", True, "softmax_xent")])
def test_vit_classifier_step(pool, posemb, rep, loss):
  from big_vision_b200.models import vit
  kw = dict(width=64, depth=2, mlp_dim=128, num_heads=1, patch_size=(16, 16), pool_type=pool,
            posemb=posemb, rep_size=rep)
  model = vit.Model(16, **kw)
  cfg = dict(depth=2, num_heads=1, pool_type=pool, posemb=posemb, rep_size=rep, num_classes=16)
  _check(model, lambda p, img, mm: O.vit_forward(p, img, cfg, mm), (4, 64, 48, 3), loss, 16,
         tokens=12 if pool == "none" else None)


def test_mlp_mixer_step():
  from big_vision_b200.models import mlp_mixer
  model = mlp_mixer.Model(16, patch_size=(16, 16), num_blocks=2, hidden_dim=64, tokens_mlp_dim=32,
                          channels_mlp_dim=128)
  cfg = dict(num_blocks=2, num_classes=16)
  # 48 x 64 image -> 12 tokens: exercises the token padding (12 -> 16) of the token-mixing GEMMs
  _check(model, lambda p, img, mm: O.mixer_forward(p, img, cfg, mm), (4, 48, 64, 3), "sigmoid_xent", 16)


def test_mlp_mixer_stochastic_depth_masks():
  """mlp_mixer.py:52,55,173-177: per-sample residual gates; the same 0/1 masks go to the oracle."""
  from big_vision_b200.models import mlp_mixer
  model = mlp_mixer.Model(16, patch_size=(16, 16), num_blocks=3, hidden_dim=64, tokens_mlp_dim=32,
                          channels_mlp_dim=128, stoch_depth=0.5)
  assert [model.drop_p(i) for i in range(3)] == [0.0, 0.25, 0.5]
  masks = np.array([[[1, 1, 1, 1], [1, 1, 1, 1]], [[1, 0, 1, 1], [0, 1, 1, 0]], [[0, 0, 1, 1], [1, 0, 1, 0]]],
                   dtype=np.float32)
  cfg = dict(num_blocks=3, num_classes=16)
  _check(model, lambda p, img, mm: O.mixer_forward(p, img, cfg, mm, masks=masks), (4, 64, 64, 3),
         "sigmoid_xent", 16, masks=torch.from_numpy(masks).cuda())
  # drawn masks: block 0 never drops; eval mode (train=False) ignores stoch_depth
  m = model.draw_masks(np.random.default_rng(0), 64, "cpu")
  assert m.shape == (3, 2, 64) and bool((m[0] == 1).all()) and 0.2 < float(1 - m[2].mean()) < 0.8


def test_scan_remat_encoder_equals_pyloop():
  """scan=True (models/vit.py:129-148: stacked `encoderblock` params, per-block remat with
  nothing_saveable) must give the pyloop model's loss and gradients: same kernels, the block is
  just recomputed from its saved input in the backward."""
  from big_vision_b200 import train, utils as u
  from big_vision_b200.models import vit
  kw = dict(width=64, depth=3, mlp_dim=128, num_heads=1, patch_size=(16, 16), pool_type="gap")
  shape = (4, 64, 64, 3)
  m_loop, m_scan = vit.Model(16, **kw), vit.Model(16, scan=True, **kw)
  P_loop = m_loop.init(0, shape, device="cuda")
  tree = _randomize_zero_inits(P_loop.numpy_tree("f"), 1)
  P_loop.load_tree(tree)
  nested = u.recover_tree(list(tree.keys()), list(tree.values()))
  stacked = dict(u.tree_flatten_with_names(vit.pyloop_to_scan(nested))[0])
  P_scan = m_scan.init(0, shape, device="cuda")
  assert set(P_scan.tree("f")) == set(stacked), set(P_scan.tree("f")) ^ set(stacked)
  assert tuple(P_scan.tree("f")["Transformer/encoderblock/MultiHeadDotProductAttention_0/query/kernel"].shape) \
      == (3, 64, 1, 64)
  P_scan.load_tree(stacked)
  rng = np.random.default_rng(2)
  image = torch.from_numpy(rng.uniform(-1, 1, size=shape).astype(np.float32)).cuda()
  labels = torch.from_numpy(np.eye(16, dtype=np.float32)[rng.integers(0, 16, size=4)]).cuda()
  l1, lg1 = train.loss_and_grads(m_loop, P_loop, image, labels, "softmax_xent")
  l2, lg2 = train.loss_and_grads(m_scan, P_scan, image, labels, "softmax_xent")
  assert torch.equal(lg1, lg2)
  assert float(l1) == float(l2)
  g1 = P_loop.numpy_tree("g")
  g1 = dict(u.tree_flatten_with_names(vit.pyloop_to_scan(u.recover_tree(list(g1.keys()), list(g1.values()))))[0])
  g2 = P_scan.numpy_tree("g")
  for k in g2:     # identical kernels on identical inputs; only atomics order differs (fp32 last bits)
    np.testing.assert_allclose(g2[k], g1[k], rtol=1e-4, atol=1e-6 * max(1.0, float(np.abs(g1[k]).max())))


def test_classifier_update_fn_runs_and_learns():
  from big_vision_b200 import optax as bv_optax, train
  from big_vision_b200.models import vit
  model = vit.Model(16, width=64, depth=2, mlp_dim=128, num_heads=1, patch_size=(16, 16),
                    pool_type="tok", rep_size=True)
  P = model.init(0, (8, 64, 64, 3), device="cuda")
  P.load_tree(_randomize_zero_inits(P.numpy_tree("f"), 1))
  config = dict(optax_name="scale_by_adam", optax=dict(mu_dtype="bfloat16"), lr=1e-3, wd=1e-4,
                grad_clip_norm=1.0, loss="sigmoid_xent", schedule=dict(decay_type="cosine", warmup_steps=0))
  tx, _ = bv_optax.make(config, P, sched_kw=dict(total_steps=100, batch_size=8, data_size=1000))
  state = {"params": P, "opt": tx.init(P)}
  fn = train.make_update_fn(model, tx, config)
  rng = np.random.default_rng(0)
  batch = {"image": torch.from_numpy(rng.uniform(-1, 1, (8, 64, 64, 3)).astype(np.float32)).cuda(),
           "labels": torch.from_numpy(np.eye(16, dtype=np.float32)[rng.integers(0, 16, 8)]).cuda()}
  losses = []
  for _ in range(10):
    state, m = fn(state, None, batch)
    losses.append(float(m["training_loss"]))
  assert losses[-1] < losses[0]


@pytest.mark.parametrize("shape,a", [((8, 64, 64, 3), 0.73), ((5, 1000), 0.5), ((1, 8), 0.9), ((3, 4), 1.0)])
def test_mixup_kernel_is_bit_exact(shape, a):
  """utils.py:1146-1158: a*x + (1-a)*roll(x, 1, axis=0), evaluated in fp32 with each product and
  the sum rounded separately -- identical bits to the numpy fp32 expression."""
  from big_vision_b200 import ops
  x = np.random.default_rng(7).standard_normal(shape).astype(np.float32)
  a32 = np.float32(a)
  ref = a32 * x + (np.float32(1) - a32) * np.roll(x, 1, axis=0)
  assert ref.dtype == np.float32
  got = ops.mixup(torch.from_numpy(x).cuda(), float(a32)).cpu().numpy()
  assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_update_fn_with_mixup_equals_step_on_mixed_batch():
  """The trainer draws a ~ Beta(p,p), a = max(a, 1-a), and mixes images and labels of this rank's
  shard before the step: its loss equals loss_and_grads on the host-mixed batch with the same a."""
  from big_vision_b200 import optax as bv_optax, train, utils as u
  from big_vision_b200.models import vit
  model = vit.Model(16, width=64, depth=2, mlp_dim=128, num_heads=1, patch_size=(16, 16),
                    pool_type="gap", rep_size=False)
  P = model.init(0, (8, 64, 64, 3), device="cuda")
  P.load_tree(_randomize_zero_inits(P.numpy_tree("f"), 1))
  config = dict(optax_name="scale_by_adam", optax=dict(mu_dtype="float32"), lr=0.0, wd=0.0,
                loss="softmax_xent", mixup=dict(p=0.2), schedule=dict(decay_type="cosine", warmup_steps=0))
  tx, _ = bv_optax.make(config, P, sched_kw=dict(total_steps=100, batch_size=8, data_size=1000))
  state = {"params": P, "opt": tx.init(P)}
  fn = train.make_update_fn(model, tx, config)
  rng = np.random.default_rng(3)
  image = rng.uniform(-1, 1, (8, 64, 64, 3)).astype(np.float32)
  labels = np.eye(16, dtype=np.float32)[rng.integers(0, 16, 8)]
  a = u.get_mixup(np.random.default_rng(11), 0.2).a
  assert 0.5 <= a <= 1.0
  a32 = np.float32(a)
  mix = lambda t: a32 * t + (np.float32(1) - a32) * np.roll(t, 1, axis=0)
  ref_loss, _ = train.loss_and_grads(model, P, torch.from_numpy(mix(image)).cuda(),
                                     torch.from_numpy(mix(labels)).cuda(), "softmax_xent")
  ref_loss = float(ref_loss)
  with pytest.raises(ValueError):
    fn(state, None, {"image": torch.from_numpy(image).cuda(), "labels": torch.from_numpy(labels).cuda()})
  state, m = fn(state, np.random.default_rng(11), {"image": torch.from_numpy(image).cuda(),
                                                    "labels": torch.from_numpy(labels).cuda()})
  # the loss kernels sum per-row losses in a fixed order (row_loss_ws + finishing pass), so two
  # evaluations of the same step agree to the bit; rel=1e-6 only guards against a future change of
  # that policy turning this into a flaky 1-ulp failure
  assert float(m["training_loss"]) == pytest.approx(ref_loss, rel=1e-6)


def test_scalar_losses_are_run_to_run_deterministic():
  """sigmoid_xent / softmax_xent / siglip loss scalars: per-row (per-block) partials and a
  fixed-order finishing pass instead of atomics -> identical bits on every evaluation."""
  from big_vision_b200 import ops
  g = torch.Generator().manual_seed(5)
  logits = (torch.randn(1000, 1000, generator=g) * 3).cuda()
  labels = torch.nn.functional.one_hot(torch.randint(0, 1000, (1000,), generator=g), 1000).float().cuda()
  for fn in (ops.sigmoid_xent, ops.softmax_xent):
    vals = set()
    for _ in range(8):
      loss = torch.zeros(1, device="cuda")
      fn(logits, labels, loss)
      vals.add(float(loss))
    assert len(vals) == 1, vals
  dots = torch.randn(512, 2048, generator=g).cuda()
  t, b = torch.tensor([2.3]).cuda(), torch.tensor([-10.0]).cuda()
  vals = set()
  for _ in range(8):
    sc = torch.zeros(3, device="cuda")
    ops.siglip_loss(dots, 512, t, b, 2048, sc[0:1], sc[1:2], sc[2:3])
    vals.add(tuple(sc.tolist()))
  assert len(vals) == 1, vals
