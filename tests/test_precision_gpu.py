"""Parity at a BASELINE config's REAL shapes through the whole model (SURVEY.md 8d config 2 / the
image tower of config 4): ViT-B/16, width 768, 12 heads, 12 blocks, 196 (+cls) tokens, batch 8, fp32
parameters, against (a) the bf16-emulating oracle (same rounding points as the CUDA path: tight) and
(b) the plain fp64 oracle (what the bf16 compute dtype costs).  The measured errors are written to
gpurun_out/r02_precision.json; DESIGN.md section 4 quotes them.

north_star asks for 1e-3 relative on logits and gradients against the reference's JAX forward/backward.
With bf16 matmul operands (`dtype_mm="bfloat16"`, what BASELINE.json's configs name) that bar is out of
reach for ANY implementation -- one bf16 rounding is 2^-9 = 2e-3 -- which is why (a) and (b) are
reported separately: (a) bounds implementation error, (b) is the precision of the dtype."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bv_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(got, ref):
  got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
  return float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30))


def _rel_l2(got, ref):
  got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
  return float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30))


@pytest.mark.parametrize("pool", ["tok", "map"])
def test_vit_b16_real_shape_logits_and_gradients(pool):
  from big_vision_b200 import train
  from big_vision_b200.models import vit
  n, C = 8, 1000
  shape = (n, 224, 224, 3)
  model = vit.Model(C, variant="B/16", rep_size=(pool == "tok"), pool_type=pool)
  P = model.init(0, shape, device="cuda")
  rng = np.random.default_rng(1)
  tree = P.numpy_tree("f")
  for k, v in tree.items():                      # zero-initialised head / cls: small values instead
    if not np.any(v):
      tree[k] = (rng.standard_normal(v.shape) * 0.02).astype(np.float32)
  P.load_tree(tree)
  image = rng.uniform(-1, 1, size=shape).astype(np.float32)
  labels = np.eye(C, dtype=np.float32)[rng.integers(0, C, size=n)]
  loss, logits = train.loss_and_grads(model, P, torch.from_numpy(image).cuda(), torch.from_numpy(labels).cuda(),
                                      "sigmoid_xent")
  cfg = dict(depth=12, num_heads=12, pool_type=pool, posemb="learn", rep_size=(pool == "tok"), num_classes=C)
  with torch.no_grad():
    ref16 = O.vit_forward(O.to_f64_tree(tree), torch.from_numpy(image), cfg, "bfloat16").numpy()
  p64 = O.to_f64_tree(tree, requires_grad=True)
  ref64 = O.vit_forward(p64, torch.from_numpy(image), cfg, "float32")
  ref_loss = O.sigmoid_xent(ref64, torch.from_numpy(labels).double())
  ref_loss.backward()
  got = logits.double().cpu().numpy()
  res = {"logits_vs_bf16_oracle_max": _rel(got, ref16), "logits_vs_fp64_oracle_max": _rel(got, ref64.detach().numpy()),
         "logits_vs_fp64_oracle_l2": _rel_l2(got, ref64.detach().numpy()),
         "loss_rel": abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))}
  grads = P.numpy_tree("g")
  worst, l2s = ("", 0.0), []
  gmax = max(float(v.grad.abs().max()) for v in p64.values() if v.grad is not None)
  for k, g in grads.items():
    ref = p64[k].grad.numpy() if p64[k].grad is not None else np.zeros_like(g)
    e = float(np.abs(g - ref).max() / (np.abs(ref).max() + 1e-3 * gmax))
    l2s.append(_rel_l2(g, ref) if np.abs(ref).max() > 1e-3 * gmax else 0.0)
    if e > worst[1]:
      worst = (k, e)
  res.update(grad_worst_tensor=worst[0], grad_worst_rel_max=worst[1], grad_median_rel_l2=float(np.median(l2s)),
             grad_max_rel_l2=float(np.max(l2s)))
  os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
  path = os.path.join(ROOT, "gpurun_out", "r02_precision.json")
  allres = json.load(open(path)) if os.path.exists(path) else {}
  allres[f"vit_b16_{pool}_n8"] = res
  json.dump(allres, open(path, "w"), indent=1)
  print(res)
  # (a) implementation error against the oracle with the SAME rounding points
  assert res["logits_vs_bf16_oracle_max"] <= 1.5e-2, res
  # (b) cost of the bf16 compute dtype against the fp64 model
  assert res["logits_vs_fp64_oracle_max"] <= 3e-2 and res["loss_rel"] <= 2e-3, res
  assert res["grad_worst_rel_max"] <= 1.2e-1 and res["grad_median_rel_l2"] <= 3e-2, res
