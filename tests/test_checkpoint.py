"""Checkpoint interchange with the reference's .npz format (host logic, CPU).

Known answers from the reference's tests (utils_test.py:104-186: _traverse_with_names,
tree_flatten_with_names, recover_tree on d1/d2) and format round trips: flat 'a/b/c' keys, the three
containers `load_params` understands, ':subtree' selection, bfloat16-as-void, pyloop<->scan
stacking, old-checkpoint fixes, hi-res position-embedding resampling, merge_params' dont_load."""
import os

import numpy as np
import pytest
import torch

import common
from big_vision_b200 import utils as u
from big_vision_b200.models import common as mcommon
from big_vision_b200.models import vit


def test_tree_helpers_known_answers():
  d1 = {"w1": 1, "w2": 2, "w34": (3, 4)}
  d2 = {"conv1": {"kernel": 0, "bias": 1}, "conv2": {"kernel": 2, "bias": 3}}
  assert list(u._walk(d1)) == [("w1", 1), ("w2", 2), ("w34/0", 3), ("w34/1", 4)]
  assert list(u._walk(d2)) == [("conv1/bias", 1), ("conv1/kernel", 0),
                                              ("conv2/bias", 3), ("conv2/kernel", 2)]
  assert list(u._walk(d2, inner=True)) == [
      ("conv1/bias", 1), ("conv1/kernel", 0), ("conv1", d2["conv1"]),
      ("conv2/bias", 3), ("conv2/kernel", 2), ("conv2", d2["conv2"]), ("", d2)]
  assert u.tree_flatten_with_names(d1)[0] == [("w1", 1), ("w2", 2), ("w34/0", 3), ("w34/1", 4)]
  assert u.recover_tree(["a/b", "a/c/x", "a/c/y", "d"], [0, 1, 2, 3]) == {"a": {"b": 0, "c": {"x": 1, "y": 2}}, "d": 3}
  assert u.tree_get({"a": 1, "b": {"c": 2, "d": 3}}, "b/c") == 2
  assert u.tree_get({"a": 1, "b": {"c": 2, "d": 3}}, "b") == {"c": 2, "d": 3}
  with pytest.raises(KeyError):
    u.tree_get({"a": 1}, "z")


def _model_and_tree(seed=0):
  from big_vision_b200.models.proj.image_text import two_towers
  model = two_towers.Model(**common.TINY)
  P = model.init(seed, common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE, device="cpu")
  flat = P.numpy_tree("f")
  return model, P, u.recover_tree(list(flat.keys()), list(flat.values()))


def _assert_tree_equal(a, b):
  fa, fb = dict(u.tree_flatten_with_names(a)[0]), dict(u.tree_flatten_with_names(b)[0])
  assert fa.keys() == fb.keys()
  for k in fa:
    np.testing.assert_array_equal(np.asarray(fa[k]), np.asarray(fb[k]), err_msg=k)


def test_npz_round_trip_containers_and_subtrees(tmp_path):
  model, P, tree = _model_and_tree()
  path = os.path.join(tmp_path, "ckpt.npz")
  u.save_checkpoint_np(tree, path)
  with np.load(path) as z:                      # the on-disk format: flat names, one array each
    assert "img/Transformer/encoderblock_0/MlpBlock_0/Dense_0/kernel" in z.files
    assert z["img/embedding/kernel"].shape == (16, 16, 3, 64) and z["t"].shape == (1,)
  _assert_tree_equal(u.load_params(path), tree)
  _assert_tree_equal(u.load_params(path + ":img"), tree["img"])
  np.testing.assert_array_equal(u.load_params(path + ":txt/head/bias"), tree["txt"]["head"]["bias"])
  for wrapped in ({"params": tree, "opt": {"count": np.zeros(())}}, {"opt": {"target": tree}}):
    u.save_checkpoint_np(wrapped, path)
    _assert_tree_equal(u.load_params(path), tree)
  with pytest.raises(ValueError):
    u.load_params("no_slash.npz")
  # back into a fresh FlatParams: bit-identical flat buffer
  P2 = model.init(1, common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE, device="cpu")
  P2.load_tree(dict(u.tree_flatten_with_names(u.load_params(path))[0]))
  assert torch.equal(P.flat, P2.flat)


def test_bfloat16_saved_as_void_is_recovered():
  x = np.array([1.0, -2.5, 3.140625, 1e-3], np.float32)
  bf = (x.view(np.uint32) >> 16).astype(np.uint16)            # truncate to bf16 bit patterns
  as_void = bf.view(np.dtype("V2"))                           # what np.save makes of jnp.bfloat16
  back = u.recover_dtype(as_void)
  assert back.dtype == np.float32
  np.testing.assert_array_equal(back.view(np.uint32), bf.astype(np.uint32) << 16)
  assert u.recover_dtype(x) is x


def test_pyloop_scan_round_trip_and_shapes():
  _, _, tree = _model_and_tree()
  img = tree["img"]
  scan = vit.pyloop_to_scan(img)
  assert "encoderblock" in scan["Transformer"] and "encoderblock_0" not in scan["Transformer"]
  depth = common.TINY["image"].get("depth", None) or len([k for k in img["Transformer"] if k.startswith("encoderblock_")])
  k = scan["Transformer"]["encoderblock"]["MlpBlock_0"]["Dense_0"]["kernel"]
  assert k.shape == (depth,) + img["Transformer"]["encoderblock_0"]["MlpBlock_0"]["Dense_0"]["kernel"].shape
  np.testing.assert_array_equal(k[1], img["Transformer"]["encoderblock_1"]["MlpBlock_0"]["Dense_0"]["kernel"])
  _assert_tree_equal(vit.scan_to_pyloop(scan), img)
  assert "encoderblock_0" in img["Transformer"]              # inputs are not modified


def test_resample_posemb():
  rng = np.random.default_rng(0)
  old = rng.standard_normal((1, 14 * 14, 8)).astype(np.float32)
  assert vit.resample_posemb(old, np.zeros((1, 196, 8))) is old
  new = vit.resample_posemb(old, np.zeros((1, 24 * 24, 8)))
  assert new.shape == (1, 576, 8)
  g_old, g_new = old.reshape(14, 14, 8), new.reshape(24, 24, 8)
  for (a, b) in (((0, 0), (0, 0)), ((0, 13), (0, 23)), ((13, 0), (23, 0)), ((13, 13), (23, 23))):
    np.testing.assert_allclose(g_new[b], g_old[a], rtol=1e-6)   # order-1 zoom keeps the corners
  assert g_new.min() >= g_old.min() - 1e-6 and g_new.max() <= g_old.max() + 1e-6   # interpolation
  const = vit.resample_posemb(np.full((1, 49, 4), 2.5, np.float32), np.zeros((1, 100, 4)))
  np.testing.assert_allclose(const, 2.5)


def test_fix_old_checkpoints():
  pe = np.arange(1 * 5 * 2, dtype=np.float32).reshape(1, 5, 2)        # 2x2 grid + cls slot
  old = {"Transformer": {"posembed_input": {"pos_embedding": pe}, "encoder_norm": {"scale": np.ones(2)}},
         "cls": np.ones((1, 1, 2), np.float32),
         "probe": np.zeros((1, 1, 2)), "MlpBlock_0": {}, "MultiHeadDotProductAttention_0": {}, "LayerNorm_0": {}}
  new = vit.fix_old_checkpoints(old)
  np.testing.assert_array_equal(new["pos_embedding"], pe[:, 1:])
  np.testing.assert_array_equal(new["cls"], 1 + pe[:, :1])
  assert "posembed_input" not in new["Transformer"] and set(new["MAPHead_0"]) == {
      "probe", "MlpBlock_0", "MultiHeadDotProductAttention_0", "LayerNorm_0"}
  assert "posembed_input" in old["Transformer"]                   # the input tree is left alone


def test_merge_params_dont_load_and_errors():
  loaded = {"a": {"kernel": np.ones(2), "bias": np.ones(1)}, "extra": np.zeros(1)}
  inited = {"a": {"kernel": np.zeros(2), "bias": np.zeros(1)}, "head": {"kernel": np.full(3, 7.0)}}
  with pytest.raises(ValueError) as e:
    mcommon.merge_params(loaded, inited)
  assert " - head/kernel" in str(e.value) and " + extra" in str(e.value)
  out = mcommon.merge_params(loaded, inited, dont_load=("head/.*", "extra", "a/bias"))
  np.testing.assert_array_equal(out["a"]["kernel"], 1)      # taken from the checkpoint
  np.testing.assert_array_equal(out["a"]["bias"], 0)        # dont_load: keeps its init value
  np.testing.assert_array_equal(out["head"]["kernel"], 7)   # missing in the checkpoint, allowed
  assert "extra" not in out
  assert mcommon.merge_params(loaded, None) is loaded


def test_two_towers_load_single_file_with_hires_posemb(tmp_path):
  """Single two-tower .npz (img, txt, t, b), scan-stacked image tower, lower-resolution posemb:
  `load` unstacks, resamples the grid and returns a tree the model's FlatParams accepts."""
  from big_vision_b200.models.proj.image_text import two_towers
  model, P, tree = _model_and_tree(seed=3)
  ckpt = dict(tree)
  ckpt["img"] = vit.pyloop_to_scan(tree["img"])
  n_new = tree["img"]["pos_embedding"].shape[1]
  gs = int(np.sqrt(n_new))
  assert gs * gs == n_new and gs > 2
  small = np.random.default_rng(1).standard_normal((1, (gs - 2) ** 2, tree["img"]["pos_embedding"].shape[2]))
  ckpt["img"]["pos_embedding"] = small.astype(np.float32)
  path = os.path.join(tmp_path, "siglip.npz")
  u.save_checkpoint_np({"params": ckpt}, path)
  _, P_init, init_tree = _model_and_tree(seed=4)
  cfg = dict(common.TINY)
  restored = two_towers.load(init_tree, path, cfg)
  np.testing.assert_array_equal(restored["t"], tree["t"])
  np.testing.assert_array_equal(restored["b"], tree["b"])
  np.testing.assert_array_equal(restored["txt"]["head"]["kernel"], tree["txt"]["head"]["kernel"])
  np.testing.assert_array_equal(restored["img"]["Transformer"]["encoderblock_1"]["LayerNorm_0"]["scale"],
                                tree["img"]["Transformer"]["encoderblock_1"]["LayerNorm_0"]["scale"])
  np.testing.assert_allclose(restored["img"]["pos_embedding"],
                             vit.resample_posemb(small.astype(np.float32), tree["img"]["pos_embedding"]))
  P_init.load_tree(dict(u.tree_flatten_with_names(restored)[0]))
  assert torch.equal(P_init.f("txt/head/kernel"), P.f("txt/head/kernel"))
