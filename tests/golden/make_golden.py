"""Generates tests/golden/siglip_tiny.npz with the CPU oracle (float64).

No JAX/flax is installable here or on the GPU box, so these vectors come from oracle/bv_oracle.py
(itself checked against torch's independent operators in tests/test_oracle.py); they pin the
oracle against drift and give the GPU tests a fixed target that does not depend on re-running it.
  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import bv_oracle as O  # noqa: E402
from big_vision_b200.models.proj.image_text import two_towers  # noqa: E402
import common  # noqa: E402


def main():
  model = two_towers.Model(**common.TINY)
  P = model.init(0, common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE, device="cpu")
  tree = P.numpy_tree("f")
  image, text = common.synthetic_batch(common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE,
                                       common.TINY["text"]["vocab_size"])
  out = {"image": image, "text": text}
  for k, v in tree.items():
    out["param:" + k] = v.astype(np.float32)
  cfg = common.oracle_cfg(common.TINY)
  for mm in ("float32", "bfloat16"):
    loss, grads, zimg, ztxt = O.siglip_value_and_grad(tree, image, text, cfg, mm)
    out[f"{mm}:loss"] = np.float64(loss)
    out[f"{mm}:zimg"] = zimg
    out[f"{mm}:ztxt"] = ztxt
    if mm == "float32":   # full gradients only for the high-precision model (keeps the file small)
      for k, g in grads.items():
        out[f"{mm}:grad:" + k] = g.astype(np.float32)
    else:
      out[f"{mm}:gradnorm"] = np.float64(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    print(mm, "loss", loss)
  np.savez_compressed(os.path.join(HERE, "siglip_tiny.npz"), **out)
  print("wrote", os.path.join(HERE, "siglip_tiny.npz"))


if __name__ == "__main__":
  main()
