"""Image-text retrieval metrics on the device.

Mirror of big_vision/evaluators/proj/image_text/image_text_retrieval.py:23-85: same function
names, arguments and returned dict.  The reference argsorts the distance matrix and checks whether
the positive is among the first k entries; `bv_retrieval_ranks` counts, for every text (column) /
image (row), how many entries sort before the positive, which is the same integer without a sort.
Recall@k is then mean(rank < k), evaluated on the host in float64 exactly as numpy's
`bool_array.mean()` does.  Ties are ordered by index (a stable argsort).
"""
import numpy as np
import torch

from big_vision_b200 import ops

RECALL_THRESHOLDS = (1, 5, 10)


def _device_inputs(dist_matrix, text_image_correspondence):
  d = dist_matrix if isinstance(dist_matrix, torch.Tensor) else torch.from_numpy(
      np.ascontiguousarray(dist_matrix, dtype=np.float32))
  d = d.to(device="cuda", dtype=torch.float32)
  corr = torch.as_tensor(np.asarray(text_image_correspondence), dtype=torch.int32)
  return d, corr


def _recalls(rank):
  r = rank.cpu().numpy()
  return {f"Recall@{k}": (r < k).mean() for k in RECALL_THRESHOLDS}


def text_to_image_retrieval_eval(dist_matrix, text_image_correspondence):
  """dist_matrix [N_IMAGES, N_TEXTS]; text j belongs to image text_image_correspondence[j]."""
  d, corr = _device_inputs(dist_matrix, text_image_correspondence)
  rank, _ = ops.retrieval_ranks(d, corr, t2i=True, i2t=False)
  return _recalls(rank)


def image_to_text_retrieval_eval(dist_matrix, text_image_correspondence):
  d, corr = _device_inputs(dist_matrix, text_image_correspondence)
  _, rank = ops.retrieval_ranks(d, corr, t2i=False, i2t=True)
  return _recalls(rank)
