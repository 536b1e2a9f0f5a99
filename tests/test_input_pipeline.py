"""Host-side logic of the prefetching input iterator (input_pipeline.py:250-270 semantics):
order preserved, every element delivered once, works for n_prefetch = 0, 1, 3."""
import numpy as np
import pytest

from big_vision_b200 import input_pipeline


@pytest.mark.parametrize("n_prefetch", [0, 1, 3])
def test_prefetch_iterator_preserves_order(n_prefetch):
  pulled = []

  def source():
    for i in range(7):
      pulled.append(i)
      yield {"image": np.full((2, 3), i, np.float32), "labels": np.full((2,), i, np.int32)}

  out = []
  for batch in input_pipeline.start_input_pipeline(source(), n_prefetch=n_prefetch, device="cpu"):
    out.append(int(batch["labels"][0]))
    assert float(batch["image"][0, 0]) == out[-1]
    # the source runs at most n_prefetch + 1 elements ahead of the consumer
    assert len(pulled) - len(out) <= n_prefetch + 1
  assert out == list(range(7))


def test_empty_source():
  assert list(input_pipeline.start_input_pipeline(iter(()), n_prefetch=2, device="cpu")) == []


@pytest.mark.gpu
@pytest.mark.parametrize("n_prefetch", [1, 2])
def test_device_prefetch_delivers_every_batch_intact(n_prefetch):
  """Device path: slots are recycled behind the consumer's kernels, so a slow consumer must still
  see each batch's own values (a premature overwrite would show the next batch's)."""
  import torch
  rng = np.random.default_rng(0)
  host = [{"image": rng.standard_normal((64, 32, 32, 3)).astype(np.float32),
           "labels": rng.integers(0, 100, (64, 16)).astype(np.int32)} for _ in range(6)]
  sums = []
  big = torch.randn(4096, 4096, device="cuda")
  for batch in input_pipeline.start_input_pipeline(iter(host), n_prefetch=n_prefetch):
    assert batch["image"].is_cuda and batch["labels"].is_cuda
    for _ in range(3):
      big = torch.tanh(big @ big * 1e-3)          # keep the consumer stream busy
    sums.append((batch["image"].double().sum(), batch["labels"].sum()))
  torch.cuda.synchronize()
  for (si, sl), h in zip(sums, host):
    assert abs(float(si) - float(h["image"].astype(np.float64).sum())) < 1e-6 * h["image"].size
    assert int(sl) == int(h["labels"].sum())
