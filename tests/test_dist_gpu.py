"""GPU, 2 ranks over NCCL: the SigLIP step on 2 GPUs (batch sharded, positives on each rank's
diagonal block, all-gather / reduce-scatter / all-reduce) must reproduce the single-GPU
global-batch loss and gradients (SURVEY.md 8e invariant).  Skipped on a 1-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

import common

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret, loss_fn="sigmoid"):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, "tests"))
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  import torch.distributed as dist
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  import common as c
  from big_vision_b200.models.proj.image_text import two_towers
  from big_vision_b200.trainers.proj.image_text import siglip
  model = two_towers.Model(**c.TINY)
  P = model.init(0, c.TINY_IMAGE_SHAPE, c.TINY_TEXT_SHAPE, device="cuda")
  image, text = c.synthetic_batch(c.TINY_IMAGE_SHAPE, c.TINY_TEXT_SHAPE, c.TINY["text"]["vocab_size"])
  n = image.shape[0] // world
  img = torch.from_numpy(image[rank * n:(rank + 1) * n]).cuda()
  txt = torch.from_numpy(text[rank * n:(rank + 1) * n]).cuda()
  loss, _ = siglip.loss_and_grads(model, P, img, txt, loss_fn=loss_fn)
  torch.cuda.synchronize()
  if rank == 0:
    ret["loss"] = float(loss)
    ret["grad"] = P.grad.cpu().numpy()
  dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("loss_fn", ["sigmoid", "chunked_sigmoid"])
def test_two_rank_step_equals_single_rank_global_batch(loss_fn):
  import torch.multiprocessing as mp
  from big_vision_b200.models.proj.image_text import two_towers
  from big_vision_b200.trainers.proj.image_text import siglip
  ctx = mp.get_context("spawn")
  ret = ctx.Manager().dict()
  port = 29700 + os.getpid() % 200 + (7 if loss_fn != "sigmoid" else 0)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, ret, loss_fn)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(300)
    assert p.exitcode == 0
  model = two_towers.Model(**common.TINY)
  P = model.init(0, common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE, device="cuda")
  image, text = common.synthetic_batch(common.TINY_IMAGE_SHAPE, common.TINY_TEXT_SHAPE,
                                       common.TINY["text"]["vocab_size"])
  loss, _ = siglip.loss_and_grads(model, P, torch.from_numpy(image).cuda(), torch.from_numpy(text).cuda())
  g1 = P.grad.cpu().numpy()
  g2 = ret["grad"]
  assert ret["loss"] == pytest.approx(float(loss), rel=1e-4)
  # same kernels on the same rows; only the fp32 summation order across ranks / split-K differs
  assert np.abs(g1 - g2).max() <= 2e-2 * np.abs(g1).max()
  assert np.linalg.norm(g1 - g2) <= 1e-2 * np.linalg.norm(g1)
