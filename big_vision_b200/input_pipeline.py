"""Host -> device input prefetch, the device-facing end of the reference's input pipeline.

Mirrors `start_input_pipeline(data, n_prefetch)` / `prefetch_iterator(it, n)`
(big_vision/input_pipeline.py:250-270, 346-349): an iterator over batches (dicts of arrays) that
keeps `n_prefetch` batches ahead of the consumer.  There the put is `jax.device_put` on a sharded
array; here each batch is copied from pinned host memory to device buffers on a side stream, so the
copy of step i+1 overlaps the kernels of step i.  Device buffers are recycled: a slot is
overwritten only after the kernels that read it were enqueued and an event recorded behind them.
"""
import collections

import numpy as np
import torch


def _as_host_tensor(x, pin):
  t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
  if pin and not t.is_pinned():
    t = t.pin_memory()
  return t


def start_input_pipeline(data, n_prefetch=1, device="cuda"):
  """Yields dicts of device tensors for the dicts of host arrays produced by `data`."""
  dev = torch.device(device)
  it = iter(data)
  if dev.type != "cuda":          # host-only use (tests of the iteration logic): plain prefetch
    queue = collections.deque()
    for elem in it:
      queue.append({k: _as_host_tensor(v, False) for k, v in elem.items()})
      if len(queue) > n_prefetch:
        yield queue.popleft()
    while queue:
      yield queue.popleft()
    return

  copy_stream = torch.cuda.Stream(device=dev)
  nslots = n_prefetch + 1
  slots = [None] * nslots                 # dict name -> device tensor
  consumed = [None] * nslots              # event: the consumer's kernels on this slot are enqueued
  queue = collections.deque()             # (slot index, ready event)

  def enqueue(i):
    try:
      elem = next(it)
    except StopIteration:
      return False
    k = i % nslots
    host = {n: _as_host_tensor(v, True) for n, v in elem.items()}
    if slots[k] is None or any(slots[k][n].shape != h.shape or slots[k][n].dtype != h.dtype
                               for n, h in host.items()):
      slots[k] = {n: torch.empty(h.shape, dtype=h.dtype, device=dev) for n, h in host.items()}
      # The caching allocator may hand back a block that main-stream kernels still queued behind the
      # host are going to touch: the first copy into a NEW slot must wait for the main stream.
      copy_stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(copy_stream):
      if consumed[k] is not None:
        copy_stream.wait_event(consumed[k])
      for n, h in host.items():
        slots[k][n].copy_(h, non_blocking=True)
      ready = torch.cuda.Event()
      ready.record(copy_stream)
    queue.append((k, ready))
    return True

  produced = 0
  for _ in range(nslots):
    if enqueue(produced):
      produced += 1
  last = None
  while queue:
    k, ready = queue.popleft()
    main = torch.cuda.current_stream(dev)
    if last is not None:                  # the previous batch's consumer kernels are enqueued by now
      ev = torch.cuda.Event()
      ev.record(main)
      consumed[last] = ev
      if enqueue(produced):
        produced += 1
    main.wait_event(ready)
    last = k
    yield slots[k]
