"""Host <-> device plumbing around a training step: the input side of the reference's loop
(``for images in dataloader: images = images.cuda()``, train.py:38-41 / train_Style_GAN.py:150-160) and the
``loss.item()`` read-backs it logs every step (train.py:80-93), both taken off the step's critical path.

* ``HostBatchPipeline``: batches go pinned host memory -> one of two device staging buffers on a dedicated copy stream;
  the copy of batch i+1 runs while step i computes.  The consumer's stream waits on the copy's event only.
* ``ScalarReadback``: a step's scalar results (losses) are copied device -> pinned host memory asynchronously and read
  one step late, so the host never drains the GPU queue between two steps.

Plain torch streams / events: no kernels of ours are involved, and there is no CPU fallback to speak of.
"""
from __future__ import annotations

import torch


class HostBatchPipeline:
    def __init__(self, shape, dtype=torch.float32, device=None, depth: int = 2):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise ValueError("HostBatchPipeline: the staging buffers live on a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.device)
        self.stage = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]        # copy into stage[i] finished
        self.consumed = [None] * depth                                 # last consumer of stage[i] is done with it
        self.head = self.tail = 0
        self.bytes_per_batch = self.stage[0].numel() * self.stage[0].element_size()

    def feed(self, host_batch: torch.Tensor):
        """Queue the copy of a (pinned) host batch; returns immediately."""
        if self.head - self.tail >= len(self.stage):
            raise RuntimeError("HostBatchPipeline: all staging buffers are in flight; take() one first")
        if not host_batch.is_pinned():
            raise ValueError("HostBatchPipeline: host batches must be in pinned memory for the copy to be asynchronous")
        i = self.head % len(self.stage)
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[i] is not None:
                self.copy_stream.wait_event(self.consumed[i])
            self.stage[i].copy_(host_batch, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.head += 1

    def take(self) -> torch.Tensor:
        """The oldest fed batch, on the device; the CURRENT stream waits for its copy.  Call ``release()`` once the work that
        reads it has been queued."""
        if self.tail >= self.head:
            raise RuntimeError("HostBatchPipeline: nothing was fed")
        i = self.tail % len(self.stage)
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        self._taken = i
        self.tail += 1
        return self.stage[i]

    def release(self):
        """Everything queued so far on the current stream is the last reader of the batch handed out by take()."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.consumed[self._taken] = ev


class ScalarReadback:
    def __init__(self, n: int = 1, depth: int = 2, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.head = self.tail = 0
        self.bytes_per_step = 4 * n

    def push(self, values: torch.Tensor):
        """Queue the device -> host copy of this step's scalar(s) behind the step, on the current stream."""
        if self.head - self.tail >= len(self.host):
            raise RuntimeError("ScalarReadback: pop() the oldest result first")
        i = self.head % len(self.host)
        self.host[i].copy_(values.detach().reshape(-1).float(), non_blocking=True)
        self.done[i].record(torch.cuda.current_stream(self.device))
        self.head += 1

    def pending(self) -> int:
        return self.head - self.tail

    def pop(self):
        """Block until the OLDEST pushed result is on the host and return it (a list of floats)."""
        if self.tail >= self.head:
            raise RuntimeError("ScalarReadback: nothing pending")
        i = self.tail % len(self.host)
        self.done[i].synchronize()
        self.tail += 1
        return self.host[i].tolist()
