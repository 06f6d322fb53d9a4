"""Whole-step CUDA-graph capture of the denoiser training step (forward + EDM_LOSS + backward + gradient
all-reduce + clip_grad_norm_ + AdamW).

The reference's step issues ~7 100 kernel launches and >= 8 host synchronisations; the eager
B200 step is CPU-launch-bound.  Every hot-path launch of this package is shape-static and sync-free (worst-case
row capacity, live counts read on the device, pointer-stable weight-prep tables), so the entire step can be
recorded once and replayed as one graph launch."""
from typing import Callable, Dict, Optional

import torch


class GraphedTrainStep:
    def __init__(self, step_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], example_batch: Dict[str, torch.Tensor],
                 warmup: int = 3):
        """step_fn(batch) runs one full training step on the given (static) input tensors and returns the loss."""
        self.step_fn = step_fn
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self.warmup = warmup
        self._stage = self._copy_stream = self._staged = self._stage_free = None

    def capture(self) -> "GraphedTrainStep":
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self.warmup):          # allocator warm-up, lazy kernel attributes, table uploads
                self.step_fn(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            self.loss = self.step_fn(self.static)
        self.graph = g
        return self

    def load(self, batch: Dict[str, torch.Tensor]) -> None:
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)

    # -- input pipeline: the next batch's host->device copy runs on a copy stream beside the current replay -------
    def prefetch(self, host_batch: Dict[str, torch.Tensor]) -> None:
        """Start the asynchronous H2D copy of the NEXT step's (pinned) host batch into a staging set."""
        if self._stage is None:
            self._stage = {k: torch.empty_like(v) for k, v in self.static.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staged, self._stage_free = torch.cuda.Event(), torch.cuda.Event()
            self._stage_free.record()
        cs = self._copy_stream
        cs.wait_event(self._stage_free)                  # the previous commit has drained the staging set
        with torch.cuda.stream(cs):
            for k, v in host_batch.items():
                self._stage[k].copy_(v, non_blocking=True)
            self._staged.record(cs)

    def commit(self) -> None:
        """Make the prefetched batch the graph's static input (device-to-device, after the H2D copy finished)."""
        main = torch.cuda.current_stream()
        main.wait_event(self._staged)
        for k, v in self._stage.items():
            self.static[k].copy_(v, non_blocking=True)
        self._stage_free.record(main)

    # -- output pipeline: the loss of step i is read on the host while step i + 1 runs ------------------------------
    def enqueue_loss_read(self) -> int:
        """After a replay: start the asynchronous D2H copy of the loss into a pinned slot (the static loss tensor is
        overwritten by the next replay, the slot is not).  Returns the handle for `result`."""
        if getattr(self, "_pin", None) is None:
            self._pin = torch.empty(4, dtype=torch.float32).pin_memory()
            self._pin_ev = [torch.cuda.Event() for _ in range(4)]
            self._pin_i = 0
        slot = self._pin_i % 4
        self._pin_i += 1
        self._pin[slot:slot + 1].copy_(self.loss.detach().reshape(1).float(), non_blocking=True)
        self._pin_ev[slot].record()
        return slot

    def result(self, handle: int) -> float:
        """The loss a previous `enqueue_loss_read` copied (waits for that copy only, not for later replays)."""
        self._pin_ev[handle].synchronize()
        return float(self._pin[handle])

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        return self.loss
