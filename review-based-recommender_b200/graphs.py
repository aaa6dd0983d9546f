"""CUDA-graph capture of one training step of a model of this package (zero_grad + forward + loss + backward).

The step's host side (Python, autograd bookkeeping, ~25 kernel launches) costs ~0.8 ms against ~1.2 ms of device time for
DeepCoNN at B=4096: fine on an idle host, the bottleneck on a busy one.  `GraphedTrainStep` captures the step ONCE through
the normal `nn.Module` API and replays it: inputs are copied into static device buffers (directly from pinned host memory
if that is where they live), gradients land in the same `.grad` tensors every step (consume them — clip, optimizer step —
before the next call, exactly as the reference loop trainer/train_deepconn_pp.py:161-168 does), and the FM dropout mask
still changes every step (its seed is a device-resident counter bumped inside the graph; torch's own nn.Dropout layers use
the graph-safe Philox offset).

    step = GraphedTrainStep(model, torch.nn.MSELoss(), example_batch, example_ratings)
    for batch, ratings in loader:                        # tensors of the example's shapes / dtypes (CPU pinned or CUDA)
        loss = step(batch, ratings)                      # static tensor, valid until the next call
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0); optimizer.step()

Shapes are fixed at capture (the reference pads every batch to the same [bz, doc_len]; a ragged last batch needs its own
step object or the eager path).  Data-parallel: pass `post_backward=lambda: parallel.allreduce_gradients(model)` — the
gradient exchange (NVLS kernel and its device-side barriers, or ncclAllReduce) is captured with the step; every rank must
construct and replay its step object in lockstep.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable, example_batch: Sequence[torch.Tensor],
                 example_ratings: torch.Tensor, restage_operands: bool = True, warmup: int = 3,
                 pool=None, device: Optional[torch.device] = None, post_backward: Optional[Callable[[], None]] = None,
                 fuse_mse: bool = True, staged: bool = False, optimizer=None, max_grad_norm: Optional[float] = None,
                 host_loss: bool = False):
        dev = device or next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        self.model, self.loss_fn, self.restage = model, loss_fn, restage_operands
        self.post_backward = post_backward      # e.g. lambda: parallel.allreduce_gradients(model) — captured with the step
        # optimizer (optim.FusedClipAdam): the update — global-norm clip + Adam, step counter on the device — is captured too:
        # one replay = the reference's whole loop body, trainer/train_deepconn_pp.py:161-168
        self.optimizer, self.max_grad_norm = optimizer, max_grad_norm
        # staged=True: the static inputs are typed views of ONE device arena — token ids as int32, masks derived
        # on the device — filled by a single H2D copy per step (staging.StagedInputs; `load_packed`)
        self.staged = None
        if staged:
            from .staging import StagedInputs
            self.staged = StagedInputs.for_model(model, example_batch, example_ratings, dev)
            self.static_batch, self.static_ratings = self.staged.batch, self.staged.ratings
        else:
            self.static_batch = [None if t is None else torch.empty(t.shape, dtype=t.dtype, device=dev) for t in example_batch]
            self.static_ratings = torch.empty(example_ratings.shape, dtype=example_ratings.dtype, device=dev)
        self._fm = getattr(model, "fm", None)
        # FM dropout seed counter: attached to the model only while this object's body runs (warm-up and capture), so eager
        # calls made between replays keep drawing fresh host seeds
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev) if self._fm is not None else None
        # nn.MSELoss() (mean) is evaluated inside the head kernel's launch when the model offers forward_loss (fused K4 + loss)
        self.fused_loss = (fuse_mse and isinstance(loss_fn, torch.nn.MSELoss) and loss_fn.reduction == "mean"
                           and hasattr(model, "forward_loss"))
        self.load(example_batch, example_ratings)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        model.zero_grad(set_to_none=True)
        # host_loss=True: the step's loss is copied to pinned host memory by the graph's LAST node (`loss_host`, valid once the
        # replay has finished), so a training loop that reads every step's loss puts nothing between two graph launches —
        # a separate D2H copy (or an event wait) between them keeps the next graph from being staged while this one runs
        # (measured: 0.77 -> 0.80 ms per DeepCoNN step for the 4-byte copy alone)
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory() if host_loss else None
        with torch.cuda.graph(self.graph, pool=pool):
            self.loss = self._body()
            if self.loss_host is not None:
                self.loss_host.copy_(self.loss.detach().reshape(1), non_blocking=True)
        self.pool = self.graph.pool()

    def _body(self) -> torch.Tensor:
        if self._seed_dev is not None:
            self._seed_dev.add_(1)                       # new FM dropout mask per replay; forward and backward read the same value
            self._fm.__dict__["_rbr_seed_dev"] = self._seed_dev
        try:
            self.model.zero_grad(set_to_none=True)
            if self.restage and self.optimizer is None and hasattr(self.model, "invalidate_operand_cache"):
                self.model.invalidate_operand_cache()    # the parameters change between replays: re-stage bf16 shadow / packed weights
                # (with a captured optimizer the update itself keeps the shadow current and drops the packed conv weights)
            if self.fused_loss:
                loss, _ = self.model.forward_loss(*self.static_batch, self.static_ratings)
            else:
                out = self.model(*self.static_batch)
                pred = out[0] if isinstance(out, tuple) else out
                loss = self.loss_fn(pred, self.static_ratings)
            loss.backward()
            if self.post_backward is not None:
                self.post_backward()
            if self.optimizer is not None:
                self.grad_norm = self.optimizer.clip_and_step(self.max_grad_norm)
        finally:
            if self._fm is not None:
                self._fm.__dict__.pop("_rbr_seed_dev", None)
        return loss

    def load(self, batch: Sequence[torch.Tensor], ratings: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Copy a batch into the static input buffers (H2D straight from pinned memory, or D2D), optionally on `stream`."""
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            if self.staged is not None:
                self.staged.load_device(batch, ratings)
                return
            for dst, src in zip(self.static_batch, batch):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            self.static_ratings.copy_(ratings, non_blocking=True)

    def load_packed(self, host: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """staged=True: ONE cudaMemcpyAsync of a step packed by `self.staged.pack(batch, ratings)` into pinned memory."""
        self.staged.upload(host, stream)

    def refresh_operands(self) -> None:
        """With a captured optimizer the graph never re-casts the bf16 shadow of the word table (the update keeps it current).
        Call this after changing parameters OUTSIDE the graph (load_state_dict, manual edits) so the next replay sees them."""
        we = getattr(self.model, "word_embeddings", None)
        if we is not None and hasattr(we, "refresh_operand_cache"):
            we.refresh_operand_cache()

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss

    def __call__(self, batch: Sequence[torch.Tensor], ratings: torch.Tensor) -> torch.Tensor:
        self.load(batch, ratings)
        return self.replay()


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
