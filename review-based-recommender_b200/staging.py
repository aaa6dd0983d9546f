"""Staged input pipeline: what reaches the GPU per step, and how.

The reference's loop builds int64 LongTensors and bool masks in collate_fn and moves SEVEN tensors to the device one by one
every step (trainer/train_deepconn_pp.py:153-159, 281-292): 36.9 MB for DeepCoNN at B=4096, of which half is the upper 32 bits
of token ids < 2^31 and 4 MB are masks that equal `ids != 0` (utils.py:30-42).  At multi-million samples/s that traffic —
through one host's PCIe root for 8 GPUs — is the end-to-end limit.

`StagedInputs` packs a step's inputs into ONE pinned host arena and uploads it with ONE cudaMemcpyAsync into a device arena
whose typed views are the model's inputs:
  * token-id tensors travel as int32, or as uint16 when the vocabulary has at most 65536 rows (the kernels take any of the
    three widths: RBR_IDS_I32 / RBR_IDS_U16);
  * the masks are not sent: the kernels derive `ids != 0` on the fly (RBR_MASK_FROM_IDS) — pass `derive_masks=False` to keep
    sending masks that differ from that rule;
  * everything else (entity ids, review ids, ratings) travels as it is.
DeepCoNN at B=4096: 16.5 MB per step (8.3 MB at vocab 50 000: uint16) instead of 36.9 MB.

    staged = StagedInputs.for_model(model, example_batch, example_ratings)
    host = staged.pack(batch, ratings)                  # in the DataLoader's collate / pin thread
    staged.upload(host, stream=copy_stream)             # one H2D copy
    loss = model.forward_loss(*staged.batch, staged.ratings)[0]

`graphs.GraphedTrainStep(..., staged=True)` uses the device arena's views as the captured graph's static inputs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

_ALIGN = 256


class StagedInputs:
    def __init__(self, example_batch: Sequence[torch.Tensor], example_ratings: torch.Tensor, device,
                 token_idx: Sequence[int] = (0, 1), mask_idx: Sequence[int] = (2, 3), derive_masks: bool = True,
                 vocab: Optional[int] = None):
        self.device = torch.device(device)
        self.token_idx, self.mask_idx = tuple(token_idx), tuple(mask_idx)
        self.derive_masks = derive_masks
        # wire type of the token ids: the narrowest the vocabulary allows (ids outside it would alias under narrowing: `pack`
        # checks the range)
        self.vocab = vocab
        self.token_dtype = torch.uint16 if (vocab is not None and vocab <= 65536) else torch.int32
        self.n_inputs = len(example_batch)
        # slot per input: (offset, shape, wire dtype) or None when the input is not sent
        self.slots: List[Optional[Tuple[int, torch.Size, torch.dtype]]] = []
        off = 0
        for i, t in enumerate(list(example_batch) + [example_ratings]):
            if i in self.mask_idx and i < self.n_inputs and (derive_masks or t is None):
                self.slots.append(None)
                continue
            dt = self.token_dtype if (i in self.token_idx and i < self.n_inputs) else (torch.uint8 if t.dtype == torch.bool else t.dtype)
            nbytes = t.numel() * torch.empty((), dtype=dt).element_size()
            self.slots.append((off, t.shape, dt))
            off += (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        self.nbytes = off
        self.dev = torch.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        views = [self._view(self.dev, s) for s in self.slots]
        self.batch = views[:-1]                  # model inputs: int32 / uint16 token ids, None masks, the rest unchanged
        self.ratings = views[-1]

    @staticmethod
    def for_model(model, example_batch, example_ratings, device=None, derive_masks: bool = True) -> "StagedInputs":
        spec = getattr(model, "staging_spec", None) or dict(tokens=(0, 1), masks=(2, 3))
        dev = device or next(model.parameters()).device
        vocab = spec.get("vocab")
        if vocab is None:
            we = getattr(model, "word_embeddings", None)
            emb = getattr(we, "embedding", None)
            vocab = int(emb.weight.shape[0]) if emb is not None else None
        return StagedInputs(example_batch, example_ratings, dev, spec["tokens"], spec["masks"], derive_masks, vocab)

    @staticmethod
    def _view(buf: torch.Tensor, slot):
        if slot is None:
            return None
        off, shape, dt = slot
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dt).element_size()
        return buf[off:off + nbytes].view(dt).view(shape)

    def new_host_buffer(self) -> torch.Tensor:
        return torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()

    def pack(self, batch: Sequence[torch.Tensor], ratings: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host side (collate / pin thread): write one step's inputs into a pinned arena in wire format."""
        out = out if out is not None else self.new_host_buffer()
        for slot, t in zip(self.slots, list(batch) + [ratings]):
            if slot is None:
                continue
            dst = self._view(out, slot)
            if t.dtype == torch.bool:
                t = t.view(torch.uint8)
            if dst.dtype == torch.uint16 and t.dtype != torch.uint16 and t.numel() and (int(t.min()) < 0 or int(t.max()) > 65535):
                raise ValueError("rbr_b200: token id outside [0, 65535] cannot travel as uint16 (vocabulary mismatch?)")
            dst.copy_(t)                        # int64 → int32 / uint16 narrowing for the token tensors happens here
        return out

    def upload(self, host: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """ONE cudaMemcpyAsync of the whole step (pinned host arena → device arena)."""
        if stream is None:
            self.dev.copy_(host, non_blocking=True)
        else:
            with torch.cuda.stream(stream):
                self.dev.copy_(host, non_blocking=True)

    def load_device(self, batch: Sequence[torch.Tensor], ratings: torch.Tensor) -> None:
        """Fill the device arena from tensors that already live on a device (or unpinned host memory)."""
        for slot, v, t in zip(self.slots, list(self.batch) + [self.ratings], list(batch) + [ratings]):
            if slot is None:
                continue
            if t.dtype == torch.bool:
                t = t.view(torch.uint8)
            if v.dtype == torch.uint16 and t.dtype != torch.uint16:
                t = t.to(torch.int32)            # (device-side int64 → uint16 goes through int32: both conversions exist everywhere)
            v.copy_(t, non_blocking=True)

    @property
    def h2d_bytes(self) -> int:
        return self.nbytes
