"""Fused global-norm clip + Adam on flat arenas, the step after the hot path.

The reference's loop runs `nn.utils.clip_grad_norm_(self.model.parameters(), 5.0)` and `torch.optim.Adam.step()` over every
parameter tensor including the dense [V, E] word-table gradient (trainer/train_deepconn_pp.py:135,167-168).  Here every
parameter of the model becomes a view of ONE flat fp32 buffer with the gradient arena's slot layout (ops.GradArena), and the
update is two kernels over flat memory (csrc/optim.cu): Σg², then clip·Adam in one pass that also rewrites the bf16 shadow of
the word table — the operand staging (`rbr_table_to_bf16`) leaves the step.

    opt = FusedClipAdam(model, lr=0.002, max_grad_norm=5.0)
    for batch, ratings in loader:
        opt.zero_grad()
        loss = loss_fn(model(*batch), ratings); loss.backward()
        gnorm = opt.clip_and_step()            # replaces clip_grad_norm_(...) + optimizer.step(); returns a device scalar

`step()` alone is Adam without clipping (for loops that keep calling clip_grad_norm_ themselves).  The step counter lives on
the device, so `graphs.GraphedTrainStep(..., optimizer=opt)` captures the whole trainer step, update included.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import lib


class FusedClipAdam:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: Optional[float] = None):
        self.model = model
        self.lr, self.betas, self.eps, self.max_grad_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), max_grad_norm
        was = torch.is_grad_enabled()
        torch.set_grad_enabled(True)
        try:
            layout = ops.GradArena.for_module(model)
        finally:
            torch.set_grad_enabled(was)
        if layout is None or layout.total == 0:
            raise ValueError("FusedClipAdam: the model has no trainable parameters")
        self.layout = layout
        params = [p for _, p in model.named_parameters() if p.requires_grad]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedClipAdam needs the model on a CUDA device (move it first: the parameters are re-homed into a flat buffer)")
        self.flat_p = torch.zeros(layout.total, dtype=torch.float32, device=dev)
        seen = set()
        for prm in params:
            if id(prm) in seen:
                continue
            seen.add(id(prm))
            off, shape = layout.slots[id(prm)]
            view = self.flat_p[off:off + shape.numel()].view(shape)
            view.copy_(prm.data)
            prm.data = view                        # the parameter now lives in the flat buffer (state_dict, forward: unchanged)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        # word table: keep its bf16 shadow current from inside the update
        self._we = getattr(model, "word_embeddings", None)
        self._table_slot = None
        if self._we is not None and self._we.embedding.weight.requires_grad and self._we.embedding.weight.shape[1] % 4 == 0:
            w = self._we.embedding.weight
            self._table_slot = (layout.slots[id(w)][0], w.shape[0], w.shape[1])

    # torch.optim-like surface -------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        self.model.zero_grad(set_to_none=True)

    @property
    def param_groups(self):
        return [{"lr": self.lr, "betas": self.betas, "eps": self.eps}]

    def state_dict(self):
        return {"step": int(self.step_dev.item()), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd):
        self.step_dev.fill_(int(sd["step"]))
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps = sd["lr"], tuple(sd["betas"]), sd["eps"]

    # --------------------------------------------------------------------------------------------------------------
    def _grad_flat(self) -> torch.Tensor:
        arena = getattr(self.model, "last_arena", None)
        flat = None if arena is None else arena.flat
        if flat is not None and flat.numel() >= self.layout.total:
            lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
            ok = True
            for prm in self.model.parameters():
                if prm.requires_grad and (prm.grad is None or not (lo <= prm.grad.data_ptr() < hi)):
                    ok = False
                    break
            if ok:
                return flat[:self.layout.total]
        # generic path (a gradient produced outside the arena, e.g. a plain torch layer): gather into a flat buffer
        g = torch.zeros_like(self.flat_p)
        for prm in self.model.parameters():
            if prm.requires_grad and prm.grad is not None:
                off, shape = self.layout.slots[id(prm)]
                g[off:off + shape.numel()].view(shape).copy_(prm.grad)
        return g

    def _run(self, max_norm: float) -> torch.Tensor:
        g = self._grad_flat()
        shadow, t_off, t_rows, emb = None, 0, 0, 0
        if self._table_slot is not None and getattr(self._we, "_shadow", None) is not None:
            w = self._we.embedding.weight
            key, sh = self._we._shadow
            if key == (w._version, w.data_ptr()):            # a shadow of the CURRENT table exists: keep it current
                shadow = sh
                t_off, t_rows, emb = self._table_slot
        lib.check(lib.rbr_clip_adam_step(self.flat_p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                         self.layout.total, self.lr, self.betas[0], self.betas[1], self.eps, float(max_norm),
                                         self.sumsq.data_ptr(), self.step_dev.data_ptr(), self.grad_norm.data_ptr(), t_off, t_rows, emb,
                                         ops._p(shadow), ops._stream(True)), "rbr_clip_adam_step")
        # the conv weights changed: their packed operand copies are re-staged by the next forward (9 us); the table's shadow was
        # rewritten in place, its cache entry stays valid
        conv = getattr(getattr(self.model, "ngram", None), "conv", None)
        if conv is not None:
            conv.invalidate_operand_cache()
        return self.grad_norm

    def clip_and_step(self, max_norm: Optional[float] = None) -> torch.Tensor:
        """clip_grad_norm_(params, max_norm) + Adam.step() in two kernels.  Returns the unclipped gradient norm (device)."""
        mn = self.max_grad_norm if max_norm is None else max_norm
        return self._run(float(mn) if mn else 0.0)

    def step(self) -> torch.Tensor:
        """Adam without clipping (the gradients are used as they are)."""
        return self._run(0.0)
