"""Host-side operators: torch.autograd.Functions whose forward/backward call the C-ABI kernels.

PyTorch provides device memory, the current CUDA stream and autograd bookkeeping; every arithmetic step
of the hot path runs in librbr_b200.so.  Tensors handed to the library must be CUDA, contiguous, fp32 /
int64 / bool — anything else raises (there is no CPU path).
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ._lib import lib

PREC_FP32, PREC_BF16 = 0, 1
ACT_RELU, ACT_TANH = 0, 1
_PREC = {"fp32": PREC_FP32, "bf16": PREC_BF16}
# per-call `flags` of the conv entry points (include/rbr_b200.h)
CONV_TC_SINGLE_CTA, CONV_TC_PAIR_ONLY, CONV_BWD_DENSE_TC, CONV_BWD_SPARSE, IDS_I32, MASK_FROM_IDS, IDS_U16 = 1, 2, 4, 8, 16, 32, 64


_STREAM_CACHE = [0, None]


def _stream(refresh: bool = False) -> int:
    """Raw handle of torch's current CUDA stream.  torch.cuda.current_stream() costs ~10 us; autograd Functions refresh it
    once on entry (refresh=True) and the launches inside reuse the handle."""
    if refresh or _STREAM_CACHE[1] is None:
        try:
            _STREAM_CACHE[0] = torch.cuda.current_stream().cuda_stream
        except Exception as e:      # no driver / no device
            raise RuntimeError(f"rbr_b200: a CUDA device is required (the hot path has no CPU implementation): {e}") from None
        _STREAM_CACHE[1] = True
    return _STREAM_CACHE[0]


_SIDE_STREAMS: Dict[str, list] = {}


def _side_streams(device, n: int) -> list:
    """n persistent auxiliary streams on `device` (created once)."""
    pool = _SIDE_STREAMS.setdefault(str(device), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"rbr_b200: `{name}` must be a CUDA tensor (the hot path has no CPU implementation)")
    if t.dtype != dtype:
        raise TypeError(f"rbr_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ids(t: torch.Tensor, name: str) -> torch.Tensor:
    """Token ids: int64 (torch.LongTensor, what the reference's collate_fn yields), or int32 / uint16 (staged input pipeline)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"rbr_b200: `{name}` must be a CUDA tensor (the hot path has no CPU implementation)")
    if t.dtype not in (torch.int64, torch.int32, torch.uint16):
        raise TypeError(f"rbr_b200: `{name}` must be int64, int32 or uint16, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _id_width_flag(ids: torch.Tensor) -> int:
    return IDS_I32 if ids.dtype == torch.int32 else (IDS_U16 if ids.dtype == torch.uint16 else 0)


def _id_flags(ids: torch.Tensor, mask: Optional[torch.Tensor], mask_from_ids: bool) -> int:
    return _id_width_flag(ids) | (MASK_FROM_IDS if (mask is None and mask_from_ids) else 0)


def _mask_u8(mask: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    if not mask.is_cuda:
        raise RuntimeError(f"rbr_b200: `{name}` must be a CUDA tensor")
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    if mask.dtype == torch.uint8:
        return mask.contiguous()
    raise TypeError(f"rbr_b200: `{name}` must be bool, got {mask.dtype}")


# ---------------------------------------------------------------------------------------------------
# Step context: one flat fp32 gradient buffer per backward pass, with a view per parameter.
# ---------------------------------------------------------------------------------------------------
class GradArena:
    """Allocates ONE zero-filled flat buffer holding the gradient of every trainable parameter of a model
    (allocated at the first backward call of a step) and hands out per-parameter views.  The kernels
    accumulate straight into these views, autograd installs them as `.grad`, and data-parallel training
    all-reduces the flat buffer with a single NCCL call (parallel.py)."""

    def __init__(self, named_params: Sequence[Tuple[str, torch.nn.Parameter]], layout: Optional["GradArena"] = None):
        if layout is not None:                  # same model, next step: reuse the slot layout, fresh buffer
            self.slots, self.total, self.signature = layout.slots, layout.total, layout.signature
        else:
            self.slots: Dict[int, Tuple[int, torch.Size]] = {}
            off = 0
            for _, prm in named_params:
                if prm.requires_grad and id(prm) not in self.slots:
                    self.slots[id(prm)] = (off, prm.shape)
                    off += (prm.numel() + 63) // 64 * 64            # 256-byte aligned slots (float4 atomics)
            self.total = off
            self.signature = GradArena.signature_of(named_params)
        self.flat: Optional[torch.Tensor] = None
        self.device = None
        self.external: Optional[torch.Tensor] = None
        self.params = named_params
        self.no_zero: set = set()       # id(param) of slots whose producer OVERWRITES every element (dense table gradient, K2c)

    @staticmethod
    def signature_of(named_params) -> tuple:
        return tuple((id(p), p.requires_grad) for _, p in named_params)

    @staticmethod
    def for_module(module: torch.nn.Module) -> Optional["GradArena"]:
        """A fresh arena for this step (None under no_grad); the slot layout is computed once per module and reused while
        the parameter objects and their requires_grad flags are unchanged (building it walks named_parameters(): ~0.15 ms
        of host time per step).  Parameters registered after the first forward are not picked up."""
        if not torch.is_grad_enabled():
            return None
        params = module.__dict__.get("_rbr_param_list")
        if params is None:
            params = list(module.named_parameters())
            module.__dict__["_rbr_param_list"] = params
            module.__dict__["_rbr_arena_layout"] = None
        layout = module.__dict__.get("_rbr_arena_layout")
        if layout is None or layout.signature != GradArena.signature_of(params):
            layout = GradArena(params)
            module.__dict__["_rbr_arena_layout"] = layout
        arena = GradArena(params, layout=layout)
        arena.external = module.__dict__.get("_rbr_arena_buffer")
        return arena

    def _ensure(self, device):
        if self.flat is None:
            ext = self.external
            if ext is not None:          # persistent buffer (symmetric memory for the NVLS all-reduce): re-zeroed, not re-allocated
                lo, hi = ext.data_ptr(), ext.data_ptr() + ext.numel() * 4
                for name, prm in self.params:
                    g = prm.grad
                    if g is not None and lo <= g.data_ptr() < hi:
                        raise RuntimeError(
                            f"rbr_b200: `{name}.grad` still aliases the persistent (NVLS symmetric-memory) gradient arena, which "
                            "is re-zeroed by every backward: call zero_grad(set_to_none=True) before each step — gradient "
                            "accumulation over micro-batches is not supported with enable_nvls_allreduce")
                self.flat = ext[:self.total]
            else:
                self.flat = torch.empty(self.total, dtype=torch.float32, device=device)
            # zero-fill, except the slots a kernel is going to overwrite completely (the 60 MB word-table gradient)
            # (only the slot's own elements are skipped: the alignment gap behind it is zeroed like everything else, so the flat
            # buffer can be summed / all-reduced / fed to the fused optimizer as it is)
            skip = sorted((self.slots[i][0], self.slots[i][0] + self.slots[i][1].numel()) for i in self.no_zero if i in self.slots)
            pos = 0
            for lo, hi in skip + [(self.total, self.total)]:
                if lo > pos:
                    self.flat[pos:lo].zero_()
                pos = max(pos, hi)

    def view(self, prm: torch.Tensor) -> Optional[torch.Tensor]:
        slot = self.slots.get(id(prm))
        if slot is None:
            return None
        self._ensure(prm.device)
        off, shape = slot
        return self.flat[off:off + shape.numel()].view(shape)


def _grad_buf(arena: Optional[GradArena], prm: torch.Tensor, needs: bool) -> Optional[torch.Tensor]:
    if not needs:
        return None
    if arena is not None:
        v = arena.view(prm)
        if v is not None:
            return v
    return torch.zeros_like(prm)


# ---------------------------------------------------------------------------------------------------
# K1 / K1b: embedding gather
# ---------------------------------------------------------------------------------------------------
def gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    table = _req(table, torch.float32, "table")
    ids = _req(ids, torch.int64, "ids")
    out = torch.empty(*ids.shape, table.shape[1], dtype=torch.float32, device=table.device)
    lib.check(lib.rbr_gather_fwd(_p(table), table.shape[0], table.shape[1], _p(ids), ids.numel(), _p(out), _stream(True)),
              "rbr_gather_fwd")
    return out


def embedding_dense_grad(ids: torch.Tensor, grad_rows: torch.Tensor, vocab: int, padding_idx: Optional[int],
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    ids = _req(ids, torch.int64, "ids")
    grad_rows = _req(grad_rows, torch.float32, "grad_rows")
    emb = grad_rows.shape[-1]
    if out is None:
        out = torch.zeros(vocab, emb, dtype=torch.float32, device=grad_rows.device)
    n = ids.numel()
    ws_bytes = lib.rbr_embgrad_workspace_bytes(n, vocab)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=grad_rows.device)
    lib.check(lib.rbr_embgrad_scatter_add(_p(ids), _p(grad_rows), n, emb, vocab, -1 if padding_idx is None else padding_idx,
                                          _p(out), _p(ws), ws_bytes, _stream(True)), "rbr_embgrad_scatter_add")
    return out


class EmbeddingFn(torch.autograd.Function):
    """nn.Embedding forward/backward (reference models/deepconn/layers.py:15,23)."""

    @staticmethod
    def forward(ctx, table, ids, padding_idx, arena):
        _stream(refresh=True)
        ctx.save_for_backward(ids)
        ctx.vocab, ctx.padding_idx, ctx.arena, ctx.table = table.shape[0], padding_idx, arena, table
        return gather_rows(table, ids)

    @staticmethod
    def backward(ctx, grad_out):
        _stream(refresh=True)
        (ids,) = ctx.saved_tensors
        buf = _grad_buf(ctx.arena, ctx.table, True)
        embedding_dense_grad(ids, grad_out.contiguous(), ctx.vocab, ctx.padding_idx, out=buf)
        return buf, None, None, None


# ---------------------------------------------------------------------------------------------------
# K0 staging: bf16 shadow table + packed conv weights
# ---------------------------------------------------------------------------------------------------
def table_to_bf16(table: torch.Tensor) -> torch.Tensor:
    table = _req(table, torch.float32, "table")
    emb_pad = lib.rbr_emb_pad(table.shape[1])
    shadow = torch.empty(table.shape[0], emb_pad, dtype=torch.bfloat16, device=table.device)
    lib.check(lib.rbr_table_to_bf16(_p(table), table.shape[0], table.shape[1], _p(shadow), _stream(True)), "rbr_table_to_bf16")
    return shadow


def conv_pack(weight: torch.Tensor) -> torch.Tensor:
    weight = _req(weight, torch.float32, "conv weight")
    h, e, k = weight.shape
    nbytes = lib.rbr_conv_pack_bytes(e, h, k)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    lib.check(lib.rbr_conv_pack(_p(weight), e, h, k, _p(packed), _stream(True)), "rbr_conv_pack")
    return packed


# ---------------------------------------------------------------------------------------------------
# K2 / K2b: fused gather → mask → conv → activation → max-over-time, for one or several token tensors
# ("sides": DeepCoNN's user and item documents share the table and the conv, deepconn.py:43-47)
# ---------------------------------------------------------------------------------------------------
class EncodeDocsFn(torch.autograd.Function):
    """feats[s] = max_t act(conv(mask(table[ids[s]])))  for every side s; all sides share table/weights.

    Replaces WordEmbedding.forward + NgramFeat.forward (reference models/deepconn/layers.py:22-24, 123-136).
    Inputs after the fixed ones: n_conv conv weights, n_conv conv biases, then per side (ids, mask-or-None).
    """

    @staticmethod
    def forward(ctx, table, cfg, *rest):
        _stream(refresh=True)
        n_conv = cfg["n_conv"]
        weights, biases = rest[:n_conv], rest[n_conv:2 * n_conv]
        sides = rest[2 * n_conv:]
        ids_l = [_ids(t, "token ids") for t in sides[0::2]]
        mask_l = [_mask_u8(m, "token mask") for m in sides[1::2]]
        table = _req(table, torch.float32, "embedding table")
        prec = _PREC[cfg["precision"]]
        mfi = bool(cfg.get("mask_from_ids", False))
        flags_l = [cfg.get("flags", 0) | _id_flags(i, m, mfi) for i, m in zip(ids_l, mask_l)]
        act, pads = cfg["act"], cfg["pads"]
        vocab, emb = table.shape
        shadow = cfg["shadow_fn"]() if prec == PREC_BF16 else None
        packed = [cfg["pack_fn"](i) for i in range(n_conv)]
        h_total = sum(w.shape[0] for w in weights)
        feats, argmaxes, scratch = [], [], []
        for ids, mask in zip(ids_l, mask_l):             # every buffer is allocated on the calling stream, before the fork
            doc_len = ids.shape[-1]
            n_docs = ids.numel() // doc_len
            if mask is not None and mask.numel() != ids.numel():
                raise ValueError("rbr_b200: mask shape does not match token ids")
            feats.append(torch.empty(n_docs, h_total, dtype=torch.float32, device=table.device))
            argmaxes.append(torch.empty(n_docs, h_total, dtype=torch.int32, device=table.device))
            ws_bytes = (max(lib.rbr_conv_fwd_workspace_bytes2(n_docs, doc_len, w.shape[2], pad) for w, pad in zip(weights, pads))
                        if prec == PREC_BF16 else 0)
            scratch.append((torch.empty(ws_bytes, dtype=torch.uint8, device=table.device) if ws_bytes else None, ws_bytes))
        # The sides are independent: side s > 0 runs on its own stream, so its pre-pass kernels (document selection, row-index
        # table: a few small launches) overlap the previous side's tensor-core kernel, which leaves room for them on every SM.
        main = torch.cuda.current_stream()
        side_streams = _side_streams(table.device, max(0, len(ids_l) - 1))
        ev_fork = main.record_event() if len(ids_l) > 1 else None
        for s, (ids, mask, fl) in enumerate(zip(ids_l, mask_l, flags_l)):
            doc_len = ids.shape[-1]
            n_docs = ids.numel() // doc_len
            feat, amax, (ws, ws_bytes) = feats[s], argmaxes[s], scratch[s]
            stream_s = main if s == 0 else side_streams[s - 1]
            if s > 0:
                stream_s.wait_event(ev_fork)
            with torch.cuda.stream(stream_s):
                sh = _stream(refresh=True)
                col = 0
                for w, b, pk, pad in zip(weights, biases, packed, pads):
                    h, _, k = w.shape
                    lib.check(lib.rbr_conv_act_maxpool_fwd(
                        prec, act, _p(table), _p(shadow), vocab, emb, _p(ids), _p(mask), None, 0, n_docs, doc_len, _p(pk),
                        _p(_req(b, torch.float32, "conv bias")), h, k, pad, feat.data_ptr() + 4 * col, amax.data_ptr() + 4 * col,
                        None, h_total, _p(ws), ws_bytes, fl, sh), "rbr_conv_act_maxpool_fwd")
                    col += h
        for s in range(1, len(ids_l)):
            main.wait_event(side_streams[s - 1].record_event())
        _stream(refresh=True)
        del scratch
        ctx.cfg, ctx.n_conv, ctx.n_sides = cfg, n_conv, len(ids_l)
        ctx.table, ctx.weights, ctx.biases = table, weights, biases
        ctx.shadow, ctx.packed = shadow, packed
        ctx.save_for_backward(*ids_l, *[m for m in mask_l if m is not None], *feats, *argmaxes)
        ctx.mask_present = [m is not None for m in mask_l]
        ctx.flags_l = flags_l
        ctx.mark_non_differentiable(*argmaxes)
        return tuple(feats) + tuple(argmaxes)          # [n_docs, H] features per side, then the int32 arg-max positions

    @staticmethod
    def backward(ctx, *out_grads):
        _stream(refresh=True)
        cfg, n_conv, ns = ctx.cfg, ctx.n_conv, ctx.n_sides
        feat_grads = out_grads[:ns]
        saved = list(ctx.saved_tensors)
        ids_l = saved[:ns]
        n_masks = sum(ctx.mask_present)
        masks_present = saved[ns:ns + n_masks]
        feats = saved[ns + n_masks:ns + n_masks + ns]
        argmaxes = saved[ns + n_masks + ns:]
        mask_l, mi = [], 0
        for present in ctx.mask_present:
            mask_l.append(masks_present[mi] if present else None)
            mi += int(present)
        table = ctx.table
        arena: Optional[GradArena] = cfg.get("arena")
        prec = _PREC[cfg["precision"]]
        vocab, emb = table.shape
        need_table = ctx.needs_input_grad[0]
        g_table = _grad_buf(arena, cfg["table_param"], need_table)
        g_w = [_grad_buf(arena, cfg["weight_params"][i], True) for i in range(n_conv)]
        g_b = [_grad_buf(arena, cfg["bias_params"][i], True) for i in range(n_conv)]
        h_total = feats[0].shape[1]
        shapes = [tuple(w.shape) for w in ctx.weights]                   # (H, E, k) per conv
        cols = [sum(sh[0] for sh in shapes[:i]) for i in range(n_conv)]
        flags0 = cfg.get("flags", 0)
        # Formulation per conv: bf16 → the dense tensor-core backward over the coefficient matrix (K2c) when the shape allows,
        # else (fp32 precision, CONV_BWD_SPARSE, unsupported shapes) the arg-max-sparse CUDA-core kernels (K2b).
        dense = [prec == PREC_BF16 and not (flags0 & CONV_BWD_SPARSE) and ctx.shadow is not None and cfg.get("dense_bwd", True)
                 and bool(lib.rbr_conv_bwd_cmat_supported(vocab, emb, sh[0], sh[2])) for sh in shapes]
        if flags0 & CONV_BWD_DENSE_TC and not all(dense):
            raise RuntimeError("rbr_b200: CONV_BWD_DENSE_TC requested but the dense tensor-core backward does not take this shape / precision")
        cm_ws = {i: _cmat_workspace(table, i, vocab, emb, shapes[i][0], shapes[i][2]) for i in range(n_conv) if dense[i]}
        # The table gradient of every side is finished first (it is by far the largest gradient: data-parallel training starts
        # its all-reduce from the `table_ready` hook while the weight-gradient kernels still run), then the weight part.
        hook = cfg.get("table_ready") if need_table else None
        # The document sides are independent (they only meet in += accumulations, all atomic): side s > 0 runs on its own
        # stream, so the small kernels of one side overlap the other side's work.
        main = torch.cuda.current_stream()
        live = [s for s in range(ns) if feat_grads[s] is not None]
        side_streams = _side_streams(table.device, max(0, len(live) - 1))
        fgs = {s: feat_grads[s].contiguous() for s in live}

        def per_side(fn):
            ev_fork = main.record_event() if len(live) > 1 else None
            for rank_s, s in enumerate(live):
                stream_s = main if rank_s == 0 else side_streams[rank_s - 1]
                if rank_s > 0:
                    stream_s.wait_event(ev_fork)
                with torch.cuda.stream(stream_s):
                    fn(s, _stream(refresh=True))
            for rank_s in range(1, len(live)):
                main.wait_event(side_streams[rank_s - 1].record_event())
            _stream(refresh=True)

        def sparse_part(do_table, do_weight):
            def run(s, sh):
                ids, mask = ids_l[s], mask_l[s]
                doc_len = ids.shape[-1]
                n_docs = ids.numel() // doc_len
                for i in range(n_conv):
                    if dense[i]:
                        continue
                    h, _, k = shapes[i]
                    gt = g_table if do_table else None
                    if gt is None and not do_weight:
                        continue
                    ws_bytes = lib.rbr_conv_bwd_workspace_bytes(n_docs, h, k, emb, vocab)
                    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=table.device)
                    lib.check(lib.rbr_conv_act_maxpool_bwd(
                        prec, cfg["act"], _p(table), _p(ctx.shadow), vocab, emb, _p(ids), _p(mask), None, 0, n_docs, doc_len,
                        _p(ctx.packed[i]), h, k, cfg["pads"][i], feats[s].data_ptr() + 4 * cols[i],
                        argmaxes[s].data_ptr() + 4 * cols[i], fgs[s].data_ptr() + 4 * cols[i], None, h_total, cfg["padding_idx"],
                        _p(g_w[i]) if do_weight else None, _p(g_b[i]) if do_weight else None, _p(gt), None, _p(ws),
                        ws_bytes, ctx.flags_l[s], sh), "rbr_conv_act_maxpool_bwd")
            return run

        def dense_scatter_chunk(i, c):
            """Block c of conv i's coefficient matrix, one document side (run per side, possibly on two streams)."""
            ws = cm_ws[i]
            h, _, k = shapes[i]

            def run(s, sh):
                ids, mask = ids_l[s], mask_l[s]
                doc_len = ids.shape[-1]
                n_docs = ids.numel() // doc_len
                lib.check(lib.rbr_conv_bwd_cmat_scatter(
                    _p(ids), _p(mask), n_docs, doc_len, vocab, emb, h, k, cfg["pads"][i], cfg["act"],
                    feats[s].data_ptr() + 4 * cols[i], argmaxes[s].data_ptr() + 4 * cols[i], fgs[s].data_ptr() + 4 * cols[i],
                    h_total, _p(g_b[i]), c, _p(ws), ws.numel(), ctx.flags_l[s], sh), "rbr_conv_bwd_cmat_scatter")
            return run

        def dense_accumulate():
            """Per conv and filter block: zero-fill the block (it then sits in L2), scatter every side into it, split it into the
            bf16 hi|lo operand while it is still there."""
            for i, ws in cm_ws.items():
                h, _, k = shapes[i]
                for c in range(lib.rbr_conv_bwd_cmat_chunks(vocab, emb, h, k)):
                    lib.check(lib.rbr_conv_bwd_cmat_begin(c, vocab, emb, h, k, _p(ws), ws.numel(), _stream()), "rbr_conv_bwd_cmat_begin")
                    per_side(dense_scatter_chunk(i, c))
                    lib.check(lib.rbr_conv_bwd_cmat_finish(1, c, 0, 0, None, _p(ctx.packed[i]), vocab, emb, h, k, cfg["padding_idx"], None, None,
                                                           _p(ws), ws.numel(), _stream()), "rbr_conv_bwd_cmat_finish")

        # the arena left the table slot un-zeroed because this backward writes every element of it (NgramFeat.encode decided)
        overwrite = bool(cfg.get("table_overwrite")) and need_table and arena is not None and id(cfg["table_param"]) in arena.no_zero
        if overwrite and not (all(dense) and n_conv == 1 and live):
            g_table.zero_()                 # nothing will overwrite it after all
            overwrite = False

        def dense_finish(what, rows=(0, 0)):
            if overwrite and (what & 2):
                what |= 8
            for i, ws in cm_ws.items():
                h, _, k = shapes[i]
                lib.check(lib.rbr_conv_bwd_cmat_finish(what, -1, rows[0], rows[1], _p(ctx.shadow), _p(ctx.packed[i]), vocab, emb, h, k,
                                                       cfg["padding_idx"], _p(g_table), _p(g_w[i]), _p(ws), ws.numel(), _stream()),
                          "rbr_conv_bwd_cmat_finish")

        any_sparse = not all(dense)
        if live:
            two_pass = hook is not None
            if any_sparse:
                per_side(sparse_part(True, not two_pass) if need_table else sparse_part(False, True))
            if cm_ws:
                dense_accumulate()
            # Data-parallel: the word-table gradient (~90 % of the exchanged bytes) CAN be produced in row slices, the hook starting
            # the all-reduce of each slice on its own stream while the next slice's GEMM runs (only when the dense path is the one
            # and only writer of the gradient).  Measured on 2 x B200 (bench.py, DeepCoNN): exposed exchange 0.164 ms with one
            # slice, 0.177 with 4, 0.230 with 8 — the table GEMM is only ~48 us long and every slice pays two cross-GPU barriers,
            # so the default stays ONE slice; RBR_TABLE_GRAD_SLICES is kept for experiments with larger vocabularies.
            n_slices = 1
            if getattr(hook, "accepts_slices", False) and cm_ws and need_table and not any_sparse and len(cm_ws) == 1:
                n_slices = max(1, min(int(os.environ.get("RBR_TABLE_GRAD_SLICES", "1")), vocab // 1024))
            if cm_ws and need_table:
                if n_slices == 1:
                    dense_finish(2)
                else:
                    step = (vocab + n_slices - 1) // n_slices
                    step = (step + 127) // 128 * 128
                    for lo in range(0, vocab, step):
                        hi = min(vocab, lo + step)
                        dense_finish(2, (lo, hi))
                        hook(g_table[lo:hi])
            if hook is not None and n_slices == 1:
                hook(g_table)
            if any_sparse and two_pass:
                per_side(sparse_part(False, True))
            if cm_ws:
                dense_finish(4)
        return (g_table, None, *g_w, *g_b, *([None] * (2 * ns)))


_CMAT_WS: Dict[tuple, torch.Tensor] = {}


def _cmat_workspace(table: torch.Tensor, conv_index: int, vocab: int, emb: int, filters: int, ksize: int) -> torch.Tensor:
    """Persistent, self-cleaning workspace of the dense tensor-core backward (K2c) for one (table, conv): zero-filled ONCE here;
    rbr_conv_bwd_cmat_finish leaves it zeroed again, so steps — eager or CUDA-graph replays — never memset its 130-200 MB."""
    key = (table.device.index, table.data_ptr(), conv_index, vocab, emb, filters, ksize)
    ws = _CMAT_WS.get(key)
    if ws is None:
        for k in [k for k in _CMAT_WS if k[:2] == key[:2] and k[2] == conv_index]:
            del _CMAT_WS[k]                        # same table storage, new shape: the old buffer is dead
        nbytes = lib.rbr_conv_bwd_cmat_workspace_bytes(vocab, emb, filters, ksize)
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("rbr_b200: run one eager step before capturing a CUDA graph (the dense backward's workspace is "
                               "allocated and zero-filled on first use)")
        while len(_CMAT_WS) >= 8:                  # a handful of (model, conv) pairs per process; drop the oldest beyond that
            del _CMAT_WS[next(iter(_CMAT_WS))]
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=table.device)
        _CMAT_WS[key] = ws
    return ws


class HierPoolFn(torch.autograd.Function):
    """pooled[s] = max_t mean_{j<k} mask(table[ids[s]])[t+j]  per document and embedding channel, for every side s —
    HierPooling.forward (reference models/deepconn/layers.py:81-98) on the masked embeddings (layers.py:131-133).
    All sides go through ONE autograd node: they share the table, whose gradient buffer is accumulated into by every side and
    handed to autograd once.  Inputs after (table, cfg): per side (ids, mask-or-None).  → [n_docs, E] per side."""

    @staticmethod
    def forward(ctx, table, cfg, *flat):
        _stream(refresh=True)
        table = _req(table, torch.float32, "embedding table")
        vocab, emb = table.shape
        ids_l = [_ids(t, "token ids") for t in flat[0::2]]
        mask_l = [_mask_u8(m, "token mask") for m in flat[1::2]]
        flags_l = [_id_flags(i, m, bool(cfg.get("mask_from_ids", False))) for i, m in zip(ids_l, mask_l)]
        outs, amaxes = [], []
        for ids, mask, fl in zip(ids_l, mask_l, flags_l):
            doc_len = ids.shape[-1]
            n_docs = ids.numel() // doc_len
            pooled = torch.empty(n_docs, emb, dtype=torch.float32, device=table.device)
            amax = torch.empty(n_docs, emb, dtype=torch.int32, device=table.device)
            lib.check(lib.rbr_hier_pool_fwd(_p(table), vocab, emb, _p(ids), _p(mask), n_docs, doc_len, cfg["ksize"], _p(pooled), _p(amax),
                                            fl, _stream()), "rbr_hier_pool_fwd")
            outs.append(pooled)
            amaxes.append(amax)
        ctx.save_for_backward(*ids_l, *amaxes, *[m for m in mask_l if m is not None])
        ctx.cfg, ctx.flags_l, ctx.mask_present, ctx.shape = cfg, flags_l, [m is not None for m in mask_l], (vocab, emb)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        _stream(refresh=True)
        cfg = ctx.cfg
        ns = len(ctx.flags_l)
        vocab, emb = ctx.shape
        if not ctx.needs_input_grad[0]:
            return (None,) * (2 + 2 * ns)
        saved = ctx.saved_tensors
        ids_l, amaxes, masks = saved[:ns], saved[ns:2 * ns], list(saved[2 * ns:])
        g_table = _grad_buf(cfg.get("arena"), cfg["table_param"], True)
        for s in range(ns):
            mask = masks.pop(0) if ctx.mask_present[s] else None
            if grads[s] is None:
                continue
            ids = ids_l[s]
            doc_len = ids.shape[-1]
            lib.check(lib.rbr_hier_pool_bwd(_p(ids), _p(mask), ids.numel() // doc_len, doc_len, vocab, emb, cfg["ksize"], cfg["padding_idx"],
                                            _p(amaxes[s]), _p(grads[s].contiguous()), _p(g_table), ctx.flags_l[s], _stream()),
                      "rbr_hier_pool_bwd")
        return (g_table, None, *([None] * (2 * ns)))


class MaskedAvgPoolFn(torch.autograd.Function):
    """out[s][n, :] = sum_t mask * table[ids[s][n, t]] / (sum_t mask + 1e-8) for every side s (K9) — the SimpleSiamese encoder
    (reference models/simple_siamese/layers.py:90-110 on the gathered embeddings).  One autograd node for all sides (shared
    table gradient).  Inputs after (table, cfg): per side (ids [n_docs, T], mask-or-None)."""

    @staticmethod
    def forward(ctx, table, cfg, *flat):
        _stream(refresh=True)
        table = _req(table, torch.float32, "embedding table")
        vocab, emb = table.shape
        ids_l = [_ids(t, "token ids") for t in flat[0::2]]
        mask_l = [_mask_u8(m, "token mask") for m in flat[1::2]]
        flags_l = [_id_flags(i, m, bool(cfg.get("mask_from_ids", False))) for i, m in zip(ids_l, mask_l)]
        outs = []
        for ids, mask, fl in zip(ids_l, mask_l, flags_l):
            doc_len = ids.shape[-1]
            n_docs = ids.numel() // doc_len
            out = torch.empty(n_docs, emb, dtype=torch.float32, device=table.device)
            lib.check(lib.rbr_masked_avg_pool_fwd(_p(table), vocab, emb, _p(ids), _p(mask), n_docs, doc_len, _p(out), fl, _stream()),
                      "rbr_masked_avg_pool_fwd")
            outs.append(out)
        ctx.save_for_backward(*ids_l, *[m for m in mask_l if m is not None])
        ctx.cfg, ctx.flags_l, ctx.mask_present, ctx.shape = cfg, flags_l, [m is not None for m in mask_l], (vocab, emb)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        _stream(refresh=True)
        cfg = ctx.cfg
        ns = len(ctx.flags_l)
        vocab, emb = ctx.shape
        if not ctx.needs_input_grad[0]:
            return (None,) * (2 + 2 * ns)
        saved = ctx.saved_tensors
        ids_l, masks = saved[:ns], list(saved[ns:])
        g_table = _grad_buf(cfg.get("arena"), cfg["table_param"], True)
        for s in range(ns):
            mask = masks.pop(0) if ctx.mask_present[s] else None
            if grads[s] is None:
                continue
            ids = ids_l[s]
            doc_len = ids.shape[-1]
            lib.check(lib.rbr_masked_avg_pool_bwd(_p(ids), _p(mask), ids.numel() // doc_len, doc_len, vocab, emb, cfg["padding_idx"],
                                                  _p(grads[s].contiguous()), _p(g_table), ctx.flags_l[s], _stream()),
                      "rbr_masked_avg_pool_bwd")
        return (g_table, None, *([None] * (2 * ns)))


# ---------------------------------------------------------------------------------------------------
# K3: NARRE attention
# ---------------------------------------------------------------------------------------------------
class NarreAttnFn(torch.autograd.Function):
    """LinearAttention.forward without dropout (reference models/narre/narre.py:40-60)."""

    @staticmethod
    def forward(ctx, feat, other_id, W_rv, W_id, h, b_1, b_2, ebd, padding_idx, arena, params):
        _stream(refresh=True)
        feat = _req(feat, torch.float32, "feat")
        other_id = _req(other_id, torch.int64, "other_id")
        B, R, H = feat.shape
        A = W_rv.shape[1]
        out = torch.empty(B, H, dtype=torch.float32, device=feat.device)
        scores = torch.empty(B, R, dtype=torch.float32, device=feat.device)
        args = [_req(t, torch.float32, n) for t, n in ((W_rv, "W_rv"), (W_id, "W_id"), (h, "h"), (b_1, "b_1"), (b_2, "b_2"),
                                                        (ebd, "ebd_vals"))]
        lib.check(lib.rbr_narre_attn_fwd(_p(feat), _p(other_id), B, R, H, A, *[_p(a) for a in args], ebd.shape[0], _p(out),
                                         _p(scores), _stream()), "rbr_narre_attn_fwd")
        ctx.save_for_backward(feat, other_id, scores, *args)
        ctx.padding_idx, ctx.arena, ctx.params = padding_idx, arena, params
        return out, scores.view(B, R, 1)

    @staticmethod
    def backward(ctx, g_out, g_scores):
        _stream(refresh=True)
        feat, other_id, scores, W_rv, W_id, h, b_1, b_2, ebd = ctx.saved_tensors
        B, R, H = feat.shape
        A = W_rv.shape[1]
        g_out = torch.zeros(B, H, device=feat.device) if g_out is None else g_out.contiguous()
        g_scores = None if g_scores is None else g_scores.contiguous()
        g_feat = torch.empty_like(feat)
        grads = [_grad_buf(ctx.arena, p, True) for p in ctx.params]
        lib.check(lib.rbr_narre_attn_bwd(_p(feat), _p(other_id), B, R, H, A, _p(W_rv), _p(W_id), _p(h), _p(b_1), _p(b_2), _p(ebd),
                                         ebd.shape[0], -1 if ctx.padding_idx is None else ctx.padding_idx, _p(scores), _p(g_out),
                                         _p(g_scores), _p(g_feat), *[_p(g) for g in grads], _stream()), "rbr_narre_attn_bwd")
        return (g_feat, None, *grads, None, None, None)


def _ptr_array(ptrs):
    import ctypes
    return (ctypes.c_void_p * len(ptrs))(*[None if p is None else int(p) for p in ptrs])


def _i64_array(vals):
    import ctypes
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


class NarreAttnPairFn(torch.autograd.Function):
    """Both LinearAttention modules of NARRE (narre.py:184-185) in ONE launch per direction on the tensor cores
    (csrc/attn_tc.cu).  Inputs: feat_u, id_u, feat_i, id_i, then the 6 parameters of each side
    (W_rv, W_id, h, b_1, b_2, ebd_vals).  Returns (out_u, scores_u [B,R,1], out_i, scores_i [B,R,1]), dropout excluded."""

    @staticmethod
    def forward(ctx, feat_u, id_u, feat_i, id_i, *rest):
        _stream(refresh=True)
        prm = [_req(t, torch.float32, "attention parameter") for t in rest[:12]]
        pads, arena, params = rest[12], rest[13], rest[14]
        feats = [_req(feat_u, torch.float32, "feat"), _req(feat_i, torch.float32, "feat")]
        ids = [_req(id_u, torch.int64, "other_id"), _req(id_i, torch.int64, "other_id")]
        B, R, H = feats[0].shape
        if feats[1].shape != feats[0].shape:
            raise ValueError("rbr_b200: the two attention sides must have the same [B, R, H] shape")
        A = prm[0].shape[1]
        outs = [torch.empty(B, H, dtype=torch.float32, device=feats[0].device) for _ in range(2)]
        scores = [torch.empty(B, R, dtype=torch.float32, device=feats[0].device) for _ in range(2)]
        side = [prm[:6], prm[6:]]
        lib.check(lib.rbr_narre_attn_pair_fwd(
            2, _ptr_array([_p(f) for f in feats]), _ptr_array([_p(i) for i in ids]), B, R, H, A,
            *[_ptr_array([_p(side[0][j]), _p(side[1][j])]) for j in range(6)],
            _i64_array([side[0][5].shape[0], side[1][5].shape[0]]), _ptr_array([_p(o) for o in outs]),
            _ptr_array([_p(s) for s in scores]), _stream()), "rbr_narre_attn_pair_fwd")
        ctx.save_for_backward(*feats, *ids, *scores, *prm)
        ctx.pads, ctx.arena, ctx.params = pads, arena, params
        return outs[0], scores[0].view(B, R, 1), outs[1], scores[1].view(B, R, 1)

    @staticmethod
    def backward(ctx, g_out_u, g_sc_u, g_out_i, g_sc_i):
        _stream(refresh=True)
        saved = ctx.saved_tensors
        feats, ids, scores, prm = saved[0:2], saved[2:4], saved[4:6], saved[6:18]
        B, R, H = feats[0].shape
        A = prm[0].shape[1]
        dev = feats[0].device
        g_outs = [torch.zeros(B, H, device=dev) if g is None else g.contiguous() for g in (g_out_u, g_out_i)]
        g_scs = [None if g is None else g.contiguous() for g in (g_sc_u, g_sc_i)]
        g_feats = [torch.empty_like(feats[0]), torch.empty_like(feats[1])]
        grads = [_grad_buf(ctx.arena, p, True) for p in ctx.params]             # 12: side u then side i
        side, gside = [prm[:6], prm[6:]], [grads[:6], grads[6:]]
        pads = [-1 if x is None else int(x) for x in ctx.pads]
        any_sc = any(g is not None for g in g_scs)
        lib.check(lib.rbr_narre_attn_pair_bwd(
            2, _ptr_array([_p(f) for f in feats]), _ptr_array([_p(i) for i in ids]), B, R, H, A,
            *[_ptr_array([_p(side[0][j]), _p(side[1][j])]) for j in range(6)],
            _i64_array([side[0][5].shape[0], side[1][5].shape[0]]), _i64_array(pads), _ptr_array([_p(s) for s in scores]),
            _ptr_array([_p(g) for g in g_outs]), _ptr_array([_p(g) for g in g_scs]) if any_sc else None,
            _ptr_array([_p(g) for g in g_feats]),
            *[_ptr_array([_p(gside[0][j]), _p(gside[1][j])]) for j in range(6)], _stream()), "rbr_narre_attn_pair_bwd")
        return (g_feats[0], None, g_feats[1], None, *grads, None, None, None)


def narre_attn_pair_supported(R: int, H: int, A: int) -> bool:
    return bool(lib.rbr_narre_attn_pair_supported(R, H, A))


# ---------------------------------------------------------------------------------------------------
# K4: LastFeat x2 + FM head
# ---------------------------------------------------------------------------------------------------
class HeadFn(torch.autograd.Function):
    """pred = FM(LastFeat_u(u_text, u_id), LastFeat_i(i_text, i_id))  (reference layers.py:156-165, 188-209)."""

    @staticmethod
    def forward(ctx, u_text, i_text, u_id, i_id, Wu, bu, ebd_u, Wi, bi, ebd_i, fm_h, user_bias, item_bias, g_bias, drop_p,
                drop_seed, padding_idx, arena, params, seed_dev=None):
        """padding_idx: one int / None for all four id tables, or a 4-tuple (LastFeat_u.ebd, LastFeat_i.ebd, FM.user_bias,
        FM.item_bias) — each nn.Embedding keeps its own padding row."""
        _stream(refresh=True)
        u_text = _req(u_text, torch.float32, "u_text")
        i_text = _req(i_text, torch.float32, "i_text")
        u_id = _req(u_id, torch.int64, "u_id")
        i_id = _req(i_id, torch.int64, "i_id")
        B, H = u_text.shape
        K = Wu.shape[1]
        fl = [_req(t, torch.float32, "head parameter") for t in (Wu, bu, ebd_u, Wi, bi, ebd_i, fm_h, user_bias, item_bias, g_bias)]
        pred = torch.empty(B, dtype=torch.float32, device=u_text.device)
        u_lat = torch.empty(B, K, dtype=torch.float32, device=u_text.device)
        i_lat = torch.empty(B, K, dtype=torch.float32, device=u_text.device)
        lib.check(lib.rbr_head_fwd(_p(u_text), _p(i_text), _p(u_id), _p(i_id), B, H, K, *[_p(t) for t in fl], ebd_u.shape[0],
                                   ebd_i.shape[0], float(drop_p), int(drop_seed), _p(seed_dev), _p(pred), _p(u_lat), _p(i_lat), None,
                                   0.0, None, None, _stream()), "rbr_head_fwd")
        ctx.save_for_backward(u_text, i_text, u_id, i_id, fl[0], fl[3], fl[6], u_lat, i_lat)
        if not isinstance(padding_idx, (tuple, list)):
            padding_idx = (padding_idx,) * 4
        padding_idx = tuple(-1 if x is None else int(x) for x in padding_idx)
        ctx.drop, ctx.padding_idx, ctx.arena, ctx.params = (float(drop_p), int(drop_seed), seed_dev), padding_idx, arena, params
        ctx.sizes = (ebd_u.shape[0], ebd_i.shape[0])
        return pred

    @staticmethod
    def backward(ctx, g_pred):
        _stream(refresh=True)
        u_text, i_text, u_id, i_id, Wu, Wi, fm_h, u_lat, i_lat = ctx.saved_tensors
        B, H = u_text.shape
        K = Wu.shape[1]
        g_pred = g_pred.contiguous()
        g_ut, g_it = torch.empty_like(u_text), torch.empty_like(i_text)
        # params order: Wu, bu, ebd_u, Wi, bi, ebd_i, fm_h, user_bias, item_bias, g_bias
        grads = [_grad_buf(ctx.arena, p, True) for p in ctx.params]
        lib.check(lib.rbr_head_bwd(_p(u_text), _p(i_text), _p(u_id), _p(i_id), B, H, K, _p(Wu), _p(Wi), _p(fm_h), _p(u_lat),
                                   _p(i_lat), ctx.drop[0], ctx.drop[1], _p(ctx.drop[2]), *ctx.padding_idx,
                                   ctx.sizes[0], ctx.sizes[1], _p(g_pred), _p(g_ut), _p(g_it), *[_p(g) for g in grads],
                                   _stream()), "rbr_head_bwd")
        return (g_ut, g_it, None, None, *grads, None, None, None, None, None, None)


class HeadLossFn(torch.autograd.Function):
    """(loss, pred) = (MSELoss(pred, ratings), pred) with pred as in HeadFn: the fused-MSE branch of rbr_head_fwd
    (reference layers.py:156-165, 188-209 + trainer/train_deepconn_pp.py:140,164: nn.MSELoss(), reduction 'mean').
    The forward launch also writes d loss / d pred = 2 (pred - rating) / B, so the backward is one rbr_head_bwd launch."""

    @staticmethod
    def forward(ctx, u_text, i_text, u_id, i_id, ratings, Wu, bu, ebd_u, Wi, bi, ebd_i, fm_h, user_bias, item_bias, g_bias,
                drop_p, drop_seed, padding_idx, arena, params, seed_dev=None):
        _stream(refresh=True)
        u_text = _req(u_text, torch.float32, "u_text")
        i_text = _req(i_text, torch.float32, "i_text")
        u_id = _req(u_id, torch.int64, "u_id")
        i_id = _req(i_id, torch.int64, "i_id")
        ratings = _req(ratings, torch.float32, "ratings").view(-1)
        B, H = u_text.shape
        if ratings.numel() != B:
            raise ValueError("rbr_b200: ratings must have one value per sample")
        K = Wu.shape[1]
        fl = [_req(t, torch.float32, "head parameter") for t in (Wu, bu, ebd_u, Wi, bi, ebd_i, fm_h, user_bias, item_bias, g_bias)]
        dev = u_text.device
        pred = torch.empty(B, dtype=torch.float32, device=dev)
        u_lat = torch.empty(B, K, dtype=torch.float32, device=dev)
        i_lat = torch.empty(B, K, dtype=torch.float32, device=dev)
        pred_grad = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.zeros((), dtype=torch.float32, device=dev)         # the kernel accumulates sum (pred - rating)^2 / B into it
        lib.check(lib.rbr_head_fwd(_p(u_text), _p(i_text), _p(u_id), _p(i_id), B, H, K, *[_p(t) for t in fl], ebd_u.shape[0],
                                   ebd_i.shape[0], float(drop_p), int(drop_seed), _p(seed_dev), _p(pred), _p(u_lat), _p(i_lat),
                                   _p(ratings), 1.0 / B, _p(loss), _p(pred_grad), _stream()), "rbr_head_fwd")
        ctx.save_for_backward(u_text, i_text, u_id, i_id, fl[0], fl[3], fl[6], u_lat, i_lat, pred_grad)
        if not isinstance(padding_idx, (tuple, list)):
            padding_idx = (padding_idx,) * 4
        padding_idx = tuple(-1 if x is None else int(x) for x in padding_idx)
        ctx.drop, ctx.padding_idx, ctx.arena, ctx.params = (float(drop_p), int(drop_seed), seed_dev), padding_idx, arena, params
        ctx.sizes = (ebd_u.shape[0], ebd_i.shape[0])
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, g_loss, g_pred):
        _stream(refresh=True)
        u_text, i_text, u_id, i_id, Wu, Wi, fm_h, u_lat, i_lat, pred_grad = ctx.saved_tensors
        B, H = u_text.shape
        K = Wu.shape[1]
        pg = pred_grad * g_loss                         # upstream d / d loss (a device scalar: no host read)
        g_ut, g_it = torch.empty_like(u_text), torch.empty_like(i_text)
        grads = [_grad_buf(ctx.arena, p, True) for p in ctx.params]
        lib.check(lib.rbr_head_bwd(_p(u_text), _p(i_text), _p(u_id), _p(i_id), B, H, K, _p(Wu), _p(Wi), _p(fm_h), _p(u_lat),
                                   _p(i_lat), ctx.drop[0], ctx.drop[1], _p(ctx.drop[2]), *ctx.padding_idx,
                                   ctx.sizes[0], ctx.sizes[1], _p(pg), _p(g_ut), _p(g_it), *[_p(g) for g in grads],
                                   _stream()), "rbr_head_bwd")
        return (g_ut, g_it, None, None, None, *grads, None, None, None, None, None, None)


def head_supported(hidden: int, latent: int) -> bool:
    """Whether the fused head kernels (K4) take this shape: both LastFeat weights, their CTA-partial gradients and a 32-sample
    tile must fit shared memory (csrc/head.cu)."""
    bwd = (2 * hidden * (latent + 1) + 2 * hidden * latent + 2 * 32 * hidden + 2 * 32 * latent) * 4
    return latent <= 128 and bwd <= 200 * 1024


def head_dropout_mask(batch: int, latent: int, drop_p: float, drop_seed: int, seed_dev: Optional[torch.Tensor],
                      device=None) -> torch.Tensor:
    """The FM dropout keep-scale [B, K] (values 0 or 1/(1-p)) the head kernels apply for this (p, seed, device counter) — a
    test / debug aid: the mask is a counter hash, not torch's Philox stream, so parity at p > 0 is checked by feeding this mask
    to the oracle."""
    dev = device or (seed_dev.device if seed_dev is not None else torch.device("cuda"))
    keep = torch.empty(batch, latent, dtype=torch.float32, device=dev)
    lib.check(lib.rbr_head_dropout_mask(batch, latent, float(drop_p), int(drop_seed), _p(seed_dev), _p(keep), _stream(True)),
              "rbr_head_dropout_mask")
    return keep


# ---------------------------------------------------------------------------------------------------
# Raw (no-autograd) entry used by the kernel-level tests and by bench.py's per-kernel roofline timing
# ---------------------------------------------------------------------------------------------------
def conv_act_maxpool(table: torch.Tensor, ids: torch.Tensor, mask: Optional[torch.Tensor], weight: torch.Tensor,
                     bias: torch.Tensor, pad: int, act: int = ACT_RELU, precision: str = "bf16",
                     shadow: Optional[torch.Tensor] = None, packed: Optional[torch.Tensor] = None, flags: int = 0,
                     mask_from_ids: bool = False, select_docs: bool = True, row_index_table: bool = True):
    """One K2 launch: returns (feat [n_docs, H] fp32, argmax [n_docs, H] int32).  `flags`: CONV_TC_* kernel selection.
    `select_docs=False`: no scratch at all; `row_index_table=False`: scratch for the document selection only, so the kernel's
    index warp resolves ids and masks itself (rbr_conv_fwd_workspace_bytes vs ..._bytes2)."""
    table = _req(table, torch.float32, "table")
    ids = _ids(ids, "ids")
    flags |= _id_flags(ids, mask, mask_from_ids)
    mask = _mask_u8(mask, "mask")
    prec = _PREC[precision]
    if prec == PREC_BF16 and shadow is None:
        shadow = table_to_bf16(table)
    if packed is None:
        packed = conv_pack(weight)
    h, emb, k = weight.shape
    doc_len = ids.shape[-1]
    n_docs = ids.numel() // doc_len
    feat = torch.empty(n_docs, h, dtype=torch.float32, device=table.device)
    amax = torch.empty(n_docs, h, dtype=torch.int32, device=table.device)
    ws_bytes = 0
    if prec == PREC_BF16 and select_docs:
        ws_bytes = (lib.rbr_conv_fwd_workspace_bytes2(n_docs, doc_len, k, pad) if row_index_table
                    else lib.rbr_conv_fwd_workspace_bytes(n_docs))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=table.device) if ws_bytes else None
    lib.check(lib.rbr_conv_act_maxpool_fwd(prec, act, _p(table), _p(shadow), table.shape[0], emb, _p(ids), _p(mask), None, 0,
                                           n_docs, doc_len, _p(packed), _p(_req(bias, torch.float32, "bias")), h, k, pad,
                                           _p(feat), _p(amax), None, h, _p(ws), ws_bytes, flags, _stream(True)),
              "rbr_conv_act_maxpool_fwd")
    return feat, amax


# ---------------------------------------------------------------------------------------------------
# K5 + gated K2: one side of the D-ATT encoder (local attention + global attention), ids → [N, l_out + 3*g_out]
# ---------------------------------------------------------------------------------------------------
class DattEncodeFn(torch.autograd.Function):
    """feats[s] = cat(LocalAttention_s(x_s), *GlobalAttention_s(x_s)) for x_s = table[ids_s], s in (user, item)
    (reference models/dual_att/layers.py:43-53, 81-89 and dual_att.py:45-50, 53-56), without materialising x, its
    permuted copy or the gated copies.  Both sides go through ONE autograd node because they share the embedding table
    (dual_att.py:22): its gradient buffer is accumulated into by both and handed to autograd once.

    Inputs after (table, cfg): per side (ids, then 12 parameters: local attn w, b; local conv w, b; global attn w, b;
    conv1 w, b; conv2 w, b; conv3 w, b)."""

    N_PRM = 12

    @staticmethod
    def forward(ctx, table, cfg, *rest):
        _stream(refresh=True)
        table = _req(table, torch.float32, "embedding table")
        stride = 1 + DattEncodeFn.N_PRM
        n_sides = len(rest) // stride
        prec = _PREC[cfg["precision"]]
        vocab, emb = table.shape
        dev = table.device
        shadow = cfg["shadow_fn"]() if prec == PREC_BF16 else None
        saved, feats, side_ctx = [], [], []
        for s in range(n_sides):
            ids = _ids(rest[s * stride], "token ids")
            idf = _id_width_flag(ids)
            prm = [_req(t, torch.float32, "D-ATT parameter") for t in rest[s * stride + 1:(s + 1) * stride]]
            la_w, la_b, lc_w, lc_b, ga_w, ga_b = prm[:6]
            g_convs = [(prm[6 + 2 * i], prm[7 + 2 * i]) for i in range(3)]
            n_docs, doc_len = ids.shape
            win = la_w.shape[2]
            ws_bytes = lib.rbr_datt_gate_workspace_bytes(n_docs, doc_len, emb, win, vocab)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            gate_l = torch.empty(n_docs, doc_len, dtype=torch.float32, device=dev)
            gate_g = torch.empty(n_docs, dtype=torch.float32, device=dev)
            lib.check(lib.rbr_datt_gate_fwd(_p(table), vocab, emb, _p(ids), n_docs, doc_len, _p(la_w), _p(la_b), win, _p(ga_w),
                                            _p(ga_b), _p(gate_l), _p(gate_g), _p(ws), ws_bytes, idf, _stream()), "rbr_datt_gate_fwd")
            convs = [(lc_w, lc_b, gate_l, 1)] + [(w, b, gate_g, 2) for w, b in g_convs]
            packed = [conv_pack(w) for w, _, _, _ in convs]
            h_total = sum(w.shape[0] for w, _, _, _ in convs)
            feat = torch.empty(n_docs, h_total, dtype=torch.float32, device=dev)
            amax = torch.empty(n_docs, h_total, dtype=torch.int32, device=dev)
            pre = torch.empty(n_docs, h_total, dtype=torch.float32, device=dev)
            col = 0
            ws_bytes = (max(lib.rbr_conv_fwd_workspace_bytes2(n_docs, doc_len, w.shape[2], 0) for w, _, _, _ in convs)
                        if prec == PREC_BF16 else 0)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
            for (w, b, gate, mode), pk in zip(convs, packed):
                h, _, k = w.shape
                lib.check(lib.rbr_conv_act_maxpool_fwd(prec, ACT_TANH, _p(table), _p(shadow), vocab, emb, _p(ids), None, _p(gate),
                                                       mode, n_docs, doc_len, _p(pk), _p(b), h, k, 0, feat.data_ptr() + 4 * col,
                                                       amax.data_ptr() + 4 * col, pre.data_ptr() + 4 * col, h_total, _p(ws), ws_bytes,
                                                       idf, _stream()), "rbr_conv_act_maxpool_fwd")
                col += h
            saved += [ids, gate_l, gate_g, feat, amax, pre]
            feats.append(feat)
            side_ctx.append((prm, packed))
        ctx.cfg, ctx.table, ctx.shadow, ctx.side_ctx, ctx.n_sides = cfg, table, shadow, side_ctx, n_sides
        ctx.save_for_backward(*saved)
        return tuple(feats)

    @staticmethod
    def backward(ctx, *feat_grads):
        _stream(refresh=True)
        cfg, table = ctx.cfg, ctx.table
        arena: Optional[GradArena] = cfg.get("arena")
        prec = _PREC[cfg["precision"]]
        vocab, emb = table.shape
        dev = table.device
        need_table = ctx.needs_input_grad[0]
        g_table = _grad_buf(arena, cfg["table_param"], need_table)
        ret = []
        # the user and item sides only meet in the (atomic) table gradient: the second side runs on its own stream
        main = torch.cuda.current_stream()
        side_streams = _side_streams(dev, max(0, ctx.n_sides - 1))
        # every buffer the side stream accumulates into is allocated AND zero-filled on the main stream before the fork event
        # (a frozen table or arena=None materialises them here for the first time)
        all_grads = [[_grad_buf(arena, p, True) for p in cfg["params"][s]] for s in range(ctx.n_sides)]
        fgs = [None if feat_grads[s] is None else feat_grads[s].contiguous() for s in range(ctx.n_sides)]
        ev_fork = main.record_event() if ctx.n_sides > 1 else None
        for s in range(ctx.n_sides):
            ids, gate_l, gate_g, feat, amax, pre = ctx.saved_tensors[6 * s:6 * s + 6]
            prm, packed = ctx.side_ctx[s]
            grads = all_grads[s]
            ret += [None, *grads]
            if fgs[s] is None:
                continue
            n_docs, doc_len = ids.shape
            fg = fgs[s]
            stream_s = main if s == 0 else side_streams[s - 1]
            if s > 0:
                stream_s.wait_event(ev_fork)
            with torch.cuda.stream(stream_s):
                _stream(refresh=True)
                DattEncodeFn._side_backward(cfg, table, ctx.shadow, prec, arena, ids, gate_l, gate_g, feat, amax, pre, prm, packed,
                                            grads, fg, g_table, vocab, emb, dev)
        for s in range(1, ctx.n_sides):
            if fgs[s] is not None:
                main.wait_event(side_streams[s - 1].record_event())
        _stream(refresh=True)
        return (g_table, None, *ret)

    @staticmethod
    def _side_backward(cfg, table, shadow, prec, arena, ids, gate_l, gate_g, feat, amax, pre, prm, packed, grads, fg, g_table, vocab,
                       emb, dev):
        if True:
            n_docs, doc_len = ids.shape
            idf = _id_width_flag(ids)
            la_w, la_b, lc_w, lc_b, ga_w, ga_b = prm[:6]
            convs = [(lc_w, lc_b, gate_l, 1, 2, 3)] + [(prm[6 + 2 * i], prm[7 + 2 * i], gate_g, 2, 6 + 2 * i, 7 + 2 * i)
                                                       for i in range(3)]
            d_gate_l = torch.zeros_like(gate_l)
            d_gate_g = torch.zeros_like(gate_g)
            h_total = feat.shape[1]
            col = 0
            for (w, b, gate, mode, wi, bi), pk in zip(convs, packed):
                h, _, k = w.shape
                ws_bytes = lib.rbr_conv_bwd_workspace_bytes(n_docs, h, k, emb, vocab)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                lib.check(lib.rbr_conv_act_maxpool_bwd(
                    prec, ACT_TANH, _p(table), _p(shadow), vocab, emb, _p(ids), None, _p(gate), mode, n_docs, doc_len, _p(pk),
                    h, k, 0, feat.data_ptr() + 4 * col, amax.data_ptr() + 4 * col, fg.data_ptr() + 4 * col,
                    pre.data_ptr() + 4 * col, h_total, cfg["padding_idx"], _p(grads[wi]), _p(grads[bi]), _p(g_table),
                    _p(d_gate_l if mode == 1 else d_gate_g), _p(ws), ws_bytes, idf, _stream()), "rbr_conv_act_maxpool_bwd")
                col += h
            win = la_w.shape[2]
            ws_bytes = lib.rbr_datt_gate_workspace_bytes(n_docs, doc_len, emb, win, vocab)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            lib.check(lib.rbr_datt_gate_bwd(_p(table), vocab, emb, _p(ids), n_docs, doc_len, _p(la_w), win, _p(ga_w), _p(gate_l),
                                            _p(gate_g), _p(d_gate_l), _p(d_gate_g), cfg["padding_idx"], _p(grads[0]), _p(grads[1]),
                                            _p(grads[4]), _p(grads[5]), _p(g_table), _p(ws), ws_bytes, idf, _stream()),
                      "rbr_datt_gate_bwd")
