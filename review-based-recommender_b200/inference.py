"""Inference scoring (BASELINE.json configs[4]) and the per-entity feature cache.

The reference has no predict entry point: scoring is `valid_one_epoch` (trainer/train_deepconn_pp.py:191-212) — `model.eval()`,
`torch.no_grad()`, the same forward — and every example carries its own two padded documents (:274), so the encoder runs
twice per (user, item) pair although a user's document is the same in every pair the user appears in.

`PairScorer.score_pairs` is that data flow on the fused kernels.  `PairScorer.build_cache` encodes each entity's document ONCE
(K2 forward over [U, L] and [I, L]) and `score_cached` then scores pairs with a row gather (K1) + the head kernel (K4) only:
10 M pairs over U + I = 32 k documents is ≈ 600x less encoder work, bit-identical scores.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .layers import fused_head


class PairScorer:
    def __init__(self, model):
        if not hasattr(model, "ngram") or hasattr(model, "user_att"):
            raise TypeError("PairScorer serves DeepCoNNpp (one document per entity); NARRE / D-ATT score through model.eval() forward")
        self.model = model.eval()
        self.user_feat: Optional[torch.Tensor] = None
        self.item_feat: Optional[torch.Tensor] = None

    @torch.no_grad()
    def score_pairs(self, u_docs, i_docs, u_masks, i_masks, u_ids, i_ids) -> torch.Tensor:
        """The reference's per-pair data flow: both documents encoded for every pair."""
        return self.model(u_docs, i_docs, u_masks, i_masks, u_ids, i_ids)

    @torch.no_grad()
    def build_cache(self, user_docs: torch.Tensor, item_docs: torch.Tensor, user_masks: Optional[torch.Tensor] = None,
                    item_masks: Optional[torch.Tensor] = None, chunk: int = 4096) -> None:
        """user_docs [U, L], item_docs [I, L]: row e = the document of entity id e (row 0 = the padding entity).
        Masks default to ids != 0 (utils.py:30-42)."""
        m = self.model

        def encode(docs, masks):
            outs = []
            for lo in range(0, docs.shape[0], chunk):
                mk = None if masks is None else masks[lo:lo + chunk]
                (f,) = m.ngram.encode(m.word_embeddings, [docs[lo:lo + chunk]], [mk])
                outs.append(f)
            return torch.cat(outs, dim=0)
        self.user_feat = encode(user_docs, user_masks)
        self.item_feat = encode(item_docs, item_masks)

    @torch.no_grad()
    def score_cached(self, u_ids: torch.Tensor, i_ids: torch.Tensor) -> torch.Tensor:
        """preds for pairs (u_ids[b], i_ids[b]) from the cached features: K1 row gather + K4 head, no encoder."""
        if self.user_feat is None:
            raise RuntimeError("PairScorer.build_cache(...) first")
        m = self.model
        u_text = ops.gather_rows(self.user_feat, u_ids)
        i_text = ops.gather_rows(self.item_feat, i_ids)
        return fused_head(m.user_feat, m.item_feat, m.fm, u_text, i_text, u_ids, i_ids, False, None).view(-1)
