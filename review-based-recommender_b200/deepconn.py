"""DeepCoNNpp — drop-in for the reference's models/deepconn/deepconn.py:10-53 (same constructor, forward
signature, parameter names and state_dict keys), running on the rbr_b200 CUDA kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .layers import FM, HotPathModule, LastFeat, NgramFeat, WordEmbedding, fused_head


class DeepCoNNpp(HotPathModule):
    staging_spec = dict(tokens=(0, 1), masks=(2, 3))      # which forward() inputs are token-id tensors / their masks (staging.StagedInputs)

    def __init__(self, user_size, item_size, vocab_size, kernel_sizes, embedding_dim, hidden_dim, latent_dim, doc_len,
                 pretrained_embeddings, dropout, arch="CNN", precision=None):
        super().__init__()
        self.user_size = user_size
        self.item_size = item_size
        self.vocab_size = vocab_size
        self.hidden_dim = hidden_dim

        self.word_embeddings = WordEmbedding(vocab_size, embedding_dim, pretrained_embeddings=pretrained_embeddings)
        self.ngram = NgramFeat(kernel_sizes, embedding_dim, hidden_dim, doc_len, arch=arch, precision=precision)
        self.user_feat = LastFeat(user_size, hidden_dim, latent_dim, padding_idx=0)
        self.item_feat = LastFeat(item_size, hidden_dim, latent_dim, padding_idx=0)
        self.fm = FM(user_size, item_size, latent_dim, dropout, user_padding_idx=0, item_padding_idx=0)
        self.last_arena = None          # flat gradient buffer of the most recent step (parallel.py all-reduces it)

    def _new_arena(self):
        arena = ops.GradArena.for_module(self)
        self.last_arena = arena
        self.ngram._arena = arena
        return arena

    def invalidate_operand_cache(self):
        """Force the bf16 table shadow and the packed conv weights to be rebuilt at the next forward — what happens
        after every optimizer step in training (parameter version counters change).  bench.py calls this every
        step so that the operand staging kernels are inside the timed region."""
        self.word_embeddings.invalidate_operand_cache()
        if hasattr(self.ngram.conv, "invalidate_operand_cache"):           # (arch="HierPooling" has no conv weights to re-pack)
            self.ngram.conv.invalidate_operand_cache()

    def forward(self, u_revs, i_revs, u_rev_masks, i_rev_masks, u_ids, i_ids):
        """u_revs/i_revs [bz, doc_len] int64 (or int32: the staged input pipeline), masks [bz, doc_len] bool (None: derived on
        the device as ids != 0, which is what collate_fn computes, utils.py:30-42), ids [bz] int64 → preds [bz]."""
        arena = self._new_arena()
        u_rev_feats, i_rev_feats = self.ngram.encode(self.word_embeddings, [u_revs, i_revs], [u_rev_masks, i_rev_masks])
        preds = fused_head(self.user_feat, self.item_feat, self.fm, u_rev_feats, i_rev_feats, u_ids, i_ids, self.training, arena)
        self._after_forward()
        return preds.view(u_revs.shape[0])

    def forward_loss(self, u_revs, i_revs, u_rev_masks, i_rev_masks, u_ids, i_ids, ratings):
        """forward + nn.MSELoss() in the head kernel's launch (trainer/train_deepconn_pp.py:162-164 as one call):
        returns (loss, preds); `loss.backward()` as usual."""
        arena = self._new_arena()
        u_rev_feats, i_rev_feats = self.ngram.encode(self.word_embeddings, [u_revs, i_revs], [u_rev_masks, i_rev_masks])
        loss, preds = fused_head(self.user_feat, self.item_feat, self.fm, u_rev_feats, i_rev_feats, u_ids, i_ids, self.training,
                                 arena, ratings=ratings)
        self._after_forward()
        return loss, preds.view(u_revs.shape[0])
