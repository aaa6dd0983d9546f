"""NARRE — drop-in for the reference's models/narre/narre.py:139-192 (same constructor, forward signature,
parameter names and state_dict keys), running on the rbr_b200 CUDA kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .layers import FM, HotPathModule, LastFeat, LinearAttention, NgramFeat, WordEmbedding, fused_head


class NARRE(HotPathModule):
    staging_spec = dict(tokens=(0, 1), masks=(2, 3))      # which forward() inputs are token-id tensors / their masks (staging.StagedInputs)

    def __init__(self, user_size, item_size, vocab_size, kernel_sizes, hidden_dim, embedding_dim, att_dim, latent_dim,
                 max_doc_num, max_doc_len, dropout, word_padding_idx, user_padding_idx, item_padding_idx,
                 pretrained_embeddings, arch, precision=None):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.hiddem_dim = hidden_dim          # (sic) the reference's attribute name, narre.py:147
        self.doc_num = max_doc_num
        self.doc_len = max_doc_len

        self.word_embeddings = WordEmbedding(vocab_size, embedding_dim, pretrained_embeddings=pretrained_embeddings)
        self.ngram = NgramFeat(kernel_sizes, embedding_dim, hidden_dim, max_doc_len, arch=arch, precision=precision)
        # a user's reviews are attended with the ITEM-id embedding and vice versa (narre.py:156-157)
        self.user_att = LinearAttention(item_size, hidden_dim, att_dim, dropout, padding_idx=item_padding_idx)
        self.item_att = LinearAttention(user_size, hidden_dim, att_dim, dropout, padding_idx=user_padding_idx)
        self.user_feat = LastFeat(user_size, hidden_dim, latent_dim, padding_idx=user_padding_idx)
        self.item_feat = LastFeat(item_size, hidden_dim, latent_dim, padding_idx=item_padding_idx)
        self.fm = FM(user_size, item_size, latent_dim, dropout, user_padding_idx=user_padding_idx,
                     item_padding_idx=item_padding_idx)
        self.last_arena = None
        self.fused_attention = True     # both attention sides in one tensor-core launch (False: the per-side K3 kernels)

    def _new_arena(self):
        arena = ops.GradArena.for_module(self)
        self.last_arena = arena
        self.ngram._arena = arena
        self.user_att._arena = arena
        self.item_att._arena = arena
        return arena

    def invalidate_operand_cache(self):
        """Force the bf16 table shadow and the packed conv weights to be rebuilt at the next forward — what happens
        after every optimizer step in training (parameter version counters change).  bench.py calls this every
        step so that the operand staging kernels are inside the timed region."""
        self.word_embeddings.invalidate_operand_cache()
        if hasattr(self.ngram.conv, "invalidate_operand_cache"):           # (arch="HierPooling" has no conv weights to re-pack)
            self.ngram.conv.invalidate_operand_cache()

    def _encode_attend(self, u_text, i_text, u_text_masks, i_text_masks, reuid, reiid):
        bz = u_text.shape[0]
        # every review is an independent doc of length T (narre.py:170-176)
        u_docs = u_text.reshape(-1, self.doc_len)
        i_docs = i_text.reshape(-1, self.doc_len)
        u_m = None if u_text_masks is None else u_text_masks.reshape(-1, self.doc_len)
        i_m = None if i_text_masks is None else i_text_masks.reshape(-1, self.doc_len)
        u_feat, i_feat = self.ngram.encode(self.word_embeddings, [u_docs, i_docs], [u_m, i_m])
        u_feat = u_feat.view(bz, self.doc_num, self.hiddem_dim)
        i_feat = i_feat.view(bz, self.doc_num, self.hiddem_dim)
        ua, ia = self.user_att, self.item_att
        if self.fused_attention and ops.narre_attn_pair_supported(self.doc_num, self.hiddem_dim, ua.W_rv.shape[1]):
            # both sides in one launch per direction on the tensor cores (csrc/attn_tc.cu); dropout as in LinearAttention.forward
            pu = [ua.W_rv, ua.W_id, ua.h, ua.b_1, ua.b_2, ua.ebd_vals.weight]
            pi = [ia.W_rv, ia.W_id, ia.h, ia.b_1, ia.b_2, ia.ebd_vals.weight]
            u_out, u_att_scores, i_out, i_att_scores = ops.NarreAttnPairFn.apply(
                u_feat, reuid, i_feat, reiid, *pu, *pi, (ua.padding_idx, ia.padding_idx), self.last_arena, pu + pi)
            return ua.dropout(u_out), ia.dropout(i_out), u_att_scores, i_att_scores
        # fallback (att_dim > 32, ...): one warp-per-sample launch per side, the item side on an auxiliary stream — forward and
        # (because autograd replays each node on its forward stream) backward
        main = torch.cuda.current_stream()
        aux = ops._side_streams(u_feat.device, 1)[0]
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            i_feat, i_att_scores = self.item_att(i_feat, reiid)    # narre.py:185
        u_feat, u_att_scores = self.user_att(u_feat, reuid)        # narre.py:184
        main.wait_stream(aux)
        # allocated on the auxiliary stream, consumed (and possibly freed) on the main one: tell the caching allocator
        i_feat.record_stream(main)
        i_att_scores.record_stream(main)
        return u_feat, i_feat, u_att_scores, i_att_scores

    def forward(self, u_text, i_text, u_text_masks, i_text_masks, u_id, i_id, reuid, reiid):
        """u_text/i_text [bz, R, T] int64 (or int32), masks [bz, R, T] bool (None: ids != 0), ids [bz], reuid/reiid [bz, R] →
        (pred [bz], u_att_scores [bz, R, 1], i_att_scores [bz, R, 1])."""
        arena = self._new_arena()
        u_feat, i_feat, u_att_scores, i_att_scores = self._encode_attend(u_text, i_text, u_text_masks, i_text_masks, reuid, reiid)
        pred = fused_head(self.user_feat, self.item_feat, self.fm, u_feat, i_feat, u_id, i_id, self.training, arena)
        self._after_forward()
        return pred.view(-1), u_att_scores, i_att_scores

    def forward_loss(self, u_text, i_text, u_text_masks, i_text_masks, u_id, i_id, reuid, reiid, ratings):
        """forward + nn.MSELoss() in the head kernel's launch (trainer/train_narre.py:165-167 as one call):
        returns (loss, (pred, u_att_scores, i_att_scores))."""
        arena = self._new_arena()
        u_feat, i_feat, u_att_scores, i_att_scores = self._encode_attend(u_text, i_text, u_text_masks, i_text_masks, reuid, reiid)
        loss, pred = fused_head(self.user_feat, self.item_feat, self.fm, u_feat, i_feat, u_id, i_id, self.training, arena,
                                ratings=ratings)
        self._after_forward()
        return loss, (pred.view(-1), u_att_scores, i_att_scores)
