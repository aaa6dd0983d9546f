"""Seeded synthetic batches and parameter sets for the review-encoder hot path.

Shapes and value rules follow the reference's preprocessing output contract:
token id 0 = <pad>, 1 = <unk>, 2 = <sep> (preprocess/divide_and_create_example_doc.py:198),
frequency-ranked vocabulary (preprocess/_tokenizer.py:53-65) → skewed ids; docs are
truncated / tail-padded with 0 to a fixed length (preprocess/_tokenizer.py:113-121);
NARRE pads missing reviews with all-zero rows and review id 0
(preprocess/divide_and_create_example_word.py:271-273); ratings are the `overall` stars 1..5.

Everything is generated on the CPU with a seeded torch.Generator so the same batch can be
fed to the CUDA path and to the CPU oracle.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

SEED_BASE = 20200616  # echoes preprocess/divide_and_create_example_doc.py:104


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def skewed_tokens(g: torch.Generator, shape, vocab: int, uniform: bool = False) -> torch.Tensor:
    """Token ids in [3, vocab) skewed toward small (frequent) ids: 3 + floor((V-3)·u³)."""
    u = torch.rand(shape, generator=g, dtype=torch.float64)
    if uniform:
        ids = 1 + (u * (vocab - 1)).floor().long()
    else:
        ids = 3 + ((vocab - 3) * u ** 3).floor().long()
    return ids.clamp_(max=vocab - 1)


def doc_batch(batch: int, doc_len: int, vocab: int, seed: int = SEED_BASE, uniform: bool = False,
              min_frac: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """[B,L] int64 token ids, tail-padded with 0 to length ~U{L*min_frac..L}, and mask = ids != 0."""
    g = _gen(seed)
    ids = skewed_tokens(g, (batch, doc_len), vocab, uniform)
    lo = max(1, int(doc_len * min_frac))
    lens = torch.randint(lo, doc_len + 1, (batch,), generator=g)
    pos = torch.arange(doc_len).unsqueeze(0)
    ids = torch.where(pos < lens.unsqueeze(1), ids, torch.zeros((), dtype=torch.long))
    return ids, ids != 0


def doc_batch_device(batch: int, doc_len: int, vocab: int, gen: torch.Generator, device, dtype=torch.int64) -> torch.Tensor:
    """doc_batch's distribution generated ON THE DEVICE (config 5 scores 10 M pairs = 80 GB of int64 ids: produced chunk by
    chunk, never materialised).  Returns ids only (mask = ids != 0)."""
    u = torch.rand(batch, doc_len, device=device, generator=gen, dtype=torch.float32)
    ids = (3 + ((vocab - 3) * u * u * u).floor()).to(torch.int64).clamp_(max=vocab - 1)
    lens = torch.randint(max(1, doc_len // 2), doc_len + 1, (batch, 1), device=device, generator=gen)
    pos = torch.arange(doc_len, device=device).unsqueeze(0)
    return torch.where(pos < lens, ids, torch.zeros((), dtype=torch.int64, device=device)).to(dtype)


def deepconn_batch(batch: int, doc_len: int, vocab: int, users: int, items: int, seed: int = SEED_BASE,
                   uniform: bool = False):
    """(u_revs, i_revs, u_masks, i_masks, u_ids, i_ids), ratings — the 7 tensors collate_fn yields
    (trainer/train_deepconn_pp.py:281-292)."""
    u_revs, u_masks = doc_batch(batch, doc_len, vocab, seed, uniform)
    i_revs, i_masks = doc_batch(batch, doc_len, vocab, seed + 7919, uniform)
    g = _gen(seed + 104729)
    u_ids = torch.randint(1, users, (batch,), generator=g)
    i_ids = torch.randint(1, items, (batch,), generator=g)
    ratings = torch.randint(1, 6, (batch,), generator=g).float()
    return (u_revs, i_revs, u_masks, i_masks, u_ids, i_ids), ratings


def narre_batch(batch: int, reviews: int, rev_len: int, vocab: int, users: int, items: int,
                seed: int = SEED_BASE, uniform: bool = False):
    """(u_text, i_text, u_masks, i_masks, u_id, i_id, reuid, reiid), ratings
    (trainer/train_narre.py:318-331).  Real-review count ~U{1..R}; the rest are all-zero with id 0."""
    g = _gen(seed)

    def side(offset: int, other_size: int):
        gg = _gen(seed + offset)
        ids = skewed_tokens(gg, (batch, reviews, rev_len), vocab, uniform)
        lens = torch.randint(min(10, rev_len), rev_len + 1, (batch, reviews), generator=gg)
        nrev = torch.randint(1, reviews + 1, (batch,), generator=gg)
        pos = torch.arange(rev_len).view(1, 1, -1)
        ridx = torch.arange(reviews).view(1, -1)
        real = ridx < nrev.unsqueeze(1)                                    # [B,R]
        keep = (pos < lens.unsqueeze(-1)) & real.unsqueeze(-1)
        ids = torch.where(keep, ids, torch.zeros((), dtype=torch.long))
        other = torch.randint(1, other_size, (batch, reviews), generator=gg)
        other = torch.where(real, other, torch.zeros((), dtype=torch.long))
        return ids, ids != 0, other

    u_text, u_masks, reuid = side(1, items)      # a user's reviews are keyed by the item they are about
    i_text, i_masks, reiid = side(2, users)
    u_id = torch.randint(1, users, (batch,), generator=g)
    i_id = torch.randint(1, items, (batch,), generator=g)
    ratings = torch.randint(1, 6, (batch,), generator=g).float()
    return (u_text, i_text, u_masks, i_masks, u_id, i_id, reuid, reiid), ratings


def dual_att_batch(batch: int, doc_len: int, vocab: int, seed: int = SEED_BASE, uniform: bool = False):
    """(u_docs, i_docs), ratings (trainer/train_dual_att.py collate)."""
    u_docs, _ = doc_batch(batch, doc_len, vocab, seed, uniform)
    i_docs, _ = doc_batch(batch, doc_len, vocab, seed + 7919, uniform)
    g = _gen(seed + 104729)
    ratings = torch.randint(1, 6, (batch,), generator=g).float()
    return (u_docs, i_docs), ratings


# ---------------------------------------------------------------------------------------
# Parameter sets keyed by the reference's state_dict names, with the reference's init rules.
# ---------------------------------------------------------------------------------------
def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def _conv_init(g, out_c, in_c, k):
    # nn.Conv1d default init: kaiming_uniform(a=sqrt(5)) → U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
    bound = 1.0 / (in_c * k) ** 0.5
    return _uniform(g, (out_c, in_c, k), bound), _uniform(g, (out_c,), bound)


def _embedding_init(g, rows, dim):
    w = torch.randn(rows, dim, generator=g)
    w[0] = 0.0                                   # nn.Embedding(padding_idx=0) zero-inits the pad row
    return w


def _head_params(g, p: Dict[str, torch.Tensor], users: int, items: int, hidden: int, latent: int):
    # LastFeat.reset_parameters (models/deepconn/layers.py:149-153), FM.reset_parameters (:181-186)
    for side, size in (("user", users), ("item", items)):
        p[f"{side}_feat.W"] = _uniform(g, (hidden, latent), 0.1)
        p[f"{side}_feat.b"] = torch.full((latent,), 0.1)
        p[f"{side}_feat.ebd.weight"] = _uniform(g, (size, latent), 0.1)
    p["fm.h"] = _uniform(g, (latent, 1), 0.1)
    p["fm.user_bias.weight"] = _uniform(g, (users, 1), 0.1)
    p["fm.item_bias.weight"] = _uniform(g, (items, 1), 0.1)
    p["fm.g_bias"] = torch.full((1,), 0.1)


def deepconn_params(users: int, items: int, vocab: int, emb: int, hidden: int, latent: int,
                    kernel_sizes=(3,), seed: int = 0) -> Dict[str, torch.Tensor]:
    g = _gen(seed)
    p: Dict[str, torch.Tensor] = {"word_embeddings.embedding.weight": _embedding_init(g, vocab, emb)}
    per = hidden // len(kernel_sizes)
    for i, k in enumerate(kernel_sizes):
        w, b = _conv_init(g, per, emb, k)
        p[f"ngram.feature_layer.0.list_of_conv1d.{i}.weight"] = w
        p[f"ngram.feature_layer.0.list_of_conv1d.{i}.bias"] = b
    _head_params(g, p, users, items, hidden, latent)
    return p


def deepconn_hier_params(users: int, items: int, vocab: int, emb: int, hidden: int, latent: int, seed: int = 0):
    """DeepCoNNpp(arch="HierPooling"): no conv weights; a Linear(emb → hidden) projection when the sizes differ
    (models/deepconn/layers.py:72-76)."""
    g = _gen(seed)
    p: Dict[str, torch.Tensor] = {"word_embeddings.embedding.weight": _embedding_init(g, vocab, emb)}
    if emb != hidden:
        b = 1.0 / emb ** 0.5
        p["ngram.feature_layer.0.proj_layer.weight"] = _uniform(g, (hidden, emb), b)
        p["ngram.feature_layer.0.proj_layer.bias"] = _uniform(g, (hidden,), b)
    _head_params(g, p, users, items, hidden, latent)
    return p


def narre_params(users: int, items: int, vocab: int, emb: int, hidden: int, att: int, latent: int,
                 kernel_sizes=(3,), seed: int = 0) -> Dict[str, torch.Tensor]:
    g = _gen(seed)
    p = deepconn_params(users, items, vocab, emb, hidden, latent, kernel_sizes, seed)
    # LinearAttention.__init__ (models/narre/narre.py:30-36); user_att's id table is sized item_size
    for side, size in (("user", items), ("item", users)):
        p[f"{side}_att.W_rv"] = _uniform(g, (hidden, att), 0.1)
        p[f"{side}_att.W_id"] = _uniform(g, (att, att), 0.1)
        p[f"{side}_att.h"] = _uniform(g, (att, 1), 0.1)
        p[f"{side}_att.b_1"] = torch.full((att,), 0.1)
        p[f"{side}_att.b_2"] = torch.full((1,), 0.1)
        p[f"{side}_att.ebd_vals.weight"] = _embedding_init(g, size, att)
    return p


def dual_att_params(vocab: int, doc_len: int, l_window: int = 5, l_out: int = 200, g_out: int = 100,
                    emb: int = 100, hidden1: int = 500, hidden2: int = 50, seed: int = 0):
    g = _gen(seed)
    p: Dict[str, torch.Tensor] = {"word_embeddings.embedding.weight": _embedding_init(g, vocab, emb)}
    for s in ("u", "i"):
        p[f"{s}_local_atten.attn.0.weight"], p[f"{s}_local_atten.attn.0.bias"] = _conv_init(g, 1, emb, l_window)
        p[f"{s}_local_atten.conv.0.weight"], p[f"{s}_local_atten.conv.0.bias"] = _conv_init(g, l_out, emb, 1)
        p[f"{s}_global_atten.attn.0.weight"], p[f"{s}_global_atten.attn.0.bias"] = _conv_init(g, 1, emb, doc_len)
        for c, k in ((1, 2), (2, 3), (3, 4)):
            w, b = _conv_init(g, g_out, emb, k)
            p[f"{s}_global_atten.conv{c}.0.weight"], p[f"{s}_global_atten.conv{c}.0.bias"] = w, b
    fc_in = l_out + 3 * g_out
    b0 = 1.0 / fc_in ** 0.5
    p["fc.0.weight"], p["fc.0.bias"] = _uniform(g, (hidden1, fc_in), b0), _uniform(g, (hidden1,), b0)
    b3 = 1.0 / hidden1 ** 0.5
    p["fc.3.weight"], p["fc.3.bias"] = _uniform(g, (hidden2, hidden1), b3), _uniform(g, (hidden2,), b3)
    return p


def simple_siamese_params(users: int, items: int, vocab: int, emb: int, latent: int, use_ui_bias: bool = True,
                          latent_transform: bool = False, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Parameter set of the reference's SimpleSiamese (models/simple_siamese/simple_siamese.py:9-36), keyed by its state_dict."""
    g = _gen(seed)
    p: Dict[str, torch.Tensor] = {"word_embedding.embedding.weight": _embedding_init(g, vocab, emb)}
    feat = latent if latent_transform else emb
    if latent_transform:
        b = 1.0 / emb ** 0.5
        p["latent_transform_layer.0.weight"], p["latent_transform_layer.0.bias"] = _uniform(g, (latent, emb), b), _uniform(g, (latent,), b)
    for side, size in (("user", users), ("item", items)):
        p[f"{side}_last_feat_layer.W"] = _uniform(g, (feat, latent), 0.1)
        p[f"{side}_last_feat_layer.b"] = torch.full((latent,), 0.1)
        p[f"{side}_last_feat_layer.ebd.weight"] = _uniform(g, (size, latent), 0.1)
    b = 1.0 / feat ** 0.5
    p["review_att_layer.proj_layer.0.weight"], p["review_att_layer.proj_layer.0.bias"] = _uniform(g, (latent, feat), b), _uniform(g, (latent,), b)
    p["review_att_layer.inner_product.weight"] = _uniform(g, (1, latent), 1.0 / latent ** 0.5)
    p["fm.h"] = _uniform(g, (latent, 1), 0.1)
    p["fm.g_bias"] = torch.full((1,), 4.0)
    if use_ui_bias:
        p["fm.user_bias.weight"] = _uniform(g, (users, 1), 0.1)
        p["fm.item_bias.weight"] = _uniform(g, (items, 1), 0.1)
    return p


def simple_siamese_batch(batch: int, reviews: int, rev_len: int, vocab: int, users: int, items: int, seed: int = SEED_BASE):
    """(u_revs, i_revs, u_word_masks, i_word_masks, u_rev_masks, i_rev_masks, u_ids, i_ids), ratings
    (trainer/train_simple_siamese.py collate): NARRE-shaped review tensors plus review-level masks (a review is real when it
    has at least one token, models/simple_siamese/utils.py get_rev_mask)."""
    (u_text, i_text, u_m, i_m, u_id, i_id, _, _), ratings = narre_batch(batch, reviews, rev_len, vocab, users, items, seed)
    return (u_text, i_text, u_m, i_m, u_m.any(dim=-1), i_m.any(dim=-1), u_id, i_id), ratings
