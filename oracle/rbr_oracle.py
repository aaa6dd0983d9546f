"""CPU oracle for the review-encoder hot path (TEST INFRASTRUCTURE — not product code).

This file restates, in plain functional PyTorch on the CPU, the arithmetic that the
reference repo H263/review-based-recommender performs on its hot path.  It is only ever
imported by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s `cpu_baseline` /
`--impl reference` legs, and only as the checker or the timed CPU baseline — the product
package `review-based-recommender_b200/` never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors for this path (SURVEY.md §4,
§8c).  The oracle is pinned instead against outputs of the unmodified reference modules run
in the build container: `tests/golden/make_golden.py` imports `/root/reference/models/*`
and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function here
against those fixtures.

The arithmetic library is torch (the reference's own arithmetic IS torch: aten::embedding,
conv1d, max_pool1d, matmul — SURVEY.md §8c "third-party arithmetic"); what is restated is
the algorithm: the convolution is written out as a sum of k shifted matmuls, the pooling
as an explicit first-arg-max, the attention and FM as explicit formulas.  Parameters are
passed as a flat dict keyed by the reference's `state_dict()` names so that fixtures,
oracle and product all share one naming.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
import time
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# a1: get_mask  (utils.py:30-42)
# --------------------------------------------------------------------------------------
def get_mask(ids: Tensor, padding_idx: int = 0) -> Tensor:
    """mask = (ids != padding_idx), bool, same shape.  utils.py:30-42."""
    return ids != padding_idx


# --------------------------------------------------------------------------------------
# a2: WordEmbedding  (models/deepconn/layers.py:9-24; narre/narre.py:9-24; dual_att/layers.py:8-23)
# --------------------------------------------------------------------------------------
def embedding_gather(table: Tensor, ids: Tensor) -> Tensor:
    """out[..., :] = table[ids[...], :]  — nn.Embedding forward, deepconn/layers.py:15,23.

    padding_idx only affects the gradient (row padding_idx receives zero), not the forward:
    row 0 is returned as stored (it is zero only because nn.Embedding zero-inits it).
    """
    flat = ids.reshape(-1)
    out = table.index_select(0, flat)
    return out.reshape(*ids.shape, table.shape[1])


def embedding_dense_grad(ids: Tensor, grad_rows: Tensor, num_rows: int, padding_idx: Optional[int] = 0) -> Tensor:
    """Dense [V,E] gradient of embedding_gather: scatter-add of row gradients, padding row skipped.

    Mirrors aten::embedding_dense_backward as invoked by autograd for deepconn/layers.py:23.
    Accumulates in float64 then rounds, so it is an order-independent reference sum.
    """
    flat = ids.reshape(-1)
    g = grad_rows.reshape(flat.shape[0], -1).to(torch.float64)
    out = torch.zeros(num_rows, g.shape[1], dtype=torch.float64)
    if padding_idx is not None:
        keep = flat != padding_idx
        flat, g = flat[keep], g[keep]
    out.index_add_(0, flat, g)
    return out.to(grad_rows.dtype)


# --------------------------------------------------------------------------------------
# a3: masked_tensor  (models/deepconn/utils.py:49-61)
# --------------------------------------------------------------------------------------
def mask_rows(x: Tensor, mask: Tensor) -> Tensor:
    """x[~mask] = 0 (out of place; blocks the gradient at masked rows).  deepconn/utils.py:58-60."""
    assert x.shape[:-1] == mask.shape
    return torch.where(mask.unsqueeze(-1), x, torch.zeros((), dtype=x.dtype))


# --------------------------------------------------------------------------------------
# a4/a5: MyConv1d + ReLU + MaxPool1d(seq_len)   (models/deepconn/layers.py:26-60, 100-136)
# --------------------------------------------------------------------------------------
def conv1d_same(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """Y[n,t,h] = bias[h] + sum_{j<k} sum_e W[h,e,j] * X[n, t+j-p, e],  p=(k-1)//2, zero padded.

    x [N,L,E] (token-major — the reference transposes to [N,E,L] first, layers.py:132),
    weight [H,E,k] (nn.Conv1d layout, layers.py:44), returns [N,L,H].
    Written as k shifted matmuls — the definition, not a call to conv1d.
    """
    n, l, e = x.shape
    h, e2, k = weight.shape
    assert e == e2 and k % 2 == 1          # layers.py:39 asserts odd kernel sizes
    p = (k - 1) // 2
    xp = torch.zeros(n, l + 2 * p, e, dtype=x.dtype, device=x.device)
    xp[:, p:p + l] = x
    y = bias.view(1, 1, h).expand(n, l, h).clone()
    for j in range(k):
        y = y + xp[:, j:j + l] @ weight[:, :, j].t()
    return y


def conv1d_valid(x: Tensor, weight: Tensor, bias: Tensor, pad: int = 0) -> Tensor:
    """General Conv1d (stride 1) with symmetric zero padding `pad`; output length L+2*pad-k+1.

    Used for the D-ATT convs (models/dual_att/layers.py:34-40, 65-79).  x [N,L,E] → [N,Lout,H].
    """
    n, l, e = x.shape
    h, _, k = weight.shape
    xp = torch.zeros(n, l + 2 * pad, e, dtype=x.dtype, device=x.device)
    xp[:, pad:pad + l] = x
    lout = l + 2 * pad - k + 1
    y = bias.view(1, 1, h).expand(n, lout, h).clone()
    for j in range(k):
        y = y + xp[:, j:j + lout] @ weight[:, :, j].t()
    return y


def first_argmax_pool(y: Tensor) -> Tuple[Tensor, Tensor]:
    """Max over dim 1 (time) of y [N,L,H] → (values [N,H], first arg-max [N,H]).

    nn.MaxPool1d(seq_len) (layers.py:109) pools over ALL positions, padded ones included, and
    its backward routes to the first maximal index (probed).
    """
    vals = y.max(dim=1).values
    idx = (y == vals.unsqueeze(1)).to(torch.uint8).argmax(dim=1)   # argmax returns the first occurrence
    return vals, idx


def ngram_feat(x: Tensor, mask: Optional[Tensor], conv_weights: Sequence[Tensor], conv_biases: Sequence[Tensor],
               return_argmax: bool = False, argmax_override: Optional[Tensor] = None):
    """NgramFeat.forward with arch="CNN" (layers.py:123-136): mask → conv(s) → ReLU → max over time.

    Several kernel sizes are concatenated on the filter dim (MyConv1d.forward, layers.py:54-58).
    Returns [N,H] (the reference returns [N,H,1]; callers `.view(bz, H)` it, deepconn.py:46).
    `argmax_override` [N,H] (tests only): pool at these positions instead of the oracle's own first arg-max — used to compare
    gradients under IDENTICAL routing when two positions tie to within rounding (each test first checks that every
    overridden position attains the oracle's max up to summation-order noise).
    """
    if mask is not None:
        x = mask_rows(x, mask)
    feats, args = [], []
    col = 0
    for w, b in zip(conv_weights, conv_biases):
        y = torch.relu(conv1d_same(x, w, b))
        if argmax_override is not None:
            idx = argmax_override[:, col:col + w.shape[0]].to(torch.long)
        else:
            idx = (y == y.max(dim=1, keepdim=True).values).to(torch.uint8).argmax(dim=1)
        col += w.shape[0]
        v = torch.gather(y, 1, idx.unsqueeze(1)).squeeze(1)       # differentiable, first-arg-max routed
        feats.append(v)
        args.append(idx)
    feat = torch.cat(feats, dim=1)
    if return_argmax:
        return feat, torch.cat(args, dim=1)
    return feat


def hier_pooling(x: Tensor, mask: Optional[Tensor], kernel_size: int, proj_w: Optional[Tensor] = None,
                 proj_b: Optional[Tensor] = None) -> Tensor:
    """NgramFeat.forward with arch="HierPooling" (layers.py:110-114, 123-136; HierPooling.forward :81-98):
    mask → avg_pool1d(kernel, stride 1) over time → max over time (first arg-max) → [Linear] → ReLU.  x [N,L,E] → [N,out]."""
    if mask is not None:
        x = mask_rows(x, mask)
    n, l, e = x.shape
    lout = l - kernel_size + 1
    y = torch.zeros(n, lout, e, dtype=x.dtype, device=x.device)
    for j in range(kernel_size):
        y = y + x[:, j:j + lout]
    y = y / kernel_size
    idx = (y == y.max(dim=1, keepdim=True).values).to(torch.uint8).argmax(dim=1)
    v = torch.gather(y, 1, idx.unsqueeze(1)).squeeze(1)
    if proj_w is not None:
        v = v @ proj_w.t() + proj_b
    return torch.relu(v)


# --------------------------------------------------------------------------------------
# a6: LinearAttention  (models/narre/narre.py:26-64)
# --------------------------------------------------------------------------------------
def linear_attention(feat: Tensor, other_id: Tensor, W_rv: Tensor, W_id: Tensor, h: Tensor, b_1: Tensor,
                     b_2: Tensor, ebd_vals: Tensor) -> Tuple[Tensor, Tensor]:
    """NARRE review-level attention, narre.py:40-64 (dropout omitted: parity runs use p=0 / eval()).

    logit = relu(feat@W_rv + ebd_vals[other_id]@W_id + b_1) @ h + b_2        (narre.py:53-55)
    score = exp(logit) / (sum_R exp(logit) + 1e-8)   — no max-subtraction, no review mask (narre.py:58)
    out   = sum_R score * feat                                                 (narre.py:60)
    Returns (out [B,H], score [B,R,1]).
    """
    e = embedding_gather(ebd_vals, other_id)                       # [B,R,A]
    hid = torch.relu(feat @ W_rv + e @ W_id + b_1)
    logit = hid @ h + b_2                                          # [B,R,1]
    ex = logit.exp()
    score = ex / (ex.sum(dim=1, keepdim=True) + 1e-8)
    out = (score * feat).sum(dim=1)
    return out, score


# --------------------------------------------------------------------------------------
# a7/a8/a9: LastFeat, FM, MSELoss
# --------------------------------------------------------------------------------------
def last_feat(text_feat: Tensor, my_id: Tensor, W: Tensor, b: Tensor, ebd: Tensor) -> Tensor:
    """text_feat @ W + b + ebd[my_id]   (models/deepconn/layers.py:163; narre/narre.py:91)."""
    return text_feat @ W + b + embedding_gather(ebd, my_id)


def fm_head(u_feat: Tensor, i_feat: Tensor, u_id: Tensor, i_id: Tensor, h: Tensor, g_bias: Tensor,
            user_bias: Tensor, item_bias: Tensor, drop_mask: Optional[Tensor] = None) -> Tensor:
    """relu(u*i) [→ dropout] @ h + user_bias[u_id] + item_bias[i_id] + g_bias → [B,1]
    (models/deepconn/layers.py:200-207).  `drop_mask` (already scaled by 1/(1-p)) stands in for
    nn.Dropout in train mode; None = eval / p=0."""
    fm = torch.relu(u_feat * i_feat)
    if drop_mask is not None:
        fm = fm * drop_mask
    return fm @ h + embedding_gather(user_bias, u_id) + embedding_gather(item_bias, i_id) + g_bias


def mse_loss(pred: Tensor, target: Tensor) -> Tensor:
    """nn.MSELoss() default reduction='mean' (trainer/train_deepconn_pp.py:140,164)."""
    d = pred - target
    return (d * d).mean()


# --------------------------------------------------------------------------------------
# a10: DeepCoNNpp.forward  (models/deepconn/deepconn.py:28-53)
# --------------------------------------------------------------------------------------
def _conv_params(p: Params, prefix: str) -> Tuple[List[Tensor], List[Tensor]]:
    ws, bs, i = [], [], 0
    while f"{prefix}.{i}.weight" in p:
        ws.append(p[f"{prefix}.{i}.weight"])
        bs.append(p[f"{prefix}.{i}.bias"])
        i += 1
    return ws, bs


def deepconn_forward(p: Params, u_revs: Tensor, i_revs: Tensor, u_masks: Tensor, i_masks: Tensor, u_ids: Tensor,
                     i_ids: Tensor, return_aux: bool = False, fm_drop_mask: Optional[Tensor] = None,
                     argmax_override: Optional[Tuple[Tensor, Tensor]] = None, hier_kernel: Optional[int] = None):
    """DeepCoNNpp.forward, deepconn.py:28-53.  Embedding + conv weights are SHARED between the
    user and the item side (deepconn.py:20-22, 43-47)."""
    table = p["word_embeddings.embedding.weight"]
    ws, bs = _conv_params(p, "ngram.feature_layer.0.list_of_conv1d")
    u_x = embedding_gather(table, u_revs)                                        # deepconn.py:43
    i_x = embedding_gather(table, i_revs)                                        # deepconn.py:44
    ao = argmax_override or (None, None)
    if hier_kernel is not None:                                                  # arch="HierPooling" (layers.py:110-114)
        pw, pb = p.get("ngram.feature_layer.0.proj_layer.weight"), p.get("ngram.feature_layer.0.proj_layer.bias")
        u_rev, i_rev = hier_pooling(u_x, u_masks, hier_kernel, pw, pb), hier_pooling(i_x, i_masks, hier_kernel, pw, pb)
        u_arg = i_arg = None
    else:
        u_rev, u_arg = ngram_feat(u_x, u_masks, ws, bs, return_argmax=True, argmax_override=ao[0])   # deepconn.py:46
        i_rev, i_arg = ngram_feat(i_x, i_masks, ws, bs, return_argmax=True, argmax_override=ao[1])   # deepconn.py:47
    u_f = last_feat(u_rev, u_ids, p["user_feat.W"], p["user_feat.b"], p["user_feat.ebd.weight"])   # :48
    i_f = last_feat(i_rev, i_ids, p["item_feat.W"], p["item_feat.b"], p["item_feat.ebd.weight"])   # :49
    pred = fm_head(u_f, i_f, u_ids, i_ids, p["fm.h"], p["fm.g_bias"], p["fm.user_bias.weight"],
                   p["fm.item_bias.weight"], drop_mask=fm_drop_mask)              # deepconn.py:51
    pred = pred.view(-1)                                                          # deepconn.py:53
    if return_aux:
        return pred, {"u_rev_feats": u_rev, "i_rev_feats": i_rev, "u_argmax": u_arg, "i_argmax": i_arg,
                      "u_feats": u_f, "i_feats": i_f}
    return pred


# --------------------------------------------------------------------------------------
# a11: NARRE.forward  (models/narre/narre.py:165-192)
# --------------------------------------------------------------------------------------
def narre_forward(p: Params, u_text: Tensor, i_text: Tensor, u_masks: Tensor, i_masks: Tensor, u_id: Tensor,
                  i_id: Tensor, reuid: Tensor, reiid: Tensor, return_aux: bool = False, fm_drop_mask: Optional[Tensor] = None,
                  argmax_override: Optional[Tuple[Tensor, Tensor]] = None):
    """NARRE.forward, narre.py:165-192.  [B,R,T] tokens are encoded as B*R independent docs
    (narre.py:170-176); the review-level masks computed at narre.py:180-181 are never used."""
    table = p["word_embeddings.embedding.weight"]
    ws, bs = _conv_params(p, "ngram.feature_layer.0.list_of_conv1d")
    b, r, t = u_text.shape
    hdim = sum(w.shape[0] for w in ws)
    u_x = embedding_gather(table, u_text).view(b * r, t, -1)
    i_x = embedding_gather(table, i_text).view(b * r, t, -1)
    ao = argmax_override or (None, None)
    u_rf = ngram_feat(u_x, u_masks.reshape(b * r, t), ws, bs, argmax_override=ao[0]).view(b, r, hdim)
    i_rf = ngram_feat(i_x, i_masks.reshape(b * r, t), ws, bs, argmax_override=ao[1]).view(b, r, hdim)
    ua = {k: p[f"user_att.{k}"] for k in ("W_rv", "W_id", "h", "b_1", "b_2")}
    ia = {k: p[f"item_att.{k}"] for k in ("W_rv", "W_id", "h", "b_1", "b_2")}
    u_feat, u_sc = linear_attention(u_rf, reuid, ebd_vals=p["user_att.ebd_vals.weight"], **ua)   # narre.py:184
    i_feat, i_sc = linear_attention(i_rf, reiid, ebd_vals=p["item_att.ebd_vals.weight"], **ia)   # narre.py:185
    u_f = last_feat(u_feat, u_id, p["user_feat.W"], p["user_feat.b"], p["user_feat.ebd.weight"])  # :187
    i_f = last_feat(i_feat, i_id, p["item_feat.W"], p["item_feat.b"], p["item_feat.ebd.weight"])  # :188
    pred = fm_head(u_f, i_f, u_id, i_id, p["fm.h"], p["fm.g_bias"], p["fm.user_bias.weight"],
                   p["fm.item_bias.weight"], drop_mask=fm_drop_mask).view(-1)                     # :190-192
    if return_aux:
        return pred, u_sc, i_sc, {"u_rev_feats": u_rf, "i_rev_feats": i_rf, "u_att_out": u_feat,
                                  "i_att_out": i_feat}
    return pred, u_sc, i_sc


# --------------------------------------------------------------------------------------
# a12: D-ATT  (models/dual_att/layers.py:25-89, dual_att.py:37-61)
# --------------------------------------------------------------------------------------
def local_attention(x: Tensor, attn_w: Tensor, attn_b: Tensor, conv_w: Tensor, conv_b: Tensor) -> Tensor:
    """LocalAttention.forward (dual_att/layers.py:43-53): gate = sigmoid(Conv1d(E→1,k=win,pad)) [N,L,1];
    out = max_t tanh(Conv1d(E→out,k=1)(gate*x)) → [N,out]."""
    win = attn_w.shape[2]
    gate = torch.sigmoid(conv1d_valid(x, attn_w, attn_b, pad=(win - 1) // 2))      # [N,L,1]
    y = torch.tanh(conv1d_valid(gate * x, conv_w, conv_b))                         # [N,L,out]
    return y.max(dim=1).values


def global_attention(x: Tensor, attn_w: Tensor, attn_b: Tensor, convs: Sequence[Tuple[Tensor, Tensor]]) -> List[Tensor]:
    """GlobalAttention.forward (dual_att/layers.py:81-89): gate = sigmoid(Conv1d(E→1,k=L)) is ONE scalar per
    doc; out_k = max_t tanh(Conv1d(E→out,k∈{2,3,4}, no padding)(gate*x))."""
    gate = torch.sigmoid(conv1d_valid(x, attn_w, attn_b))                          # [N,1,1]
    gx = gate * x
    return [torch.tanh(conv1d_valid(gx, w, b)).max(dim=1).values for w, b in convs]


def dual_att_forward(p: Params, u_docs: Tensor, i_docs: Tensor, return_aux: bool = False):
    """DualAtt.forward (dual_att.py:37-61), dropout omitted (eval / p=0).  The FC stack is shared
    between the two sides (dual_att.py:31-35, 51, 57)."""
    table = p["word_embeddings.embedding.weight"]

    def side(ids: Tensor, s: str) -> Tensor:
        x = embedding_gather(table, ids)
        loc = local_attention(x, p[f"{s}_local_atten.attn.0.weight"], p[f"{s}_local_atten.attn.0.bias"],
                              p[f"{s}_local_atten.conv.0.weight"], p[f"{s}_local_atten.conv.0.bias"])
        glo = global_attention(x, p[f"{s}_global_atten.attn.0.weight"], p[f"{s}_global_atten.attn.0.bias"],
                               [(p[f"{s}_global_atten.conv{c}.0.weight"], p[f"{s}_global_atten.conv{c}.0.bias"])
                                for c in (1, 2, 3)])
        return torch.cat([loc] + glo, dim=1)                                       # dual_att.py:49

    def fc(f: Tensor) -> Tensor:
        hid = torch.relu(f @ p["fc.0.weight"].t() + p["fc.0.bias"])
        return hid @ p["fc.3.weight"].t() + p["fc.3.bias"]

    u_cat, i_cat = side(u_docs, "u"), side(i_docs, "i")
    u_f, i_f = fc(u_cat), fc(i_cat)
    rating = (u_f * i_f).sum(dim=1).view(-1)                                       # dual_att.py:59-61
    if return_aux:
        return rating, {"u_cat": u_cat, "i_cat": i_cat, "u_feat": u_f, "i_feat": i_f}
    return rating


# --------------------------------------------------------------------------------------
# SimpleSiamese  (models/simple_siamese/simple_siamese.py:38-88, layers.py:90-110, 171-197)
# --------------------------------------------------------------------------------------
def masked_avg_pool(x: Tensor, mask: Tensor) -> Tensor:
    """MaskedAvgPooling1d (simple_siamese/layers.py:94-110) in token-major layout: x [N,T,E], mask [N,T] → [N,E]."""
    m = mask.to(x.dtype).unsqueeze(-1)
    return (x * m).sum(dim=1) / (m.sum(dim=1) + 1e-8)


def simple_siamese_forward(p: Params, u_revs: Tensor, i_revs: Tensor, u_wm: Tensor, i_wm: Tensor, u_rm: Tensor, i_rm: Tensor,
                           u_ids: Tensor, i_ids: Tensor) -> Tensor:
    """SimpleSiamese.forward with every dropout at 0 / eval (simple_siamese.py:55-88)."""
    table = p["word_embedding.embedding.weight"]
    b, r, t = u_revs.shape

    def side(revs, wm, rm):
        pooled = masked_avg_pool(embedding_gather(table, revs).view(-1, t, table.shape[1]), wm.reshape(-1, t)).view(revs.shape[0], revs.shape[1], -1)
        if "latent_transform_layer.0.weight" in p:
            pooled = torch.tanh(pooled @ p["latent_transform_layer.0.weight"].t() + p["latent_transform_layer.0.bias"])
        hid = torch.tanh(pooled @ p["review_att_layer.proj_layer.0.weight"].t() + p["review_att_layer.proj_layer.0.bias"])
        logits = hid @ p["review_att_layer.inner_product.weight"].t()
        logits = torch.where(rm.unsqueeze(2), logits, torch.full((), -1e8, dtype=logits.dtype, device=logits.device))
        return (torch.softmax(logits, dim=1) * pooled).sum(dim=1)

    u_f = last_feat(side(u_revs, u_wm, u_rm), u_ids, p["user_last_feat_layer.W"], p["user_last_feat_layer.b"], p["user_last_feat_layer.ebd.weight"])
    i_f = last_feat(side(i_revs, i_wm, i_rm), i_ids, p["item_last_feat_layer.W"], p["item_last_feat_layer.b"], p["item_last_feat_layer.ebd.weight"])
    if "fm.user_bias.weight" in p:
        return fm_head(u_f, i_f, u_ids, i_ids, p["fm.h"], p["fm.g_bias"], p["fm.user_bias.weight"], p["fm.item_bias.weight"]).view(-1)
    return (torch.relu(u_f * i_f) @ p["fm.h"] + p["fm.g_bias"]).view(-1)


# --------------------------------------------------------------------------------------
# loss + all parameter gradients (what `loss.backward()` leaves in `.grad`,
# trainer/train_deepconn_pp.py:161-165)
# --------------------------------------------------------------------------------------
_PADDED_TABLES = ("word_embeddings.embedding.weight", "word_embedding.embedding.weight", "user_last_feat_layer.ebd.weight",
                  "item_last_feat_layer.ebd.weight", "user_feat.ebd.weight", "item_feat.ebd.weight",
                  "fm.user_bias.weight", "fm.item_bias.weight", "user_att.ebd_vals.weight",
                  "item_att.ebd_vals.weight")


def loss_and_grads(model: str, p: Params, batch: Sequence[Tensor], ratings: Tensor, **fwd_kw):
    """Run forward + MSE + backward with torch autograd over the functional forward above.

    Returns (pred, loss, grads) with grads keyed like `p`.  Rows `padding_idx=0` of every
    nn.Embedding table receive zero gradient (nn.Embedding(padding_idx=0) semantics).
    fwd_kw: `fm_drop_mask`, `argmax_override` of deepconn_forward / narre_forward.
    """
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    if model == "deepconn":
        pred = deepconn_forward(leaves, *batch, **fwd_kw)
    elif model == "narre":
        pred = narre_forward(leaves, *batch, **fwd_kw)[0]
    elif model == "dual_att":
        pred = dual_att_forward(leaves, *batch)
    elif model == "simple_siamese":
        pred = simple_siamese_forward(leaves, *batch)
    else:
        raise ValueError(model)
    loss = mse_loss(pred, ratings)
    loss.backward()
    grads = {}
    for k, v in leaves.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        if k in _PADDED_TABLES:
            g = g.clone()
            g[0] = 0
        grads[k] = g
    return pred.detach(), loss.detach(), grads


# --------------------------------------------------------------------------------------
# ReLU decision margins (test aid for full-size comparisons)
# --------------------------------------------------------------------------------------
def relu_margins(model: str, p: Params, batch: Sequence[Tensor], chunk: int = 512) -> Tensor:
    """[B] per sample: the smallest |pre-activation| over every ReLU of the path that the sample owns — the pooled conv
    features `max_t conv(x) + bias` (layers.py:108-109), NARRE's attention hidden units (narre.py:55) and the FM interaction
    `u * i` (layers.py:200).  A ReLU whose input is within rounding of 0 may legitimately open on one implementation and
    close on another (fp32 summation order), which moves that unit's whole gradient contribution; full-size parity tests
    replace samples whose margin is below a threshold far above fp32 noise, then compare at the strict tolerance."""
    assert model in ("deepconn", "narre")
    table = p["word_embeddings.embedding.weight"]
    ws, bs = _conv_params(p, "ngram.feature_layer.0.list_of_conv1d")

    def pooled_pre(ids: Tensor, mask: Tensor) -> Tensor:                 # [N, L] → [N, H] (before the ReLU)
        x = mask_rows(embedding_gather(table, ids), mask)
        return torch.cat([conv1d_same(x, w, b).max(dim=1).values for w, b in zip(ws, bs)], dim=1)

    bias_all = torch.cat(list(bs))

    def conv_margin(pre: Tensor) -> Tensor:
        """|pre| per unit, except units that pooled an all-zero window (fully padded document): their value is EXACTLY the
        bias in every implementation (0 + bias), so their ReLU decision cannot differ however small the bias is."""
        m = pre.abs()
        return torch.where(pre == bias_all.view(*([1] * (pre.dim() - 1)), -1), torch.full_like(m, float("inf")), m)

    n = batch[0].shape[0]
    out = []
    with torch.no_grad():
        for lo in range(0, n, chunk):
            sl = slice(lo, lo + chunk)
            if model == "deepconn":
                u_revs, i_revs, u_m, i_m, u_ids, i_ids = [t[sl] for t in batch]
                u_pre, i_pre = pooled_pre(u_revs, u_m), pooled_pre(i_revs, i_m)
                margin = torch.minimum(conv_margin(u_pre).min(dim=1).values, conv_margin(i_pre).min(dim=1).values)
                u_txt, i_txt = torch.relu(u_pre), torch.relu(i_pre)
            else:
                u_text, i_text, u_m, i_m, u_ids, i_ids, reuid, reiid = [t[sl] for t in batch]
                b, r, t = u_text.shape
                u_pre = pooled_pre(u_text.reshape(b * r, t), u_m.reshape(b * r, t)).view(b, r, -1)
                i_pre = pooled_pre(i_text.reshape(b * r, t), i_m.reshape(b * r, t)).view(b, r, -1)
                margin = torch.minimum(conv_margin(u_pre).flatten(1).min(dim=1).values, conv_margin(i_pre).flatten(1).min(dim=1).values)
                feats = []
                for side, pre, oid in (("user", u_pre, reuid), ("item", i_pre, reiid)):
                    f = torch.relu(pre)
                    e = embedding_gather(p[f"{side}_att.ebd_vals.weight"], oid)
                    hid = f @ p[f"{side}_att.W_rv"] + e @ p[f"{side}_att.W_id"] + p[f"{side}_att.b_1"]
                    margin = torch.minimum(margin, hid.abs().flatten(1).min(dim=1).values)
                    out_s, _ = linear_attention(f, oid, p[f"{side}_att.W_rv"], p[f"{side}_att.W_id"], p[f"{side}_att.h"],
                                                p[f"{side}_att.b_1"], p[f"{side}_att.b_2"], p[f"{side}_att.ebd_vals.weight"])
                    feats.append(out_s)
                u_txt, i_txt = feats
            u_f = last_feat(u_txt, u_ids, p["user_feat.W"], p["user_feat.b"], p["user_feat.ebd.weight"])
            i_f = last_feat(i_txt, i_ids, p["item_feat.W"], p["item_feat.b"], p["item_feat.ebd.weight"])
            margin = torch.minimum(margin, (u_f * i_f).abs().min(dim=1).values)
            out.append(margin)
    return torch.cat(out)


# --------------------------------------------------------------------------------------
# Timed CPU baseline (bench.py `cpu_baseline` / `--impl reference`)
# --------------------------------------------------------------------------------------
def time_fwd_bwd(model: str, p: Params, batch: Sequence[Tensor], ratings: Tensor, steps: int = 3,
                 warmup: int = 1) -> float:
    """Seconds per fwd+loss+bwd step of the CPU oracle on the host cores (torch threads)."""
    for _ in range(warmup):
        loss_and_grads(model, p, batch, ratings)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss_and_grads(model, p, batch, ratings)
    return (time.perf_counter() - t0) / max(steps, 1)
